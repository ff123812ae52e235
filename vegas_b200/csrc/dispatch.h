// dispatch.h -- launchers of the templated engine kernel, one translation unit per integrand
// family so the instantiations compile in parallel.
#pragma once
#include "engine.cuh"
#include "integrands.cuh"

// Each launcher picks the instantiation for the padded dimension D >= dim, sizes the persistent
// grid from the occupancy calculator, launches on `stream` and returns the grid size (>0) or a
// negative error.  `functor` points to the host copy of the family's functor struct.
struct LaunchCfg {
    int sm_count;
    int cap;          // staged samples per tile
    size_t smem;      // dynamic shared memory bytes
    int blocks_per_sm_out;
};

int launch_fused_poly(const EngineP& p, const void* functor, LaunchCfg& cfg, int max_grid, cudaStream_t st);
int launch_fused_gaussmix(const EngineP& p, const void* functor, LaunchCfg& cfg, int max_grid, cudaStream_t st);
int launch_fused_ridge(const EngineP& p, const void* functor, LaunchCfg& cfg, int max_grid, cudaStream_t st);
int launch_fused_genz(const EngineP& p, const void* functor, LaunchCfg& cfg, int max_grid, cudaStream_t st);
int launch_fused_pathint(const EngineP& p, const void* functor, int nx0, LaunchCfg& cfg, int max_grid, cudaStream_t st);
int launch_buffer(const EngineP& p, int nf, LaunchCfg& cfg, int max_grid, cudaStream_t st);

// dry-run variants: only compute the grid size (blocks/SM * SMs) so the caller can size scratch
// before the real launch.  Implemented by passing st == (cudaStream_t)-1.
#define VB_DRYRUN ((cudaStream_t)(intptr_t)-1)

template <class Src>
static int launch_engine(const EngineP& p, const Src& src, LaunchCfg& cfg, int max_grid, cudaStream_t st)
{
    auto kern = k_engine<Src>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
    if (e != cudaSuccess) return -(int)e - 1000;
    int bps = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, VB_ENT, cfg.smem);
    if (e != cudaSuccess) return -(int)e - 1000;
    if (bps < 1) bps = 1;
    cfg.blocks_per_sm_out = bps;
    int grid = bps * cfg.sm_count;
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    if (st == VB_DRYRUN) return grid;
    kern<<<grid, VB_ENT, cfg.smem, st>>>(p, src);
    e = cudaGetLastError();
    if (e != cudaSuccess) return -(int)e - 1000;
    return grid;
}

// padded-dimension dispatch
#define VB_DISPATCH_D(F, fobj, LIST_MACRO)                                                     \
    do {                                                                                       \
        const int dim_ = p.map.dim;                                                            \
        LIST_MACRO(F, fobj)                                                                    \
        return -22;                                                                            \
    } while (0)
#define VB_CASE_D(F, fobj, DD)                                                                 \
    if (dim_ <= DD) { FusedSrc<F, DD> s_{fobj}; return launch_engine(p, s_, cfg, max_grid, st); }
