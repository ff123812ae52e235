// dispatch.h -- launchers of the templated engine kernel, one translation unit per integrand
// family so the instantiations compile in parallel.
#pragma once
#include <stdlib.h>

#include <map>

#include "engine.cuh"
#include "integrands.cuh"

// Each launcher picks the instantiation for the padded dimension D >= dim (and the light or heavy
// geometry), sizes shared memory, histogram windows and the persistent grid for it, launches on
// `stream` and returns the grid size (>0) or a negative error.  `functor` points to the host copy
// of the family's functor struct.
struct LaunchCfg {
    // in
    int sm_count;
    size_t smem_per_sm, smem_optin;   // device limits
    bool light;                       // use the light geometry (fused sources only)
    bool very_light;                  // the integrand is a handful of flops: the sampler is the whole kernel
    // work items of the launch's chunk range, per geometry ([0]: VB_CH-cube chunks, [1]: VB_LCH-cube
    // chunks, whole range only); item_off == nullptr: one item per chunk
    const int64_t* item_off[2];
    int64_t item_begin[2], item_end[2];
    // out
    int nt, ch;                       // CTA size and cubes per chunk of the chosen instantiation
    int cap;                          // staged samples per tile
    size_t smem;                      // dynamic shared memory bytes
    int blocks_per_sm;
    int wtot;                         // bins of the shared-memory histogram windows
    int64_t nchunks;                  // chunks of `ch` cubes in the launch
};

int launch_fused_poly(const EngineP& p, const void* functor, LaunchCfg& cfg, cudaStream_t st);
int launch_fused_gaussmix(const EngineP& p, const void* functor, LaunchCfg& cfg, cudaStream_t st);
int launch_fused_ridge(const EngineP& p, const void* functor, LaunchCfg& cfg, cudaStream_t st);
int launch_fused_genz(const EngineP& p, const void* functor, LaunchCfg& cfg, cudaStream_t st);
int launch_fused_pathint(const EngineP& p, const void* functor, int nx0, LaunchCfg& cfg, cudaStream_t st);
int launch_buffer(const EngineP& p, int nf, LaunchCfg& cfg, cudaStream_t st);
int launch_reduce(const EngineP& p, int nf, LaunchCfg& cfg, cudaStream_t st);   // k_reduce<NF> (reduce.cuh): TMA-staged rows
bool reduce_bulk_ok(const EngineP& p);

// dry-run variants: only compute the geometry (cfg outputs, grid size) so the caller can size
// scratch before the real launch.  Implemented by passing st == (cudaStream_t)-1.
#define VB_DRYRUN ((cudaStream_t)(intptr_t)-1)

static inline int vb_env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// Windows of the training histogram (and, for GRIDW sources, of the map's grid) kept in shared
// memory (engine.cuh, HistW).  On axis d a chunk of `ch` consecutive hypercubes spans at most `nd`
// strata; the window must hold their bins.  Axes are admitted smallest window first until
// `budget_bins` is used up (the others fall back to global memory); what is left upgrades partial
// windows to the full axis, fastest-running axis first, so they are never flushed.
static inline void vb_plan_windows(EngineP& p, int ch, long long budget_bins, bool always)
{
    const int dim = p.map.dim;
    p.wtot = 0;
    for (int d = 0; d < VB_MAXD; ++d) { p.wcap[d] = 0; p.woff[d] = 0; }
    if ((!always && !(p.flags & (VBF_TRAIN | VBF_TRAIN_ERRORS))) || budget_bins <= 0) return;
    long long need[VB_MAXD];
    int order[VB_MAXD];
    for (int d = 0; d < dim; ++d) {
        const long long ns = p.st.nstrat[d], ni = p.map.ninc[d], cs = p.cstride[d];
        long long nd;
        if (cs >= ch) nd = (cs % ch == 0) ? 1 : 2;
        else nd = (ch + cs - 1) / cs + ((ch % cs == 0) ? 0 : 1);
        need[d] = (nd >= ns) ? ni : (nd * ni + ns - 1) / ns + 1;
        if (need[d] > ni) need[d] = ni;
        order[d] = d;
    }
    for (int i = 1; i < dim; ++i)                                   // insertion sort by need (stable)
        for (int j = i; j > 0 && need[order[j]] < need[order[j - 1]]; --j) { int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
    long long used = 0;
    for (int i = 0; i < dim; ++i) {
        const int d = order[i];
        if (used + need[d] <= budget_bins) { p.wcap[d] = (int)need[d]; used += need[d]; }
    }
    for (int d = 0; d < dim; ++d) {
        const long long ni = p.map.ninc[d];
        if (p.wcap[d] > 0 && p.wcap[d] < ni && used + (ni - p.wcap[d]) <= budget_bins) { used += ni - p.wcap[d]; p.wcap[d] = (int)ni; }
    }
    int off = 0;
    for (int d = 0; d < dim; ++d) { p.woff[d] = off; off += p.wcap[d]; }
    p.wtot = off;
}

template <class Src>
static int launch_engine(const EngineP& p_in, const Src& src, LaunchCfg& cfg, cudaStream_t st)
{
    constexpr int NF = Src::NF, NT = Src::NT, CH = Src::CH;
    constexpr int DIGB = (int)sizeof(typename Src::dig_t);
    auto kern = k_engine<Src>;
    EngineP p = p_in;
    const int dim = p.map.dim;
    cfg.nt = NT; cfg.ch = CH;
    // staging capacity: 16 samples per thread (heavy); light: 20 (a 512-cube chunk then usually is one tile: fewer
    // barriers) unless the integrand is only a handful of flops -- then 8, which leaves room for the histogram and
    // grid windows of ALL axes of the 8-D benchmark (N = 1 ridge: 5.96 ms against 6.4 ms; N = 30: 11.3 against 10.7)
    int cap = vb_env_int(CH == VB_CH ? "VB200_CAP" : "VB200_LCAP", (CH != VB_CH ? (cfg.very_light ? 8 : 20) : 16) * NT);
    const int lim = ((CH == VB_CH ? (NF > 4 ? 48 : 32) : 64) * 1024) / (8 * NF);   // many components: fewer, larger tiles (measured)
    if (cap > lim) cap = lim;
    if (cap < 256) cap = 256;
    if (cap > 32 * NT) cap = 32 * NT;                               // k_engine's start-bit words: one per thread
    if (CH != VB_CH) {
        // super-chunks: only whole-range launches (the fused path); slabs must be whole chunks
        if (p.chunk_begin != 0 || p.chunk_off != nullptr) return -23;
        if (p.st.world > 1 && p.st.slab % CH != 0) return -23;
        p.chunk_end = (p.st.nlocal + CH - 1) / CH;
    }
    cfg.nchunks = p.chunk_end - p.chunk_begin;
    const int gi = CH == VB_CH ? 0 : 1;
    p.item_off = cfg.item_off[gi]; p.item_begin = cfg.item_begin[gi]; p.item_end = cfg.item_end[gi];
    if (p.item_off == nullptr) { p.item_begin = p.chunk_begin; p.item_end = p.chunk_end; }
    const int64_t nwork = p.item_end - p.item_begin;                // claimable work items
    const int max_grid = (int)(nwork < 0x7fffffff ? nwork : 0x7fffffff);
    // The kernel's attributes and its residency per shared-memory size are asked of the driver once
    // per (device, instantiation): these queries run twice per iteration otherwise, a visible share
    // of the ~0.5 ms an iteration costs at the reference's everyday sizes (neval = 1e4).
    struct KernelInfo { bool ready = false; cudaFuncAttributes fa; std::map<size_t, int> bps; };
    static thread_local std::map<int, KernelInfo> info_by_device;
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return -(int)e - 1000;
    KernelInfo& ki = info_by_device[device];
    if (!ki.ready) {
        e = cudaFuncGetAttributes(&ki.fa, kern);
        if (e != cudaSuccess) return -(int)e - 1000;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((long long)cfg.smem_optin - (long long)ki.fa.sharedSizeBytes));
        if (e != cudaSuccess) return -(int)e - 1000;
        ki.ready = true;
    }
    const cudaFuncAttributes& fa = ki.fa;
    const long long dyn_max = (long long)cfg.smem_optin - (long long)fa.sharedSizeBytes;
    auto residency = [&](size_t smem, int& out) -> cudaError_t {
        auto it = ki.bps.find(smem);
        if (it != ki.bps.end()) { out = it->second; return cudaSuccess; }
        cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&out, kern, NT, smem);
        if (err == cudaSuccess) ki.bps[smem] = out;
        return err;
    };
    int bps = 0;
    const int cap_bins = vb_env_int("VB200_HIST_BINS", 1 << 30);
    const int cap_first = cap;
    bool giving_up = false;
    for (;;) {
        cfg.cap = p.cap = cap;
        // pass 1: residency without the windows (registers / staging buffer decide)
        vb_plan_windows(p, CH, 0, false);
        size_t smem0 = engine_layout(p, NF, cap, CH, dim, Src::GRIDW, DIGB);
        e = residency(smem0, bps);
        if (e != cudaSuccess) return -(int)e - 1000;
        if (bps < 1) return -24;                                    // does not fit at all
        // pass 2: give the windows the shared memory that this residency leaves unused
        long long per_cta = (long long)cfg.smem_per_sm / bps - 1024;    // 1 KB reserved per CTA
        per_cta -= (long long)fa.sharedSizeBytes;
        if (per_cta > dyn_max) per_cta = dyn_max;
        const long long per_bin = sizeof(double) + sizeof(unsigned) + (Src::GRIDW ? sizeof(double) : 0);
        long long budget = (per_cta - (long long)smem0 - 256) / per_bin;
        if (budget > cap_bins) budget = cap_bins;
        vb_plan_windows(p, CH, budget, Src::GRIDW);
        cfg.wtot = p.wtot;
        cfg.smem = engine_layout(p, NF, cap, CH, dim, Src::GRIDW, DIGB);
        // Sources that read the map's grid from the windows fall back to global memory for EVERY sample as soon
        // as one axis has no window (FusedSrc::sample) -- a cliff: the N = 1 ridge took 8.4 ms instead of 6.0 at
        // the strata of neval = 4e8, where the windows missed the budget by 3 %.  Trade staging capacity for
        // window bins until every axis fits (down to 4 samples per thread).
        // (only for the cheapest functors: with a term loop around, larger tiles matter more -- N = 30: 10.8 ms with
        //  20 samples per thread and the fallback, 12.0 ms with windows on every axis and 8 per thread)
        if (!Src::GRIDW || !cfg.very_light || p.win_all || giving_up) break;
        if (cap <= 4 * NT) { cap = cap_first; giving_up = true; continue; }     // no capacity makes them fit: keep the first choice
        cap -= NT;
    }
    e = residency(cfg.smem, bps);
    if (e != cudaSuccess) return -(int)e - 1000;
    if (bps < 1) return -24;
    cfg.blocks_per_sm = bps;
    int grid = bps * cfg.sm_count;
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    if (st == VB_DRYRUN) return grid;
    kern<<<grid, NT, cfg.smem, st>>>(p, src);
    e = cudaGetLastError();
    if (e != cudaSuccess) return -(int)e - 1000;
    return grid;
}

#ifndef VB_LGW_MAXD
#define VB_LGW_MAXD 10     // light geometry: map-grid windows in shared memory up to this many dimensions
#endif
// padded-dimension dispatch; light geometry first when asked for and compiled in
#define VB_DISPATCH_D(F, fobj, LIST_MACRO)                                                     \
    do {                                                                                       \
        const int dim_ = p.map.dim;                                                            \
        LIST_MACRO(F, fobj)                                                                    \
        return -22;                                                                            \
    } while (0)
#define VB_CASE_D(F, fobj, DD)                                                                 \
    if (dim_ <= DD) { FusedSrc<F, DD> s_{fobj}; return launch_engine(p, s_, cfg, st); }
#define VB_CASE_L(F, fobj, DD)                                                                 \
    if (cfg.light && dim_ <= DD) { FusedSrc<F, DD, true, (DD <= VB_LGW_MAXD)> s_{fobj}; return launch_engine(p, s_, cfg, st); }
// exact dimension: no axis predication (FusedSrc::sample)
#define VB_CASE_DX(F, fobj, DD)                                                                \
    if (dim_ == DD) { FusedSrc<F, DD, false, false, true> s_{fobj}; return launch_engine(p, s_, cfg, st); }
#define VB_CASE_LX(F, fobj, DD)                                                                \
    if (cfg.light && dim_ == DD) { FusedSrc<F, DD, true, (DD <= VB_LGW_MAXD), true> s_{fobj}; return launch_engine(p, s_, cfg, st); }


// ---------------------------------------------------------------------------------------------
// The built-in functors on HBM buffers: f[rows][NF] = F(x[rows][dim]) -- the same device code the
// fused kernel inlines, for callers that need the values themselves (vb200_eval_integrand: the
// stratification profile of restratify, DeviceIntegrand as a device batch callback).
// ---------------------------------------------------------------------------------------------
template <class F, int D>
__global__ void __launch_bounds__(256) k_eval(const __grid_constant__ F f, int dim, const double* __restrict__ x, int64_t rows,
                                              double* __restrict__ out)
{
    vb_exp_init();
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x) {
        double xv[D], fx[F::NF];
#pragma unroll
        for (int d = 0; d < D; ++d) xv[d] = d < dim ? x[i * dim + d] : 0.0;
        f(xv, dim, fx);
#pragma unroll
        for (int s = 0; s < F::NF; ++s) out[i * F::NF + s] = fx[s];
    }
}

template <class F, int D>
static int launch_eval(const F& f, int dim, const double* x, int64_t rows, double* out, int sm_count, cudaStream_t st)
{
    int64_t g = (rows + 255) / 256, cap = (int64_t)sm_count * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    k_eval<F, D><<<(int)g, 256, 0, st>>>(f, dim, x, rows, out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e - 1000;
}
#define VB_EVAL_D(F, fobj, DD) if (dim <= DD) return launch_eval<F, DD>(fobj, dim, x, rows, out, sm_count, st);

int eval_poly(const void* functor, int dim, const double* x, int64_t rows, double* out, int sm_count, cudaStream_t st);
int eval_gaussmix(const void* functor, int dim, const double* x, int64_t rows, double* out, int sm_count, cudaStream_t st);
int eval_ridge(const void* functor, int dim, const double* x, int64_t rows, double* out, int sm_count, cudaStream_t st);
int eval_genz(const void* functor, int dim, const double* x, int64_t rows, double* out, int sm_count, cudaStream_t st);
int eval_pathint(const void* functor, int nx0, int dim, const double* x, int64_t rows, double* out, int sm_count, cudaStream_t st);
