// reduce_buffer.cu -- the per-hypercube reduce of the callback path (vb200_reduce): f, wgt and the training
// bins come from HBM buffers.  Two kernels:
//   k_reduce<NF>            (reduce.cuh) rows staged by TMA bulk copies, double-buffered on mbarriers -- the
//                           default whenever the training bins were written by the sampler (or no training);
//   k_engine<BufferSrc<NF>> (engine.cuh) per-thread global loads, Philox replay of the training bins -- kept
//                           for callers that pass no bins buffer.
#include "dispatch.h"
#include "reduce.cuh"

int launch_buffer(const EngineP& p, int nf, LaunchCfg& cfg, cudaStream_t st)
{
    switch (nf) {
#define C_(N) case N: { BufferSrc<N> s_; return launch_engine(p, s_, cfg, st); }
    C_(1) C_(2) C_(3) C_(4) C_(5) C_(6) C_(7) C_(8)
#undef C_
    default: return -22;
    }
}

template <int NF>
static int launch_reduce_nf(const EngineP& p_in, LaunchCfg& cfg, cudaStream_t st)
{
    typedef ReduceGeom<NF> G;
    auto kern = k_reduce<NF>;
    EngineP p = p_in;
    const int dim = p.map.dim;
    const bool train = (p.flags & VBF_TRAIN) != 0 && p.bins != nullptr;
    cfg.nt = G::NT; cfg.ch = G::CH;
    cfg.nchunks = p.chunk_end - p.chunk_begin;
    p.item_off = cfg.item_off[0]; p.item_begin = cfg.item_begin[0]; p.item_end = cfg.item_end[0];
    if (p.item_off == nullptr) { p.item_begin = p.chunk_begin; p.item_end = p.chunk_end; }
    const int64_t nwork = p.item_end - p.item_begin;
    const int max_grid = (int)(nwork < 0x7fffffff ? nwork : 0x7fffffff);
    static thread_local std::map<int, cudaFuncAttributes> fa_by_device;
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return -(int)e - 1000;
    auto it = fa_by_device.find(device);
    if (it == fa_by_device.end()) {
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, kern);
        if (e != cudaSuccess) return -(int)e - 1000;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((long long)cfg.smem_optin - (long long)fa.sharedSizeBytes));
        if (e != cudaSuccess) return -(int)e - 1000;
        it = fa_by_device.emplace(device, fa).first;
    }
    const cudaFuncAttributes& fa = it->second;
    const int bps = G::MINB;
    long long per_cta = (long long)cfg.smem_per_sm / bps - 1024 - (long long)fa.sharedSizeBytes;
    const long long dyn_max = (long long)cfg.smem_optin - (long long)fa.sharedSizeBytes;
    if (per_cta > dyn_max) per_cta = dyn_max;
    // a stage holds cap + 16 rows of f, wgt and (training) bins; two stages.  About half of the CTA's shared
    // memory goes to the stages (enough bytes in flight per SM to cover the HBM latency), the rest to the
    // histogram windows.
    const int rowb = 8 * NF + 8 + (train ? 2 * dim : 0);
    const long long fixed = sizeof(long long) * (G::CH + 1) + sizeof(int) * G::CH + sizeof(int) * (G::MAXT + 1) + 256;
    long long stage_budget = (per_cta - fixed) / 2;
    if (!(p.flags & (VBF_TRAIN | VBF_TRAIN_ERRORS))) stage_budget = per_cta - fixed - 2048;
    if (stage_budget > 144 * 1024) stage_budget = 144 * 1024;
    int cap = (int)(stage_budget / 2 / rowb) - 16;
    cap = vb_env_int("VB200_RCAP", cap);
    // cap = 3 * 2^k (k_reduce bins the cubes by their first row in buckets of 2/3 cap rows, a power of two)
    int k = 6;
    while (3 * (2 << k) <= cap && k < 9) ++k;
    if (3 * (1 << k) > cap) return -24;
    cap = 3 * (1 << k);
    cfg.cap = p.cap = cap;
    vb_plan_windows(p, G::CH, 0, false);
    const size_t smem0 = reduce_layout(p, NF, cap + 16, G::CH, dim, G::MAXT, train);
    long long budget = (per_cta - (long long)smem0 - 256) / (long long)(sizeof(double) + sizeof(unsigned));
    vb_plan_windows(p, G::CH, budget, false);
    cfg.wtot = p.wtot;
    cfg.smem = reduce_layout(p, NF, cap + 16, G::CH, dim, G::MAXT, train);
    if ((long long)cfg.smem > dyn_max) return -24;
    cfg.blocks_per_sm = bps;
    int grid = bps * cfg.sm_count;
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    if (st == VB_DRYRUN) return grid;
    kern<<<grid, G::NT, cfg.smem, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return -(int)e - 1000;
    return grid;
}

// usable when the rows can be fetched in 16-byte units from 16-byte aligned buffers and the training points
// (if any) come from the sampler's bins
bool reduce_bulk_ok(const EngineP& p)
{
    const bool wants_bins = (p.flags & (VBF_TRAIN | VBF_TRAIN_ERRORS)) != 0;
    if (wants_bins && p.bins == nullptr) return false;
    if (((uintptr_t)p.fbuf | (uintptr_t)p.wbuf | (uintptr_t)p.bins) & 15u) return false;
    return vb_env_int("VB200_REDUCE_BULK", 1) != 0;
}

int launch_reduce(const EngineP& p, int nf, LaunchCfg& cfg, cudaStream_t st)
{
    switch (nf) {
#define C_(N) case N: return launch_reduce_nf<N>(p, cfg, st);
    C_(1) C_(2) C_(3) C_(4) C_(5) C_(6) C_(7) C_(8)
#undef C_
    default: return -22;
    }
}
