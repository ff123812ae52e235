// engine.cuh -- the vegas+ iteration engine: one persistent kernel per iteration (or per batch).
//
// Work decomposition ("hypercube tiling"):
//   chunk  = VB_CH consecutive hypercubes of this rank's share.
//   item   = the unit the persistent CTAs claim from an atomic counter (so they all finish together):
//            a whole chunk, or -- when the vegas+ allocation piled more than VB_ITEM samples onto one
//            chunk (k_plan) -- one of m runs of its cubes holding ~1/m of its samples each.
//   tile   = run of cubes inside an item whose samples fit the shared-memory staging buffer (cap
//            samples; an item is cut into the fewest tiles of about equal size).  Cubes larger
//            than cap are staged through global scratch.
//   phase 1 (thread per SAMPLE): Philox -> stratified y -> AdaptiveMap -> integrand -> w*f staged
//            in shared memory; training-histogram adds are issued here because
//            fdv2 = (J f dv_y)^2 does not depend on the cube sums.
//   phase 2 (thread per CUBE up to VB_WARP_CUBE samples; larger cubes shared by all warps of the
//            CTA): the reference's two-pass mean/variance with its EPSILON clamp -- for the small
//            cubes in the reference's summation order -- from the staged values; sigf[h] update;
//            per-thread fp64 accumulators.
// Reference: Integrator._random_batch (_vegas.pyx:1692-1759) + Integrator.__call__
// (_vegas.pyx:2136-2197).
#pragma once
#include <type_traits>
#include "common.cuh"

#define VB_LARGE_MAX 128      // large cubes (> VB_WARP_CUBE samples) one tile can hold: 8192 / 65
#define VB_LARGE_G 8          // large cubes reduced per round of phase 2b
#define VBF_UPDATE_SIGF   1   // adaptive stratification: write sigf, accumulate sum_sigf
#define VBF_TRAIN         2   // add (J f dv_y)^2 to the map's training histogram
#define VBF_TRAIN_ERRORS  4   // adapt_to_errors: one training point per cube carrying its variance
#define VBF_CORRELATE     8   // accumulate the full covariance of multi-component integrands

struct EngineP {
    MapP map;
    StrataP st;
    AllocP al;
    PhiloxKey key;
    uint32_t itn;
    int flags;
    double dv_y;               // 1 / nhcube
    double beta_half;          // beta / 2
    double* sigf_out;          // [nlocal]
    double* sum_f;             // [dim][hstride]
    unsigned long long* n_f;   // [dim][hstride]
    int hstride;
    int cap;                   // samples staged per tile
    int* status;               // [0] != 0 => integrand returned NaN
    double* partials;          // [gridDim.x][NF + NF(NF+1)/2 + 1]
    double* scratch;           // [gridDim.x][NF][scratch_stride]
    int64_t scratch_stride;
    int64_t chunk_begin, chunk_end;   // local chunk range of this launch
    unsigned long long* work_counter; // zeroed before the launch: next work item to claim
    // work items: item j of the launch is global item item_begin + j; chunk lc (CH cubes of the launched
    // geometry) owns items [item_off[lc], item_off[lc+1]) (item_off == nullptr: item j == chunk j)
    const int64_t* item_off;          // [nchunks+1]
    int64_t item_begin, item_end;
    int64_t cstride[VB_MAXD];  // cstride[d] = prod_{e<d} nstrat[e]
    // shared-memory windows of the training histogram (0 bins on an axis: global atomics there)
    int wcap[VB_MAXD];         // bins of axis d's window
    int woff[VB_MAXD];         // offset of axis d's window in the shared arrays
    int wtot;                  // total window bins (0: no shared histogram)
    double dni[VB_MAXD];       // (double) map.ninc[d]
    // layout of the kernel's dynamic shared memory (engine_layout, filled by the launcher): byte offsets
    // of the arrays, and per axis the element index of its window's first bin in the fp64 sums
    // (vb_smem as double[]), the u32 counts (vb_smem as unsigned[]) and the grid nodes (double[])
    int win_all;               // every axis has a window (GRIDW sources then read the map's grid from shared memory)
    int o_ex, o_hsum, o_gw, o_dvn, o_n, o_hcnt, o_y0;
    int hs_idx[VB_MAXD], hc_idx[VB_MAXD], gw_idx[VB_MAXD];
    // unfused path
    const double* fbuf;        // [rows][nf]
    const double* wbuf;        // [rows]
    const uint16_t* bins;      // [rows][dim] training bins written by the sampler (nullptr: replay Philox)
    const int64_t* chunk_off;  // [nchunks+1] exclusive scan of samples per chunk
    int64_t row0;              // chunk_off[chunk_begin]
    // k_reduce (reduce.cuh): rows of the launch's batch buffers, and the layout of one shared-memory stage
    int64_t batch_rows;
    int st_w, st_b, st_bytes;  // byte offsets of wgt / bins inside a stage, stage size
};

__device__ __forceinline__ int tri(int s, int t) { return s * (s + 1) / 2 + t; }

// y-space bin of sample (h,k) on axis d for the training histogram; -1 when y is on the boundary
// (AdaptiveMap.add_training_data skips y<=0 and y>=1, _vegas.pyx:460).
__device__ __forceinline__ int bin_of(const EngineP& p, int d, uint32_t y0, double u, double* y_out)
{
    double y = div_exact((double)y0 + u, p.st.dns[d], p.st.rns[d]);
    if (y_out) *y_out = y;
    int iy = min(__double2int_rd(__dmul_rn(y, (double)p.map.ninc[d])), p.map.ninc[d] - 1);
    return (y > 0.0 && y < 1.0) ? iy : -1;
}

// first training bin that samples of stratum `digit` on axis d can fall into.  Same arithmetic as
// bin_of with u = 0, and bin_of is monotonic in (y0 + u): every sample of strata [a, b] on this
// axis lands in bins [bin_floor(a), bin_floor(b + 1)].
__device__ __forceinline__ int bin_floor(const EngineP& p, int d, int digit)
{
    double y = div_exact((double)digit, p.st.dns[d], p.st.rns[d]);
    return min(__double2int_rd(__dmul_rn(y, (double)p.map.ninc[d])), p.map.ninc[d] - 1);
}

// The training histogram (AdaptiveMap.add_training_data, _vegas.pyx:421-464) is accumulated in
// shared memory: a chunk of consecutive hypercubes only touches a window of bins on each axis
// (all of them on the fastest-running axes, one or two strata on the others).  The CTA keeps one
// window per axis, [lo[d], lo[d] + wcap[d]), and flushes it to the global histogram with fp64 /
// u64 atomics only when a newly claimed chunk needs a different window.  Bins outside the window
// (axes whose window did not fit, chunks that wrap around an axis) go straight to global memory.
extern __shared__ __align__(16) double vb_smem[];   // dynamic shared memory of the engine kernel (16-byte aligned: bulk-copy destination)
static __shared__ int vb_wlo_s[VB_MAXD];   // first bin of the current window of each axis
// (file-scope declarations so that every access compiles to LDS / ATOMS: through generic pointers
//  carried in a struct the compiler falls back to generic loads and the slower generic ATOM forms)

struct HistW {
    double* sum;        // [wtot]
    unsigned* cnt;      // [wtot]
};

#define VB_NO_SLOT 0xffffffffu

__device__ __forceinline__ void hist_global(const EngineP& p, int d, int bin, double v)
{
    atomicAdd(p.sum_f + (size_t)d * p.hstride + bin, v);
    atomicAdd(p.n_f + (size_t)d * p.hstride + bin, 1ull);
}

// One training point on axis d.  `code` < 2^31: slot of the axis' window; VB_NO_SLOT: none (y on the
// boundary); else 2^31 | bin: outside the window, straight to the global histogram.
// There is no native shared-memory fp64 add; on an address the compiler can prove to be shared
// (vb_smem itself, not a pointer carried in a struct) atomicAdd(double) becomes the 4-instruction loop
// LDS.64 / DADD / ATOMS.CAST.SPIN.64 / BRA -- measured (tools/atomics_bench.cu, B200, 8 axes per
// sample, 125-bin windows): 6.8 SM-cycles per sample against 10.4 for round 1's hand-written PTX
// loop running 4 compare-and-swaps in lock-step, 17.9 for a 128-bit CAS on {sum, count}, 15.2 with
// __match_any_sync pre-aggregation, and a floor of 4.2 for the same updates without atomicity.
__device__ __forceinline__ void hist_add_code(const EngineP& p, int d, unsigned code, double v)
{
    if (code < 0x80000000u) {
        atomicAdd((unsigned*)vb_smem + (p.hc_idx[d] + (int)code), 1u);
        atomicAdd(vb_smem + (p.hs_idx[d] + (int)code), v);
    } else if (code != VB_NO_SLOT) hist_global(p, d, (int)(code & 0x7fffffffu), v);
}

// training code of bin `bin` (>= 0) on axis d
__device__ __forceinline__ unsigned hist_code(const EngineP& p, int d, int bin)
{
    const unsigned r = (unsigned)(bin - vb_wlo_s[d]);
    return r < (unsigned)p.wcap[d] ? r : (0x80000000u | (unsigned)bin);
}

__device__ __forceinline__ void hist_add(const EngineP& p, const HistW&, int d, int bin, double v)
{
    if (bin >= 0) hist_add_code(p, d, hist_code(p, d, bin), v);
}

__device__ __forceinline__ void hist_add_divergent(const EngineP& p, const HistW& H, int d, int bin, double v)
{
    hist_add(p, H, d, bin, v);
}

// add the windows of the axes with need[d] != 0 (all axes when need == nullptr) to the global
// histogram and clear them; called by the whole CTA between barriers
template <int NT>
__device__ __forceinline__ void hist_flush(const EngineP& p, const HistW& H, const int* need)
{
    for (int d = 0; d < p.map.dim; ++d) {
        const int cap = p.wcap[d];
        if (cap == 0 || (need && !need[d])) continue;
        const int lo = vb_wlo_s[d], off = p.woff[d];
        for (int i = threadIdx.x; i < cap; i += NT) {
            const unsigned c = H.cnt[off + i];
            if (c) {
                atomicAdd(p.sum_f + (size_t)d * p.hstride + lo + i, H.sum[off + i]);
                atomicAdd(p.n_f + (size_t)d * p.hstride + lo + i, (unsigned long long)c);
                H.sum[off + i] = 0.0;
                H.cnt[off + i] = 0u;
            }
        }
    }
}

// whole CTA: bring the windows to the strata the CH cubes starting at global cube h0 touch; the windows
// that have to move (or all of them when `force`) are flushed to the global histogram first.
// wneed_s[d] tells the caller which ones moved (to wnew_s[d]).  Contains barriers.
template <int NT>
__device__ __forceinline__ void hist_move_windows(const EngineP& p, const HistW& H, int64_t h0, int CH, bool force,
                                                  int* wnew_s, int* wneed_s)
{
    const int tid = threadIdx.x, dim = p.map.dim;
    int* const wlo_s = vb_wlo_s;
    if (tid < dim) {
        const int d = tid;
        int need = 0, lo_bin = 0;
        if (p.wcap[d] > 0) {
            const int64_t a = h0 / p.cstride[d], b = (h0 + CH - 1) / p.cstride[d];
            const int64_t ns = p.st.nstrat[d];
            int dlo = 0, dhi = (int)ns - 1;
            if (b - a + 1 < ns) {
                dlo = (int)(a % ns);
                const int e = (int)(b % ns);
                if (e >= dlo) dhi = e;                     // else the chunk wraps: keep [dlo, ns-1]
            }
            lo_bin = bin_floor(p, d, dlo);
            int hi_bin = bin_floor(p, d, dhi + 1);
            if (hi_bin > lo_bin + p.wcap[d] - 1) hi_bin = lo_bin + p.wcap[d] - 1;
            const int cur = wlo_s[d];
            need = (force || lo_bin < cur || hi_bin >= cur + p.wcap[d]) ? 1 : 0;
        }
        wneed_s[d] = need;
        wnew_s[d] = lo_bin;
    }
    __syncthreads();
    hist_flush<NT>(p, H, wneed_s);
    __syncthreads();
    if (tid < dim && wneed_s[tid]) wlo_s[tid] = wnew_s[tid];
}

// training point of a whole cube (adapt_to_errors, _vegas.pyx:2187-2193): y of its LAST sample
template <class dig_t>
static __device__ __noinline__ void train_cube(const EngineP& p, const HistW& H, int64_t h, uint32_t klast,
                                               const dig_t* y0, double v, int64_t row_last)
{
    if (p.bins != nullptr) {        // unfused path with the sampler's bins (uniforms injected by the caller: no Philox replay)
        const uint16_t* b = p.bins + row_last * p.map.dim;
        for (int d = 0; d < p.map.dim; ++d)
            if (b[d] != 0xffffu) hist_add_divergent(p, H, d, (int)b[d], fabs(v));
        return;
    }
    for (int pr = 0; 2 * pr < p.map.dim; ++pr) {
        double ua, ub;
        philox_pair(p.key, p.itn, h, klast, pr, ua, ub);
        hist_add_divergent(p, H, 2 * pr, bin_of(p, 2 * pr, y0[2 * pr], ua, nullptr), fabs(v));
        if (2 * pr + 1 < p.map.dim)
            hist_add_divergent(p, H, 2 * pr + 1, bin_of(p, 2 * pr + 1, y0[2 * pr + 1], ub, nullptr), fabs(v));
    }
}

// ---------------------------------------------------------------------------------------------
// sample sources
// ---------------------------------------------------------------------------------------------
// Fused: everything from the Philox counter to w*f happens in registers.
// Two geometries:
//   heavy (LIGHT = false): 128-thread CTAs, several per SM, 256-cube chunks, wide register budget --
//                          for integrands whose own loop dominates (the N = 1000 ridge);
//   light (LIGHT = true):  TWO 256-thread CTAs per SM (VB_LNT threads, VB_LCH = 512-cube chunks, 128
//                          registers), each with ~100 KB of shared memory for its histogram windows AND
//                          a copy of the map's grid nodes for those windows -- for cheap integrands, where
//                          the sampler itself is the work.  (One 512-thread CTA per SM shares bigger
//                          windows but idles the whole SM at every barrier: 7.7 -> 7.4 ms on the N = 1 ridge.)
#ifndef VB_LNT
#define VB_LNT 256
#endif
#ifndef VB_LCH
#define VB_LCH 512
#endif
// what a source needs to know about one sample: its cube (n samples, dvn = dv_y / n, global index h,
// stratum digits y0), its index k in the cube and its row in the batch buffers (unfused path)
template <class dig_t>
struct SampleRef {
    int n;
    double dvn;
    int64_t h;
    uint32_t k;
    int64_t row;
    const dig_t* y0;
};

template <class F, int D, bool LIGHT = false, bool GW = LIGHT, bool EXACT = false>
struct FusedSrc {
    static constexpr int NF = F::NF;
    static constexpr int NT = LIGHT ? VB_LNT : VB_ENT;             // threads per CTA
    static constexpr int CH = LIGHT ? VB_LCH : VB_CH;              // hypercubes per chunk
#ifndef VB_LMINB
#define VB_LMINB 2
#endif
#ifndef VB_HMINB
#define VB_HMINB 4          // heavy: 4 CTAs of 128 threads per SM at 128 registers (no spills; 173 ms against 178 ms at 3 x 165)
#endif
    static constexpr int MINB = LIGHT ? VB_LMINB : (F::NF == 1 ? VB_HMINB : 2);   // resident CTAs per SM the register budget is set for
    static constexpr bool GRIDW = GW;                              // grid windows in shared memory (light, D <= 10)
    static constexpr bool USES_EXP = true;
    static constexpr bool CLAIM = !LIGHT;                          // phase 1: warps claim 32-sample words dynamically
    // stratum digits of a cube (light: narrow, to leave the shared memory to the histogram windows;
    // the host falls back to the heavy geometry when a digit does not fit)
    typedef typename std::conditional<LIGHT, typename std::conditional<(D > 10), uint8_t, uint16_t>::type, uint32_t>::type dig_t;
    F f;
    // Philox -> stratified y -> AdaptiveMap (pyx:310-360) for all axes of one sample.  WIN: every axis
    // reads its grid nodes from the shared-memory window (callers guarantee the bins are inside);
    // otherwise from global memory.  Returns false when WIN met a bin outside its window (the
    // results are then meaningless and the caller redoes the sample with WIN = false).
    template <bool WIN>
    __device__ __forceinline__ bool map_axes(const EngineP& p, const SampleRef<dig_t>& r, int dim, double (&x)[D],
                                             unsigned (&code)[D], double& jac_out) const
    {
        double jac = 1.0;
        bool ok = true;
        // Above 10 dimensions the axis loops stay rolled (x[], code[] then live in local memory, which
        // is lane-interleaved and L1-resident): fully unrolled, the 20-D kernels were bound by
        // instruction fetch (ncu: stall_no_instruction on top).
#ifndef VB_ROLL_UNR
#define VB_ROLL_UNR 1
#endif
        constexpr int UNR = D > 10 ? VB_ROLL_UNR : (D + 1) / 2;
#pragma unroll UNR
        for (int pr = 0; pr < (D + 1) / 2; ++pr) {
            if (EXACT || 2 * pr < dim) {
                double u[2];
                philox_pair(p.key, p.itn, r.h, r.k, pr, u[0], u[1]);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int d = 2 * pr + e;
                    if (d < D && (EXACT || d < dim)) {
                        const int ni = p.map.ninc[d];
                        const double yy = (double)r.y0[d] + u[e];
                        const double y = div_exact(yy, p.st.dns[d], p.st.rns[d]);
                        const double t = __dmul_rn(y, p.dni[d]);
                        const int iy = __double2int_rd(t);
                        const int ic = min(iy, ni - 1);
                        const unsigned w = (unsigned)(ic - vb_wlo_s[d]);
                        const bool inw = w < (unsigned)p.wcap[d];
                        double g0, g1;
                        if (WIN) {
                            const double* gw = vb_smem + (p.gw_idx[d] + (int)(inw ? w : 0u));
                            g0 = gw[0]; g1 = gw[1];
                            ok &= inw;
                        } else {
                            const double* gp = p.map.grid + ((size_t)d * p.map.gstride + ic);
                            g0 = __ldg(gp); g1 = __ldg(gp + 1);
                        }
                        const double inc = g1 - g0;
                        const double xin = __dadd_rn(g0, __dmul_rn(inc, __dsub_rn(t, (double)iy)));   // no FMA: bit-identical to pyx:354
                        x[d] = iy < ni ? xin : g1;                                                     // pyx:357-359
                        jac *= inc * p.dni[d];
                        // pyx:460 trains only 0 < y < 1.  y > 0 <=> y0 + u > 0; y < 1 <=> iy < ninc (y <= 1 - 2^-53
                        // cannot round up to ninc in y * ninc, and y == 1 gives iy == ninc exactly)
                        code[d] = (iy < ni && yy > 0.0) ? ((WIN || inw) ? w : (0x80000000u | (unsigned)ic)) : VB_NO_SLOT;
                    }
                }
            }
        }
        jac_out = jac;
        return ok;
    }

    // EXACT: the run-time dimension equals D, so no axis is predicated and the compiler is free to
    // interleave the D independent axis chains (Philox pair -> y -> bin -> x)
    __device__ __forceinline__ void sample(const EngineP& p, const SampleRef<dig_t>& r, double (&wf)[NF]) const
    {
        const int dim = EXACT ? D : p.map.dim;
        double x[D];
        unsigned code[D];     // training slot of axis d, see hist_add_code
        double jac;
        // GRIDW sources: windows first; a sample with a bin outside them (an axis without a window, a
        // chunk wrapping around an axis) is redone from global memory -- a real, rarely taken branch
        // instead of predicated global-memory code on every axis
        if (!GRIDW || !p.win_all || !__builtin_expect(map_axes<true>(p, r, dim, x, code, jac), 1))
            map_axes<false>(p, r, dim, x, code, jac);
        double fx[NF];
        f(x, dim, fx);
        const double wgt = jac * r.dvn;
        bool bad = false;
#pragma unroll
        for (int s = 0; s < NF; ++s) { wf[s] = wgt * fx[s]; bad |= isnan(fx[s]); }
        if (bad) p.status[0] = 1;
        if (p.flags & VBF_TRAIN) {
            const double a = wf[0] * (double)r.n;
            const double fdv2 = __dmul_rn(a, a);
            {
                constexpr int UNRH = D > 10 ? 1 : D;
#pragma unroll UNRH
                for (int d = 0; d < D; ++d)
                    if (EXACT || d < dim) hist_add_code(p, d, code[d], fdv2);
            }
        }
    }
};

// Unfused: w and f come from HBM buffers filled by k_sample and the user's batch integrand;
// y (for the training bins) is re-derived from the Philox counter instead of round-tripping HBM.
template <int NF_>
struct BufferSrc {
    static constexpr int NF = NF_;
    static constexpr int NT = VB_ENT, CH = VB_CH;
#ifndef VB_BUF_MINB_HI
#define VB_BUF_MINB_HI 3
#endif
#ifndef VB_BUF_MINB_LO
#define VB_BUF_MINB_LO 4
#endif
    static constexpr int MINB = NF_ <= 4 ? VB_BUF_MINB_LO : VB_BUF_MINB_HI;
    static constexpr bool GRIDW = false;
    static constexpr bool USES_EXP = false;
    static constexpr bool CLAIM = true;
    typedef uint32_t dig_t;
    __device__ __forceinline__ void sample(const EngineP& p, const SampleRef<dig_t>& r, double (&wf)[NF]) const
    {
        const int dim = p.map.dim;
        const int64_t row = r.row;
        const double wgt = p.wbuf[row];
        double fx[NF];
#pragma unroll
        for (int s = 0; s < NF; ++s) fx[s] = p.fbuf[row * NF + s];
        if (p.flags & VBF_TRAIN) {
            const double a = (wgt * fx[0]) * (double)r.n;
            const double fdv2 = a * a;
            if (p.bins != nullptr) {
                const uint16_t* b = p.bins + row * dim;
                for (int d = 0; d < dim; ++d) {
                    const unsigned bv = b[d];
                    if (bv != 0xffffu) hist_add_code(p, d, hist_code(p, d, (int)bv), fdv2);
                }
            } else {
                HistW H{};
                for (int pr = 0; 2 * pr < dim; ++pr) {
                    double ua, ub;
                    philox_pair(p.key, p.itn, r.h, r.k, pr, ua, ub);
                    hist_add(p, H, 2 * pr, bin_of(p, 2 * pr, r.y0[2 * pr], ua, nullptr), fdv2);
                    if (2 * pr + 1 < dim)
                        hist_add(p, H, 2 * pr + 1, bin_of(p, 2 * pr + 1, r.y0[2 * pr + 1], ub, nullptr), fdv2);
                }
            }
        }
        bool bad = false;
#pragma unroll
        for (int s = 0; s < NF; ++s) { bad |= isnan(fx[s]); wf[s] = wgt * fx[s]; }
        if (bad) p.status[0] = 1;
    }
};

// ---------------------------------------------------------------------------------------------
// per-cube statistics (reference two-pass, _vegas.pyx:2142-2186)
// ---------------------------------------------------------------------------------------------
template <int NF>
struct CubeAcc {
    static constexpr int NV = NF * (NF + 1) / 2;
    double mean[NF];
    double var[NV];
    double sum_sigf;
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int s = 0; s < NF; ++s) mean[s] = 0.0;
#pragma unroll
        for (int v = 0; v < NV; ++v) var[v] = 0.0;
        sum_sigf = 0.0;
    }
};

// second-pass contribution of one sample
template <int NF>
__device__ __forceinline__ void pass2_sample(const double (&w)[NF], const double (&m)[NF], bool correlate,
                                             double (&sd)[NF], double (&q)[NF * (NF + 1) / 2])
{
    double d[NF];
#pragma unroll
    for (int s = 0; s < NF; ++s) {
        d[s] = w[s] - m[s];
        double thr = VB_EPSILON * fabs(m[s]);
        if (fabs(d[s]) < thr) {
            q[tri(s, s)] += thr * thr;
            d[s] = 0.0;
        } else {
            q[tri(s, s)] += d[s] * d[s];
        }
        sd[s] += d[s];
        if (correlate) {
#pragma unroll
            for (int t = 0; t < s; ++t) q[tri(s, t)] += d[s] * d[t];
        }
    }
}

// fold one finished cube into the accumulators; returns sigf2 = |var_00|
template <int NF>
__device__ __forceinline__ double cube_finish(CubeAcc<NF>& A, int n, const double (&S)[NF],
                                              const double (&sd)[NF], const double (&q)[NF * (NF + 1) / 2],
                                              bool correlate)
{
    const double dn = (double)n, dn1 = dn - 1.0;
#pragma unroll
    for (int s = 0; s < NF; ++s) {
        A.mean[s] += S[s] + sd[s];
        if (correlate) {
#pragma unroll
            for (int t = 0; t <= s; ++t) A.var[tri(s, t)] += (dn * q[tri(s, t)] - sd[s] * sd[t]) / dn1;
        } else {
            A.var[tri(s, s)] += (dn * q[tri(s, s)] - sd[s] * sd[s]) / dn1;
        }
    }
    return fabs((dn * q[0] - sd[0] * sd[0]) / dn1);
}

template <int NF, class dig_t>
__device__ __forceinline__ void cube_epilogue(const EngineP& p, const HistW& H, CubeAcc<NF>& A, double sigf2,
                                              int64_t lh, int64_t h, int n, const dig_t* y0, int64_t row_last)
{
    if (p.flags & VBF_UPDATE_SIGF) {
        // sigf = sigf2^(beta/2) (pyx:2183) as exp(b log x): 2-3 times fewer instructions than pow(); relative error
        // <= (2 + |b ln x|) ulp, far inside the 1e-12 the stratification is compared at
        const double sg = sigf2 == 0.0 ? 0.0 : exp(p.beta_half * log(sigf2));
        p.sigf_out[lh] = sg;
        A.sum_sigf += sg;
    }
    if (p.flags & VBF_TRAIN_ERRORS) train_cube(p, H, h, (uint32_t)(n - 1), y0, sigf2, row_last);
}

// ---------------------------------------------------------------------------------------------
// work items and chunk set-up, shared by the engine, the samplers and the restratify profile
// ---------------------------------------------------------------------------------------------
// which chunk does work item g (numbered over the whole plan) belong to, and which of its parts?
// (one thread; item_off == nullptr: item g is chunk g)
__device__ __forceinline__ void locate_item(const EngineP& p, long long g, long long& lc, int& sub, int& nsub)
{
    if (p.item_off == nullptr) { lc = g; sub = 0; nsub = 1; return; }
    int64_t lo = p.chunk_begin, hi = p.chunk_end;                  // item_off[lo] <= g < item_off[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (p.item_off[mid] <= g) lo = mid; else hi = mid;
    }
    lc = lo;
    sub = (int)(g - p.item_off[lo]);
    nsub = (int)(p.item_off[lo + 1] - p.item_off[lo]);
}

// whole CTA (NT threads): samples per cube (n_s), their exclusive scan (ex_s[0..CH], ex_s[CH] = total)
// and the stratum digits (y0_s[c*dim + d], mixed radix with 32-bit carries) of the CH cubes starting
// at local cube lh0 / global cube h0.  Returns the chunk's sample total; ends with a barrier.
template <int NT, int CH, class dig_t>
__device__ __forceinline__ long long chunk_setup(const EngineP& p, int64_t lh0, int64_t h0, long long* ex_s, int* n_s,
                                                 dig_t* y0_s, uint32_t* base_s, long long* scan_s, double* dvn_s = nullptr)
{
    constexpr int CPT = CH / NT;                                  // cubes per thread
    static_assert(CH % NT == 0, "chunk must be a multiple of the CTA size");
    const int tid = threadIdx.x, dim = p.map.dim;
    if (tid < dim) base_s[tid] = (uint32_t)((h0 / p.cstride[tid]) % p.st.nstrat[tid]);
    int n_mine[CPT];
    long long mine = 0;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const int c = tid * CPT + i;
        n_mine[i] = (lh0 + c < p.st.nlocal) ? alloc_neval(p.al, lh0 + c) : 0;
        mine += n_mine[i];
    }
    long long total;
    long long ex = block_exscan<NT>(mine, scan_s, &total);         // contains __syncthreads
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const int c = tid * CPT + i;
        ex_s[c] = ex;
        n_s[c] = n_mine[i];
        if (dvn_s) dvn_s[c] = p.dv_y / (double)n_mine[i];           // weight factor of the cube's samples (pyx:1746-1752), once per cube
        ex += n_mine[i];
        uint32_t carry = (uint32_t)c;                              // digits of cube h0+c: base digits plus c, with carries
        if (y0_s != nullptr)
        for (int d = 0; d < dim; ++d) {
            const uint32_t v = base_s[d] + carry, ns = (uint32_t)p.st.nstrat[d];
            const uint32_t qd = digit_div(p.st, d, v);                // v < ns + CH
            y0_s[c * dim + d] = (dig_t)(v - qd * ns);
            carry = qd;
        }
    }
    if (tid == NT - 1) ex_s[CH] = total;
    __syncthreads();
    return total;
}

// cubes [c0, cend) of part `sub` of `nsub` of a chunk: those whose first sample lies in that share
// of the chunk's samples (a cube is never split; broadcast reads of the shared prefix array)
__device__ __forceinline__ void item_cubes(const long long* ex_s, int ch, long long total, int sub, int nsub, int& c0, int& cend)
{
    c0 = 0; cend = ch;
    if (nsub <= 1) return;
    const long long b0 = total * sub / nsub, b1 = total * (sub + 1) / nsub;
    int lo = -1, hi = ch;                                          // first c with ex_s[c] >= b0
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (ex_s[mid] >= b0) hi = mid; else lo = mid; }
    c0 = hi;
    lo = c0 - 1; hi = ch;                                          // first c with ex_s[c] >= b1
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (ex_s[mid] >= b1) hi = mid; else lo = mid; }
    cend = hi;
}

// ---------------------------------------------------------------------------------------------
// the engine kernel
// ---------------------------------------------------------------------------------------------
// dynamic shared memory of k_engine<Src>: the launcher lays it out (byte offsets and the per-axis
// window indices in p) and sizes the launch with the same function
__host__ inline size_t engine_layout(EngineP& p, int nf, int cap, int ch, int dim, bool gridw, int digbytes)
{
    const int wtot = p.wtot;
    size_t b = sizeof(double) * (size_t)nf * cap;           // wf_s   staged w*f
    p.o_ex = (int)b;    b += sizeof(long long) * (size_t)(ch + 1);      // ex_s   exclusive scan of the cubes' sample counts
    p.o_hsum = (int)b;  b += sizeof(double) * (size_t)wtot;             // H.sum
    p.o_gw = (int)b;    b += gridw ? sizeof(double) * (size_t)(wtot + dim) : 0;   // grid nodes of the windows
    p.o_dvn = (int)b;   b += sizeof(double) * (size_t)ch;               // dvn_s  dv_y / n per cube
    p.o_n = (int)b;     b += sizeof(int) * (size_t)ch;                  // n_s
    p.o_hcnt = (int)b;  b += sizeof(unsigned) * (size_t)wtot;           // H.cnt
    p.o_y0 = (int)b;    b += (size_t)digbytes * ch * dim;               // y0_s   stratum digits
    for (int d = 0; d < VB_MAXD; ++d) {
        p.hs_idx[d] = p.o_hsum / 8 + p.woff[d];
        p.hc_idx[d] = p.o_hcnt / 4 + p.woff[d];
        p.gw_idx[d] = p.o_gw / 8 + p.woff[d] + d;
    }
    p.win_all = wtot > 0;
    for (int d = 0; d < dim; ++d) if (p.wcap[d] == 0) p.win_all = 0;
    return (b + 15) & ~(size_t)15;
}

template <class Src>
__global__ void __launch_bounds__(Src::NT, Src::MINB) k_engine(const __grid_constant__ EngineP p, const __grid_constant__ Src src)
{
    typedef typename Src::dig_t dig_t;
    constexpr int NF = Src::NF;
    constexpr int NV = NF * (NF + 1) / 2;
    constexpr int NT = Src::NT;
    constexpr int CH = Src::CH;
    constexpr int NW = NT / 32;
    static_assert(CH % VB_CH == 0, "chunk must be a multiple of the ABI chunk");
    char* const smem_b = (char*)vb_smem;
    double* wf_s = vb_smem;                                       // [NF][cap]
    long long* ex_s = (long long*)(smem_b + p.o_ex);              // [CH + 1]
    HistW H;
    H.sum = (double*)(smem_b + p.o_hsum);                         // [wtot]
    double* gw_s = (double*)(smem_b + p.o_gw);                    // [wtot + dim] (GRIDW)
    double* dvn_s = (double*)(smem_b + p.o_dvn);                  // [CH]
    int* n_s = (int*)(smem_b + p.o_n);                            // [CH]
    H.cnt = (unsigned*)(smem_b + p.o_hcnt);                       // [wtot]
    dig_t* y0_s = (dig_t*)(smem_b + p.o_y0);                      // [CH][dim]
    __shared__ long long scan_s[NW];
    __shared__ double red_s[NW];
    __shared__ uint32_t base_s[VB_MAXD];
    __shared__ long long next_s;
    __shared__ int sub_s[2];
    __shared__ int nlarge_s, large_s[VB_LARGE_MAX];
    __shared__ double p1_s[VB_LARGE_G * NW * NF], p2_s[VB_LARGE_G * NW * (NF + NV)];
    __shared__ int wnew_s[VB_MAXD], wneed_s[VB_MAXD];
    // cube of a sample without searching: bit j of sb_s[w] is set when a cube starts at sample 32 w + j
    // of the tile, pc_s[w] counts the starts before word w
    __shared__ uint32_t sb_s[NT + 1];
    __shared__ int pc_s[NT + 1];
    __shared__ int word_s;                                          // next unclaimed word of the tile (phase 1)
    int* const wlo_s = vb_wlo_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dim = p.map.dim;
    const bool correlate = (p.flags & VBF_CORRELATE) != 0;
    CubeAcc<NF> A;
    A.clear();
    if (Src::USES_EXP) vb_exp_init();                             // visible after the first barrier of the chunk loop
    for (int i = tid; i < p.wtot; i += NT) { H.sum[i] = 0.0; H.cnt[i] = 0u; }
    if (tid < VB_MAXD) wlo_s[tid] = -0x40000000;                   // no window yet: the first chunk installs them
    long long since_flush = 0;                                    // samples added since the last full flush

    // Work is claimed dynamically (one atomic per item) so that CTAs finish together -- and one item AHEAD:
    // thread 0 asks for item k+1 when item k starts and looks at the answer when item k is done, so the
    // atomic's round trip to L2 and the chunk search hide behind the item's work.
    auto resolve = [&](long long g) {
        if (g >= p.item_end) { next_s = p.chunk_end; return; }
        long long c; int sub, nsub;
        locate_item(p, g, c, sub, nsub);
        next_s = c; sub_s[0] = sub; sub_s[1] = nsub;
    };
    if (tid == 0) resolve(p.item_begin + (long long)atomicAdd(p.work_counter, 1ull));
    for (;;) {
        __syncthreads();                       // previous item fully consumed, next_s written
        const int64_t lc = next_s;
        if (lc >= p.chunk_end) break;
        long long g_ahead = 0;
        if (tid == 0) g_ahead = p.item_begin + (long long)atomicAdd(p.work_counter, 1ull);
        const int sub = sub_s[0], nsub = sub_s[1];                 // this CTA does part `sub` of `nsub` of the chunk
        const int64_t lh0 = lc * CH;
        const int64_t h0 = local_to_global(p.st, lh0);
        if (p.wtot > 0) {
            // ---- move the windows to this chunk's strata (flush the histogram of the ones that change)
            const bool force = since_flush > 0x40000000LL;         // keep the u32 counts far from overflow
            hist_move_windows<NT>(p, H, h0, CH, force, wnew_s, wneed_s);
            if (Src::GRIDW) {
                // grid nodes lo .. lo + wcap of the moved windows (nodes past the axis end repeat the last one)
                for (int d = 0; d < dim; ++d) {
                    if (!wneed_s[d]) continue;
                    const int lo = wnew_s[d], ni = p.map.ninc[d];
                    const double* g = p.map.grid + (size_t)d * p.map.gstride;
                    double* w = gw_s + p.woff[d] + d;
                    for (int i = tid; i <= p.wcap[d]; i += NT) w[i] = __ldg(g + min(lo + i, ni));
                }
            }
            if (force) since_flush = 0;
        }
        const long long total = chunk_setup<NT, CH, dig_t>(p, lh0, h0, ex_s, n_s, y0_s, base_s, scan_s, dvn_s);
        const int64_t chunk_row = p.chunk_off ? p.chunk_off[lc] - p.row0 : 0;

        int c0, cend;
        item_cubes(ex_s, CH, total, sub, nsub, c0, cend);
        since_flush += ex_s[cend] - ex_s[c0];
        while (c0 < cend) {
            const long long base = ex_s[c0];
            if (base >= total) break;                              // only empty cubes remain
            // c1 = one past the last cube of this tile.  The item's remaining samples are cut into the
            // fewest tiles that fit the staging buffer, of about equal size (not one full tile and a
            // sliver: every tile costs the same barriers), so the limit is remaining / ntiles <= cap.
            int c1;
            {
                const long long rem = ex_s[cend] - base;
                const long long ntile = (rem + p.cap - 1) / p.cap;
                long long lim = ntile > 1 ? (rem + ntile - 1) / ntile : (long long)p.cap;
                if (lim > (long long)p.cap) lim = p.cap;
                int lo = c0, hi = cend + 1;                     // ex_s[lo]-base <= lim < ex_s[hi]-base (virtual)
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (ex_s[mid] - base <= lim) lo = mid; else hi = mid;
                }
                c1 = lo;
                if (c1 == c0 && ex_s[c0 + 1] - base <= (long long)p.cap) c1 = c0 + 1;   // one cube above the even share but within the buffer
            }
            if (c1 == c0) {
                // ---- giant cube c0: staged through global scratch, reduced by the whole CTA
                const int n = n_s[c0];
                const int64_t h = h0 + c0;
                double* gs = p.scratch + (size_t)blockIdx.x * NF * p.scratch_stride;
                double S[NF];
#pragma unroll
                for (int s = 0; s < NF; ++s) S[s] = 0.0;
                for (int kb = 0; kb < n; kb += NT) {
                    const int k = kb + tid;
                    if (k < n) {
                        double w[NF];
                        const SampleRef<dig_t> sr{n, dvn_s[c0], h, (uint32_t)k, chunk_row + base + k, y0_s + c0 * dim};
                        src.sample(p, sr, w);
#pragma unroll
                        for (int s = 0; s < NF; ++s) { gs[s * p.scratch_stride + k] = w[s]; S[s] += w[s]; }
                    }
                    __syncwarp();
                }
                // both passes end in ONE block-wide reduction of all their sums (warp sums -> shared
                // memory -> fixed-order sum over the warps): 3 barriers per giant cube
                double m[NF];
#pragma unroll
                for (int s = 0; s < NF; ++s) {
                    const double t = warp_sum(S[s]);
                    if (lane == 0) p1_s[warp * NF + s] = t;
                }
                __syncthreads();
#pragma unroll
                for (int s = 0; s < NF; ++s) {
                    double t = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) t += p1_s[w * NF + s];
                    S[s] = t;
                    m[s] = t / (double)n;
                }
                double sd[NF], q[NV];
#pragma unroll
                for (int s = 0; s < NF; ++s) sd[s] = 0.0;
#pragma unroll
                for (int v = 0; v < NV; ++v) q[v] = 0.0;
                for (int k = tid; k < n; k += NT) {
                    double w[NF];
#pragma unroll
                    for (int s = 0; s < NF; ++s) w[s] = gs[s * p.scratch_stride + k];
                    pass2_sample<NF>(w, m, correlate, sd, q);
                }
#pragma unroll
                for (int s = 0; s < NF; ++s) {
                    const double t = warp_sum(sd[s]);
                    if (lane == 0) p2_s[warp * (NF + NV) + s] = t;
                }
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const double t = warp_sum(q[v]);
                    if (lane == 0) p2_s[warp * (NF + NV) + NF + v] = t;
                }
                __syncthreads();
                if (tid == 0) {
#pragma unroll
                    for (int s = 0; s < NF; ++s) sd[s] = 0.0;
#pragma unroll
                    for (int v = 0; v < NV; ++v) q[v] = 0.0;
                    for (int w = 0; w < NW; ++w) {
#pragma unroll
                        for (int s = 0; s < NF; ++s) sd[s] += p2_s[w * (NF + NV) + s];
#pragma unroll
                        for (int v = 0; v < NV; ++v) q[v] += p2_s[w * (NF + NV) + NF + v];
                    }
                    double sigf2 = cube_finish<NF>(A, n, S, sd, q, correlate);
                    cube_epilogue<NF, dig_t>(p, H, A, sigf2, lh0 + c0, h, n, y0_s + c0 * dim, chunk_row + base + n - 1);
                }
                __syncthreads();
                c0 += 1;
                continue;
            }
            const int Tt = (int)(ex_s[c1] - base);

            // ---- which cube does sample i of the tile belong to?  Start bits + their running count per
            // 32-sample word replace a per-sample binary search of the prefix array (round 1: 8 % of the
            // light kernel's instructions).  Cubes with samples are contiguous (only the padding cubes
            // after the rank's last one are empty), so the j-th start is cube c0 + j.
            const int nword = (Tt + 31) >> 5;                      // <= NT (launcher: cap <= 32 NT)
            if (tid <= nword) sb_s[tid] = 0u;
            if (tid == 0) word_s = 0;
            __syncthreads();
            for (int c = c0 + tid; c < c1; c += NT)
                if (n_s[c] > 0) {
                    const int o = (int)(ex_s[c] - base);
                    atomicOr(&sb_s[o >> 5], 1u << (o & 31));
                }
            __syncthreads();
            {
                long long tot_unused;
                const long long ex = block_exscan<NT>((long long)(tid < nword ? __popc(sb_s[tid]) : 0), scan_s, &tot_unused);
                if (tid < nword) pc_s[tid] = (int)ex;
            }
            __syncthreads();

            // ---- phase 1: one thread per sample.  Warps CLAIM the tile's 32-sample words one at a time
            // (a shared counter) instead of striding over them: a static split leaves the warps that run
            // faster -- their sub-partition is shared with other CTAs' warps -- waiting at the barrier
            // below for the slowest one (ncu, N = 1000 ridge: 12 % of all warp samples sat there, the
            // FP64 pipe 83 % busy); with claiming they all finish within one word of each other.
            // (Sampler-bound kernels keep the static stride: N = 1 ridge 6.1 ms against 6.8 ms with claiming.)
            for (int wd_static = warp;; wd_static += NW) {
                int wd = wd_static;
                if (Src::CLAIM) {
                    if (lane == 0) wd = atomicAdd(&word_s, 1);
                    wd = __shfl_sync(0xffffffffu, wd, 0);
                }
                if (wd >= nword) break;
                const int i = (wd << 5) + lane;
                if (i < Tt) {
                    const uint32_t m = sb_s[i >> 5] & (0xffffffffu >> (31 - lane));    // starts at or before this lane (word = the warp's 32 samples)
                    const int c = c0 + pc_s[i >> 5] + __popc(m) - 1;
                    const int k = i - (int)((unsigned)ex_s[c] - (unsigned)base);       // offsets inside a tile fit 32 bits
                    double w[NF];
                    const SampleRef<dig_t> sr{n_s[c], dvn_s[c], h0 + c, (uint32_t)k, chunk_row + base + i, y0_s + c * dim};
                    src.sample(p, sr, w);
#pragma unroll
                    for (int s = 0; s < NF; ++s) wf_s[(size_t)s * p.cap + i] = w[s];
                }
                __syncwarp();                                      // lanes leave the histogram loops at different times
            }
            if (tid == 0) nlarge_s = 0;
            __syncthreads();

            // ---- phase 2a: one thread per small cube, serial in the reference's order
            for (int c = c0 + tid; c < c1; c += NT) {
                const int n = n_s[c];
                if (n > 0 && n <= VB_WARP_CUBE) {
                    const int o = (int)(ex_s[c] - base);
                    double S[NF], m[NF], sd[NF], q[NV];
#pragma unroll
                    for (int s = 0; s < NF; ++s) { S[s] = 0.0; sd[s] = 0.0; }
#pragma unroll
                    for (int v = 0; v < NV; ++v) q[v] = 0.0;
                    for (int k = 0; k < n; ++k)
#pragma unroll
                        for (int s = 0; s < NF; ++s) S[s] += wf_s[(size_t)s * p.cap + o + k];
#pragma unroll
                    for (int s = 0; s < NF; ++s) m[s] = S[s] / (double)n;
                    for (int k = 0; k < n; ++k) {
                        double w[NF];
#pragma unroll
                        for (int s = 0; s < NF; ++s) w[s] = wf_s[(size_t)s * p.cap + o + k];
                        pass2_sample<NF>(w, m, correlate, sd, q);
                    }
                    double sigf2 = cube_finish<NF>(A, n, S, sd, q, correlate);
                    cube_epilogue<NF, dig_t>(p, H, A, sigf2, lh0 + c, h0 + c, n, y0_s + c * dim, chunk_row + ex_s[c] + n - 1);
                } else if (n > VB_WARP_CUBE) {
                    large_s[atomicAdd(&nlarge_s, 1)] = c;          // at most cap / (VB_WARP_CUBE + 1) per tile
                }
            }
            __syncthreads();
            // ---- phase 2b: large cubes, each shared by ALL warps (warp w reduces the w-th slice of its
            // samples; partial sums meet in shared memory).  One warp per cube left the other warps
            // idle at the barrier whenever a tile held fewer large cubes than warps -- the common
            // case once the vegas+ allocation concentrates samples.
            const int nlarge = nlarge_s;
            for (int g0 = 0; g0 < nlarge; g0 += VB_LARGE_G) {
                const int ng = min(VB_LARGE_G, nlarge - g0);
                for (int j = 0; j < ng; ++j) {
                    const int c = large_s[g0 + j], n = n_s[c], o = (int)(ex_s[c] - base);
                    const int lo = (int)((long long)n * warp / NW), hi = (int)((long long)n * (warp + 1) / NW);
                    double S[NF];
#pragma unroll
                    for (int s = 0; s < NF; ++s) S[s] = 0.0;
                    for (int k = lo + lane; k < hi; k += 32)
#pragma unroll
                        for (int s = 0; s < NF; ++s) S[s] += wf_s[(size_t)s * p.cap + o + k];
#pragma unroll
                    for (int s = 0; s < NF; ++s) {
                        const double t = warp_sum(S[s]);
                        if (lane == 0) p1_s[(j * NW + warp) * NF + s] = t;
                    }
                }
                __syncthreads();
                for (int j = 0; j < ng; ++j) {
                    const int c = large_s[g0 + j], n = n_s[c], o = (int)(ex_s[c] - base);
                    const int lo = (int)((long long)n * warp / NW), hi = (int)((long long)n * (warp + 1) / NW);
                    double m[NF], sd[NF], q[NV];
#pragma unroll
                    for (int s = 0; s < NF; ++s) {
                        double t = 0.0;
#pragma unroll
                        for (int w = 0; w < NW; ++w) t += p1_s[(j * NW + w) * NF + s];
                        m[s] = t / (double)n;
                        sd[s] = 0.0;
                    }
#pragma unroll
                    for (int v = 0; v < NV; ++v) q[v] = 0.0;
                    for (int k = lo + lane; k < hi; k += 32) {
                        double w[NF];
#pragma unroll
                        for (int s = 0; s < NF; ++s) w[s] = wf_s[(size_t)s * p.cap + o + k];
                        pass2_sample<NF>(w, m, correlate, sd, q);
                    }
#pragma unroll
                    for (int s = 0; s < NF; ++s) {
                        const double t = warp_sum(sd[s]);
                        if (lane == 0) p2_s[(j * NW + warp) * (NF + NV) + s] = t;
                    }
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const double t = warp_sum(q[v]);
                        if (lane == 0) p2_s[(j * NW + warp) * (NF + NV) + NF + v] = t;
                    }
                }
                __syncthreads();
                if (tid < ng) {
                    const int c = large_s[g0 + tid], n = n_s[c];
                    double S[NF], sd[NF], q[NV];
#pragma unroll
                    for (int s = 0; s < NF; ++s) { S[s] = 0.0; sd[s] = 0.0; }
#pragma unroll
                    for (int v = 0; v < NV; ++v) q[v] = 0.0;
                    for (int w = 0; w < NW; ++w) {
#pragma unroll
                        for (int s = 0; s < NF; ++s) { S[s] += p1_s[(tid * NW + w) * NF + s]; sd[s] += p2_s[(tid * NW + w) * (NF + NV) + s]; }
#pragma unroll
                        for (int v = 0; v < NV; ++v) q[v] += p2_s[(tid * NW + w) * (NF + NV) + NF + v];
                    }
                    double sigf2 = cube_finish<NF>(A, n, S, sd, q, correlate);
                    cube_epilogue<NF, dig_t>(p, H, A, sigf2, lh0 + c, h0 + c, n, y0_s + c * dim, chunk_row + ex_s[c] + n - 1);
                }
                __syncthreads();
            }
            __syncthreads();
            c0 = c1;
        }
        if (tid == 0) resolve(g_ahead);        // (every thread read next_s / sub_s of this item long ago)
    }

    if (p.wtot > 0) hist_flush<NT>(p, H, nullptr);               // the loop exits through a barrier

    // ---- per-CTA partial sums (fixed tree inside the CTA), finished by k_finalize (fixed order over the CTAs)
    constexpr int NACC = NF + NV + 1;
    double* out = p.partials + (size_t)blockIdx.x * NACC;
#pragma unroll
    for (int s = 0; s < NF; ++s) { double t = block_sum<NT>(A.mean[s], red_s); if (tid == 0) out[s] = t; }
#pragma unroll
    for (int v = 0; v < NV; ++v) { double t = block_sum<NT>(A.var[v], red_s); if (tid == 0) out[NF + v] = t; }
    { double t = block_sum<NT>(A.sum_sigf, red_s); if (tid == 0) out[NF + NV] = t; }
}
