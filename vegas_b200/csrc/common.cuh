// common.cuh -- shared device helpers for the vegas_b200 engine (sm_100a).
//
// Philox4x32-10 counter-based RNG, exact fp64 division by a stratum count, the parameter
// blocks passed to kernels by value, and small warp/block reduction helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define VB_MAXD 32          // compile-time bound on dimensions carried in kernel parameters
#define VB_NT 256           // threads per CTA of the engine kernels
#define VB_CH 256           // hypercubes per chunk
#ifndef VB_ENT
#define VB_ENT 128          // threads per CTA of the engine kernel (VB_CH / VB_ENT cubes per thread in set-up)
#endif
#define VB_ITEM 4096        // samples per work item: chunks holding more are cut into several items (k_plan)
#define VB_MAXITEMS 4096    // most items one chunk is cut into
#define VB_WARP_CUBE 64     // cubes with more samples than this are reduced by a whole warp
#define VB_EPSILON (2.220446049250313e-16 * 1e4)   // reference EPSILON, _vegas.pyx:36

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  key = seed; counter = (k, itn<<8 | pair, h_lo, h_hi).
// One call yields two 52-bit uniforms in [0,1): u = ((hi:lo) >> 12) * 2^-52, built by
// or-ing the mantissa under the exponent of 1.0 and subtracting 1.0 (exact).
// ---------------------------------------------------------------------------------------------
struct PhiloxKey { uint32_t k[20]; };   // per-round keys, precomputed on the host (uniform)

__host__ __device__ inline void philox_make_key(uint64_t seed, PhiloxKey& K)
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        K.k[2 * r] = k0; K.k[2 * r + 1] = k1;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

__device__ __forceinline__ void philox4x32_10(const PhiloxKey& K, uint32_t c0, uint32_t c1,
                                              uint32_t c2, uint32_t c3, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ K.k[2 * r];
        uint32_t n2 = hi0 ^ c3 ^ K.k[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double u52(uint32_t lo, uint32_t hi)
{
    // ((hi:lo) >> 12) or-ed under the exponent of 1.0 -> [1,2); minus 1.0 is exact
    uint32_t mlo = (lo >> 12) | (hi << 20);
    uint32_t mhi = (hi >> 12) | 0x3FF00000u;
    return __hiloint2double((int)mhi, (int)mlo) - 1.0;
}

// two uniforms for (cube h, sample k, iteration itn, axis pair p)
__device__ __forceinline__ void philox_pair(const PhiloxKey& K, uint32_t itn, int64_t h, uint32_t k,
                                            int p, double& ua, double& ub)
{
    uint32_t r[4];
    philox4x32_10(K, k, ((itn & 0xFFFFFFu) << 8) | (uint32_t)p, (uint32_t)(uint64_t)h,
                  (uint32_t)((uint64_t)h >> 32), r);
    ua = u52(r[0], r[1]);
    ub = u52(r[2], r[3]);
}

// Correctly rounded a / b for b a small positive integer (as double), rb = RN(1/b).
// q = RN(a*rb) is within 1 ulp; one FMA residual step then gives RN(a/b) (Markstein).
__device__ __forceinline__ double div_exact(double a, double b, double rb)
{
    double q = __dmul_rn(a, rb);
    double e = __fma_rn(-q, b, a);
    return __fma_rn(e, rb, q);
}


// ---------------------------------------------------------------------------------------------
// exp() for W independent arguments evaluated in lock-step, so that the W dependent chains
// interleave in the FP64 pipe.  Table-driven: x * 128/ln2 = 128 k + j + f, |f| <= 1/2;
// r = x - (128 k + j) ln2/128 (two-step Cody-Waite, |r| <= ln2/256);
// exp(x) = 2^k * T[j] * P(r) with T[j] = 2^(j/128) correctly rounded and P a degree-5 near-minimax
// polynomial (relative error 1.7e-20).  10 FP64 instructions (8 DFMA, 1 DADD, 1 DMUL) against 16 in
// CUDA's exp(); error <= 1.5 ulp.  Constants: tools/exp_table.py.
// |x| >= 708 (overflow / underflow / denormal results, inf, NaN) takes the library exp().
//
// The table lives in shared memory (per-lane indices would serialise in the constant cache) in
// VB_EXP_COPIES interleaved copies: entry j of copy l at [j * COPIES + l], lane L reads copy
// L % COPIES.  With 16 copies (16 KB) the 16 lanes of a half-warp -- the unit an 8-byte shared load
// is served in -- always hit 16 different bank pairs, whatever their j: no bank conflicts (a single
// copy: 62 % of the kernel's shared-memory wavefronts were conflict replays, ncu round 1).
// ---------------------------------------------------------------------------------------------
#include "exp_table.inc"
#ifndef VB_EXP_COPIES
#define VB_EXP_COPIES 16
#endif
__constant__ double vb_exp_c[6] = VB_EXP_POLY;
__device__ const double vb_exp_tab_g[128] = VB_EXP_TABLE;
static __shared__ double vb_exp_tab_s[128 * VB_EXP_COPIES];

// called by all threads of a CTA before the first vb_exp*(); needs a barrier after.
// The shared copy holds T[j] with j << 13 subtracted from its high word: the reader adds n << 13 =
// (128 k + j) << 13 to it and gets T[j] * 2^k in one integer multiply-add (no masking of n).
__device__ __forceinline__ void vb_exp_init()
{
    for (int i = threadIdx.x; i < 128 * VB_EXP_COPIES; i += blockDim.x) {
        const int j = i / VB_EXP_COPIES;
        const double t = vb_exp_tab_g[j];
        vb_exp_tab_s[i] = __hiloint2double(__double2hiint(t) - (j << 13), __double2loint(t));
    }
}

template <int W>
__device__ __forceinline__ void vb_exp_n(const double (&x)[W], double (&e)[W])
{
    const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52: rounds to nearest integer
    constexpr int CSHIFT = VB_EXP_COPIES == 16 ? 7 : (VB_EXP_COPIES == 1 ? 3 : -1);   // log2(8 * COPIES)
    static_assert(CSHIFT > 0, "VB_EXP_COPIES must be 1 or 16");
    // this lane's copy of the table, as a shared-state-space address
    const unsigned tab = (unsigned)__cvta_generic_to_shared(vb_exp_tab_s) + (threadIdx.x & (VB_EXP_COPIES - 1)) * 8u;
    double t[W], r[W], p[W];
#pragma unroll
    for (int j = 0; j < W; ++j) t[j] = __fma_rn(x[j], VB_EXP_INVL, MAGIC);
#pragma unroll
    for (int j = 0; j < W; ++j) {
        double kf = t[j] - MAGIC;
        r[j] = __fma_rn(kf, -VB_EXP_LHEAD, x[j]);
        r[j] = __fma_rn(kf, -VB_EXP_LTAIL, r[j]);
    }
#pragma unroll
    for (int j = 0; j < W; ++j) p[j] = __fma_rn(vb_exp_c[0], r[j], vb_exp_c[1]);
#pragma unroll
    for (int i = 2; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j < W; ++j) p[j] = __fma_rn(p[j], r[j], vb_exp_c[i]);
    }
    // |x| >= 708 shows in the high word of x: as an unsigned number for negative x, as a signed one for positive x
    unsigned umax = 0u;
    int smax = 0;
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const int n = __double2loint(t[j]);                  // 128 k + j (two's complement in the low word)
        // 3 integer instructions and one shared load per argument: table address (mask, multiply-add), scaled entry
        unsigned sa, thi, tlo;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(sa) : "r"((unsigned)n & 127u), "n"(1 << CSHIFT), "r"(tab));
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(tlo), "=r"(thi) : "r"(sa));
        asm("mad.lo.u32 %0, %1, 8192, %2;" : "=r"(thi) : "r"((unsigned)n), "r"(thi));
        e[j] = __dmul_rn(p[j], __hiloint2double((int)thi, (int)tlo));
        umax = max(umax, (unsigned)__double2hiint(x[j]));
        smax = max(smax, __double2hiint(x[j]));
    }
    if (umax >= 0xC0862000u || smax >= 0x40862000) {        // some |x| >= 708, inf or NaN: one branch for all W
#pragma unroll
        for (int j = 0; j < W; ++j)
            if ((__double2hiint(x[j]) & 0x7fffffff) >= 0x40862000) e[j] = exp(x[j]);
    }
}

__device__ __forceinline__ double vb_exp(double x)
{
    double a[1] = {x}, e[1];
    vb_exp_n<1>(a, e);
    return e[0];
}

// ---------------------------------------------------------------------------------------------
// parameter blocks
// ---------------------------------------------------------------------------------------------
struct MapP {                 // AdaptiveMap on the device: grid[d*gstride + i], i = 0..ninc[d]
    const double* grid;
    int dim;
    int gstride;
    int ninc[VB_MAXD];
};

struct StrataP {              // stratification of y-space + this rank's share of the hypercubes
    int64_t nhcube;           // global number of hypercubes
    int64_t nlocal;           // hypercubes owned by this rank (dense local index space)
    int64_t slab;             // block-cyclic slab size in cubes (multiple of VB_CH)
    int rank, world;
    int nstrat[VB_MAXD];
    double dns[VB_MAXD];      // (double) nstrat[d]
    double rns[VB_MAXD];      // RN(1 / nstrat[d])
    uint32_t nsm[VB_MAXD];    // ceil(2^32 / nstrat[d]) for 2 <= nstrat[d] < 32768, else 0 (see digit_div)
};

// v / nstrat[d] for v < nstrat[d] + 65536 / 2 (a stratum digit plus a carry of at most a chunk): a
// multiply-high by the precomputed reciprocal (exact while v * nstrat < 2^32), 0 or 1 for large strata
// counts, v itself for a single stratum -- instead of the ~24-instruction 32-bit division
__device__ __forceinline__ uint32_t digit_div(const StrataP& st, int d, uint32_t v)
{
    const uint32_t ns = (uint32_t)st.nstrat[d], m = st.nsm[d];
    return m ? __umulhi(v, m) : (ns == 1u ? v : (v >= ns ? 1u : 0u));
}

struct AllocP {               // vegas+ allocation of samples to hypercubes (_vegas.pyx:1692-1706)
    const double* sigf;       // [nlocal] or nullptr when not adaptive
    double neval_sigf;
    int min_neval_hcube;
    int max_neval_hcube;
    int uniform_neval;        // used when sigf == nullptr
};

// Block-cyclic sharding with a rotated deal: round ls of `world` consecutive slabs goes to the ranks
// in the order rotated by slab_rot(ls), so rank r's ls-th slab is slab ls * world + (r + rot) % world.
// (A plain round-robin deal resonates with the stratification when world and the strata counts
// share structure -- round 1 measured 3 % rank skew at 4 GPUs against 0.2 % at 2 and 8.)
// Host mirrors: vb200_set_strata (nlocal) and vegas_b200/_integrator.py::_local_cubes.
__host__ __device__ inline int slab_rot(int64_t ls, int world)
{
    return (int)((((uint64_t)ls * 0x9E3779B97F4A7C15ull) >> 40) % (uint64_t)world);
}

__device__ __forceinline__ int64_t local_to_global(const StrataP& s, int64_t lh)
{
    if (s.world == 1) return lh;
    const int64_t ls = lh / s.slab;
    const int j = (s.rank + slab_rot(ls, s.world)) % s.world;
    return (ls * s.world + j) * s.slab + (lh - ls * s.slab);
}

__device__ __forceinline__ int alloc_neval(const AllocP& a, int64_t lh)
{
    if (a.sigf == nullptr) return a.uniform_neval;
    double p = __dmul_rn(a.sigf[lh], a.neval_sigf);
    long long n = (long long)__double2int_rz(p) + a.min_neval_hcube;   // cvt saturates; NaN -> 0
    if (n > a.max_neval_hcube) n = a.max_neval_hcube;
    return (int)n;
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// fixed-tree block sum over NT threads; result valid in thread 0.  red: >= NT/32 doubles of smem.
template <int NT = VB_NT>
__device__ __forceinline__ double block_sum(double v, double* red)
{
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) t += red[w];
    }
    return t;
}

// block-wide exclusive scan of one value per thread (NT threads); returns the exclusive prefix,
// the total through *total.  scratch: NT/32 long longs of shared memory.
template <int NT = VB_NT>
__device__ __forceinline__ long long block_exscan(long long v, long long* scratch, long long* total)
{
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();
    if (lane == 31) scratch[w] = x;
    __syncthreads();
    long long base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) {
        long long s = scratch[i];
        if (i < w) base += s;
        tot += s;
    }
    *total = tot;
    return base + x - v;
}
