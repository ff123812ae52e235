// engine.cuh -- the vegas+ iteration engine: one persistent kernel per iteration (or per batch).
//
// Work decomposition ("hypercube tiling"):
//   chunk  = VB_CH consecutive hypercubes of this rank's share; the persistent CTAs claim chunks
//            from an atomic counter, so they all finish together whatever the sample counts.
//   tile   = maximal run of cubes inside a chunk whose samples fit the shared-memory staging
//            buffer (cap samples).  Cubes larger than cap are staged through global scratch.
//   phase 1 (thread per SAMPLE): Philox -> stratified y -> AdaptiveMap -> integrand -> w*f staged
//            in shared memory; training-histogram adds are issued here because
//            fdv2 = (J f dv_y)^2 does not depend on the cube sums.
//   phase 2 (thread per CUBE, warp per cube above VB_WARP_CUBE samples): the reference's two-pass
//            mean/variance with its EPSILON clamp, in the reference's summation order, from the
//            staged values; sigf[h] update; per-thread fp64 accumulators.
// Reference: Integrator._random_batch (_vegas.pyx:1692-1759) + Integrator.__call__
// (_vegas.pyx:2136-2197).
#pragma once
#include "common.cuh"

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
    unsigned long long* work_counter; // zeroed before the launch: next chunk to claim
    int64_t cstride[VB_MAXD];  // cstride[d] = prod_{e<d} nstrat[e]
    // unfused path
    const double* fbuf;        // [rows][nf]
    const double* wbuf;        // [rows]
    const int64_t* chunk_off;  // [nchunks+1] exclusive scan of samples per chunk
    int64_t row0;              // chunk_off[chunk_begin]
};

__device__ __forceinline__ int tri(int s, int t) { return s * (s + 1) / 2 + t; }

// y-space bin of sample (h,k) on axis d for the training histogram; -1 when y is on the boundary
// (AdaptiveMap.add_training_data skips y<=0 and y>=1, _vegas.pyx:460).
__device__ __forceinline__ int bin_of(const EngineP& p, int d, uint32_t y0, double u, double* y_out)
{
    double y = div_exact((double)y0 + u, p.st.dns[d], p.st.rns[d]);
    if (y_out) *y_out = y;
    int iy = __double2int_rd(__dmul_rn(y, (double)p.map.ninc[d]));
    return (y > 0.0 && y < 1.0) ? iy : -1;
}

__device__ __forceinline__ void hist_add(const EngineP& p, int d, int bin, double v)
{
    if (bin >= 0) {
        atomicAdd(p.sum_f + (size_t)d * p.hstride + bin, v);
        atomicAdd(p.n_f + (size_t)d * p.hstride + bin, 1ull);
    }
}

// training point of a whole cube (adapt_to_errors, _vegas.pyx:2187-2193): y of its LAST sample
static __device__ __noinline__ void train_cube(const EngineP& p, int64_t h, uint32_t klast, const uint32_t* y0, double v)
{
    for (int pr = 0; 2 * pr < p.map.dim; ++pr) {
        double ua, ub;
        philox_pair(p.key, p.itn, h, klast, pr, ua, ub);
        hist_add(p, 2 * pr, bin_of(p, 2 * pr, y0[2 * pr], ua, nullptr), fabs(v));
        if (2 * pr + 1 < p.map.dim)
            hist_add(p, 2 * pr + 1, bin_of(p, 2 * pr + 1, y0[2 * pr + 1], ub, nullptr), fabs(v));
    }
}

// ---------------------------------------------------------------------------------------------
// sample sources
// ---------------------------------------------------------------------------------------------
// Fused: everything from the Philox counter to w*f happens in registers.
template <class F, int D>
struct FusedSrc {
    static constexpr int NF = F::NF;
    F f;
    __device__ __forceinline__ void sample(const EngineP& p, int n, int64_t h, uint32_t k,
                                           int64_t /*row*/, const uint32_t* y0, double (&wf)[NF]) const
    {
        const int dim = p.map.dim;
        double x[D];
        int bin[D];
        double jac = 1.0;
#pragma unroll
        for (int pr = 0; pr < (D + 1) / 2; ++pr) {
            if (2 * pr < dim) {
                double u[2];
                philox_pair(p.key, p.itn, h, k, pr, u[0], u[1]);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int d = 2 * pr + e;
                    if (d < D && d < dim) {
                        double y;
                        int b = bin_of(p, d, y0[d], u[e], &y);
                        const int ni = p.map.ninc[d];
                        const double* g = p.map.grid + (size_t)d * p.map.gstride;
                        double t = __dmul_rn(y, (double)ni);
                        int iy = __double2int_rd(t);
                        if (iy < ni) {
                            double g0 = __ldg(g + iy), g1 = __ldg(g + iy + 1);
                            double inc = g1 - g0;
                            x[d] = __dadd_rn(g0, __dmul_rn(inc, __dsub_rn(t, (double)iy)));   // no FMA: bit-identical to pyx:354
                            jac *= inc * (double)ni;
                        } else {
                            double g0 = __ldg(g + ni - 1), g1 = __ldg(g + ni);
                            x[d] = g1;
                            jac *= (g1 - g0) * (double)ni;
                        }
                        bin[d] = b;
                    }
                }
            }
        }
        double fx[NF];
        f(x, dim, fx);
        double wgt = jac * (p.dv_y / (double)n);
        bool bad = false;
#pragma unroll
        for (int s = 0; s < NF; ++s) { wf[s] = wgt * fx[s]; bad |= isnan(fx[s]); }
        if (bad) p.status[0] = 1;
        if (p.flags & VBF_TRAIN) {
            double a = wf[0] * (double)n;
            double fdv2 = a * a;
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (d < dim) hist_add(p, d, bin[d], fdv2);
        }
    }
};

// Unfused: w and f come from HBM buffers filled by k_sample and the user's batch integrand;
// y (for the training bins) is re-derived from the Philox counter instead of round-tripping HBM.
template <int NF_>
struct BufferSrc {
    static constexpr int NF = NF_;
    __device__ __forceinline__ void sample(const EngineP& p, int n, int64_t h, uint32_t k,
                                           int64_t row, const uint32_t* y0, double (&wf)[NF]) const
    {
        double wgt = p.wbuf[row];
        bool bad = false;
#pragma unroll
        for (int s = 0; s < NF; ++s) {
            double fx = p.fbuf[row * NF + s];
            bad |= isnan(fx);
            wf[s] = wgt * fx;
        }
        if (bad) p.status[0] = 1;
        if (p.flags & VBF_TRAIN) {
            double a = wf[0] * (double)n;
            double fdv2 = a * a;
            for (int pr = 0; 2 * pr < p.map.dim; ++pr) {
                double ua, ub;
                philox_pair(p.key, p.itn, h, k, pr, ua, ub);
                hist_add(p, 2 * pr, bin_of(p, 2 * pr, y0[2 * pr], ua, nullptr), fdv2);
                if (2 * pr + 1 < p.map.dim)
                    hist_add(p, 2 * pr + 1, bin_of(p, 2 * pr + 1, y0[2 * pr + 1], ub, nullptr), fdv2);
            }
        }
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

template <int NF>
__device__ __forceinline__ void cube_epilogue(const EngineP& p, CubeAcc<NF>& A, double sigf2, int64_t lh,
                                              int64_t h, int n, const uint32_t* y0)
{
    if (p.flags & VBF_UPDATE_SIGF) {
        double sg = pow(sigf2, p.beta_half);
        p.sigf_out[lh] = sg;
        A.sum_sigf += sg;
    }
    if (p.flags & VBF_TRAIN_ERRORS) train_cube(p, h, (uint32_t)(n - 1), y0, sigf2);
}

// ---------------------------------------------------------------------------------------------
// the engine kernel
// ---------------------------------------------------------------------------------------------
template <class Src>
__global__ void __launch_bounds__(VB_ENT) k_engine(const __grid_constant__ EngineP p, const __grid_constant__ Src src)
{
    constexpr int NF = Src::NF;
    constexpr int NV = NF * (NF + 1) / 2;
    constexpr int NT = VB_ENT;
    constexpr int NW = NT / 32;
    constexpr int CPT = VB_CH / NT;                               // cubes per thread in set-up
    static_assert(VB_CH % NT == 0, "chunk must be a multiple of the CTA size");
    extern __shared__ double smem[];
    double* wf_s = smem;                                          // [NF][cap]
    long long* ex_s = (long long*)(wf_s + (size_t)NF * p.cap);    // [VB_CH + 1]
    int* n_s = (int*)(ex_s + VB_CH + 1);                          // [VB_CH]
    uint32_t* y0_s = (uint32_t*)(n_s + VB_CH);                    // [VB_CH][dim]
    __shared__ long long scan_s[NW];
    __shared__ double red_s[NW];
    __shared__ double bc_s[NF];
    __shared__ uint32_t base_s[VB_MAXD];
    __shared__ long long next_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dim = p.map.dim;
    const bool correlate = (p.flags & VBF_CORRELATE) != 0;
    CubeAcc<NF> A;
    A.clear();

    // chunks are claimed dynamically (one atomic per chunk) so that CTAs finish together
    for (;;) {
        __syncthreads();                       // previous chunk fully consumed
        if (tid == 0) next_s = p.chunk_begin + (long long)atomicAdd(p.work_counter, 1ull);
        __syncthreads();
        const int64_t lc = next_s;
        if (lc >= p.chunk_end) break;
        const int64_t lh0 = lc * VB_CH;
        const int64_t h0 = local_to_global(p.st, lh0);
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
        long long ex = block_exscan<NT>(mine, scan_s, &total);     // contains __syncthreads
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            const int c = tid * CPT + i;
            ex_s[c] = ex;
            n_s[c] = n_mine[i];
            ex += n_mine[i];
            // mixed-radix digits of cube h0+c: base digits plus c, with carries
            uint32_t carry = (uint32_t)c;
            for (int d = 0; d < dim; ++d) {
                uint32_t v = base_s[d] + carry, ns = (uint32_t)p.st.nstrat[d];
                uint32_t qd = v / ns;
                y0_s[c * dim + d] = v - qd * ns;
                carry = qd;
            }
        }
        if (tid == NT - 1) ex_s[VB_CH] = total;
        __syncthreads();
        const int64_t chunk_row = p.chunk_off ? p.chunk_off[lc] - p.row0 : 0;

        int c0 = 0;
        while (c0 < VB_CH) {
            const long long base = ex_s[c0];
            if (base >= total) break;                              // only empty cubes remain
            // c1 = one past the last cube whose samples still fit the staging buffer (all threads
            // search the same shared array: broadcast reads, no barrier)
            int c1;
            {
                int lo = c0, hi = VB_CH + 1;                       // ex_s[lo]-base <= cap < ex_s[hi]-base (virtual)
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (ex_s[mid] - base <= (long long)p.cap) lo = mid; else hi = mid;
                }
                c1 = lo;
            }
            if (c1 == c0) {
                // ---- giant cube c0: staged through global scratch, reduced by the whole CTA
                const int n = n_s[c0];
                const int64_t h = h0 + c0;
                double* gs = p.scratch + (size_t)blockIdx.x * NF * p.scratch_stride;
                double S[NF];
#pragma unroll
                for (int s = 0; s < NF; ++s) S[s] = 0.0;
                for (int k = tid; k < n; k += NT) {
                    double w[NF];
                    src.sample(p, n, h, (uint32_t)k, chunk_row + base + k, y0_s + c0 * dim, w);
#pragma unroll
                    for (int s = 0; s < NF; ++s) { gs[s * p.scratch_stride + k] = w[s]; S[s] += w[s]; }
                }
                double m[NF];
#pragma unroll
                for (int s = 0; s < NF; ++s) {
                    double t = block_sum<NT>(S[s], red_s);
                    if (tid == 0) bc_s[s] = t;
                }
                __syncthreads();
#pragma unroll
                for (int s = 0; s < NF; ++s) { S[s] = bc_s[s]; m[s] = S[s] / (double)n; }
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
                for (int s = 0; s < NF; ++s) sd[s] = block_sum<NT>(sd[s], red_s);
#pragma unroll
                for (int v = 0; v < NV; ++v) q[v] = block_sum<NT>(q[v], red_s);
                if (tid == 0) {
                    double sigf2 = cube_finish<NF>(A, n, S, sd, q, correlate);
                    cube_epilogue<NF>(p, A, sigf2, lh0 + c0, h, n, y0_s + c0 * dim);
                }
                __syncthreads();
                c0 += 1;
                continue;
            }
            const int Tt = (int)(ex_s[c1] - base);

            // ---- phase 1: one thread per sample
            for (int i = tid; i < Tt; i += NT) {
                int lo = c0, hi = c1;
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if ((int)(ex_s[mid] - base) <= i) lo = mid; else hi = mid;
                }
                const int c = lo;
                const int k = i - (int)(ex_s[c] - base);
                double w[NF];
                src.sample(p, n_s[c], h0 + c, (uint32_t)k, chunk_row + base + i, y0_s + c * dim, w);
#pragma unroll
                for (int s = 0; s < NF; ++s) wf_s[(size_t)s * p.cap + i] = w[s];
            }
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
                    cube_epilogue<NF>(p, A, sigf2, lh0 + c, h0 + c, n, y0_s + c * dim);
                }
            }
            // ---- phase 2b: one warp per large cube
            for (int c = c0 + warp; c < c1; c += NW) {
                const int n = n_s[c];
                if (n <= VB_WARP_CUBE) continue;
                const int o = (int)(ex_s[c] - base);
                double S[NF], m[NF], sd[NF], q[NV];
#pragma unroll
                for (int s = 0; s < NF; ++s) { S[s] = 0.0; sd[s] = 0.0; }
#pragma unroll
                for (int v = 0; v < NV; ++v) q[v] = 0.0;
                for (int k = lane; k < n; k += 32)
#pragma unroll
                    for (int s = 0; s < NF; ++s) S[s] += wf_s[(size_t)s * p.cap + o + k];
#pragma unroll
                for (int s = 0; s < NF; ++s) { S[s] = warp_sum(S[s]); m[s] = S[s] / (double)n; }
                for (int k = lane; k < n; k += 32) {
                    double w[NF];
#pragma unroll
                    for (int s = 0; s < NF; ++s) w[s] = wf_s[(size_t)s * p.cap + o + k];
                    pass2_sample<NF>(w, m, correlate, sd, q);
                }
#pragma unroll
                for (int s = 0; s < NF; ++s) sd[s] = warp_sum(sd[s]);
#pragma unroll
                for (int v = 0; v < NV; ++v) q[v] = warp_sum(q[v]);
                if (lane == 0) {
                    double sigf2 = cube_finish<NF>(A, n, S, sd, q, correlate);
                    cube_epilogue<NF>(p, A, sigf2, lh0 + c, h0 + c, n, y0_s + c * dim);
                }
            }
            __syncthreads();
            c0 = c1;
        }
    }

    // ---- per-CTA partial sums (fixed tree inside the CTA), finished by k_finalize in CTA order
    constexpr int NACC = NF + NV + 1;
    double* out = p.partials + (size_t)blockIdx.x * NACC;
#pragma unroll
    for (int s = 0; s < NF; ++s) { double t = block_sum<NT>(A.mean[s], red_s); if (tid == 0) out[s] = t; }
#pragma unroll
    for (int v = 0; v < NV; ++v) { double t = block_sum<NT>(A.var[v], red_s); if (tid == 0) out[NF + v] = t; }
    { double t = block_sum<NT>(A.sum_sigf, red_s); if (tid == 0) out[NF + NV] = t; }
}
