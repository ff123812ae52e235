// profile.cu -- built-in functors on buffers (vb200_eval_integrand) and the stratification profile of
// vegas.restratify (vb200_dy_profile; src/vegas/__init__.py:1314-1419).
#include "ctx.h"

// ---------------------------------------------------------------------------------------------
// built-in functor on buffers
// ---------------------------------------------------------------------------------------------
extern "C" int vb200_eval_integrand(vb200_ctx* c, const double* x_dev, int64_t rows, double* f_dev, void* stream)
{
    if (!c || !x_dev || !f_dev) return fail(-1, "vb200_eval_integrand: null argument");
    if (c->fid < 0 || !c->have_map) return fail(-1, "vb200_eval_integrand: no integrand / map set");
    if (rows <= 0) return 0;
    CK(cudaSetDevice(c->device));
    const void* f = c->functor.data();
    const int dim = c->map.dim;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    switch (c->fid) {
    case VB200_F_POLY: rc = eval_poly(f, dim, x_dev, rows, f_dev, c->sm_count, st); break;
    case VB200_F_GAUSS_MIX: rc = eval_gaussmix(f, dim, x_dev, rows, f_dev, c->sm_count, st); break;
    case VB200_F_RIDGE: rc = eval_ridge(f, dim, x_dev, rows, f_dev, c->sm_count, st); break;
    case VB200_F_PATHINT: rc = eval_pathint(f, c->nx0, dim, x_dev, rows, f_dev, c->sm_count, st); break;
    default: rc = eval_genz(f, dim, x_dev, rows, f_dev, c->sm_count, st); break;
    }
    if (rc == -22) return fail(-4, "vb200_eval_integrand: no kernel compiled for dim=%d integrand=%d", dim, c->fid);
    if (rc) return fail(-2, "vb200_eval_integrand: launch failed (%s)", cudaGetErrorString((cudaError_t)(-(rc + 1000))));
    c->launches += 1;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Stratification profile (vegas.restratify, __init__.py:1314-1419): the auxiliary integrand there
// has components dI[mu][i] = f(x) * [yst[i] <= y_mu <= yst[i+1]] (one-hot in each axis' y-bin), and
// the iteration computes mean and variance of each through the usual per-hypercube two-pass
// (pyx:2142-2186, correlate_integrals=False).  Here the D*ndy components are never materialised:
// this kernel re-derives y from the Philox counter, takes w*f from the callback path's buffers and
// accumulates, per hypercube, axis and OCCUPIED bin, exactly those two passes (a component that
// is zero on every sample of a cube contributes nothing).  acc[(mu*ndy + i)*2 + {0,1}] += {mean, var}.
//   small cubes (<= 32 samples): one thread per cube, samples in the reference's order;
//   larger cubes: one warp per cube.
// Sums are kept per warp in shared memory (only lane 0 of a warp writes its array), added to acc
// with fp64 atomics at the end.
// ---------------------------------------------------------------------------------------------
#define VB_DY_MAX 32
struct DyP {
    int ndy;
    double yst[VB_DY_MAX + 1];     // numpy.linspace(0, 1, ndy + 1)
    const double* f; int fstride;  // f[row * fstride]  (component 0 of the integrand)
    const double* w;               // wgt[row]
    double* acc;                   // [dim][ndy][2]
};

// bins y belongs to (closed intervals as in the reference: a y on a boundary is in two bins)
__device__ __forceinline__ uint32_t dy_mask(const DyP& q, double y)
{
    int i = __double2int_rd(y * (double)q.ndy);
    i = max(0, min(i, q.ndy - 1));
    uint32_t m = 0;
#pragma unroll
    for (int j = -1; j <= 1; ++j) {
        const int b = i + j;
        if (b >= 0 && b < q.ndy && q.yst[b] <= y && y <= q.yst[b + 1]) m |= 1u << b;
    }
    return m;
}

__global__ void __launch_bounds__(VB_ENT) k_dy_profile(const __grid_constant__ EngineP p, const __grid_constant__ DyP q)
{
    constexpr int NT = VB_ENT, NW = NT / 32, CH = VB_CH;
    __shared__ long long ex_s[CH + 1];
    __shared__ int n_s[CH];
    __shared__ long long scan_s[NW];
    __shared__ uint32_t base_s[VB_MAXD];
    __shared__ long long next_s;
    __shared__ int sub_s[2];
    extern __shared__ double dy_dyn[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dim = p.map.dim, ndy = q.ndy;
    const int nacc = dim * ndy * 2;
    double* accw = dy_dyn + (size_t)warp * nacc;                        // [dim][ndy][2] of this warp
    uint32_t* y0_s = (uint32_t*)(dy_dyn + (size_t)NW * nacc);           // [CH][dim]
    for (int i = tid; i < NW * nacc; i += NT) dy_dyn[i] = 0.0;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const long long g = p.item_begin + (long long)atomicAdd(p.work_counter, 1ull);
            if (g >= p.item_end) next_s = p.chunk_end;
            else {
                long long c; int sb, ns;
                locate_item(p, g, c, sb, ns);
                next_s = c; sub_s[0] = sb; sub_s[1] = ns;
            }
        }
        __syncthreads();
        const int64_t lc = next_s;
        if (lc >= p.chunk_end) break;
        const int sub = sub_s[0], nsub = sub_s[1];
        const int64_t lh0 = lc * CH, h0 = local_to_global(p.st, lh0);
        const long long total = chunk_setup<NT, CH, uint32_t>(p, lh0, h0, ex_s, n_s, y0_s, base_s, scan_s);
        const int64_t chunk_row = p.chunk_off[lc] - p.row0;
        int c0, cend;
        item_cubes(ex_s, CH, total, sub, nsub, c0, cend);

        // ---- small cubes: one thread per cube (warp-uniform loop: the warp reduces together)
        for (int cb = c0; cb < cend; cb += NT) {
            const int c = cb + tid;
            const int n = (c < cend && n_s[c] <= 32) ? n_s[c] : 0;
            const int64_t row = chunk_row + (c < CH ? ex_s[c] : 0);
            const int64_t h = h0 + c;
            double wf[32];
            uint32_t mk[2][32];
            for (int k = 0; k < n; ++k) wf[k] = q.w[row + k] * q.f[(row + k) * q.fstride];
            for (int pr = 0; 2 * pr < dim; ++pr) {
                for (int k = 0; k < n; ++k) {
                    double ua, ub;
                    philox_pair(p.key, p.itn, h, (uint32_t)k, pr, ua, ub);
                    mk[0][k] = dy_mask(q, div_exact((double)y0_s[c * dim + 2 * pr] + ua, p.st.dns[2 * pr], p.st.rns[2 * pr]));
                    mk[1][k] = (2 * pr + 1 < dim)
                        ? dy_mask(q, div_exact((double)y0_s[c * dim + 2 * pr + 1] + ub, p.st.dns[2 * pr + 1], p.st.rns[2 * pr + 1])) : 0u;
                }
                for (int e = 0; e < 2 && 2 * pr + e < dim; ++e) {
                    const int mu = 2 * pr + e;
                    uint32_t um = 0;
                    for (int k = 0; k < n; ++k) um |= mk[e][k];
                    const uint32_t wum = __reduce_or_sync(0xffffffffu, um);
                    for (int i = 0; i < ndy; ++i) {
                        const uint32_t bit = 1u << i;
                        if (!(wum & bit)) continue;                     // warp-uniform
                        double madd = 0.0, vadd = 0.0;
                        if (um & bit) {
                            double S = 0.0;
                            for (int k = 0; k < n; ++k) if (mk[e][k] & bit) S += wf[k];
                            const double mS = S / (double)n, thr = VB_EPSILON * fabs(mS);
                            double sd = 0.0, qq = 0.0;
                            for (int k = 0; k < n; ++k) {
                                double d = ((mk[e][k] & bit) ? wf[k] : 0.0) - mS;
                                if (fabs(d) < thr) { qq += thr * thr; d = 0.0; } else qq += d * d;
                                sd += d;
                            }
                            madd = S + sd;
                            vadd = ((double)n * qq - sd * sd) / ((double)n - 1.0);
                        }
                        madd = warp_sum(madd);
                        vadd = warp_sum(vadd);
                        if (lane == 0) { accw[(mu * ndy + i) * 2] += madd; accw[(mu * ndy + i) * 2 + 1] += vadd; }
                    }
                }
            }
        }
        // ---- larger cubes: one warp per cube
        for (int c = c0 + warp; c < cend; c += NW) {
            const int n = n_s[c];
            if (n <= 32) continue;
            const int64_t row = chunk_row + ex_s[c];
            const int64_t h = h0 + c;
            for (int mu = 0; mu < dim; ++mu) {
                const int pr = mu >> 1;
                const uint32_t y0 = y0_s[c * dim + mu];
                uint32_t um = 0;
                for (int k = lane; k < n; k += 32) {
                    double ua, ub;
                    philox_pair(p.key, p.itn, h, (uint32_t)k, pr, ua, ub);
                    um |= dy_mask(q, div_exact((double)y0 + ((mu & 1) ? ub : ua), p.st.dns[mu], p.st.rns[mu]));
                }
                um = __reduce_or_sync(0xffffffffu, um);
                for (int i = 0; i < ndy; ++i) {
                    const uint32_t bit = 1u << i;
                    if (!(um & bit)) continue;
                    double S = 0.0;
                    for (int k = lane; k < n; k += 32) {
                        double ua, ub;
                        philox_pair(p.key, p.itn, h, (uint32_t)k, pr, ua, ub);
                        if (dy_mask(q, div_exact((double)y0 + ((mu & 1) ? ub : ua), p.st.dns[mu], p.st.rns[mu])) & bit)
                            S += q.w[row + k] * q.f[(row + k) * q.fstride];
                    }
                    S = warp_sum(S);
                    const double mS = S / (double)n, thr = VB_EPSILON * fabs(mS);
                    double sd = 0.0, qq = 0.0;
                    for (int k = lane; k < n; k += 32) {
                        double ua, ub;
                        philox_pair(p.key, p.itn, h, (uint32_t)k, pr, ua, ub);
                        const bool in = dy_mask(q, div_exact((double)y0 + ((mu & 1) ? ub : ua), p.st.dns[mu], p.st.rns[mu])) & bit;
                        double d = (in ? q.w[row + k] * q.f[(row + k) * q.fstride] : 0.0) - mS;
                        if (fabs(d) < thr) { qq += thr * thr; d = 0.0; } else qq += d * d;
                        sd += d;
                    }
                    sd = warp_sum(sd);
                    qq = warp_sum(qq);
                    if (lane == 0) {
                        accw[(mu * ndy + i) * 2] += S + sd;
                        accw[(mu * ndy + i) * 2 + 1] += ((double)n * qq - sd * sd) / ((double)n - 1.0);
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < nacc; i += NT) {
        double t = 0.0;
        for (int w = 0; w < NW; ++w) t += dy_dyn[(size_t)w * nacc + i];
        if (t != 0.0) atomicAdd(q.acc + i, t);
    }
}

extern "C" int vb200_dy_profile(vb200_ctx* c, uint32_t itn, int64_t chunk_begin, int64_t chunk_end, const double* f_dev,
                                int fstride, const double* wgt_dev, int ndy, const double* yst_host, double* acc_dev, void* stream)
{
    if (!c || !f_dev || !wgt_dev || !yst_host || !acc_dev) return fail(-1, "vb200_dy_profile: null argument");
    if (!c->have_map || !c->have_strata || !c->have_plan) return fail(-1, "vb200_dy_profile: map/strata/plan not set");
    if (ndy < 1 || ndy > VB_DY_MAX) return fail(-1, "vb200_dy_profile: ndy=%d outside 1..%d", ndy, VB_DY_MAX);
    if (fstride < 1) return fail(-1, "vb200_dy_profile: fstride < 1");
    if (chunk_begin < 0 || chunk_end > c->nchunks || chunk_begin > chunk_end) return fail(-1, "vb200_dy_profile: bad chunk range");
    CK(cudaSetDevice(c->device));
    if (chunk_begin == chunk_end) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    EngineP p;
    memset(&p, 0, sizeof p);
    p.map = c->map; p.st = c->st; p.al = c->al; p.key = c->key;
    p.itn = itn;
    for (int d = 0; d < VB_MAXD; ++d) p.cstride[d] = c->cstride[d];
    p.chunk_begin = chunk_begin; p.chunk_end = chunk_end;
    p.chunk_off = (const int64_t*)c->chunk_off.p;
    int rc = vb_fetch_chunk_off(c, st);
    if (rc) return rc;
    p.row0 = c->chunk_off_host[(size_t)chunk_begin];
    ItemsSel it;
    rc = vb_set_items(c, chunk_begin, chunk_end, it, st);
    if (rc) return rc;
    p.item_off = it.off[0]; p.item_begin = it.begin[0]; p.item_end = it.end[0];
    CK(c->counter.ensure(sizeof(unsigned long long)));
    CK(cudaMemsetAsync(c->counter.p, 0, sizeof(unsigned long long), st));
    p.work_counter = (unsigned long long*)c->counter.p;
    DyP q;
    memset(&q, 0, sizeof q);
    q.ndy = ndy;
    for (int i = 0; i <= ndy; ++i) q.yst[i] = yst_host[i];
    q.f = f_dev; q.fstride = fstride; q.w = wgt_dev; q.acc = acc_dev;
    const int dim = c->map.dim;
    size_t smem = sizeof(double) * (size_t)(VB_ENT / 32) * dim * ndy * 2 + sizeof(uint32_t) * (size_t)VB_CH * dim;
    CK(cudaFuncSetAttribute(k_dy_profile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t g = (int64_t)c->sm_count * 4, nitems = p.item_end - p.item_begin;
    if (g > nitems) g = nitems;
    k_dy_profile<<<(int)g, VB_ENT, smem, st>>>(p, q);
    c->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

