/*
 * vegas_oracle.c -- CPU restatement of the gplepage/vegas (v6.4.1) hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under vegas_b200/ links, imports or executes this
 * file; it is the checker that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg compare the CUDA path against.  It is pinned against the unmodified reference module
 * (oracle/_ref, built by oracle/Makefile) and the reference's own known-answer tests by
 * tests/test_oracle_vs_reference.py and tests/golden/.
 *
 * Every function restates one piece of /root/reference/src/vegas/_vegas.pyx ("pyx:N"),
 * in plain scalar C, compiled with -ffp-contract=off so that every multiply and add
 * rounds separately exactly as the reference's (non-FMA) C does.
 *
 * Layouts: grid[d*gstride + i] (i = 0..ninc[d]), inc is never stored (inc[d,i] ==
 * grid[d,i+1]-grid[d,i] bit for bit, pyx:590-592), samples are row-major [n][dim].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VO_TINY    1e-257  /* 10**(min_10_exp+50)  pyx:34 */
#define VO_EPSILON (2.220446049250313e-16 * 1e4) /* pyx:36 */

/* ---------------------------------------------------------------- AdaptiveMap.map  pyx:310-360 */
void vo_map(const double *grid, const int64_t *ninc, int64_t gstride, int dim,
            const double *y, double *x, double *jac, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        double j = 1.0;
        for (int d = 0; d < dim; ++d) {
            const double *g = grid + d * gstride;
            int64_t ni = ninc[d];
            double t = y[i * dim + d] * (double)ni;
            int64_t iy = (int64_t)(int)floor(t);
            double dy = t - (double)iy;
            if (iy < ni) {
                double inc = g[iy + 1] - g[iy];
                x[i * dim + d] = g[iy] + inc * dy;
                j *= inc * (double)ni;
            } else {
                x[i * dim + d] = g[ni];
                j *= (g[ni] - g[ni - 1]) * (double)ni;
            }
        }
        jac[i] = j;
    }
}

/* ---------------------------------------------------------------- AdaptiveMap.jac1d  pyx:265-295 */
void vo_jac1d(const double *grid, const int64_t *ninc, int64_t gstride, int dim,
              const double *y, double *jac1d, int64_t n)
{
    for (int64_t i = 0; i < n; ++i)
        for (int d = 0; d < dim; ++d) {
            const double *g = grid + d * gstride;
            int64_t ni = ninc[d];
            int64_t iy = (int64_t)(int)floor(y[i * dim + d] * (double)ni);
            if (iy >= ni) iy = ni - 1;
            jac1d[i * dim + d] = (g[iy + 1] - g[iy]) * (double)ni;
        }
}

/* ---------------------------------------------------------------- AdaptiveMap.invmap  pyx:362-416
 * numpy.searchsorted(grid[d,:], x, side='right') over the WHOLE padded row (pyx:404), which
 * for an axis with ninc[d] < max ninc includes padding; callers pass rows without garbage
 * padding (we search only the meaningful ninc[d]+1 nodes, which is what the reference
 * means and does whenever all axes share ninc). */
void vo_invmap(const double *grid, const int64_t *ninc, int64_t gstride, int dim,
               const double *x, double *y, double *jac, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) jac[i] = 1.0;
    for (int d = 0; d < dim; ++d) {
        const double *g = grid + d * gstride;
        int64_t ni = ninc[d];
        for (int64_t i = 0; i < n; ++i) {
            double xv = x[i * dim + d];
            /* first index with g[idx] > xv, in [0, ni+1] */
            int64_t lo = 0, hi = ni + 1;
            while (lo < hi) {
                int64_t mid = (lo + hi) >> 1;
                if (g[mid] <= xv) lo = mid + 1; else hi = mid;
            }
            int64_t iy = lo;
            if (iy > 0 && iy <= ni) {
                int64_t k = iy - 1;
                double inc = g[k + 1] - g[k];
                y[i * dim + d] = ((double)k + (xv - g[k]) / inc) / (double)ni;
                jac[i] *= inc * (double)ni;
            } else if (iy <= 0) {
                y[i * dim + d] = 0.0;
                jac[i] *= (g[1] - g[0]) * (double)ni;
            } else {
                y[i * dim + d] = 1.0;
                jac[i] *= (g[ni] - g[ni - 1]) * (double)ni;
            }
        }
    }
}

/* ---------------------------------------------------------------- add_training_data  pyx:421-464
 * sum_f, n_f are [dim][hstride]; the caller initialises n_f to TINY (pyx:452). */
void vo_add_training_data(const int64_t *ninc, int dim, int64_t hstride,
                          const double *y, const double *f, int64_t n,
                          double *sum_f, double *n_f)
{
    for (int d = 0; d < dim; ++d) {
        int64_t ni = ninc[d];
        for (int64_t i = 0; i < n; ++i) {
            double yv = y[i * dim + d];
            if (yv > 0 && yv < 1) {
                int64_t iy = (int64_t)(int)floor(yv * (double)ni);
                sum_f[d * hstride + iy] += fabs(f[i]);
                n_f[d * hstride + iy] += 1;
            }
        }
    }
}

/* ---------------------------------------------------------------- AdaptiveMap.adapt  pyx:467-594
 * One call regrids every axis.  have_data: sum_f/n_f valid.  avg_f is ONE scratch row shared
 * by all axes (pyx:534) -- entries beyond an axis' own ninc, and whole rows when there is
 * no training data, deliberately carry over from the previous axis as in the reference.
 * new_grid rows have stride ngstride; nodes a regrid loop never reaches (the reference
 * leaves numpy.empty garbage there, pyx:576-587) are set to NaN.
 * The max(new_ninc)==1 early-out (pyx:519-530) is handled by the Python wrapper. */
void vo_adapt(const double *grid, const int64_t *ninc, int64_t gstride, int dim,
              int have_data, const double *sum_f, const double *n_f, int64_t hstride,
              double alpha, const int64_t *new_ninc, double *new_grid, int64_t ngstride)
{
    int64_t maxold = 0;
    for (int d = 0; d < dim; ++d) if (ninc[d] > maxold) maxold = ninc[d];
    double *avg_f = (double *)malloc(sizeof(double) * (size_t)(maxold > 0 ? maxold : 1));
    double *tmp_f = (double *)malloc(sizeof(double) * (size_t)(maxold > 0 ? maxold : 1));
    for (int64_t i = 0; i < maxold; ++i) avg_f[i] = 1.0;
    for (int d = 0; d < dim; ++d) {
        const double *g = grid + d * gstride;
        double *ng = new_grid + d * ngstride;
        int64_t old = ninc[d], nn = new_ninc[d];
        for (int64_t i = 0; i <= nn; ++i) ng[i] = NAN;
        if (alpha != 0 && old > 1) {
            if (have_data)
                for (int64_t i = 0; i < old; ++i) {
                    double nf = n_f[d * hstride + i];
                    avg_f[i] = nf > 0 ? sum_f[d * hstride + i] / nf : 0.0;
                }
            if (alpha > 0) {
                tmp_f[0] = fabs(7. * avg_f[0] + avg_f[1]) / 8.;
                tmp_f[old - 1] = fabs(7. * avg_f[old - 1] + avg_f[old - 2]) / 8.;
                double s = tmp_f[0] + tmp_f[old - 1];
                for (int64_t i = 1; i < old - 1; ++i) {
                    tmp_f[i] = fabs(6. * avg_f[i] + avg_f[i - 1] + avg_f[i + 1]) / 8.;
                    s += tmp_f[i];
                }
                if (s > 0) for (int64_t i = 0; i < old; ++i) avg_f[i] = tmp_f[i] / s + VO_TINY;
                else       for (int64_t i = 0; i < old; ++i) avg_f[i] = VO_TINY;
                for (int64_t i = 0; i < old; ++i)
                    if (avg_f[i] > 0 && avg_f[i] <= 0.99999999)
                        avg_f[i] = pow(-(1 - avg_f[i]) / log(avg_f[i]), alpha);
            }
        }
        ng[0] = g[0];
        ng[nn] = g[old];
        double f_ninc = 0.0;
        for (int64_t i = 0; i < old; ++i) f_ninc += avg_f[i];
        f_ninc /= (double)nn;
        int64_t j = -1;
        double acc = 0.0;
        for (int64_t i = 1; i < nn; ++i) {
            int ran_out = 0;
            while (acc < f_ninc) {
                ++j;
                if (j < old) acc += avg_f[j];
                else { ran_out = 1; break; }
            }
            if (ran_out) break;
            acc -= f_ninc;
            ng[i] = g[j + 1] - (acc / avg_f[j]) * (g[j + 1] - g[j]);
        }
    }
    free(avg_f);
    free(tmp_f);
}

/* ---------------------------------------------------------------- allocation  pyx:1692-1706
 * neval_hcube[h] = min(max_nh, <int>(sigf[h]*neval_sigf) + min_nh).  The reference's <int>
 * is undefined for products >= 2^31; like the product code we saturate there (documented
 * divergence).  Returns the total; range[0..1] updated like pyx:1699-1702
 * (caller presets both to min_nh, pyx:1682). */
int64_t vo_alloc_neval(const double *sigf, int64_t nhcube, double neval_sigf,
                       int64_t min_nh, int64_t max_nh, int64_t *neval_hcube, int64_t *range)
{
    int64_t total = 0;
    for (int64_t h = 0; h < nhcube; ++h) {
        double p = sigf[h] * neval_sigf;
        int64_t n = (p >= 2147483647.0 ? 2147483647 : (int64_t)(int)p) + min_nh;
        if (n > max_nh) n = max_nh;
        if (n < range[0]) range[0] = n;
        else if (n > range[1]) range[1] = n;
        neval_hcube[h] = n;
        total += n;
    }
    return total;
}

/* ---------------------------------------------------------------- stratify  pyx:1733-1742
 * y[i,d] = (y0[d] + yran[i,d]) / nstrat[d], y0 = mixed-radix digits of the hypercube index,
 * axis 0 least significant.  hcube0 = first hypercube of the batch. */
void vo_stratify(const int64_t *nstrat, int dim, int64_t hcube0, int64_t nhcube_batch,
                 const int64_t *neval_hcube, const double *yran, double *y, int64_t *hcube_of)
{
    int64_t i = 0;
    for (int64_t c = 0; c < nhcube_batch; ++c) {
        int64_t t = hcube0 + c, y0[64];
        for (int d = 0; d < dim; ++d) { y0[d] = t % nstrat[d]; t = (t - y0[d]) / nstrat[d]; }
        for (int64_t k = 0; k < neval_hcube[c]; ++k, ++i) {
            for (int d = 0; d < dim; ++d)
                y[i * dim + d] = ((double)y0[d] + yran[i * dim + d]) / (double)nstrat[d];
            if (hcube_of) hcube_of[i] = hcube0 + c;
        }
    }
}

/* ---------------------------------------------------------------- weights  pyx:1746-1752 */
void vo_weights(double *jac, const int64_t *neval_hcube, int64_t nhcube_batch, double dv_y)
{
    int64_t i = 0;
    for (int64_t c = 0; c < nhcube_batch; ++c) {
        double w = dv_y / (double)neval_hcube[c];
        for (int64_t k = 0; k < neval_hcube[c]; ++k, ++i) jac[i] *= w;
    }
}

/* ---------------------------------------------------------------- per-hypercube reduce  pyx:2136-2186
 * Consumes one batch: wgt[n], fx[n][nf], neval_hcube[nhcube_batch] (samples are contiguous
 * per hypercube, in hypercube order).  Accumulates mean[nf] and var (correlate: [nf][nf]
 * lower triangle, else [nf]) and fills fdv2[n]; writes sigf[c] = |var00|**(beta/2) and adds
 * it to *sum_sigf when update_sigf; sigf2_out[c] (nullable) gets |var00| per hypercube
 * (needed for adapt_to_errors, pyx:2187-2193). */
void vo_reduce_batch(const double *wgt, const double *fx, int64_t nf,
                     const int64_t *neval_hcube, int64_t nhcube_batch,
                     int correlate, int update_sigf, double beta,
                     double *mean, double *var, double *sigf, double *sum_sigf,
                     double *fdv2, double *sigf2_out)
{
    double *sum_wf = (double *)malloc(sizeof(double) * (size_t)nf);
    double *sum_dwf = (double *)malloc(sizeof(double) * (size_t)nf);
    double *dwf = (double *)malloc(sizeof(double) * (size_t)nf);
    double *sum_dwf2 = (double *)malloc(sizeof(double) * (size_t)(nf * nf));
    int64_t j = 0;
    for (int64_t c = 0; c < nhcube_batch; ++c) {
        int64_t n = neval_hcube[c];
        double dn = (double)n;
        memset(sum_wf, 0, sizeof(double) * (size_t)nf);
        memset(sum_dwf, 0, sizeof(double) * (size_t)nf);
        memset(sum_dwf2, 0, sizeof(double) * (size_t)(nf * nf));
        for (int64_t k = 0; k < n; ++k)
            for (int64_t s = 0; s < nf; ++s) sum_wf[s] += wgt[j + k] * fx[(j + k) * nf + s];
        for (int64_t k = 0; k < n; ++k, ++j) {
            for (int64_t s = 0; s < nf; ++s) {
                double m = sum_wf[s] / dn;
                dwf[s] = wgt[j] * fx[j * nf + s] - m;
                if (fabs(dwf[s]) < VO_EPSILON * fabs(m)) {
                    double e = VO_EPSILON * fabs(m);
                    sum_dwf2[s * nf + s] += e * e;
                    dwf[s] = 0.0;
                } else {
                    sum_dwf2[s * nf + s] += dwf[s] * dwf[s];
                }
                sum_dwf[s] += dwf[s];
                if (correlate)
                    for (int64_t t = 0; t < s; ++t) sum_dwf2[s * nf + t] += dwf[s] * dwf[t];
            }
            double a = wgt[j] * fx[j * nf] * dn;
            fdv2[j] = a * a;
        }
        for (int64_t s = 0; s < nf; ++s) {
            mean[s] += sum_wf[s] + sum_dwf[s];
            if (correlate) {
                for (int64_t t = 0; t <= s; ++t)
                    var[s * nf + t] += (dn * sum_dwf2[s * nf + t] - sum_dwf[s] * sum_dwf[t]) / (dn - 1.);
            } else {
                var[s] += (dn * sum_dwf2[s * nf + s] - sum_dwf[s] * sum_dwf[s]) / (dn - 1.);
            }
        }
        double sigf2 = fabs((dn * sum_dwf2[0] - sum_dwf[0] * sum_dwf[0]) / (dn - 1.));
        if (sigf2_out) sigf2_out[c] = sigf2;
        if (update_sigf) {
            sigf[c] = pow(sigf2, beta / 2.);
            *sum_sigf += sigf[c];
        }
    }
    free(sum_wf); free(sum_dwf); free(dwf); free(sum_dwf2);
}

/* ---------------------------------------------------------------- Philox4x32-10
 * Published algorithm (Salmon, Moraes, Dror, Shaw, SC'11; Random123 philox.h).  This is the
 * counter-based stream the CUDA engine uses in place of the reference's PCG64
 * (pyx:1676-1680): the reference lets any uniform source be injected through its
 * ran_array_generator hook (pyx:1081-1086), which is how parity runs feed both sides the
 * same numbers.
 *   key     = (seed_lo, seed_hi)
 *   counter = (k, (itn << 8) | pair, hcube_lo, hcube_hi)      k = sample index inside the cube
 *   u[2*pair]   = ((r1:r0) >> 12) * 2^-52,  u[2*pair+1] = ((r3:r2) >> 12) * 2^-52
 * (52-bit uniforms: the mantissa is or-ed under the exponent of 1.0 on the device, then 1.0 is
 * subtracted -- exact, and the same value as this integer form.)
 */
static inline void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void vo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* uniforms for hypercubes [hcube0, hcube0+nhcube_batch): yran[i][d], samples in cube order */
void vo_philox_uniforms(uint64_t seed, uint32_t itn, int dim, int64_t hcube0,
                        int64_t nhcube_batch, const int64_t *neval_hcube, double *yran)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    int64_t i = 0;
    for (int64_t c = 0; c < nhcube_batch; ++c) {
        uint64_t h = (uint64_t)(hcube0 + c);
        for (int64_t k = 0; k < neval_hcube[c]; ++k, ++i)
            for (int p = 0; 2 * p < dim; ++p) {
                uint32_t ctr[4] = {(uint32_t)k, ((itn & 0xFFFFFFu) << 8) | (uint32_t)p,
                                   (uint32_t)h, (uint32_t)(h >> 32)};
                uint32_t r[4];
                vo_philox4x32_10(ctr, key, r);
                uint64_t a = ((uint64_t)r[1] << 32) | r[0];
                uint64_t b = ((uint64_t)r[3] << 32) | r[2];
                yran[i * dim + 2 * p] = (double)(a >> 12) * 0x1.0p-52;
                if (2 * p + 1 < dim) yran[i * dim + 2 * p + 1] = (double)(b >> 12) * 0x1.0p-52;
            }
    }
}
