/*
 * vegas_b200.h -- C ABI of libvegas_b200.so, the B200 (sm_100a) vegas / vegas+ sampling engine.
 *
 * The reference (gplepage/vegas 6.4.1) has no C ABI: its boundary is the Python API of the
 * Cython module src/vegas/_vegas.pyx ("pyx:N" below).  This library implements the body of one
 * iteration of Integrator.__call__ (pyx:2086-2217) -- allocate -> sample -> map -> evaluate ->
 * per-hypercube reduce -> train -- plus the AdaptiveMap array methods, behind plain-C entry
 * points that the Python host layer (vegas_b200/_lib.py, ctypes) binds.  INTEGRATION.md shows
 * the stub a maintainer of the reference would add to call it from _vegas.pyx.
 *
 * Conventions: every function returns 0 on success and a negative code on error (message via
 * vb200_last_error()).  Pointers named *_dev are DEVICE pointers owned by the caller (torch
 * tensors in the Python layer); *_host are host pointers.  Calls are asynchronous on `stream`
 * (a cudaStream_t passed as void*) unless stated.  One host thread per context.
 */
#ifndef VEGAS_B200_H
#define VEGAS_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB200_ABI_VERSION 3
#define VB200_MAXDIM 32          /* largest number of integration dimensions */
#define VB200_CHUNK 256          /* hypercubes per work chunk; slab sizes are multiples of this */

typedef struct vb200_ctx vb200_ctx;

/* flags of vb200_iterate_fused / vb200_reduce */
#define VB200_UPDATE_SIGF  1     /* adaptive_strat at pyx:2052-2055: write sigf, sum sum_sigf   */
#define VB200_TRAIN        2     /* pyx:2196-2197: map.add_training_data(y, fdv2)               */
#define VB200_TRAIN_ERRORS 4     /* pyx:2187-2193: adapt_to_errors training                     */
#define VB200_CORRELATE    8     /* pyx:2170-2172: covariances between integrand components     */

/* built-in device integrands ("device functors compiled into the library") */
#define VB200_F_POLY          0  /* c0 + sum_d c[d]*x[d]**p[d]                                   */
#define VB200_F_GAUSS_MIX     1  /* norm * sum_p exp(-a |x-c_p|^2)   examples/simple.py, doc/eg6 */
#define VB200_F_RIDGE         2  /* norm * mean_k exp(-a sum_d (x_d-x0_k)^2)  examples/ridge.py  */
#define VB200_F_GENZ_OSC      3
#define VB200_F_GENZ_PRODPEAK 4
#define VB200_F_GENZ_CORNER   5
#define VB200_F_GENZ_GAUSS    6
#define VB200_F_GENZ_C0       7
#define VB200_F_GENZ_DISC     8
#define VB200_F_PATHINT       9  /* examples/path_integrand.pyx:88-142                           */

/* parameter blocks for vb200_set_integrand (host memory, copied) */
typedef struct { double c0; double c[VB200_MAXDIM]; int32_t p[VB200_MAXDIM]; } vb200_poly_t;
typedef struct { int32_t npeak; int32_t pad; double a, norm; const double* centers_host; /* [npeak][dim] */ } vb200_gaussmix_t;
typedef struct { int32_t n; int32_t mode; /* 0: axis-order sum as in ridge.py; 1: shifted-mean identity */ double a, norm; const double* x0_host; /* [n] */ } vb200_ridge_t;
typedef struct { double a[VB200_MAXDIM]; double u[VB200_MAXDIM]; } vb200_genz_t;
typedef struct { double T, m, xscale, c2, c4; int32_t nx0; int32_t pad; double x0list[7]; } vb200_pathint_t;

int         vb200_abi_version(void);
const char* vb200_last_error(void);

/* context: owns the device copy of the map, scratch buffers and the integrand parameters */
int  vb200_create(vb200_ctx** out, int device);
void vb200_destroy(vb200_ctx* ctx);
int  vb200_set_seed(vb200_ctx* ctx, uint64_t seed);                  /* Philox key (replaces gvar.RNG, pyx:1676-1680) */

/* AdaptiveMap state: grid_host[d*gstride + i], i = 0..ninc[d]        (pyx:112-137; inc is derived) */
int  vb200_set_map(vb200_ctx* ctx, const double* grid_host, const int64_t* ninc, int dim, int64_t gstride);

/* stratification (pyx:1346-1407 computes nstrat on the host) + this rank's block-cyclic share of
 * the hypercube index range: slabs of `slab` cubes dealt round-robin to `world` ranks. */
int  vb200_set_strata(vb200_ctx* ctx, const int64_t* nstrat, int dim, int64_t slab, int rank, int world,
                      int64_t* nlocal_out);

int  vb200_set_integrand(vb200_ctx* ctx, int id, const void* params, size_t nbytes, int* nf_out);

/* vegas+ allocation (pyx:1657-1662, 1692-1706): neval_hcube[h] = min(max_nh, (int)(sigf[h]*neval_sigf)
 * + min_nh) for this rank's cubes (sigf_dev == NULL: uniform_neval everywhere).  Writes the
 * per-cube counts to neval_hcube_dev if non-NULL, records the samples per 256-cube chunk and into
 * how many work items each chunk is cut (chunks the allocation piled more than 4096 samples onto are
 * shared by several CTAs), and returns stats_host = {sum, min, max, nchunks}.  The row offsets used by
 * vb200_sample / vb200_reduce / vb200_dy_profile are derived on their first use.  Synchronous. */
int  vb200_plan(vb200_ctx* ctx, const double* sigf_dev, double neval_sigf, int64_t min_neval_hcube,
                int64_t max_neval_hcube, int64_t uniform_neval, int32_t* neval_hcube_dev,
                int64_t stats_host[4], void* stream);
int  vb200_chunk_offsets(vb200_ctx* ctx, int64_t* out_host, int64_t count);   /* count <= nchunks+1 */
/* The same pre-pass for the NEXT iteration without a host round trip, launched right after an
 * iteration's kernels: neval_sigf = neval_scaled / *sum_sigf_dev is formed on the device from the
 * sum_sigf the iteration has just produced (acc_dev[nf + nf(nf+1)/2], after the all-reduce when the
 * hypercube range is sharded), and the six statistics {sum, min (stored as INT64_MAX - min, so that all
 * six start from zero), max, largest chunk, work items, work items of the 512-cube geometry} go to
 * stats_dev -- opaque words the caller copies to the host together with the iteration's results and hands
 * to vb200_plan_commit.  Asynchronous; the context has no valid plan until vb200_plan_commit
 * installs those statistics (host values) with the neval_sigf the host computed from the same sum_sigf
 * (pyx:1657-1662: identical double arithmetic), or vb200_plan replaces the pre-pass. */
int  vb200_plan_ahead(vb200_ctx* ctx, const double* sigf_dev, const double* sum_sigf_dev, double neval_scaled,
                      int64_t min_neval_hcube, int64_t max_neval_hcube, int64_t uniform_neval,
                      int64_t* stats_dev, void* stream);
int  vb200_plan_commit(vb200_ctx* ctx, double neval_sigf, const int64_t stats_host[6], int64_t stats_out[4]);

/* One fused iteration over this rank's cubes with the built-in integrand (pyx:2096-2197 in one
 * kernel).  acc_dev (+=): mean[nf], var lower triangle [nf(nf+1)/2] (row-major s>=t), sum_sigf.
 * sum_f_dev / n_f_dev (+=): training histogram [dim][hstride].  status_dev[0] != 0: NaN seen. */
int  vb200_iterate_fused(vb200_ctx* ctx, uint32_t itn, double beta, int flags, double* sigf_dev,
                         double* acc_dev, double* sum_f_dev, uint64_t* n_f_dev, int64_t hstride,
                         int32_t* status_dev, void* stream);

/* Unfused path, stage 1 (pyx:1732-1759, Integrator.random_batch): samples of local chunks
 * [chunk_begin, chunk_end) in hypercube order.  x_dev[rows][dim] (or [dim][rows] if x_transposed),
 * wgt_dev[rows]; y_dev, jac1d_dev ([rows][dim]) and hcube_dev are optional (NULL).
 * bins_dev (optional): [rows][dim] uint16, the increment
 * floor(y*ninc) each sample trains (pyx:460-462; 0xffff: y on the boundary, skipped) -- hand it to
 * vb200_reduce to spare it the Philox replay that otherwise re-derives y. */
int  vb200_sample(vb200_ctx* ctx, uint32_t itn, int64_t chunk_begin, int64_t chunk_end, double* x_dev,
                  double* wgt_dev, double* y_dev, double* jac1d_dev, int64_t* hcube_dev,
                  uint16_t* bins_dev, int x_transposed, void* stream);
/* The same with the uniforms supplied by the caller, u_dev[rows][dim] in hypercube order, instead of
 * the engine's Philox stream: the reference's documented injection hook Integrator.ran_array_generator
 * (pyx:1081-1086, 1676-1680, 1732).  Pass bins_dev on to vb200_reduce: with injected uniforms the
 * reduce kernel cannot re-derive y from the Philox counter. */
int  vb200_sample_from_uniforms(vb200_ctx* ctx, uint32_t itn, int64_t chunk_begin, int64_t chunk_end,
                                const double* u_dev, double* x_dev, double* wgt_dev, double* y_dev,
                                double* jac1d_dev, int64_t* hcube_dev, uint16_t* bins_dev,
                                int x_transposed, void* stream);
/* Unfused path, stage 2 (pyx:2136-2197): reduce f_dev[rows][nf] of the same chunk range.
 * bins_dev: the array vb200_sample wrote for the same range and iteration, or NULL. */
int  vb200_reduce(vb200_ctx* ctx, uint32_t itn, double beta, int flags, int64_t chunk_begin,
                  int64_t chunk_end, const double* f_dev, int nf, const double* wgt_dev,
                  double* sigf_dev, double* acc_dev, double* sum_f_dev, uint64_t* n_f_dev,
                  int64_t hstride, const uint16_t* bins_dev, int32_t* status_dev, void* stream);

/* The built-in functor on buffers: f_dev[rows][nf] = F(x_dev[rows][dim]) (same device code the fused
 * kernel inlines; stands where the reference calls fcn.eval(x), pyx:2103-2131). */
int  vb200_eval_integrand(vb200_ctx* ctx, const double* x_dev, int64_t rows, double* f_dev, void* stream);

/* Stratification profile for vegas.restratify (src/vegas/__init__.py:1314-1419): per axis mu and
 * y-bin i (yst_host = numpy.linspace(0, 1, ndy+1), ndy <= 32), mean and variance of the component
 * dI[mu][i] = f * [yst[i] <= y_mu <= yst[i+1]] of the reference's auxiliary integrand
 * (__init__.py:1390-1419), accumulated per hypercube with the two-pass of pyx:2142-2186
 * (correlate_integrals=False) from the callback path's buffers of the same chunk range:
 * f_dev[row*fstride] (component 0 of the integrand) and wgt_dev[row].  y is re-derived from the
 * Philox counter.  acc_dev[(mu*ndy + i)*2 + {0: mean, 1: var}] += ... */
int  vb200_dy_profile(vb200_ctx* ctx, uint32_t itn, int64_t chunk_begin, int64_t chunk_end,
                      const double* f_dev, int fstride, const double* wgt_dev, int ndy,
                      const double* yst_host, double* acc_dev, void* stream);

/* PDFIntegrator's change of variables on device buffers (src/vegas/__init__.py:599-627, _f_lbatch):
 * theta_dev[rows][dim] -> p_dev[rows][dim] = mean + chiv . vec_sig with chiv = scale * tan(theta), and
 * w_dev[rows] = prod_i scale (tan^2 theta_i + 1) * dp_dchiv * pdf, pdf = the parameters' Gaussian
 * prod_i exp(-chiv_i^2/2)/sqrt(2 pi) / dp_dchiv when `gaussian`, else 1 (the caller multiplies by its own
 * pdf(p)).  mean_dev[dim], vec_sig_dev[dim][dim] (row i = principal axis i scaled by its sigma: gvar.PDF). */
int  vb200_pdf_map(vb200_ctx* ctx, const double* theta_dev, int64_t rows, int dim, double scale, double dp_dchiv,
                   int gaussian, const double* mean_dev, const double* vec_sig_dev, double* p_dev, double* w_dev,
                   void* stream);
/* rows of PDFIntegrator's integrand (__init__.py:617-640): out_dev[rows][nfp+1] = [w | fp * w] when
 * pdf_first (adapt_to_pdf=True), else [fp * w | w]; fp_dev[rows][nfp] (NULL when nfp == 0) */
int  vb200_pdf_weight(vb200_ctx* ctx, const double* fp_dev, int nfp, const double* w_dev, int64_t rows,
                      int pdf_first, double* out_dev, void* stream);

/* AdaptiveMap array methods on device buffers (pyx:310-360, 362-416, 265-295, 421-464) */
int  vb200_map(vb200_ctx* ctx, const double* y_dev, double* x_dev, double* jac_dev, int64_t n, void* stream);
int  vb200_invmap(vb200_ctx* ctx, const double* x_dev, double* y_dev, double* jac_dev, int64_t n, void* stream);
int  vb200_jac1d(vb200_ctx* ctx, const double* y_dev, double* jac1d_dev, int64_t n, void* stream);
int  vb200_add_training_data(vb200_ctx* ctx, const double* y_dev, const double* f_dev, int64_t n,
                             double* sum_f_dev, uint64_t* n_f_dev, int64_t hstride, void* stream);

/* AdaptiveMap.adapt (pyx:467-594): host-side smoothing / damping / regrid of every axis, called
 * once per iteration.  All pointers are HOST pointers; sum_f_host/n_f_host may be NULL (no
 * training data: regrid only).  new_grid_host[d*ngstride + i], i = 0..new_ninc[d]. */
int  vb200_map_adapt(const double* grid_host, const int64_t* ninc, int dim, int64_t gstride,
                     const double* sum_f_host, const double* n_f_host, int64_t hstride, double alpha,
                     const int64_t* new_ninc, double* new_grid_host, int64_t ngstride);

/* One fused iteration in one call (everyday sizes: the binding calls cost more than the kernels): zero the
 * iteration buffer buf_dev (nwords 8-byte words: fp64 [mean, cov, sum_sigf | sum_f | ...] then, from word nf64 on,
 * int64 [n_f | NaN flag | 6 pre-pass statistics]), vb200_iterate_fused, vb200_map_adapt_device (alpha_adapt > 0),
 * vb200_plan_ahead (plan_neval_scaled > 0), then head_host[0 .. nacc) = the fp64 head and head_host[nacc .. nacc+7)
 * = NaN flag + statistics (int64 bit patterns), synchronised.  Replaces pyx:2096-2230 for one iteration. */
int  vb200_iteration(vb200_ctx* ctx, uint32_t itn, double beta, int flags, double* sigf_dev, void* buf_dev,
                     int64_t nacc, int64_t nh, int64_t hstride, int64_t nf64, int64_t nwords, double alpha_adapt,
                     double plan_neval_scaled, int64_t plan_min, int64_t plan_max, int64_t plan_uniform,
                     double* head_host, void* stream);
/* The same in two halves, so that the caller's bookkeeping of iteration i-1 overlaps the kernels of iteration i:
 * _begin launches everything (the head lands in pinned host memory owned by the context, written by the last
 * kernel of the chain: no device-to-host copies) and returns; _end synchronises and copies the nacc + 7 words out.
 * vb200_iteration == _begin + _end. */
int  vb200_iteration_begin(vb200_ctx* ctx, uint32_t itn, double beta, int flags, double* sigf_dev, void* buf_dev,
                           int64_t nacc, int64_t nh, int64_t hstride, int64_t nf64, int64_t nwords, double alpha_adapt,
                           double plan_neval_scaled, int64_t plan_min, int64_t plan_max, int64_t plan_uniform,
                           void* stream);
int  vb200_iteration_end(vb200_ctx* ctx, double* head_host, int64_t nhead, void* stream);

/* AdaptiveMap.adapt on the device (pyx:467-594 for alpha > 0, training data on every axis, ninc unchanged): the
 * context's grid is adapted in place from the iteration's histogram (sum_f_dev[dim][hstride]; counts as u64 or,
 * after a sharded run's all-reduce, as fp64 -- exactly one of the two pointers non-NULL); skipped when
 * status_dev[0] != 0 (NaN seen).  vb200_get_map copies the context's grid back (host_grid[dim][gstride]). */
int  vb200_map_adapt_device(vb200_ctx* ctx, const double* sum_f_dev, const uint64_t* n_f_u64_dev,
                            const double* n_f_f64_dev, int64_t hstride, double alpha, const int32_t* status_dev,
                            void* stream);
int  vb200_get_map(vb200_ctx* ctx, double* grid_host, int64_t gstride, void* stream);

/* engine uniforms for testing: u_dev[rows][dim] of local chunks [chunk_begin, chunk_end) */
int  vb200_uniforms(vb200_ctx* ctx, uint32_t itn, int64_t chunk_begin, int64_t chunk_end, double* u_dev, void* stream);

/* measurement helpers: FP64 FMA throughput of the device (TFLOP/s, FMA = 2 flops) and kernel
 * launch count since context creation. */
int     vb200_fp64_peak(int device, int iters, double* tflops_out, double* ms_out);
int64_t vb200_launch_count(vb200_ctx* ctx);
/* geometry of the most recent engine launch: {CTAs, CTAs per SM, dynamic shared memory bytes per
 * CTA, bins of the shared-memory training-histogram windows, threads per CTA, hypercubes per chunk} */
int     vb200_last_launch(vb200_ctx* ctx, int64_t out[6]);

#ifdef __cplusplus
}
#endif
#endif
