// mapops.cu -- AdaptiveMap array methods on device buffers (pyx:310-360, 362-416, 265-295, 421-464)
// and the once-per-iteration host step AdaptiveMap.adapt (pyx:467-594).
#include "ctx.h"

// ---------------------------------------------------------------------------------------------
// AdaptiveMap array methods
// ---------------------------------------------------------------------------------------------
__global__ void k_map(const MapP m, const double* y, double* x, double* jac, int64_t n, double* jac1d)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double j = 1.0;
        for (int d = 0; d < m.dim; ++d) {
            const int ni = m.ninc[d];
            const double* g = m.grid + (size_t)d * m.gstride;
            double t = __dmul_rn(y[i * m.dim + d], (double)ni);
            int iy = (int)floor(t);
            double j1;
            if (iy < ni) {
                // reference (pyx:351-356) would index out of bounds for y < 0; clamp like y == 0 side
                if (iy < 0) iy = 0;
                double g0 = g[iy], inc = g[iy + 1] - g0;
                if (x) x[i * m.dim + d] = __dadd_rn(g0, __dmul_rn(inc, __dsub_rn(t, (double)iy)));
                j1 = __dmul_rn(inc, (double)ni);
            } else {
                if (x) x[i * m.dim + d] = g[ni];
                j1 = __dmul_rn(g[ni] - g[ni - 1], (double)ni);
            }
            j = __dmul_rn(j, j1);
            if (jac1d) jac1d[i * m.dim + d] = j1;
        }
        if (jac) jac[i] = j;
    }
}

__global__ void k_invmap(const MapP m, const double* x, double* y, double* jac, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double j = 1.0;
        for (int d = 0; d < m.dim; ++d) {
            const int ni = m.ninc[d];
            const double* g = m.grid + (size_t)d * m.gstride;
            const double xv = x[i * m.dim + d];
            int lo = 0, hi = ni + 1;               // first index with g[idx] > xv (searchsorted right)
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (g[mid] <= xv) lo = mid + 1; else hi = mid;
            }
            if (lo > 0 && lo <= ni) {
                int k = lo - 1;
                double inc = g[k + 1] - g[k];
                y[i * m.dim + d] = __ddiv_rn(__dadd_rn((double)k, __ddiv_rn(__dsub_rn(xv, g[k]), inc)), (double)ni);
                j = __dmul_rn(j, __dmul_rn(inc, (double)ni));
            } else if (lo <= 0) {
                y[i * m.dim + d] = 0.0;
                j = __dmul_rn(j, __dmul_rn(g[1] - g[0], (double)ni));
            } else {
                y[i * m.dim + d] = 1.0;
                j = __dmul_rn(j, __dmul_rn(g[ni] - g[ni - 1], (double)ni));
            }
        }
        jac[i] = j;
    }
}

__global__ void k_add_training(const MapP m, const double* y, const double* f, int64_t n, double* sum_f,
                               unsigned long long* n_f, int hstride)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double fv = fabs(f[i]);
        for (int d = 0; d < m.dim; ++d) {
            double yv = y[i * m.dim + d];
            if (yv > 0.0 && yv < 1.0) {
                int iy = (int)floor(__dmul_rn(yv, (double)m.ninc[d]));
                atomicAdd(sum_f + (size_t)d * hstride + iy, fv);
                atomicAdd(n_f + (size_t)d * hstride + iy, 1ull);
            }
        }
    }
}

static int grid_for(vb200_ctx* c, int64_t n)
{
    int64_t g = (n + 255) / 256, cap = (int64_t)c->sm_count * 16;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

extern "C" int vb200_map(vb200_ctx* c, const double* y, double* x, double* jac, int64_t n, void* stream)
{
    if (!c || !y || !x || !jac) return fail(-1, "vb200_map: null argument");
    if (!c->have_map) return fail(-1, "vb200_map: no map set");
    if (n <= 0) return 0;
    CK(cudaSetDevice(c->device));
    k_map<<<grid_for(c, n), 256, 0, (cudaStream_t)stream>>>(c->map, y, x, jac, n, nullptr);
    c->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int vb200_jac1d(vb200_ctx* c, const double* y, double* jac1d, int64_t n, void* stream)
{
    if (!c || !y || !jac1d) return fail(-1, "vb200_jac1d: null argument");
    if (!c->have_map) return fail(-1, "vb200_jac1d: no map set");
    if (n <= 0) return 0;
    CK(cudaSetDevice(c->device));
    k_map<<<grid_for(c, n), 256, 0, (cudaStream_t)stream>>>(c->map, y, nullptr, nullptr, n, jac1d);
    c->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int vb200_invmap(vb200_ctx* c, const double* x, double* y, double* jac, int64_t n, void* stream)
{
    if (!c || !y || !x || !jac) return fail(-1, "vb200_invmap: null argument");
    if (!c->have_map) return fail(-1, "vb200_invmap: no map set");
    if (n <= 0) return 0;
    CK(cudaSetDevice(c->device));
    k_invmap<<<grid_for(c, n), 256, 0, (cudaStream_t)stream>>>(c->map, x, y, jac, n);
    c->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int vb200_add_training_data(vb200_ctx* c, const double* y, const double* f, int64_t n, double* sum_f,
                                       uint64_t* n_f, int64_t hstride, void* stream)
{
    if (!c || !y || !f || !sum_f || !n_f) return fail(-1, "vb200_add_training_data: null argument");
    if (!c->have_map) return fail(-1, "vb200_add_training_data: no map set");
    for (int d = 0; d < c->map.dim; ++d)
        if (c->map.ninc[d] > hstride) return fail(-1, "vb200_add_training_data: hstride too small");
    if (n <= 0) return 0;
    CK(cudaSetDevice(c->device));
    k_add_training<<<grid_for(c, n), 256, 0, (cudaStream_t)stream>>>(c->map, y, f, n, sum_f, (unsigned long long*)n_f, (int)hstride);
    c->launches += 1;
    CK(cudaGetLastError());
    return 0;
}


// ---------------------------------------------------------------------------------------------
// AdaptiveMap.adapt (pyx:467-594): the once-per-iteration host step, O(dim * ninc).
// Smooth the per-increment training averages, damp with alpha, then move the nodes so every new
// increment holds an equal share.  `work` (one row, carried from axis to axis exactly like the
// reference's avg_f array) starts at 1.  Nodes the walk never reaches are NaN.
// ---------------------------------------------------------------------------------------------
namespace {
const double kTiny = 1e-257;     // 10**(min_10_exp + 50), pyx:34

void smooth_and_damp(std::vector<double>& w, std::vector<double>& tmp, int64_t n, double alpha)
{
    tmp[0] = fabs(7. * w[0] + w[1]) / 8.;
    tmp[n - 1] = fabs(7. * w[n - 1] + w[n - 2]) / 8.;
    double total = tmp[0] + tmp[n - 1];
    for (int64_t i = 1; i < n - 1; ++i) {
        tmp[i] = fabs(6. * w[i] + w[i - 1] + w[i + 1]) / 8.;
        total += tmp[i];
    }
    for (int64_t i = 0; i < n; ++i) {
        double a = total > 0 ? tmp[i] / total + kTiny : kTiny;
        if (a > 0 && a <= 0.99999999) {
            const double x = -(1 - a) / log(a);
            // pyx:575 raises to the power alpha; the default alpha = 0.5 is a square root (correctly rounded, and
            // several times cheaper than pow -- this loop is most of an iteration's host time at small neval)
            a = alpha == 0.5 ? sqrt(x) : (alpha == 1.0 ? x : pow(x, alpha));
        }
        w[i] = a;
    }
}

void regrid_axis(const double* g, int64_t n_old, const std::vector<double>& w, int64_t n_new, double* out)
{
    for (int64_t i = 0; i <= n_new; ++i) out[i] = NAN;
    out[0] = g[0];
    out[n_new] = g[n_old];
    double share = 0.;
    for (int64_t i = 0; i < n_old; ++i) share += w[i];
    share /= (double)n_new;
    int64_t j = -1;
    double acc = 0.;
    for (int64_t i = 1; i < n_new; ++i) {
        while (acc < share) {
            if (++j >= n_old) return;              // ran out of old increments
            acc += w[j];
        }
        acc -= share;
        out[i] = g[j + 1] - (acc / w[j]) * (g[j + 1] - g[j]);
    }
}
}  // namespace

extern "C" int vb200_map_adapt(const double* grid_host, const int64_t* ninc, int dim, int64_t gstride,
                               const double* sum_f_host, const double* n_f_host, int64_t hstride, double alpha,
                               const int64_t* new_ninc, double* new_grid_host, int64_t ngstride)
{
    if (!grid_host || !ninc || !new_ninc || !new_grid_host) return fail(-1, "vb200_map_adapt: null argument");
    int64_t widest = 1;
    for (int d = 0; d < dim; ++d) {
        if (ninc[d] < 1 || new_ninc[d] < 1 || ninc[d] + 1 > gstride || new_ninc[d] + 1 > ngstride)
            return fail(-1, "vb200_map_adapt: bad ninc on axis %d", d);
        if (ninc[d] > widest) widest = ninc[d];
    }
    const bool have = sum_f_host && n_f_host;
    std::vector<double> w((size_t)widest, 1.0), tmp((size_t)widest);
    for (int d = 0; d < dim; ++d) {
        const int64_t n_old = ninc[d];
        if (alpha != 0 && n_old > 1) {
            if (have)
                for (int64_t i = 0; i < n_old; ++i) {
                    double cnt = n_f_host[d * hstride + i];
                    w[i] = cnt > 0 ? sum_f_host[d * hstride + i] / cnt : 0.;
                }
            if (alpha > 0) smooth_and_damp(w, tmp, n_old, alpha);
        }
        regrid_axis(grid_host + d * gstride, n_old, w, new_ninc[d], new_grid_host + d * ngstride);
    }
    return 0;
}



// ---------------------------------------------------------------------------------------------
// AdaptiveMap.adapt on the device (pyx:467-594 for the common case: alpha > 0, training data on every
// axis, the same number of increments before and after).  One CTA per axis; the grid never leaves HBM,
// so an iteration's host epilogue shrinks to reading [mean, cov, sum_sigf] -- at the reference's
// everyday sizes (neval = 1e4) the host adapt, the histogram D2H and the grid H2D were most of an
// iteration.  Element-wise steps (averages, smoothing, damping) are the host's operations.  The sums
// are block reductions and the regrid is a prefix sum + one binary search per new node instead of the
// host's sequential walk (a chain of n dependent fp64 adds is ~40 us on one GPU thread): new node i
// sits where the running sum of the weights passes i * share.  The nodes agree with the host's to
// ~1e-13 of an increment (summation order; log / pow / sqrt are the device's instead of glibc's).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double adapt_block_sum(double v, double* red)       // all threads get the total
{
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    return t;
}

__global__ void __launch_bounds__(256) k_map_adapt(MapP m, double* grid, const double* __restrict__ sum_f,
                                                   const unsigned long long* __restrict__ n_u64,
                                                   const double* __restrict__ n_f64, int hstride, double alpha,
                                                   const int* __restrict__ status)
{
    extern __shared__ double ad_s[];           // w [n] | tmp -> prefix sums [n] | old nodes [n + 1]
    __shared__ double red[8];
    const int d = blockIdx.x, n = m.ninc[d], tid = threadIdx.x, NT = blockDim.x;
    if (status != nullptr && status[0] != 0) return;          // the integrand returned NaN: the caller raises, the map stays
    double* w = ad_s;
    double* tmp = ad_s + n;
    double* g = ad_s + 2 * n;
    double* row = grid + (size_t)d * m.gstride;
    for (int i = tid; i <= n; i += NT) g[i] = row[i];
    for (int i = tid; i < n; i += NT) {
        const double cnt = n_u64 ? (double)n_u64[(size_t)d * hstride + i] : n_f64[(size_t)d * hstride + i];
        w[i] = cnt > 0 ? sum_f[(size_t)d * hstride + i] / cnt : 0.;
    }
    __syncthreads();
    double part = 0.;
    for (int i = tid; i < n; i += NT) {
        double t;
        if (i == 0) t = fabs(7. * w[0] + w[1]) / 8.;
        else if (i == n - 1) t = fabs(7. * w[n - 1] + w[n - 2]) / 8.;
        else t = fabs(6. * w[i] + w[i - 1] + w[i + 1]) / 8.;
        tmp[i] = t;
        part += t;
    }
    const double total = adapt_block_sum(part, red);           // (contains the barriers that order tmp / w)
    const double tiny = 1e-257;
    __syncthreads();
    part = 0.;
    for (int i = tid; i < n; i += NT) {
        double a = total > 0 ? tmp[i] / total + tiny : tiny;
        if (a > 0 && a <= 0.99999999) {
            const double x = -(1 - a) / log(a);
            a = alpha == 0.5 ? sqrt(x) : (alpha == 1.0 ? x : pow(x, alpha));
        }
        w[i] = a;
        part += a;
    }
    const double share = adapt_block_sum(part, red) / (double)n;
    // inclusive prefix sums of w into tmp: contiguous segments per thread, then the segments' offsets
    const int per = (n + NT - 1) / NT, lo = min(tid * per, n), hi = min(lo + per, n);
    double run = 0.;
    for (int i = lo; i < hi; ++i) { run += w[i]; tmp[i] = run; }
    double scan = run;                                         // exclusive scan of the segments' sums: warp scans + the warps' totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, scan, o);
        if ((tid & 31) >= o) scan += y;
    }
    __syncthreads();                                           // (red is still being read by adapt_block_sum's tail)
    if ((tid & 31) == 31) red[tid >> 5] = scan;
    __syncthreads();
    double off = scan - run;
    for (int w = 0; w < (tid >> 5); ++w) off += red[w];
    for (int i = lo; i < hi; ++i) tmp[i] += off;
    __syncthreads();
    for (int i = 1 + tid; i < n; i += NT) {
        const double target = (double)i * share;
        int a = -1, b = n - 1;                                  // first j with tmp[j] >= target (clamped to the last increment)
        while (b - a > 1) {
            const int mid = (a + b) >> 1;
            if (tmp[mid] >= target) b = mid; else a = mid;
        }
        const int j = b;
        const double acc = tmp[j] - target;
        row[i] = g[j + 1] - (acc / w[j]) * (g[j + 1] - g[j]);
    }
    // (row[0] and row[n] keep the old end points; the padding beyond n repeats the last node)
}

extern "C" int vb200_map_adapt_device(vb200_ctx* c, const double* sum_f_dev, const uint64_t* n_f_u64_dev,
                                      const double* n_f_f64_dev, int64_t hstride, double alpha, const int32_t* status_dev,
                                      void* stream)
{
    if (!c || !sum_f_dev || (!n_f_u64_dev && !n_f_f64_dev)) return fail(-1, "vb200_map_adapt_device: null argument");
    if (!c->have_map) return fail(-1, "vb200_map_adapt_device: no map set");
    if (!(alpha > 0)) return fail(-1, "vb200_map_adapt_device: alpha must be positive (use vb200_map_adapt)");
    int widest = 0;
    for (int d = 0; d < c->map.dim; ++d) {
        if (c->map.ninc[d] < 2) return fail(-1, "vb200_map_adapt_device: axis %d has a single increment (use vb200_map_adapt)", d);
        if (c->map.ninc[d] > hstride) return fail(-1, "vb200_map_adapt_device: hstride too small");
        if (c->map.ninc[d] > widest) widest = c->map.ninc[d];
    }
    const size_t smem = sizeof(double) * (3 * (size_t)widest + 1);
    if (smem > 200 * 1024) return fail(-1, "vb200_map_adapt_device: %d increments exceed the kernel's shared memory", widest);
    CK(cudaSetDevice(c->device));
    static thread_local int attr_set_for = -1;
    if (smem > 48 * 1024 && attr_set_for != c->device) {
        CK(cudaFuncSetAttribute(k_map_adapt, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set_for = c->device;
    }
    k_map_adapt<<<c->map.dim, 256, smem, (cudaStream_t)stream>>>(c->map, (double*)c->grid.p, sum_f_dev,
                                                                 (const unsigned long long*)n_f_u64_dev, n_f_f64_dev,
                                                                 (int)hstride, alpha, (const int*)status_dev);
    c->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int vb200_get_map(vb200_ctx* c, double* grid_host, int64_t gstride, void* stream)
{
    if (!c || !grid_host) return fail(-1, "vb200_get_map: null argument");
    if (!c->have_map) return fail(-1, "vb200_get_map: no map set");
    if (gstride != c->map.gstride) return fail(-1, "vb200_get_map: gstride %lld != %d", (long long)gstride, c->map.gstride);
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(grid_host, c->grid.p, sizeof(double) * (size_t)c->map.dim * (size_t)gstride, cudaMemcpyDeviceToHost,
                       (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
