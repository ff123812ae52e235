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

