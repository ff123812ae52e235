// pdfmap.cu -- PDFIntegrator's change of variables on HBM buffers (reference src/vegas/__init__.py:599-627,
// PDFIntegrator._f_lbatch): the integration variables theta in (-atan(limit/scale), atan(limit/scale))^dim
// become parameters p = mean + chiv . vec_sig with chiv = scale * tan(theta), the unit-normal variables along
// the principal axes of the parameters' correlation matrix (gvar.PDF), together with the weight
//     w = dp/dtheta * pdf,   dp/dtheta = prod_i scale (tan^2 theta_i + 1) * dp_dchiv,
//     pdf = prod_i exp(-chiv_i^2 / 2) / sqrt(2 pi) / dp_dchiv     (the parameters' own Gaussian; else 1:
//                                                                  the caller multiplies by its pdf(p))
// and k_pdf_weight assembles the integrand's rows [pdf | f(p) pdf] (or [f(p) pdf | pdf]) from f(p) and w.
#include "ctx.h"

#define VB_PDF_NT 128

__global__ void __launch_bounds__(VB_PDF_NT) k_pdf_map(const double* __restrict__ theta, int64_t rows, int dim, double scale,
                                                       double dp_dchiv, int gaussian, const double* __restrict__ mean,
                                                       const double* __restrict__ vec_sig, double* __restrict__ p_out,
                                                       double* __restrict__ w_out)
{
    extern __shared__ double pm_s[];                // vec_sig [dim][dim] | mean [dim] | chiv [NT][dim + 1]
    double* vs_s = pm_s;
    double* mean_s = pm_s + dim * dim;
    double* c_s = mean_s + dim + (size_t)threadIdx.x * (dim + 1);
    for (int i = threadIdx.x; i < dim * dim; i += blockDim.x) vs_s[i] = vec_sig[i];
    for (int i = threadIdx.x; i < dim; i += blockDim.x) mean_s[i] = mean[i];
    __syncthreads();
    const double rs2pi = 0.3989422804014327;        // 1 / sqrt(2 pi)
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        double dp = 1.0, g = 1.0;
        for (int i = 0; i < dim; ++i) {
            const double t = tan(theta[r * dim + i]);
            const double c = scale * t;
            c_s[i] = c;
            dp *= scale * (t * t + 1.0);
            if (gaussian) g *= exp(-(c * c) / 2.0) * rs2pi;
        }
        dp *= dp_dchiv;
        for (int j = 0; j < dim; ++j) {
            double a = 0.0;
            for (int i = 0; i < dim; ++i) a = fma(c_s[i], vs_s[i * dim + j], a);
            p_out[r * dim + j] = mean_s[j] + a;
        }
        w_out[r] = gaussian ? dp * (g / dp_dchiv) : dp;
    }
}

__global__ void __launch_bounds__(256) k_pdf_weight(const double* __restrict__ fp, int nfp, const double* __restrict__ w,
                                                    int64_t rows, int pdf_first, double* __restrict__ out)
{
    const int nf = nfp + 1;
    const int64_t total = rows * nf;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / nf;
        const int c = (int)(i - r * nf);
        const double wr = w[r];
        const int cf = pdf_first ? c - 1 : c;       // column of f(p); -1 or nfp: the pdf column
        out[i] = (cf < 0 || cf >= nfp) ? wr : fp[r * nfp + cf] * wr;
    }
}

extern "C" int vb200_pdf_map(vb200_ctx* c, const double* theta_dev, int64_t rows, int dim, double scale, double dp_dchiv,
                             int gaussian, const double* mean_dev, const double* vec_sig_dev, double* p_dev, double* w_dev,
                             void* stream)
{
    if (!c || !theta_dev || !mean_dev || !vec_sig_dev || !p_dev || !w_dev) return fail(-1, "vb200_pdf_map: null argument");
    if (dim < 1 || dim > VB_MAXD) return fail(-1, "vb200_pdf_map: dim %d outside 1..%d", dim, VB_MAXD);
    if (rows <= 0) return 0;
    CK(cudaSetDevice(c->device));
    const size_t smem = sizeof(double) * ((size_t)dim * dim + dim + (size_t)VB_PDF_NT * (dim + 1));
    int64_t g = (rows + VB_PDF_NT - 1) / VB_PDF_NT, cap = (int64_t)c->sm_count * 8;
    if (g > cap) g = cap;
    k_pdf_map<<<(int)g, VB_PDF_NT, smem, (cudaStream_t)stream>>>(theta_dev, rows, dim, scale, dp_dchiv, gaussian, mean_dev,
                                                                 vec_sig_dev, p_dev, w_dev);
    c->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int vb200_pdf_weight(vb200_ctx* c, const double* fp_dev, int nfp, const double* w_dev, int64_t rows, int pdf_first,
                                double* out_dev, void* stream)
{
    if (!c || !w_dev || !out_dev || (nfp > 0 && !fp_dev)) return fail(-1, "vb200_pdf_weight: null argument");
    if (nfp < 0) return fail(-1, "vb200_pdf_weight: nfp < 0");
    if (rows <= 0) return 0;
    CK(cudaSetDevice(c->device));
    int64_t g = (rows * (nfp + 1) + 255) / 256, cap = (int64_t)c->sm_count * 16;
    if (g > cap) g = cap;
    k_pdf_weight<<<(int)g, 256, 0, (cudaStream_t)stream>>>(fp_dev, nfp, w_dev, rows, pdf_first, out_dev);
    c->launches += 1;
    CK(cudaGetLastError());
    return 0;
}
