// sampler.cu -- the samplers of the callback path (Integrator.random_batch, pyx:1601-1634 / 1732-1759):
// k_sample_x (x, wgt, training bins) and the generic k_sample (y, jac1d, hcube, raw uniforms).
#include "ctx.h"

// ---------------------------------------------------------------------------------------------
// unfused stage 1: write the samples (Integrator.random_batch, pyx:1732-1759)
// ---------------------------------------------------------------------------------------------
struct SampleOut {
    double* x; double* wgt; double* y; double* jac1d; int64_t* hcube;
    int x_transposed;
    int64_t rows;      // rows in this batch (for the transposed layout)
    double* u;         // raw uniforms (testing)
    uint16_t* bins;    // [rows][dim] training bin of every sample (0xffff: none) for vb200_reduce
    const double* u_in; // [rows][dim] uniforms supplied by the caller (Integrator.ran_array_generator, pyx:1676-1680, 1732) instead of Philox
};

// the two uniforms of axis pair pr of a sample: the caller's (row-major u_in) or Philox
__device__ __forceinline__ void uniforms_of(const EngineP& p, const SampleOut& o, int64_t h, uint32_t k, int64_t row, int pr, int dim,
                                            double& ua, double& ub)
{
    if (o.u_in) {
        ua = o.u_in[row * dim + 2 * pr];
        ub = 2 * pr + 1 < dim ? o.u_in[row * dim + 2 * pr + 1] : 0.0;
    } else philox_pair(p.key, p.itn, h, k, pr, ua, ub);
}

__global__ void __launch_bounds__(VB_NT) k_sample(const __grid_constant__ EngineP p, const __grid_constant__ SampleOut o)
{
    __shared__ long long ex_s[VB_CH + 1];
    __shared__ int n_s[VB_CH];
    __shared__ long long scan_s[VB_NT / 32];
    __shared__ uint32_t base_s[VB_MAXD];
    extern __shared__ uint32_t y0_s[];          // [VB_CH][dim]
    __shared__ long long item_s[3];
    const int tid = threadIdx.x;
    const int dim = p.map.dim;
    // work items as in k_engine (a chunk, or one of the nsub parts of a chunk the allocation piled
    // samples onto), dealt round-robin: items are bounded in size, so this balances
    for (int64_t it = p.item_begin + blockIdx.x; it < p.item_end; it += gridDim.x) {
        __syncthreads();
        if (tid == 0) {
            long long c; int sb, ns;
            locate_item(p, it, c, sb, ns);
            item_s[0] = c; item_s[1] = sb; item_s[2] = ns;
        }
        __syncthreads();
        const int64_t lc = item_s[0];
        const long long sub = item_s[1], nsub = item_s[2];
        const int64_t lh0 = lc * VB_CH;
        const int64_t h0 = local_to_global(p.st, lh0);
        const long long total = chunk_setup<VB_NT, VB_CH, uint32_t>(p, lh0, h0, ex_s, n_s, y0_s, base_s, scan_s);
        const int64_t chunk_row = p.chunk_off[lc] - p.row0;
        long long i0 = 0, i1 = total;                  // rows of this item
        if (nsub > 1) { i0 = total * sub / nsub; i1 = total * (sub + 1) / nsub; }   // by rows: no per-cube state here
        for (long long i = i0 + tid; i < i1; i += VB_NT) {
            int lo = 0, hi = VB_CH;
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (ex_s[mid] <= i) lo = mid; else hi = mid;
            }
            const int c = lo;
            const uint32_t k = (uint32_t)(i - ex_s[c]);
            const int n = n_s[c];
            const int64_t h = h0 + c, row = chunk_row + i;
            const uint32_t* y0 = y0_s + c * dim;
            double jac = 1.0;
            for (int pr = 0; 2 * pr < dim; ++pr) {
                double u[2];
                uniforms_of(p, o, h, k, row, pr, dim, u[0], u[1]);
                for (int e = 0; e < 2; ++e) {
                    const int d = 2 * pr + e;
                    if (d >= dim) break;
                    if (o.u) { o.u[row * dim + d] = u[e]; continue; }
                    double y = div_exact((double)y0[d] + u[e], p.st.dns[d], p.st.rns[d]);
                    if (o.bins) {
                        const int ib = min(__double2int_rd(__dmul_rn(y, (double)p.map.ninc[d])), p.map.ninc[d] - 1);
                        o.bins[row * dim + d] = (y > 0.0 && y < 1.0) ? (uint16_t)ib : (uint16_t)0xffff;   // pyx:460
                    }
                    const int ni = p.map.ninc[d];
                    const double* g = p.map.grid + (size_t)d * p.map.gstride;
                    double t = __dmul_rn(y, (double)ni);
                    int iy = __double2int_rd(t);
                    double xv, j1;
                    if (iy < ni) {
                        double g0 = __ldg(g + iy), g1 = __ldg(g + iy + 1);
                        double inc = g1 - g0;
                        xv = __dadd_rn(g0, __dmul_rn(inc, __dsub_rn(t, (double)iy)));   // no FMA: bit-identical to pyx:354
                        j1 = inc * (double)ni;
                    } else {
                        double g0 = __ldg(g + ni - 1), g1 = __ldg(g + ni);
                        xv = g1;
                        j1 = (g1 - g0) * (double)ni;
                    }
                    jac *= j1;
                    if (o.x_transposed) o.x[(int64_t)d * o.rows + row] = xv;
                    else o.x[row * dim + d] = xv;
                    if (o.y) o.y[row * dim + d] = y;
                    if (o.jac1d) o.jac1d[row * dim + d] = j1;
                }
            }
            if (o.u) continue;
            o.wgt[row] = jac * (p.dv_y / (double)n);
            if (o.hcube) o.hcube[row] = h;
        }
    }
}

// The integration path's sampler: x[rows][dim] (or [dim][rows]), wgt[rows] and, optionally, the
// samples' training bins.  Same items and arithmetic as k_sample; the differences are mechanical:
// the axis loop is unrolled for D <= 10 (grid loads of all axes in flight together), and a warp's
// 32 rows of x -- one contiguous block of the row-major array -- are staged in shared memory and
// written out with full-line stores instead of 32 strided 8-byte stores per axis.
template <int D, bool XT>
__global__ void __launch_bounds__(VB_NT) k_sample_x(const __grid_constant__ EngineP p, const __grid_constant__ SampleOut o)
{
    __shared__ long long ex_s[VB_CH + 1];
    __shared__ int n_s[VB_CH];
    __shared__ long long scan_s[VB_NT / 32];
    __shared__ uint32_t base_s[VB_MAXD];
    __shared__ long long item_s[3];
    extern __shared__ double sx_dyn[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dim = p.map.dim;
    const int S = dim | 1;                                       // odd row stride: conflict-free tile rows
    double* tile = sx_dyn + (size_t)warp * 32 * S;               // [32][S] (row-major x only)
    uint32_t* y0_s = (uint32_t*)(sx_dyn + (XT ? 0 : (size_t)(VB_NT / 32) * 32 * S));   // [VB_CH][dim]
    uint16_t* btile = (uint16_t*)(y0_s + VB_CH * dim) + (size_t)warp * 32 * dim;       // [32][dim]
    for (;;) {
        __syncthreads();
        if (tid == 0) {                                  // items are claimed: CTAs finish together
            const long long g = p.item_begin + (long long)atomicAdd(p.work_counter, 1ull);
            if (g >= p.item_end) item_s[0] = -1;
            else {
                long long c; int sb, ns;
                locate_item(p, g, c, sb, ns);
                item_s[0] = c; item_s[1] = sb; item_s[2] = ns;
            }
        }
        __syncthreads();
        const int64_t lc = item_s[0];
        if (lc < 0) break;
        const long long sub = item_s[1], nsub = item_s[2];
        const int64_t lh0 = lc * VB_CH;
        const int64_t h0 = local_to_global(p.st, lh0);
        const long long total = chunk_setup<VB_NT, VB_CH, uint32_t>(p, lh0, h0, ex_s, n_s, y0_s, base_s, scan_s);
        const int64_t chunk_row = p.chunk_off[lc] - p.row0;
        long long i0 = 0, i1 = total;
        if (nsub > 1) { i0 = total * sub / nsub; i1 = total * (sub + 1) / nsub; }   // by rows: no per-cube state here
        for (long long ib = i0; ib < i1; ib += VB_NT) {           // warp-uniform trip count
            const long long i = ib + tid;
            const bool live = i < i1;
            const int64_t row = chunk_row + i;
            if (live) {
                int lo = 0, hi = VB_CH;
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (ex_s[mid] <= i) lo = mid; else hi = mid;
                }
                const int c = lo;
                const uint32_t k = (uint32_t)(i - ex_s[c]);
                const int64_t h = h0 + c;
                const uint32_t* y0 = y0_s + c * dim;
                double jac = 1.0;
                constexpr int UNR = D > 10 ? 1 : (D + 1) / 2;
#pragma unroll UNR
                for (int pr = 0; pr < (D + 1) / 2; ++pr) {
                    if (2 * pr < dim) {
                        double u[2];
                        uniforms_of(p, o, h, k, row, pr, dim, u[0], u[1]);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int d = 2 * pr + e;
                            if (d < D && d < dim) {
                                const int ni = p.map.ninc[d];
                                const double y = div_exact((double)y0[d] + u[e], p.st.dns[d], p.st.rns[d]);
                                const double t = __dmul_rn(y, p.dni[d]);
                                const int iy = __double2int_rd(t);
                                const int ic = min(iy, ni - 1);
                                const double* gp = p.map.grid + (size_t)d * p.map.gstride + ic;
                                const double g0 = __ldg(gp), g1 = __ldg(gp + 1);
                                const double inc = g1 - g0;
                                const double xin = __dadd_rn(g0, __dmul_rn(inc, __dsub_rn(t, (double)iy)));   // no FMA: pyx:354
                                const double xv = iy < ni ? xin : g1;                                         // pyx:357-359
                                jac *= inc * p.dni[d];
                                if (XT) o.x[(int64_t)d * o.rows + row] = xv;
                                else tile[lane * S + d] = xv;
                                if (o.bins) btile[lane * dim + d] = (y > 0.0 && y < 1.0) ? (uint16_t)ic : (uint16_t)0xffff;   // pyx:460
                            }
                        }
                    }
                }
                o.wgt[row] = jac * (p.dv_y / (double)n_s[c]);
            }
            __syncwarp();
            // the warp's rows are consecutive: one contiguous block of x (and of bins)
            const long long wfirst = ib + (tid - lane);
            const int nlive = (int)(i1 - wfirst < 32 ? (i1 - wfirst > 0 ? i1 - wfirst : 0) : 32);
            const int64_t wrow = chunk_row + wfirst;
            const int nel = nlive * dim;
            if (!XT) {
                double* dst = o.x + wrow * dim;
                for (int e = lane; e < nel; e += 32) {
                    const int r = e / dim;
                    dst[e] = tile[r * S + (e - r * dim)];
                }
            }
            if (o.bins) {
                uint16_t* dst = o.bins + wrow * dim;
                for (int e = lane; e < nel; e += 32) dst[e] = btile[e];
            }
            __syncwarp();
        }
    }
}

template <int D>
static int launch_sample_x(const EngineP& p, const SampleOut& o, int grid, cudaStream_t st)
{
    const int dim = p.map.dim;
    const size_t tile = o.x_transposed ? 0 : sizeof(double) * (size_t)(VB_NT / 32) * 32 * (dim | 1);
    size_t smem = tile + sizeof(uint32_t) * (size_t)VB_CH * dim + (o.bins ? sizeof(uint16_t) * (size_t)VB_NT * dim : 0);
    smem = (smem + 15) & ~(size_t)15;
    cudaError_t e;
    if (o.x_transposed) {
        e = cudaFuncSetAttribute(k_sample_x<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(int)e - 1000;
        k_sample_x<D, true><<<grid, VB_NT, smem, st>>>(p, o);
    } else {
        e = cudaFuncSetAttribute(k_sample_x<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -(int)e - 1000;
        k_sample_x<D, false><<<grid, VB_NT, smem, st>>>(p, o);
    }
    return 0;
}

static int sample_common(vb200_ctx* c, uint32_t itn, int64_t chunk_begin, int64_t chunk_end, const SampleOut& o0, void* stream)
{
    if (!c->have_map || !c->have_strata) return fail(-1, "sample: map/strata not set");
    if (!c->have_plan) return fail(-1, "sample: call vb200_plan first");
    if (chunk_begin < 0 || chunk_end > c->nchunks || chunk_begin > chunk_end) return fail(-1, "sample: bad chunk range");
    CK(cudaSetDevice(c->device));
    if (chunk_begin == chunk_end) return 0;
    EngineP p;
    memset(&p, 0, sizeof p);
    p.map = c->map; p.st = c->st; p.al = c->al; p.key = c->key;
    p.itn = itn;
    p.dv_y = 1.0 / (double)c->st.nhcube;
    for (int d = 0; d < VB_MAXD; ++d) p.cstride[d] = c->cstride[d];
    p.chunk_begin = chunk_begin; p.chunk_end = chunk_end;
    p.chunk_off = (const int64_t*)c->chunk_off.p;
    int rc = vb_fetch_chunk_off(c, (cudaStream_t)stream);
    if (rc) return rc;
    long long r[2] = {c->chunk_off_host[(size_t)chunk_begin], c->chunk_off_host[(size_t)chunk_end]};
    p.row0 = r[0];
    SampleOut o = o0;
    o.rows = r[1] - r[0];
    ItemsSel it;
    rc = vb_set_items(c, chunk_begin, chunk_end, it, (cudaStream_t)stream);
    if (rc) return rc;
    p.item_off = it.off[0]; p.item_begin = it.begin[0]; p.item_end = it.end[0];
    const int64_t nch = p.item_end - p.item_begin;
    int64_t g = (int64_t)c->sm_count * 8;
    if (g > nch) g = nch;
    for (int d = 0; d < VB_MAXD; ++d) p.dni[d] = (double)c->map.ninc[d];
    CK(c->counter.ensure(sizeof(unsigned long long)));
    CK(cudaMemsetAsync(c->counter.p, 0, sizeof(unsigned long long), (cudaStream_t)stream));
    p.work_counter = (unsigned long long*)c->counter.p;
    if (o.x && o.wgt && !o.y && !o.jac1d && !o.hcube && !o.u) {
        // the integration path: x, wgt (+ training bins)
        const int dim = c->map.dim;
        int e = dim <= 4 ? launch_sample_x<4>(p, o, (int)g, (cudaStream_t)stream)
              : dim <= 8 ? launch_sample_x<8>(p, o, (int)g, (cudaStream_t)stream)
              : dim <= 10 ? launch_sample_x<10>(p, o, (int)g, (cudaStream_t)stream)
                          : launch_sample_x<VB_MAXD>(p, o, (int)g, (cudaStream_t)stream);
        if (e) return fail(-2, "sample: launch set-up failed (%s)", cudaGetErrorString((cudaError_t)(-(e + 1000))));
    } else {
        size_t smem = sizeof(uint32_t) * (size_t)VB_CH * c->map.dim;
        k_sample<<<(int)g, VB_NT, smem, (cudaStream_t)stream>>>(p, o);
    }
    c->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

static int sample_entry(vb200_ctx* c, uint32_t itn, int64_t chunk_begin, int64_t chunk_end, const double* u_dev, double* x_dev, double* wgt_dev,
                        double* y_dev, double* jac1d_dev, int64_t* hcube_dev, uint16_t* bins_dev, int x_transposed, void* stream)
{
    if (!c || !x_dev || !wgt_dev) return fail(-1, "vb200_sample: null argument");
    if (bins_dev)
        for (int d = 0; d < c->map.dim; ++d)
            if (c->map.ninc[d] > 0xffff) return fail(-1, "vb200_sample: training bins need ninc <= 65535");
    SampleOut o;
    memset(&o, 0, sizeof o);
    o.x = x_dev; o.wgt = wgt_dev; o.y = y_dev; o.jac1d = jac1d_dev; o.hcube = hcube_dev; o.x_transposed = x_transposed;
    o.bins = bins_dev;
    o.u_in = u_dev;
    return sample_common(c, itn, chunk_begin, chunk_end, o, stream);
}

extern "C" int vb200_sample(vb200_ctx* c, uint32_t itn, int64_t chunk_begin, int64_t chunk_end, double* x_dev, double* wgt_dev,
                            double* y_dev, double* jac1d_dev, int64_t* hcube_dev, uint16_t* bins_dev, int x_transposed, void* stream)
{
    return sample_entry(c, itn, chunk_begin, chunk_end, nullptr, x_dev, wgt_dev, y_dev, jac1d_dev, hcube_dev, bins_dev, x_transposed, stream);
}

extern "C" int vb200_sample_from_uniforms(vb200_ctx* c, uint32_t itn, int64_t chunk_begin, int64_t chunk_end, const double* u_dev,
                                          double* x_dev, double* wgt_dev, double* y_dev, double* jac1d_dev, int64_t* hcube_dev,
                                          uint16_t* bins_dev, int x_transposed, void* stream)
{
    if (!u_dev) return fail(-1, "vb200_sample_from_uniforms: u_dev is NULL");
    return sample_entry(c, itn, chunk_begin, chunk_end, u_dev, x_dev, wgt_dev, y_dev, jac1d_dev, hcube_dev, bins_dev, x_transposed, stream);
}

extern "C" int vb200_uniforms(vb200_ctx* c, uint32_t itn, int64_t chunk_begin, int64_t chunk_end, double* u_dev, void* stream)
{
    if (!c || !u_dev) return fail(-1, "vb200_uniforms: null argument");
    SampleOut o;
    memset(&o, 0, sizeof o);
    o.u = u_dev;
    return sample_common(c, itn, chunk_begin, chunk_end, o, stream);
}

