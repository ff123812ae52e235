// vegas_b200.cu -- C ABI of libvegas_b200.so (see include/vegas_b200.h): context, integrand registry,
// allocation pre-pass (work items), engine launches, FP64 peak probe.  The samplers are in
// sampler.cu, the AdaptiveMap methods in mapops.cu, the restratify profile in profile.cu.
#include "ctx.h"

static_assert(VB200_MAXDIM == VB_MAXD, "header/kernels disagree on MAXDIM");
static_assert(VB200_CHUNK == VB_CH, "header/kernels disagree on chunk size");
static_assert(VB200_UPDATE_SIGF == VBF_UPDATE_SIGF && VB200_TRAIN == VBF_TRAIN &&
              VB200_TRAIN_ERRORS == VBF_TRAIN_ERRORS && VB200_CORRELATE == VBF_CORRELATE, "flag mismatch");

static thread_local std::string g_err;
int vb_fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

extern "C" int vb200_abi_version(void) { return VB200_ABI_VERSION; }
extern "C" const char* vb200_last_error(void) { return g_err.c_str(); }

extern "C" int vb200_create(vb200_ctx** out, int device)
{
    if (!out) return fail(-1, "vb200_create: out is NULL");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(-1, "vb200_create: no CUDA device %d (have %d)", device, ndev);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(-3, "vb200_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    vb200_ctx* c = new vb200_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_per_sm = prop.sharedMemPerMultiprocessor;
    c->smem_per_block_optin = prop.sharedMemPerBlockOptin;
    philox_make_key(0, c->key);
    memset(&c->map, 0, sizeof c->map);
    memset(&c->st, 0, sizeof c->st);
    memset(&c->al, 0, sizeof c->al);
    *out = c;
    return 0;
}

extern "C" void vb200_destroy(vb200_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    c->grid.release(); c->chunk_tot.release(); c->chunk_off.release(); c->chunk_items.release(); c->item_off.release(); c->super_items.release(); c->super_item_off.release(); c->stats.release();
    if (c->head_host) cudaFreeHost(c->head_host);
    c->fparams.release(); c->partials.release(); c->scratch.release(); c->counter.release(); c->ecounter.release(); c->sigf_shadow.release();
    delete c;
}

extern "C" int vb200_set_seed(vb200_ctx* c, uint64_t seed)
{
    if (!c) return fail(-1, "null context");
    c->seed = seed;
    philox_make_key(seed, c->key);
    return 0;
}

extern "C" int vb200_set_map(vb200_ctx* c, const double* grid_host, const int64_t* ninc, int dim, int64_t gstride)
{
    if (!c || !grid_host || !ninc) return fail(-1, "vb200_set_map: null argument");
    if (dim < 1 || dim > VB_MAXD) return fail(-1, "vb200_set_map: dim=%d outside 1..%d", dim, VB_MAXD);
    for (int d = 0; d < dim; ++d)
        if (ninc[d] < 1 || ninc[d] + 1 > gstride || ninc[d] > 0x3fffffff)
            return fail(-1, "vb200_set_map: bad ninc[%d]=%lld (gstride %lld)", d, (long long)ninc[d], (long long)gstride);
    CK(cudaSetDevice(c->device));
    size_t bytes = sizeof(double) * (size_t)dim * (size_t)gstride;
    CK(c->grid.ensure(bytes));
    CK(cudaMemcpy(c->grid.p, grid_host, bytes, cudaMemcpyHostToDevice));
    c->map.grid = (const double*)c->grid.p;
    c->map.dim = dim;
    c->map.gstride = (int)gstride;
    for (int d = 0; d < VB_MAXD; ++d) c->map.ninc[d] = d < dim ? (int)ninc[d] : 1;
    c->have_map = true;
    return 0;
}

extern "C" int vb200_set_strata(vb200_ctx* c, const int64_t* nstrat, int dim, int64_t slab, int rank, int world,
                                int64_t* nlocal_out)
{
    if (!c || !nstrat) return fail(-1, "vb200_set_strata: null argument");
    if (dim < 1 || dim > VB_MAXD) return fail(-1, "vb200_set_strata: dim=%d outside 1..%d", dim, VB_MAXD);
    if (world < 1 || rank < 0 || rank >= world) return fail(-1, "vb200_set_strata: bad rank %d / world %d", rank, world);
    if (slab < VB_CH || slab % VB_CH) return fail(-1, "vb200_set_strata: slab must be a positive multiple of %d", VB_CH);
    long double prod = 1;
    int64_t nh = 1;
    for (int d = 0; d < dim; ++d) {
        if (nstrat[d] < 1 || nstrat[d] > 0x3fffffff) return fail(-1, "vb200_set_strata: bad nstrat[%d]", d);
        prod *= (long double)nstrat[d];
        if (prod > 4.0e18L) return fail(-1, "vb200_set_strata: too many hypercubes");
        c->cstride[d] = nh;
        nh *= nstrat[d];
    }
    StrataP& s = c->st;
    s.nhcube = nh; s.slab = slab; s.rank = rank; s.world = world;
    for (int d = 0; d < VB_MAXD; ++d) {
        s.nstrat[d] = d < dim ? (int)nstrat[d] : 1;
        s.dns[d] = (double)s.nstrat[d];
        s.rns[d] = 1.0 / s.dns[d];
        s.nsm[d] = (s.nstrat[d] >= 2 && s.nstrat[d] < 32768) ? (uint32_t)((0xffffffffull / (uint64_t)s.nstrat[d]) + 1ull) : 0u;
        if (d >= dim) c->cstride[d] = nh;
    }
    // dense local index space: my slabs in global order; only the globally last slab is partial
    // (rounds of `world` slabs, each dealt in rotated rank order: common.cuh, slab_rot)
    const int64_t nslab = (nh + slab - 1) / slab;
    const int64_t rounds = nslab / world, rem = nslab % world;
    const bool extra = rem > 0 && (rank + slab_rot(rounds, world)) % world < rem;     // a slab of the incomplete last round
    const int64_t mine = rounds + (extra ? 1 : 0);
    int64_t nlocal = mine * slab;
    const int64_t last_round = (nslab - 1) / world;
    if (mine > 0 && (rank + slab_rot(last_round, world)) % world == (nslab - 1) % world) nlocal -= nslab * slab - nh;
    s.nlocal = nlocal;
    c->nchunks = (nlocal + VB_CH - 1) / VB_CH;
    c->have_strata = true;
    c->have_plan = false;
    if (nlocal_out) *nlocal_out = nlocal;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// integrand registry
// ---------------------------------------------------------------------------------------------
extern "C" int vb200_set_integrand(vb200_ctx* c, int id, const void* params, size_t nbytes, int* nf_out)
{
    if (!c || !params) return fail(-1, "vb200_set_integrand: null argument");
    if (!c->have_map) return fail(-1, "vb200_set_integrand: call vb200_set_map first (dim is needed)");
    CK(cudaSetDevice(c->device));
    const int dim = c->map.dim;
    c->fid = -1;
    c->light_hint = false;
    c->very_light = false;
    switch (id) {
    case VB200_F_POLY: {
        if (nbytes != sizeof(vb200_poly_t)) return fail(-1, "poly: params size %zu != %zu", nbytes, sizeof(vb200_poly_t));
        const vb200_poly_t* q = (const vb200_poly_t*)params;
        FPoly f;
        f.c0 = q->c0;
        for (int d = 0; d < VB_MAXD; ++d) { f.c[d] = q->c[d]; f.p[d] = q->p[d]; }
        c->functor.assign((char*)&f, (char*)&f + sizeof f);
        c->nf = 1;
        c->light_hint = c->very_light = true;
        break;
    }
    case VB200_F_GAUSS_MIX: {
        if (nbytes != sizeof(vb200_gaussmix_t)) return fail(-1, "gaussmix: bad params size");
        const vb200_gaussmix_t* q = (const vb200_gaussmix_t*)params;
        if (q->npeak < 1 || !q->centers_host) return fail(-1, "gaussmix: npeak < 1 or no centers");
        size_t bytes = sizeof(double) * (size_t)q->npeak * dim;
        CK(c->fparams.ensure(bytes));
        CK(cudaMemcpy(c->fparams.p, q->centers_host, bytes, cudaMemcpyHostToDevice));
        FGaussMix f;
        f.centers = (const double*)c->fparams.p; f.npeak = q->npeak; f.a = q->a; f.norm = q->norm;
        c->functor.assign((char*)&f, (char*)&f + sizeof f);
        c->nf = 1;
        c->light_hint = q->npeak <= 4;
        c->very_light = q->npeak <= 2;
        break;
    }
    case VB200_F_RIDGE: {
        if (nbytes != sizeof(vb200_ridge_t)) return fail(-1, "ridge: bad params size");
        const vb200_ridge_t* q = (const vb200_ridge_t*)params;
        if (q->n < 1 || !q->x0_host) return fail(-1, "ridge: n < 1 or no x0");
        size_t bytes = sizeof(double) * (size_t)q->n;
        CK(c->fparams.ensure(2 * bytes));
        std::vector<double> xs((size_t)q->n);
        for (int k = 0; k < q->n; ++k) xs[k] = sqrt(q->a * (double)dim) * q->x0_host[k];
        CK(cudaMemcpy(c->fparams.p, q->x0_host, bytes, cudaMemcpyHostToDevice));
        CK(cudaMemcpy((char*)c->fparams.p + bytes, xs.data(), bytes, cudaMemcpyHostToDevice));
        FRidge f;
        memset(&f, 0, sizeof f);
        f.x0 = (const double*)c->fparams.p; f.xs = f.x0 + q->n; f.n = q->n; f.mode = q->mode; f.a = q->a; f.norm = q->norm;
        f.scale = q->norm / (double)q->n;
        // up to VB_RIDGE_PMAX centres ride in the kernel parameters (padded with the last one to whole lock-step groups)
        f.npar = (q->n <= VB_RIDGE_PMAX && vb_env_int("VB200_RIDGE_PAR", 1)) ? q->n : 0;
        if (f.npar)
            for (int k = 0; k < VB_RIDGE_PMAX + 8; ++k) {
                const int kk = k < q->n ? k : q->n - 1;
                f.cpar[k] = q->mode == 1 ? xs[kk] : q->x0_host[kk];
            }
        c->functor.assign((char*)&f, (char*)&f + sizeof f);
        c->nf = 1;
        c->light_hint = q->n <= vb_env_int("VB200_RIDGE_LIGHT_N", 128) && q->mode == 0;   // FRidgeLight: 4-wide lock-step (measured: N = 100 light 22.1 ms, heavy 24.7; N = 1000 light 180.8, heavy 173.0)
        c->very_light = c->light_hint && q->n <= 4;
        break;
    }
    case VB200_F_GENZ_OSC: case VB200_F_GENZ_PRODPEAK: case VB200_F_GENZ_CORNER:
    case VB200_F_GENZ_GAUSS: case VB200_F_GENZ_C0: case VB200_F_GENZ_DISC: {
        if (nbytes != sizeof(vb200_genz_t)) return fail(-1, "genz: bad params size");
        const vb200_genz_t* q = (const vb200_genz_t*)params;
        FGenz f;
        f.kind = id;
        for (int d = 0; d < VB_MAXD; ++d) { f.a[d] = q->a[d]; f.u[d] = q->u[d]; }
        c->functor.assign((char*)&f, (char*)&f + sizeof f);
        c->nf = 1;
        c->light_hint = c->very_light = true;
        break;
    }
    case VB200_F_PATHINT: {
        if (nbytes != sizeof(vb200_pathint_t)) return fail(-1, "pathint: bad params size");
        const vb200_pathint_t* q = (const vb200_pathint_t*)params;
        if (q->nx0 != 0 && q->nx0 != 6) return fail(-1, "pathint: nx0=%d not compiled in (0 or 6)", q->nx0);
        if (dim < 3) return fail(-1, "pathint: needs dim >= 3");
        const double PI = 3.14159265358979323846;
        double norm = pow(q->m * dim / 2. / PI / q->T, dim / 2.);
        if (q->nx0 == 0) {
            FPathInt<0> f;
            f.T = q->T; f.m = q->m; f.xscale = q->xscale; f.c2 = q->c2; f.c4 = q->c4;
            f.norm = norm; f.norm_x0 = norm / PI; f.x0list[0] = 0;
            c->functor.assign((char*)&f, (char*)&f + sizeof f);
        } else {
            FPathInt<6> f;
            f.T = q->T; f.m = q->m; f.xscale = q->xscale; f.c2 = q->c2; f.c4 = q->c4;
            f.norm = norm; f.norm_x0 = norm / PI;
            for (int i = 0; i < 6; ++i) f.x0list[i] = q->x0list[i];
            c->functor.assign((char*)&f, (char*)&f + sizeof f);
        }
        c->nx0 = q->nx0;
        c->nf = 1 + q->nx0;
        break;
    }
    default:
        return fail(-1, "vb200_set_integrand: unknown integrand id %d", id);
    }
    c->fid = id;
    if (nf_out) *nf_out = c->nf;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// allocation pre-pass + chunk offsets
// ---------------------------------------------------------------------------------------------
// stats: [0] sum  [1] VB_STAT_MIN_TOP - min  [2] max  [3] largest chunk total  [4] items  [5] items of the light geometry
// (every word starts at zero -- the minimum is kept as a maximum of its complement -- so the iteration buffer's one
// memset is also the initialisation of the statistics that live in it)
// chunk_items[lc] = work items chunk lc is cut into: 1, or ceil(total / item_samples) when the vegas+
// allocation piled more than item_samples samples onto its cubes (engine.cuh, "items").  The engine
// never splits a cube (items that no cube starts in are empty); the samplers split by rows.
// sum_sigf_dev != nullptr (planning ahead, vb200_plan_ahead): neval_sigf = neval_scaled / *sum_sigf_dev is
// formed here, with the same correctly rounded division the host performs once it has read sum_sigf back
#define VB_STAT_MIN_TOP 0x7fffffffffffffffLL
__global__ void __launch_bounds__(VB_NT) k_plan(StrataP st, AllocP al, int64_t nchunks, int32_t* neval_out,
                                                long long* chunk_tot, long long* chunk_items, long long item_samples,
                                                long long* stats, const double* sum_sigf_dev, double neval_scaled)
{
    if (sum_sigf_dev) al.neval_sigf = __ddiv_rn(neval_scaled, *sum_sigf_dev);
    __shared__ long long red[VB_NT / 32];
    __shared__ int rmin[VB_NT / 32], rmax[VB_NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    long long my_sum = 0, my_maxc = 0, my_items = 0;
    int my_min = 0x7fffffff, my_max = 0;
    for (int64_t lc = blockIdx.x; lc < nchunks; lc += gridDim.x) {
        int64_t lh = lc * VB_CH + tid;
        int n = 0;
        if (lh < st.nlocal) {
            n = alloc_neval(al, lh);
            if (neval_out) neval_out[lh] = n;
            my_min = min(my_min, n);
            my_max = max(my_max, n);
        }
        long long s = n;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        __syncthreads();
        if (lane == 0) red[w] = s;
        __syncthreads();
        if (tid == 0) {
            long long t = 0;
            for (int i = 0; i < VB_NT / 32; ++i) t += red[i];
            chunk_tot[lc] = t;
            long long m = (t + item_samples - 1) / item_samples;
            m = m < 1 ? 1 : (m > VB_MAXITEMS ? VB_MAXITEMS : m);
            chunk_items[lc] = m;
            my_items += m;
            my_sum += t;
            if (t > my_maxc) my_maxc = t;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_min = min(my_min, __shfl_xor_sync(0xffffffffu, my_min, o));
        my_max = max(my_max, __shfl_xor_sync(0xffffffffu, my_max, o));
    }
    if (lane == 0) { rmin[w] = my_min; rmax[w] = my_max; }
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < VB_NT / 32; ++i) { my_min = min(my_min, rmin[i]); my_max = max(my_max, rmax[i]); }
        atomicAdd((unsigned long long*)&stats[0], (unsigned long long)my_sum);
        atomicMax(&stats[1], VB_STAT_MIN_TOP - (long long)my_min);
        atomicMax(&stats[2], (long long)my_max);
        atomicMax(&stats[3], my_maxc);
        atomicAdd((unsigned long long*)&stats[4], (unsigned long long)my_items);
    }
}

// exclusive scan of chunk_tot[n] into chunk_off[n+1]; one CTA, sequential over 1024-wide segments
__global__ void __launch_bounds__(1024) k_scan(const long long* tot, int64_t n, long long* off)
{
    __shared__ long long wsum[32];
    __shared__ long long carry_s;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        int64_t i = base + tid;
        long long v = i < n ? tot[i] : 0, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        long long b = 0, t = 0;
        for (int j = 0; j < 32; ++j) { if (j < w) b += wsum[j]; t += wsum[j]; }
        long long carry = carry_s;
        if (i < n) off[i] = carry + b + x - v;
        __syncthreads();
        if (tid == 0) carry_s = carry + t;
        __syncthreads();
    }
    if (tid == 0) off[n] = carry_s;
}

// items of the light geometry's chunks (group consecutive VB_CH-cube chunks each)
__global__ void k_super_items(const long long* chunk_tot, int64_t nchunks, int group, long long item_samples,
                              int max_items, int64_t nsuper, long long* out, long long* stats)
{
    long long mine = 0;
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nsuper; s += (int64_t)gridDim.x * blockDim.x) {
        long long t = 0;
        for (int g = 0; g < group; ++g) if (s * group + g < nchunks) t += chunk_tot[s * group + g];
        long long m = (t + item_samples - 1) / item_samples;
        m = m < 1 ? 1 : (m > max_items ? max_items : m);
        out[s] = m;
        mine += m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd((unsigned long long*)&stats[5], (unsigned long long)mine);
}

// launch the allocation pre-pass; statistics (6 int64, see above) go to stats_dev.  No synchronisation.
static int plan_launch(vb200_ctx* c, const double* sigf_dev, double neval_sigf, const double* sum_sigf_dev, double neval_scaled,
                       int64_t min_nh, int64_t max_nh, int64_t uniform_neval, int32_t* neval_hcube_dev, long long* stats_dev,
                       cudaStream_t st, bool stats_zeroed = false)
{
    if (!c->have_strata) return fail(-1, "vb200_plan: call vb200_set_strata first");
    if (min_nh < 1 || min_nh > 0x7fffffff || uniform_neval > 0x7fffffff)
        return fail(-1, "vb200_plan: min_neval_hcube / uniform_neval out of int32 range");
    if (max_nh < min_nh) max_nh = min_nh;                       // pyx:1667-1669
    if (max_nh > 0x7fffffff) max_nh = 0x7fffffff;
    if (!sigf_dev && uniform_neval < 1) return fail(-1, "vb200_plan: uniform_neval < 1");
    CK(cudaSetDevice(c->device));
    c->al.sigf = sigf_dev;
    c->al.neval_sigf = neval_sigf;
    c->al.min_neval_hcube = (int)min_nh;
    c->al.max_neval_hcube = (int)max_nh;
    c->al.uniform_neval = (int)uniform_neval;
    c->have_plan = false;                                       // until vb200_plan / vb200_plan_commit install the statistics
    const int64_t nch = c->nchunks;
    CK(c->chunk_tot.ensure(sizeof(long long) * (size_t)(nch + 1)));
    CK(c->chunk_off.ensure(sizeof(long long) * (size_t)(nch + 1)));
    CK(c->chunk_items.ensure(sizeof(long long) * (size_t)(nch + 1)));
    CK(c->item_off.ensure(sizeof(long long) * (size_t)(nch + 1)));
    // samples per work item: VB_ITEM when there are chunks to spare; with few chunks (the reference's everyday sizes:
    // neval = 1e4 is 5 chunks on a 148-SM GPU) finer items, so that more than a handful of CTAs have work
    // (config 1: 0.21 -> 0.19 ms per iteration at neval = 1e4, 0.29 -> 0.26 at 1e6)
    long long item_samples = vb_env_int("VB200_ITEM", 0);
    if (item_samples <= 0) item_samples = (long long)VB_ITEM * nch / (32LL * c->sm_count);
    if (item_samples > VB_ITEM && !vb_env_int("VB200_ITEM", 0)) item_samples = VB_ITEM;
    if (item_samples < 256) item_samples = 256;
    // the light geometry's chunks are planned only where run_engine can use them (same test as there)
    const int force_light = vb_env_int("VB200_LIGHT", -1);
    const bool plan_light = c->light_hint && nch > 0 && force_light != 0
                            && (force_light == 1 || c->st.nlocal >= (int64_t)VB_LCH * 4 * c->sm_count);
    if (!stats_zeroed) CK(cudaMemsetAsync(stats_dev, 0, sizeof(long long) * 6, st));
    if (nch > 0) {
        int grid = (int)(nch < (int64_t)c->sm_count * 8 ? nch : (int64_t)c->sm_count * 8);
        k_plan<<<grid, VB_NT, 0, st>>>(c->st, c->al, nch, neval_hcube_dev, (long long*)c->chunk_tot.p,
                                       (long long*)c->chunk_items.p, item_samples, stats_dev, sum_sigf_dev, neval_scaled);
        c->launches += 1;
        if (plan_light) {
            const int group = VB_LCH / VB_CH;
            c->nsuper = (nch + group - 1) / group;
            CK(c->super_items.ensure(sizeof(long long) * (size_t)(c->nsuper + 1)));
            CK(c->super_item_off.ensure(sizeof(long long) * (size_t)(c->nsuper + 1)));
            k_super_items<<<(int)((c->nsuper + 255) / 256 < 1024 ? (c->nsuper + 255) / 256 : 1024), 256, 0, st>>>(
                (const long long*)c->chunk_tot.p, nch, group, item_samples * group, VB_LCH, c->nsuper, (long long*)c->super_items.p,
                stats_dev);
            c->launches += 1;
        }
        CK(cudaGetLastError());
    }
    c->plan_light = plan_light;
    return 0;
}

// install the statistics of the launched pre-pass (host values) in the context
static void plan_install(vb200_ctx* c, const long long out_in[6])
{
    long long out[6];
    memcpy(out, out_in, sizeof out);
    out[1] = VB_STAT_MIN_TOP - out[1];
    const int64_t nch = c->nchunks;
    c->plan_super_items = (c->plan_light && nch > 0) ? out[5] : -1;
    if (nch == 0) out[1] = 0;
    c->plan_total = out[0]; c->plan_min = out[1]; c->plan_max = out[2]; c->plan_max_chunk = out[3];
    c->plan_items = out[4];
    c->have_plan = true;
    c->chunk_off_valid = c->item_off_valid = c->super_off_valid = false;
    c->chunk_off_host.clear();
    c->item_off_host.clear();
}

extern "C" int vb200_plan(vb200_ctx* c, const double* sigf_dev, double neval_sigf, int64_t min_nh, int64_t max_nh,
                          int64_t uniform_neval, int32_t* neval_hcube_dev, int64_t stats_host[4], void* stream)
{
    if (!c) return fail(-1, "null context");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    CK(c->stats.ensure(sizeof(long long) * 6));
    int rc = plan_launch(c, sigf_dev, neval_sigf, nullptr, 0.0, min_nh, max_nh, uniform_neval, neval_hcube_dev, (long long*)c->stats.p, st);
    if (rc) return rc;
    long long out[6];
    CK(cudaMemcpyAsync(out, c->stats.p, sizeof out, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    plan_install(c, out);
    if (stats_host) { stats_host[0] = c->plan_total; stats_host[1] = c->plan_min; stats_host[2] = c->plan_max; stats_host[3] = c->nchunks; }
    return 0;
}

extern "C" int vb200_plan_ahead(vb200_ctx* c, const double* sigf_dev, const double* sum_sigf_dev, double neval_scaled,
                                int64_t min_nh, int64_t max_nh, int64_t uniform_neval, int64_t* stats_dev, void* stream)
{
    if (!c || !sigf_dev || !sum_sigf_dev || !stats_dev) return fail(-1, "vb200_plan_ahead: null argument");
    return plan_launch(c, sigf_dev, 0.0, sum_sigf_dev, neval_scaled, min_nh, max_nh, uniform_neval, nullptr, (long long*)stats_dev,
                       (cudaStream_t)stream);
}

extern "C" int vb200_plan_commit(vb200_ctx* c, double neval_sigf, const int64_t stats_host[6], int64_t stats_out[4])
{
    if (!c || !stats_host) return fail(-1, "vb200_plan_commit: null argument");
    if (c->have_plan) return fail(-1, "vb200_plan_commit: no pre-pass is waiting (call vb200_plan_ahead first)");
    c->al.neval_sigf = neval_sigf;
    long long out[6];
    for (int i = 0; i < 6; ++i) out[i] = stats_host[i];
    plan_install(c, out);
    if (stats_out) { stats_out[0] = c->plan_total; stats_out[1] = c->plan_min; stats_out[2] = c->plan_max; stats_out[3] = c->nchunks; }
    return 0;
}

// row offsets of the chunks (exclusive scan of the chunk totals), made on first use after a plan
int vb_ensure_chunk_off(vb200_ctx* c, cudaStream_t st)
{
    if (c->chunk_off_valid) return 0;
    CK(cudaSetDevice(c->device));
    if (c->nchunks > 0) {
        k_scan<<<1, 1024, 0, st>>>((const long long*)c->chunk_tot.p, c->nchunks, (long long*)c->chunk_off.p);
        c->launches += 1;
        CK(cudaGetLastError());
    } else {
        CK(cudaMemsetAsync(c->chunk_off.p, 0, sizeof(long long), st));
    }
    c->chunk_off_valid = true;
    return 0;
}

int vb_fetch_chunk_off(vb200_ctx* c, cudaStream_t st)
{
    if (!c->chunk_off_host.empty()) return 0;
    int rc0 = vb_ensure_chunk_off(c, st);
    if (rc0) return rc0;
    CK(cudaStreamSynchronize(st));
    CK(cudaSetDevice(c->device));
    c->chunk_off_host.resize((size_t)c->nchunks + 1);
    CK(cudaMemcpy(c->chunk_off_host.data(), c->chunk_off.p, sizeof(long long) * (size_t)(c->nchunks + 1), cudaMemcpyDeviceToHost));
    return 0;
}

// work items of the local chunk range [chunk_begin, chunk_end) for the heavy geometry (and, for
// whole-range launches, of the light geometry's chunks)
int vb_set_items(vb200_ctx* c, int64_t chunk_begin, int64_t chunk_end, ItemsSel& it, cudaStream_t st)
{
    it.off[0] = it.off[1] = nullptr;
    it.begin[0] = chunk_begin; it.end[0] = chunk_end;
    it.begin[1] = 0; it.end[1] = -1;                   // light: unavailable unless set below
    const bool whole = chunk_begin == 0 && chunk_end == c->nchunks;
    if (whole && c->plan_super_items >= 0) {
        it.end[1] = c->nsuper;
        if (c->plan_super_items != c->nsuper) {
            if (!c->super_off_valid) {
                k_scan<<<1, 1024, 0, st>>>((const long long*)c->super_items.p, c->nsuper, (long long*)c->super_item_off.p);
                c->launches += 1;
                CK(cudaGetLastError());
                c->super_off_valid = true;
            }
            it.off[1] = (const int64_t*)c->super_item_off.p; it.end[1] = c->plan_super_items;
        }
    }
    if (c->plan_items == c->nchunks) return 0;         // nothing was split: item j == chunk j
    if (!c->item_off_valid) {
        k_scan<<<1, 1024, 0, st>>>((const long long*)c->chunk_items.p, c->nchunks, (long long*)c->item_off.p);
        c->launches += 1;
        CK(cudaGetLastError());
        c->item_off_valid = true;
    }
    it.off[0] = (const int64_t*)c->item_off.p;
    if (whole) { it.begin[0] = 0; it.end[0] = c->plan_items; return 0; }
    if (c->item_off_host.empty()) {
        CK(cudaSetDevice(c->device));
        CK(cudaStreamSynchronize(st));
        c->item_off_host.resize((size_t)c->nchunks + 1);
        CK(cudaMemcpy(c->item_off_host.data(), c->item_off.p, sizeof(long long) * (size_t)(c->nchunks + 1), cudaMemcpyDeviceToHost));
    }
    it.begin[0] = c->item_off_host[(size_t)chunk_begin];
    it.end[0] = c->item_off_host[(size_t)chunk_end];
    return 0;
}

extern "C" int vb200_chunk_offsets(vb200_ctx* c, int64_t* out_host, int64_t count)
{
    if (!c || !out_host) return fail(-1, "null argument");
    if (!c->have_plan) return fail(-1, "vb200_chunk_offsets: call vb200_plan first");
    if (count < 0 || count > c->nchunks + 1) return fail(-1, "vb200_chunk_offsets: count out of range");
    int rc = vb_fetch_chunk_off(c);
    if (rc) return rc;
    memcpy(out_host, c->chunk_off_host.data(), sizeof(int64_t) * (size_t)count);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// engine launches
// ---------------------------------------------------------------------------------------------
// acc[j] += sum over CTAs of partials[cta][j]: lane l adds CTAs l, l + 32, ... in order, then a fixed shuffle tree
// (deterministic for a given grid; one thread walking 592 partials was 5-10 us of a 0.1 ms iteration)
__global__ void __launch_bounds__(1024) k_finalize(const double* partials, int nblocks, int nacc, double* acc,
                                                   unsigned long long* work_counter)
{
    const int lane = threadIdx.x & 31;
    for (int j = threadIdx.x >> 5; j < nacc; j += (int)(blockDim.x >> 5)) {       // one warp per accumulator
        double t = 0.0;
        for (int b = lane; b < nblocks; b += 32) t += partials[(size_t)b * nacc + j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) acc[j] += t;
    }
    if (threadIdx.x == 0) *work_counter = 0ull;      // the engine's item counter starts the next launch at zero
}

static int fill_engine(vb200_ctx* c, EngineP& p, uint32_t itn, double beta, int flags, double* sigf, double* sum_f,
                       uint64_t* n_f, int64_t hstride, int32_t* status)
{
    if (!c->have_map || !c->have_strata) return fail(-1, "engine: map/strata not set");
    if (!c->have_plan) return fail(-1, "engine: call vb200_plan first");
    if (c->map.dim != 0 && c->st.nhcube <= 0) return fail(-1, "engine: bad strata");
    if ((flags & (VB200_TRAIN | VB200_TRAIN_ERRORS)) && (!sum_f || !n_f)) return fail(-1, "engine: training buffers are NULL");
    if ((flags & VB200_UPDATE_SIGF) && !sigf) return fail(-1, "engine: sigf is NULL");
    if (!status) return fail(-1, "engine: status is NULL");
    for (int d = 0; d < c->map.dim; ++d)
        if (c->map.ninc[d] > hstride && (flags & (VB200_TRAIN | VB200_TRAIN_ERRORS)))
            return fail(-1, "engine: hstride %lld < ninc[%d]", (long long)hstride, d);
    memset(&p, 0, sizeof p);
    p.map = c->map; p.st = c->st; p.al = c->al; p.key = c->key;
    p.itn = itn; p.flags = flags;
    p.dv_y = 1.0 / (double)c->st.nhcube;
    p.beta_half = beta / 2.;
    p.sigf_out = sigf; p.sum_f = sum_f; p.n_f = (unsigned long long*)n_f; p.hstride = (int)hstride;
    p.status = status;
    for (int d = 0; d < VB_MAXD; ++d) { p.cstride[d] = c->cstride[d]; p.dni[d] = (double)c->map.ninc[d]; }
    return 0;
}

static int do_launch_fused(vb200_ctx* c, const EngineP& p, LaunchCfg& cfg, cudaStream_t st)
{
    const void* f = c->functor.data();
    switch (c->fid) {
    case VB200_F_POLY: return launch_fused_poly(p, f, cfg, st);
    case VB200_F_GAUSS_MIX: return launch_fused_gaussmix(p, f, cfg, st);
    case VB200_F_RIDGE: return launch_fused_ridge(p, f, cfg, st);
    case VB200_F_PATHINT: cfg.light = false; return launch_fused_pathint(p, f, c->nx0, cfg, st);
    default: return launch_fused_genz(p, f, cfg, st);
    }
}

static int run_engine(vb200_ctx* c, EngineP& p, int nf, bool fused, double* acc, cudaStream_t st)
{
    if (p.chunk_end - p.chunk_begin <= 0) return 0;
    LaunchCfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.sm_count = c->sm_count;
    cfg.smem_per_sm = c->smem_per_sm;
    cfg.smem_optin = c->smem_per_block_optin;
    // light geometry (two 256-thread CTAs per SM): cheap integrand, digits fit 16 bits, and enough big chunks
    // to keep every SM busy; VB200_LIGHT=0/1 overrides the work-size test (developer switch)
    bool light = fused && c->light_hint;
    for (int d = 0; d < c->map.dim; ++d) if (c->st.nstrat[d] > (c->map.dim > 10 ? 255 : 65535)) light = false;   // FusedSrc::dig_t
    const int force = vb_env_int("VB200_LIGHT", -1);
    if (force == 0) light = false;
    if (force != 1 && c->st.nlocal < (int64_t)VB_LCH * 4 * c->sm_count) light = false;
    ItemsSel it;
    int rc_items = vb_set_items(c, p.chunk_begin, p.chunk_end, it, st);
    if (rc_items) return rc_items;
    if (it.end[1] < 0) light = false;                  // light chunks were not planned (set_integrand after plan)
    cfg.light = light;
    cfg.very_light = c->very_light;
    for (int g = 0; g < 2; ++g) { cfg.item_off[g] = it.off[g]; cfg.item_begin[g] = it.begin[g]; cfg.item_end[g] = it.end[g]; }
    // k_reduce (TMA-staged rows) for up to 4 outputs; with more, the per-thread covariance accumulators make it a
    // one-CTA-per-SM kernel and k_engine<BufferSrc> is the faster one (B200, 7 outputs: 3.8 against 4.1 ms per 8.4M rows)
    const int bulk_mode = vb_env_int("VB200_REDUCE_BULK", 1);     // 0: never, 1: as measured, 2: whenever possible (developer switch)
    const bool bulk = !fused && bulk_mode != 0 && (nf <= 4 || bulk_mode == 2) && reduce_bulk_ok(p);
    auto launch = [&](cudaStream_t s) { return fused ? do_launch_fused(c, p, cfg, s) : (bulk ? launch_reduce(p, nf, cfg, s) : launch_buffer(p, nf, cfg, s)); };
    int grid = launch(VB_DRYRUN);
    if (grid == -22) return fail(-4, "engine: no kernel compiled for dim=%d nf=%d integrand=%d", p.map.dim, nf, c->fid);
    if (grid < 0) return fail(-2, "engine: occupancy query failed (%d)", grid);
    const int nacc = nf + nf * (nf + 1) / 2 + 1;
    CK(c->partials.ensure(sizeof(double) * (size_t)grid * nacc));
    p.partials = (double*)c->partials.p;
    if (c->plan_max > cfg.cap) {
        p.scratch_stride = c->plan_max;
        CK(c->scratch.ensure(sizeof(double) * (size_t)grid * nf * (size_t)c->plan_max));
        p.scratch = (double*)c->scratch.p;
    }
    if (!c->ecounter.p) {                              // zeroed once; k_finalize puts it back to zero after every launch
        CK(c->ecounter.ensure(sizeof(unsigned long long)));
        CK(cudaMemsetAsync(c->ecounter.p, 0, sizeof(unsigned long long), st));
    }
    p.work_counter = (unsigned long long*)c->ecounter.p;
    // sigf is updated in place and also drives the allocation: while a chunk is shared by several
    // CTAs (items), one of them must not see the new sigf of a cube another has already finished
    // -> write to a shadow buffer and copy the launch's cube range back afterwards
    double* const sigf_user = p.sigf_out;
    const bool shadow = (it.off[0] != nullptr || it.off[1] != nullptr) && (p.flags & VBF_UPDATE_SIGF) && sigf_user != nullptr;
    if (shadow) {
        CK(c->sigf_shadow.ensure(sizeof(double) * (size_t)c->st.nlocal));
        p.sigf_out = (double*)c->sigf_shadow.p;
    }
    int g2 = launch(st);
    if (g2 < 0) return fail(-2, "engine: launch failed (%d: %s)", g2, cudaGetErrorString((cudaError_t)(-(g2 + 1000))));
    if (shadow) {
        const int64_t lo = p.chunk_begin * VB_CH;
        const int64_t hi = p.chunk_end * VB_CH < c->st.nlocal ? p.chunk_end * VB_CH : c->st.nlocal;
        CK(cudaMemcpyAsync(sigf_user + lo, (const double*)c->sigf_shadow.p + lo, sizeof(double) * (size_t)(hi - lo),
                           cudaMemcpyDeviceToDevice, st));
    }
    c->last_grid = g2; c->last_bps = cfg.blocks_per_sm; c->last_smem = (int64_t)cfg.smem; c->last_wtot = cfg.wtot;
    c->last_nt = cfg.nt; c->last_ch = cfg.ch;
    k_finalize<<<1, 32 * (nacc < 32 ? nacc : 32), 0, st>>>(p.partials, g2, nacc, acc, p.work_counter);
    c->launches += 2;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int vb200_iterate_fused(vb200_ctx* c, uint32_t itn, double beta, int flags, double* sigf_dev, double* acc_dev,
                                   double* sum_f_dev, uint64_t* n_f_dev, int64_t hstride, int32_t* status_dev, void* stream)
{
    if (!c || !acc_dev) return fail(-1, "vb200_iterate_fused: null argument");
    if (c->fid < 0) return fail(-1, "vb200_iterate_fused: no integrand set");
    CK(cudaSetDevice(c->device));
    EngineP p;
    int rc = fill_engine(c, p, itn, beta, flags, sigf_dev, sum_f_dev, n_f_dev, hstride, status_dev);
    if (rc) return rc;
    p.chunk_begin = 0; p.chunk_end = c->nchunks;
    p.chunk_off = nullptr; p.row0 = 0;
    return run_engine(c, p, c->nf, true, acc_dev, (cudaStream_t)stream);
}

extern "C" int vb200_reduce(vb200_ctx* c, uint32_t itn, double beta, int flags, int64_t chunk_begin, int64_t chunk_end,
                            const double* f_dev, int nf, const double* wgt_dev, double* sigf_dev, double* acc_dev,
                            double* sum_f_dev, uint64_t* n_f_dev, int64_t hstride, const uint16_t* bins_dev,
                            int32_t* status_dev, void* stream)
{
    if (!c || !acc_dev || !f_dev || !wgt_dev) return fail(-1, "vb200_reduce: null argument");
    if (nf < 1 || nf > 8) return fail(-4, "vb200_reduce: nf=%d not compiled in (1..8)", nf);
    CK(cudaSetDevice(c->device));
    EngineP p;
    int rc = fill_engine(c, p, itn, beta, flags, sigf_dev, sum_f_dev, n_f_dev, hstride, status_dev);
    if (rc) return rc;
    if (chunk_begin < 0 || chunk_end > c->nchunks || chunk_begin > chunk_end) return fail(-1, "vb200_reduce: bad chunk range");
    p.chunk_begin = chunk_begin; p.chunk_end = chunk_end;
    p.chunk_off = (const int64_t*)c->chunk_off.p;
    rc = vb_fetch_chunk_off(c, (cudaStream_t)stream);
    if (rc) return rc;
    p.row0 = c->chunk_off_host[(size_t)chunk_begin];
    p.batch_rows = c->chunk_off_host[(size_t)chunk_end] - p.row0;
    p.fbuf = f_dev; p.wbuf = wgt_dev; p.bins = bins_dev;
    return run_engine(c, p, nf, false, acc_dev, (cudaStream_t)stream);
}


// ---------------------------------------------------------------------------------------------
// One fused iteration in ONE call (the everyday sizes, where a handful of binding calls cost more than
// the kernels): zero the iteration buffer, run the engine, adapt the map on the device, launch the next
// iteration's allocation pre-pass, copy the small head of the buffer back and synchronise.
// buf_dev layout (8-byte words; the caller's, see Integrator.__call__):
//   fp64  [0, nacc)                 mean, covariance, sum_sigf
//         [nacc, nacc + nh)         sum_f [dim][hstride]
//         [nacc + nh, nacc + 2 nh)  (counts as fp64: sharded runs only)      then `tail` words
//   int64 [nf64, nf64 + nh)         n_f [dim][hstride]
//         [nf64 + nh]               NaN flag (low 32 bits)
//         [nf64 + nh + 1, + 7)      statistics of the pre-pass
// head_host receives words [0, nacc) and the 7 words from the NaN flag on (nacc + 7 doubles).
// ---------------------------------------------------------------------------------------------
// the head of the iteration buffer, gathered into pinned host memory the device writes directly (two small
// device-to-host copies into pageable memory cost ~20 us of a 0.1 ms iteration)
__global__ void k_head(const double* __restrict__ f, int nacc, const long long* __restrict__ tail, double* __restrict__ out)
{
    for (int i = threadIdx.x; i < nacc + 7; i += blockDim.x) out[i] = i < nacc ? f[i] : __longlong_as_double(tail[i - nacc]);
}

extern "C" int vb200_iteration_begin(vb200_ctx* c, uint32_t itn, double beta, int flags, double* sigf_dev, void* buf_dev,
                                     int64_t nacc, int64_t nh, int64_t hstride, int64_t nf64, int64_t nwords, double alpha_adapt,
                                     double plan_neval_scaled, int64_t plan_min, int64_t plan_max, int64_t plan_uniform,
                                     void* stream)
{
    if (!c || !buf_dev) return fail(-1, "vb200_iteration: null argument");
    if (nacc < 1 || nacc + 7 > VB_HEAD_WORDS) return fail(-1, "vb200_iteration: nacc=%lld outside 1..%d", (long long)nacc, VB_HEAD_WORDS - 7);
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    if (!c->head_host) {
        CK(cudaHostAlloc(&c->head_host, sizeof(double) * VB_HEAD_WORDS, cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer(&c->head_dev, c->head_host, 0));
    }
    c->head_pending = 0;
    double* f = (double*)buf_dev;
    int64_t* iw = (int64_t*)buf_dev + nf64;
    CK(cudaMemsetAsync(buf_dev, 0, sizeof(double) * (size_t)nwords, st));
    int rc = vb200_iterate_fused(c, itn, beta, flags, sigf_dev, f, f + nacc, (uint64_t*)iw, hstride, (int32_t*)(iw + nh), stream);
    if (rc) return rc;
    if (alpha_adapt > 0) {
        rc = vb200_map_adapt_device(c, f + nacc, (const uint64_t*)iw, nullptr, hstride, alpha_adapt, (const int32_t*)(iw + nh), stream);
        if (rc) return rc;
    }
    if (plan_neval_scaled > 0) {
        if (!sigf_dev) return fail(-1, "vb200_iteration: planning ahead needs sigf");
        rc = plan_launch(c, sigf_dev, 0.0, f + nacc - 1, plan_neval_scaled, plan_min, plan_max, plan_uniform, nullptr,
                         (long long*)(iw + nh + 1), st, true);          // (the statistics were zeroed with the buffer)
        if (rc) return rc;
    }
    k_head<<<1, 64, 0, st>>>(f, (int)nacc, (const long long*)(iw + nh), (double*)c->head_dev);
    c->launches += 1;
    CK(cudaGetLastError());
    c->head_pending = nacc + 7;
    return 0;
}

extern "C" int vb200_iteration_end(vb200_ctx* c, double* head_host, int64_t nhead, void* stream)
{
    if (!c || !head_host) return fail(-1, "vb200_iteration_end: null argument");
    if (c->head_pending == 0 || nhead != c->head_pending)
        return fail(-1, "vb200_iteration_end: %lld words asked, %lld in flight", (long long)nhead, (long long)c->head_pending);
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    memcpy(head_host, c->head_host, sizeof(double) * (size_t)nhead);
    c->head_pending = 0;
    return 0;
}

extern "C" int vb200_iteration(vb200_ctx* c, uint32_t itn, double beta, int flags, double* sigf_dev, void* buf_dev,
                               int64_t nacc, int64_t nh, int64_t hstride, int64_t nf64, int64_t nwords, double alpha_adapt,
                               double plan_neval_scaled, int64_t plan_min, int64_t plan_max, int64_t plan_uniform,
                               double* head_host, void* stream)
{
    if (!head_host) return fail(-1, "vb200_iteration: null argument");
    int rc = vb200_iteration_begin(c, itn, beta, flags, sigf_dev, buf_dev, nacc, nh, hstride, nf64, nwords, alpha_adapt,
                                   plan_neval_scaled, plan_min, plan_max, plan_uniform, stream);
    if (rc) return rc;
    return vb200_iteration_end(c, head_host, nacc + 7, stream);
}

extern "C" int64_t vb200_launch_count(vb200_ctx* c) { return c ? c->launches : 0; }

extern "C" int vb200_last_launch(vb200_ctx* c, int64_t out[6])
{
    if (!c || !out) return fail(-1, "vb200_last_launch: null argument");
    out[0] = c->last_grid; out[1] = c->last_bps; out[2] = c->last_smem; out[3] = c->last_wtot;
    out[4] = c->last_nt; out[5] = c->last_ch;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// FP64 FMA throughput probe (the roofline denominator for the fused kernel)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b)
{
    double v0 = threadIdx.x * 1e-9, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
            v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
        }
    }
    double s = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
    if (s == 123.456) out[0] = s;    // keep the chain alive
}

extern "C" int vb200_fp64_peak(int device, int iters, double* tflops_out, double* ms_out)
{
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    double* out;
    CK(cudaMalloc((void**)&out, 8));
    const int grid = prop.multiProcessorCount * 8, nt = 256;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k_fp64_peak<<<grid, nt>>>(out, iters / 8 + 1, 0.999999, 1e-9);     // warm-up
    CK(cudaEventRecord(e0));
    k_fp64_peak<<<grid, nt>>>(out, iters, 0.999999, 1e-9);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * 8 * 16 * (double)iters * (double)grid * nt;
    if (tflops_out) *tflops_out = flops / (ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    return 0;
}

