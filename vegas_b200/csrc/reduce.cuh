// reduce.cuh -- k_reduce<NF>: the per-hypercube reduce of the callback path (vb200_reduce; reference
// Integrator.__call__, _vegas.pyx:2136-2197) as an HBM-streaming kernel.
//
// Inputs are the batch buffers in HBM: f[rows][NF] (the integrand's values, from the user's device
// callback), wgt[rows] and the training bins[rows][dim] (uint16) the sampler wrote.  Algorithmic
// traffic: 8 NF + 8 + 2 dim bytes per row, read once.
//
// Round 1 reduced these buffers with the fused engine's kernel and a source that loaded each row
// with per-thread global loads.  Here the rows arrive by TMA:
//   * persistent CTAs claim work items (chunks of 256 hypercubes, or parts of chunks the vegas+
//     allocation piled samples onto) as k_engine does -- one item AHEAD, so the claim's round trip to
//     L2, the chunk search and the chunk's row offset hide behind the previous item;
//   * an item's rows are cut into TILES of whole hypercubes (cubes binned by their first row in
//     buckets of 2/3 CAP rows; ballot + popc ranks, no serial walk); each tile is fetched with three
//     1-D bulk copies (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes: f, wgt, bins)
//     into one of two shared-memory stages, completion signalled on the stage's mbarrier; the copy of
//     tile t+1 is in flight while tile t is reduced;
//   * phase A (thread per row): w*f in place, NaN check, training-histogram adds from the bins;
//     phase B: the reference's two-pass mean / variance per hypercube from the staged tile (thread per
//     cube up to VB_RWARP_CUBE rows, a warp per larger cube, claimed from a counter);
//   * hypercubes larger than a tile are streamed from HBM twice by the whole CTA (second pass hits L2).
// The accumulators, the histogram windows and the sigf update are the engine's (engine.cuh).
//
// What the measurements say (B200, one 8.4M-row batch): with one output the kernel
// takes 0.77 ms with training, 0.25 ms without -- the dim shared-memory fp64 adds per row (a CAS loop each:
// there is no native shared fp64 add) are two thirds of it, the same cost the fused kernel pays.  Two
// restructurings were measured and dropped: strips of consecutive rows per thread with per-cube sums
// through shared atomics (balanced, but 4 extra CAS adds per row: 0.83 ms), and warp-autonomous pipelines
// with one pair of stages per warp (no CTA barriers inside an item, but half-empty warps: 0.91 ms).
#pragma once
#include "engine.cuh"

// cubes with more rows than this are reduced by a whole warp
#ifndef VB_RWARP_CUBE
#define VB_RWARP_CUBE 64     // measured on B200 (8-D, one output, vegas+ allocation 2..1271 rows per cube): 24 -> 0.85 ms, 40 -> 0.81, 64 -> 0.77 per 8.4M rows
#endif

template <int NF>
struct ReduceGeom {
    static constexpr int NT = 256;
    static constexpr int CH = VB_CH;
    static constexpr int MINB = NF <= 4 ? 2 : 1;                           // resident CTAs per SM the register budget is set for
    static constexpr int MAXT = VB_CH + 8;                                 // tiles per item (worst case: one cube each)
    // rows per tile (p.cap, chosen by the launcher): a stage holds cap + 16 rows (alignment slack on both sides)
};

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+; SASS: SYNCS.*, UBLKCP) ------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// 1-D bulk copy global -> shared; src, dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, unsigned bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// dynamic shared memory of k_reduce<NF> (launcher and kernel use the same function)
__host__ inline size_t reduce_layout(EngineP& p, int nf, int capa, int ch, int dim, int maxt, bool stage_bins)
{
    const int wtot = p.wtot;
    // one stage: f [capa][nf] fp64 | wgt [capa] fp64 | bins [capa][dim] u16
    p.st_w = (int)(sizeof(double) * (size_t)capa * nf);
    p.st_b = p.st_w + (int)(sizeof(double) * (size_t)capa);
    p.st_bytes = (p.st_b + (stage_bins ? 2 * capa * dim : 0) + 15) & ~15;
    size_t b = 2 * (size_t)p.st_bytes;
    p.o_ex = (int)b;    b += sizeof(long long) * (size_t)(ch + 1);
    p.o_hsum = (int)b;  b += sizeof(double) * (size_t)wtot;
    p.o_n = (int)b;     b += sizeof(int) * (size_t)ch;
    p.o_hcnt = (int)b;  b += sizeof(unsigned) * (size_t)wtot;
    p.o_y0 = (int)b;    b += sizeof(int) * (size_t)(maxt + 1);     // (here: the tile list)
    p.o_gw = p.o_dvn = 0;
    for (int d = 0; d < VB_MAXD; ++d) {
        p.hs_idx[d] = p.o_hsum / 8 + p.woff[d];
        p.hc_idx[d] = p.o_hcnt / 4 + p.woff[d];
        p.gw_idx[d] = 0;
    }
    p.win_all = 0;
    return (b + 15) & ~(size_t)15;
}

template <int NF>
__global__ void __launch_bounds__(256, ReduceGeom<NF>::MINB) k_reduce(const __grid_constant__ EngineP p)
{
    typedef ReduceGeom<NF> G;
    constexpr int NT = G::NT, CH = G::CH, NW = NT / 32, NV = NF * (NF + 1) / 2;
    static_assert(CH == NT, "one cube per thread in the set-up and the tile list");
    const int CAP = p.cap;
    char* const smem_b = (char*)vb_smem;
    long long* ex_s = (long long*)(smem_b + p.o_ex);
    HistW H;
    H.sum = (double*)(smem_b + p.o_hsum);
    int* n_s = (int*)(smem_b + p.o_n);
    H.cnt = (unsigned*)(smem_b + p.o_hcnt);
    int* tl_s = (int*)(smem_b + p.o_y0);                          // tile t = cubes [tl_s[t], tl_s[t+1])
    __shared__ __align__(8) unsigned long long mbar_s[2];
    __shared__ long long scan_s[NW];
    __shared__ double red_s[NW];
    __shared__ uint32_t base_s[VB_MAXD];
    __shared__ long long next_s, row_s;
    __shared__ int wcnt_s[NW];
    __shared__ int sub_s[2], ntile_s, lclaim_s;
    __shared__ int nlarge_s, large_s[VB_LARGE_MAX];
    __shared__ double p1_s[NW * NF], p2_s[NW * (NF + NV)];              // block reduction of a giant cube's sums
    __shared__ int wnew_s[VB_MAXD], wneed_s[VB_MAXD];
    __shared__ uint32_t sb_s[NT + 1];
    __shared__ int pc_s[NT + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dim = p.map.dim;
    const bool correlate = (p.flags & VBF_CORRELATE) != 0;
    const bool train = (p.flags & VBF_TRAIN) != 0 && p.bins != nullptr;
    const uint32_t smem_sa = (uint32_t)__cvta_generic_to_shared(vb_smem);
    const uint32_t bar_sa = (uint32_t)__cvta_generic_to_shared(mbar_s);
    CubeAcc<NF> A;
    A.clear();
    for (int i = tid; i < p.wtot; i += NT) { H.sum[i] = 0.0; H.cnt[i] = 0u; }
    if (tid < VB_MAXD) vb_wlo_s[tid] = -0x40000000;
    if (tid == 0) {
        mbar_init(bar_sa, 1);
        mbar_init(bar_sa + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned phbits = 0u;                                         // bit s: parity stage s's barrier completes next
    long long since_flush = 0;
    const int64_t rows_bulk = p.batch_rows & ~(int64_t)7;         // rows [0, rows_bulk) can be fetched in 16-byte units

    // thread 0: fetch rows [r0, r1) of the batch into stage s; returns through *ra_out the row the stage starts at.
    // The aligned part goes by bulk copies, a ragged end of the batch (< 8 rows) by plain loads of the whole CTA.
    auto issue = [&](int s, int64_t r0, int64_t r1) {
        const int64_t ra = r0 & ~(int64_t)7;
        int64_t rb = (r1 + 7) & ~(int64_t)7;
        if (rb > rows_bulk) rb = rows_bulk;
        if (rb > ra && tid == 0) {
            const unsigned nrow = (unsigned)(rb - ra);
            const uint32_t st = smem_sa + (uint32_t)s * (uint32_t)p.st_bytes, bar = bar_sa + 8u * (uint32_t)s;
            const unsigned bf = nrow * 8u * NF, bw = nrow * 8u, bb = train ? nrow * 2u * (unsigned)dim : 0u;
            fence_proxy_async();                                  // the stage was last written / read through the generic proxy
            mbar_expect_tx(bar, bf + bw + bb);
            bulk_g2s(st, p.fbuf + ra * NF, bf, bar);
            bulk_g2s(st + (uint32_t)p.st_w, p.wbuf + ra, bw, bar);
            if (bb) bulk_g2s(st + (uint32_t)p.st_b, p.bins + ra * dim, bb, bar);
        }
    };
    // whole CTA: wait for stage s (rows [r0, r1)) and copy what the bulk copies could not fetch
    auto arrive = [&](int s, int64_t r0, int64_t r1) {
        const int64_t ra = r0 & ~(int64_t)7;
        int64_t rb = (r1 + 7) & ~(int64_t)7;
        if (rb > rows_bulk) rb = rows_bulk;
        char* st = smem_b + (size_t)s * p.st_bytes;
        if (rb < r1) {                                            // ragged end of the batch
            const int64_t t0 = rb > ra ? rb : ra;
            double* fs = (double*)st; double* ws = (double*)(st + p.st_w); uint16_t* bs = (uint16_t*)(st + p.st_b);
            for (int64_t r = t0 + tid; r < r1; r += NT) {
                for (int q = 0; q < NF; ++q) fs[(r - ra) * NF + q] = p.fbuf[r * NF + q];
                ws[r - ra] = p.wbuf[r];
                if (train) for (int d = 0; d < dim; ++d) bs[(r - ra) * dim + d] = p.bins[r * dim + d];
            }
        }
        if (rb > ra) { mbar_wait(bar_sa + 8u * (uint32_t)s, (phbits >> s) & 1u); phbits ^= 1u << s; }
        __syncthreads();
    };

    // Work items are claimed one ahead: thread 0 asks for item k+1 when item k starts and looks at the
    // answer (the atomic's round trip to L2, the chunk search, the chunk's row offset) when item k is done.
    const long long extra_items = p.item_off ? (p.item_end - p.item_begin) - (p.chunk_end - p.chunk_begin) : 0;
    auto resolve = [&](long long g) {                             // thread 0: which chunk is item g, where do its rows start?
        if (g >= p.item_end) { next_s = p.chunk_end; return; }
        long long c; int sub = 0, nsub = 1;
        if (p.item_off == nullptr) c = g;
        else {
            // chunk j holds at least one item, so chunk(g) <= chunk_begin + (g - item_begin), and it is at
            // most `extra_items` (the number of additional parts of split chunks) below that
            long long hi = p.chunk_begin + (g - p.item_begin) + 1;
            if (hi > p.chunk_end) hi = p.chunk_end;
            long long lo = hi - 1 - extra_items;
            if (lo < p.chunk_begin) lo = p.chunk_begin;             // item_off[lo] <= g < item_off[hi]
            while (hi - lo > 1) {
                const long long mid = (lo + hi) >> 1;
                if (p.item_off[mid] <= g) lo = mid; else hi = mid;
            }
            c = lo;
            sub = (int)(g - p.item_off[lo]);
            nsub = (int)(p.item_off[lo + 1] - p.item_off[lo]);
        }
        next_s = c; sub_s[0] = sub; sub_s[1] = nsub;
        row_s = p.chunk_off[c] - p.row0;
    };
    if (tid == 0) resolve(p.item_begin + (long long)atomicAdd(p.work_counter, 1ull));
    for (;;) {
        __syncthreads();                       // previous item fully consumed; barriers initialised; next_s written
        const int64_t lc = next_s;
        if (lc >= p.chunk_end) break;
        long long g_ahead = 0;
        if (tid == 0) g_ahead = p.item_begin + (long long)atomicAdd(p.work_counter, 1ull);
        const int64_t chunk_row = row_s;
        const int sub = sub_s[0], nsub = sub_s[1];
        const int64_t lh0 = lc * CH;
        const int64_t h0 = local_to_global(p.st, lh0);
        if (p.wtot > 0) {
            const bool force = since_flush > 0x40000000LL;
            hist_move_windows<NT>(p, H, h0, CH, force, wnew_s, wneed_s);
            if (force) since_flush = 0;
        }
        const long long total = chunk_setup<NT, CH, uint32_t>(p, lh0, h0, ex_s, n_s, nullptr, base_s, scan_s);
        int c0, cend;
        item_cubes(ex_s, CH, total, sub, nsub, c0, cend);
        since_flush += ex_s[cend] - ex_s[c0];
        // ---- tile list: a tile is a run of whole cubes staged together.  Cubes are binned by the row
        // they start at (buckets of Q = 2/3 CAP rows, a power of two); a cube of more than CAP/3 rows is a
        // tile of its own (above CAP rows: a giant, streamed from HBM).  A tile then holds < Q + CAP/3 =
        // CAP rows, and every thread can tell on its own whether its cube starts one (no serial walk).
        {
            const int c = tid;
            const int bigthr = CAP / 3, qshift = 31 - __clz(2 * bigthr);
            int flag = 0;
            if (c >= c0 && c < cend && n_s[c] > 0) {
                const long long b0 = ex_s[c0];
                if (c == c0 || n_s[c] > bigthr || n_s[c - 1] > bigthr) flag = 1;
                else flag = (((ex_s[c] - b0) >> qshift) != ((ex_s[c - 1] - b0) >> qshift)) ? 1 : 0;
            }
            // ranks of the flags: ballot + popc inside a warp, the warps' counts through shared memory
            const unsigned bal = __ballot_sync(0xffffffffu, flag);
            if (lane == 0) wcnt_s[warp] = __popc(bal);
            __syncthreads();
            int rk = __popc(bal & ((1u << lane) - 1u)), ntl = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) { const int v = wcnt_s[w]; if (w < warp) rk += v; ntl += v; }
            if (flag) tl_s[rk] = c;
            if (tid == 0) {
                int e = cend;                                      // one past the last cube with samples
                while (e > c0 && n_s[e - 1] == 0) --e;
                tl_s[ntl] = e;
                ntile_s = ntl;
            }
        }
        __syncthreads();
        const int ntile = ntile_s;
        auto tile_rows = [&](int t, int64_t& r0, int64_t& r1, bool& giant) {
            const int a = tl_s[t], b = tl_s[t + 1];
            r0 = chunk_row + ex_s[a];
            r1 = chunk_row + ex_s[b];
            giant = r1 - r0 > CAP;
        };
        if (ntile > 0) {
            int64_t r0, r1; bool giant;
            tile_rows(0, r0, r1, giant);
            if (!giant) issue(0, r0, r1);
        }
        for (int t = 0; t < ntile; ++t) {
            const int s = t & 1;
            int64_t r0, r1; bool giant;
            tile_rows(t, r0, r1, giant);
            if (t + 1 < ntile) {                                   // prefetch the next tile into the other stage (free since the barrier that ended tile t-1)
                int64_t q0, q1; bool g2;
                tile_rows(t + 1, q0, q1, g2);
                if (!g2) issue(s ^ 1, q0, q1);
            }
            const int ca = tl_s[t], cb = tl_s[t + 1];
            if (giant) {
                // ---- one cube of more than CAP rows: streamed from HBM twice by the whole CTA
                const int n = n_s[ca];
                double S[NF], m[NF], sd[NF], q[NV];
#pragma unroll
                for (int u = 0; u < NF; ++u) { S[u] = 0.0; sd[u] = 0.0; }
#pragma unroll
                for (int v = 0; v < NV; ++v) q[v] = 0.0;
                bool bad = false;
                for (int64_t r = r0 + tid; r < r1; r += NT) {
                    const double w = p.wbuf[r];
                    double wf0 = 0.0;
#pragma unroll
                    for (int u = 0; u < NF; ++u) {
                        const double fx = p.fbuf[r * NF + u];
                        bad |= isnan(fx);
                        S[u] += w * fx;
                        if (u == 0) wf0 = w * fx;
                    }
                    if (train) {
                        const double a = wf0 * (double)n, fdv2 = a * a;
                        for (int d = 0; d < dim; ++d) {
                            const unsigned bv = p.bins[r * dim + d];
                            if (bv != 0xffffu) hist_add_code(p, d, hist_code(p, d, (int)bv), fdv2);
                        }
                    }
                }
                if (bad) p.status[0] = 1;
#pragma unroll
                for (int u = 0; u < NF; ++u) { const double tt = warp_sum(S[u]); if (lane == 0) p1_s[warp * NF + u] = tt; }
                __syncthreads();
#pragma unroll
                for (int u = 0; u < NF; ++u) {
                    double tt = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) tt += p1_s[w * NF + u];
                    S[u] = tt;
                    m[u] = tt / (double)n;
                }
                for (int64_t r = r0 + tid; r < r1; r += NT) {
                    const double w = p.wbuf[r];
                    double wv[NF];
#pragma unroll
                    for (int u = 0; u < NF; ++u) wv[u] = w * p.fbuf[r * NF + u];
                    pass2_sample<NF>(wv, m, correlate, sd, q);
                }
#pragma unroll
                for (int u = 0; u < NF; ++u) { const double tt = warp_sum(sd[u]); if (lane == 0) p2_s[warp * (NF + NV) + u] = tt; }
#pragma unroll
                for (int v = 0; v < NV; ++v) { const double tt = warp_sum(q[v]); if (lane == 0) p2_s[warp * (NF + NV) + NF + v] = tt; }
                __syncthreads();
                if (tid == 0) {
#pragma unroll
                    for (int u = 0; u < NF; ++u) sd[u] = 0.0;
#pragma unroll
                    for (int v = 0; v < NV; ++v) q[v] = 0.0;
                    for (int w = 0; w < NW; ++w) {
#pragma unroll
                        for (int u = 0; u < NF; ++u) sd[u] += p2_s[w * (NF + NV) + u];
#pragma unroll
                        for (int v = 0; v < NV; ++v) q[v] += p2_s[w * (NF + NV) + NF + v];
                    }
                    const double sigf2 = cube_finish<NF>(A, n, S, sd, q, correlate);
                    cube_epilogue<NF, uint32_t>(p, H, A, sigf2, lh0 + ca, h0 + ca, n, nullptr, r1 - 1);
                }
                __syncthreads();
                continue;
            }
            // ---- a tile of whole cubes, staged in shared memory
            const int Tt = (int)(r1 - r0);
            const int off = (int)(r0 - (r0 & ~(int64_t)7));       // row r0 sits at stage row `off`
            const long long base = ex_s[ca];
            // start bits of the tile's cubes (as in k_engine) while the copy lands
            const int nword = (Tt + 31) >> 5;
            if (tid <= nword) sb_s[tid] = 0u;
            if (tid == 0) { nlarge_s = 0; lclaim_s = 0; }
            __syncthreads();
            for (int c = ca + tid; c < cb; c += NT)
                if (n_s[c] > 0) {
                    const int o = (int)(ex_s[c] - base);
                    atomicOr(&sb_s[o >> 5], 1u << (o & 31));
                }
            __syncthreads();
            {
                long long tot_unused;
                const long long ex = block_exscan<NT>((long long)(tid < nword ? __popc(sb_s[tid]) : 0), scan_s, &tot_unused);
                if (tid < nword) pc_s[tid] = (int)ex;
            }
            arrive(s, r0, r1);                                     // (ends with a barrier)
            double* const fs = (double*)(smem_b + (size_t)s * p.st_bytes) + (size_t)off * NF;
            const double* const ws = (const double*)(smem_b + (size_t)s * p.st_bytes + p.st_w) + off;
            const uint16_t* const bs = (const uint16_t*)(smem_b + (size_t)s * p.st_bytes + p.st_b) + (size_t)off * dim;
            // ---- phase A: one thread per row
            {
                bool bad = false;
                for (int i = tid; i < Tt; i += NT) {
                    const double w = ws[i];
                    double wf0 = 0.0;
#pragma unroll
                    for (int u = 0; u < NF; ++u) {
                        const double fx = fs[(size_t)i * NF + u];
                        bad |= isnan(fx);
                        const double wf = w * fx;
                        fs[(size_t)i * NF + u] = wf;
                        if (u == 0) wf0 = wf;
                    }
                    if (train) {
                        const uint32_t mk = sb_s[i >> 5] & (0xffffffffu >> (31 - (i & 31)));
                        const int c = ca + pc_s[i >> 5] + __popc(mk) - 1;
                        const double a = wf0 * (double)n_s[c], fdv2 = a * a;
                        for (int d = 0; d < dim; ++d) {
                            const unsigned bv = bs[(size_t)i * dim + d];
                            if (bv != 0xffffu) hist_add_code(p, d, hist_code(p, d, (int)bv), fdv2);
                        }
                    }
                }
                if (bad) p.status[0] = 1;
            }
            __syncthreads();
            // ---- phase B: small cubes, one thread each, in the reference's order (pyx:2142-2186)
            for (int c = ca + tid; c < cb; c += NT) {
                const int n = n_s[c];
                if (n > 0 && n <= VB_RWARP_CUBE) {
                    const double* wfp = fs + (size_t)(ex_s[c] - base) * NF;
                    double S[NF], m[NF], sd[NF], q[NV];
#pragma unroll
                    for (int u = 0; u < NF; ++u) { S[u] = 0.0; sd[u] = 0.0; }
#pragma unroll
                    for (int v = 0; v < NV; ++v) q[v] = 0.0;
                    for (int k = 0; k < n; ++k)
#pragma unroll
                        for (int u = 0; u < NF; ++u) S[u] += wfp[(size_t)k * NF + u];
#pragma unroll
                    for (int u = 0; u < NF; ++u) m[u] = S[u] / (double)n;
                    for (int k = 0; k < n; ++k) {
                        double w[NF];
#pragma unroll
                        for (int u = 0; u < NF; ++u) w[u] = wfp[(size_t)k * NF + u];
                        pass2_sample<NF>(w, m, correlate, sd, q);
                    }
                    const double sigf2 = cube_finish<NF>(A, n, S, sd, q, correlate);
                    cube_epilogue<NF, uint32_t>(p, H, A, sigf2, lh0 + c, h0 + c, n, nullptr, chunk_row + ex_s[c] + n - 1);
                } else if (n > VB_RWARP_CUBE) {
                    large_s[atomicAdd(&nlarge_s, 1)] = c;          // at most CAP / (VB_WARP_CUBE + 1) <= VB_LARGE_MAX per tile
                }
            }
            __syncthreads();
            // ---- larger cubes: one warp each
            const int nlarge = nlarge_s;                           // (claimed from a counter: their sizes differ)
            for (;;) {
                int j = 0;
                if (lane == 0) j = atomicAdd(&lclaim_s, 1);
                j = __shfl_sync(0xffffffffu, j, 0);
                if (j >= nlarge) break;
                const int c = large_s[j], n = n_s[c];
                const double* wfp = fs + (size_t)(ex_s[c] - base) * NF;
                double S[NF], m[NF], sd[NF], q[NV];
#pragma unroll
                for (int u = 0; u < NF; ++u) { S[u] = 0.0; sd[u] = 0.0; }
#pragma unroll
                for (int v = 0; v < NV; ++v) q[v] = 0.0;
                for (int k = lane; k < n; k += 32)
#pragma unroll
                    for (int u = 0; u < NF; ++u) S[u] += wfp[(size_t)k * NF + u];
#pragma unroll
                for (int u = 0; u < NF; ++u) { S[u] = warp_sum(S[u]); m[u] = S[u] / (double)n; }
                for (int k = lane; k < n; k += 32) {
                    double w[NF];
#pragma unroll
                    for (int u = 0; u < NF; ++u) w[u] = wfp[(size_t)k * NF + u];
                    pass2_sample<NF>(w, m, correlate, sd, q);
                }
#pragma unroll
                for (int u = 0; u < NF; ++u) sd[u] = warp_sum(sd[u]);
#pragma unroll
                for (int v = 0; v < NV; ++v) q[v] = warp_sum(q[v]);
                if (lane == 0) {
                    const double sigf2 = cube_finish<NF>(A, n, S, sd, q, correlate);
                    cube_epilogue<NF, uint32_t>(p, H, A, sigf2, lh0 + c, h0 + c, n, nullptr, chunk_row + ex_s[c] + n - 1);
                }
            }
            __syncthreads();                                       // the stage may be refilled now
        }
        if (tid == 0) resolve(g_ahead);                            // (every thread read next_s / sub_s / row_s of this item long ago)
    }

    if (p.wtot > 0) hist_flush<NT>(p, H, nullptr);
    constexpr int NACC = NF + NV + 1;
    double* out = p.partials + (size_t)blockIdx.x * NACC;
#pragma unroll
    for (int u = 0; u < NF; ++u) { double t = block_sum<NT>(A.mean[u], red_s); if (tid == 0) out[u] = t; }
#pragma unroll
    for (int v = 0; v < NV; ++v) { double t = block_sum<NT>(A.var[v], red_s); if (tid == 0) out[NF + v] = t; }
    { double t = block_sum<NT>(A.sum_sigf, red_s); if (tid == 0) out[NF + NV] = t; }
}
