// atomics_bench.cu -- throughput of the shared-memory histogram update forms the engine can use
// (design probe for engine.cuh's training histogram; not product code).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/atomics_bench tools/atomics_bench.cu
// Workload per "sample": D = 8 axes, one bin per axis drawn uniformly from a window of W bins,
// count += 1 and sum += v on each.  Reported: ns per sample over the whole GPU and SM-cycles per
// sample per SM at the clock given on the command line (default 1965 MHz).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
#define D 8

extern __shared__ double sm[];

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

__device__ __forceinline__ void cas4(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, double v, uint32_t dummy)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p0, p1, p2, p3, t0, t1, t2, t3, q;\n\t"
        ".reg .b32 a0, a1, a2, a3;\n\t"
        ".reg .b64 o0, o1, o2, o3, s0, s1, s2, s3;\n\t"
        ".reg .f64 f0, f1, f2, f3;\n\t"
        "mov.b32 a0, %0; mov.b32 a1, %1; mov.b32 a2, %2; mov.b32 a3, %3;\n\t"
        "setp.ne.u32 p0, a0, 0xffffffff; setp.ne.u32 p1, a1, 0xffffffff; setp.ne.u32 p2, a2, 0xffffffff; setp.ne.u32 p3, a3, 0xffffffff;\n\t"
        "ld.shared.b64 o0, [a0]; ld.shared.b64 o1, [a1]; ld.shared.b64 o2, [a2]; ld.shared.b64 o3, [a3];\n"
        "L_CAS4:\n\t"
        "mov.b64 f0, o0; mov.b64 f1, o1; mov.b64 f2, o2; mov.b64 f3, o3;\n\t"
        "add.rn.f64 f0, f0, %4; add.rn.f64 f1, f1, %4; add.rn.f64 f2, f2, %4; add.rn.f64 f3, f3, %4;\n\t"
        "mov.b64 s0, f0; mov.b64 s1, f1; mov.b64 s2, f2; mov.b64 s3, f3;\n\t"
        "atom.shared.cas.b64 s0, [a0], o0, s0;\n\t"
        "atom.shared.cas.b64 s1, [a1], o1, s1;\n\t"
        "atom.shared.cas.b64 s2, [a2], o2, s2;\n\t"
        "atom.shared.cas.b64 s3, [a3], o3, s3;\n\t"
        "setp.ne.b64 t0, s0, o0; setp.ne.b64 t1, s1, o1; setp.ne.b64 t2, s2, o2; setp.ne.b64 t3, s3, o3;\n\t"
        "and.pred p0, p0, t0; and.pred p1, p1, t1; and.pred p2, p2, t2; and.pred p3, p3, t3;\n\t"
        "mov.b64 o0, s0; mov.b64 o1, s1; mov.b64 o2, s2; mov.b64 o3, s3;\n\t"
        "selp.b32 a0, a0, %5, p0; selp.b32 a1, a1, %5, p1; selp.b32 a2, a2, %5, p2; selp.b32 a3, a3, %5, p3;\n\t"
        "or.pred q, p0, p1; or.pred q, q, p2; or.pred q, q, p3;\n\t"
        "@q bra L_CAS4;\n\t"
        "}\n"
        :: "r"(a0), "r"(a1), "r"(a2), "r"(a3), "d"(v), "r"(dummy) : "memory");
}

__device__ __forceinline__ bool cas128(uint32_t sa, unsigned long long& o0, unsigned long long& o1, unsigned long long n0, unsigned long long n1)
{
    unsigned long long r0, r1;
    asm volatile("{\n\t.reg .b128 c, n, r;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 n, {%4, %5};\n\t"
                 "atom.shared.cas.b128 r, [%6], c, n;\n\tmov.b128 {%0, %1}, r;\n\t}"
                 : "=l"(r0), "=l"(r1) : "l"(o0), "l"(o1), "l"(n0), "l"(n1), "r"(sa) : "memory");
    bool ok = (r0 == o0) && (r1 == o1);
    o0 = r0; o1 = r1;
    return ok;
}

// MODE 0: rng only   1: u32 red only   2: f64 atomicAdd only (ATOMS.CAST.SPIN loop)   3: 2 + 1
//      4: lock-step-4 PTX CAS + u32 red (round-1 engine)   5: 128-bit CAS on {sum, count}
//      6: plain LDS/DADD/STS + LDS/IADD/STS (not atomic: the floor)   7: match_any leader + f64 atomicAdd + u32 red
//      8: f64 atomicAdd + u32 red, axes interleaved in pairs by hand-unrolling (2 independent spin loops back to back)
template <int MODE>
__global__ void __launch_bounds__(256) k_hist(double* out, int W, int iters)
{
    double* ssum = sm;                                   // [D][W]
    unsigned* scnt = (unsigned*)(sm + D * W);            // [D][W]
    for (int i = threadIdx.x; i < D * W * (MODE == 5 ? 2 : 1); i += blockDim.x) ssum[i] = 0;
    if (MODE != 5) for (int i = threadIdx.x; i < D * W; i += blockDim.x) scnt[i] = 0;
    __syncthreads();
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    const uint32_t sum_sa = (uint32_t)__cvta_generic_to_shared(ssum);
    double acc = 0;
    for (int it = 0; it < iters; ++it) {
        int b[D];
#pragma unroll
        for (int d = 0; d < D; ++d) b[d] = d * W + (int)(((unsigned long long)rng(s) * (unsigned)W) >> 32);
        const double v = 1.0 + (double)(s & 1023) * 1e-6;
        if (MODE == 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) acc += b[d];
        } else if (MODE == 1) {
#pragma unroll
            for (int d = 0; d < D; ++d) atomicAdd(scnt + b[d], 1u);
        } else if (MODE == 2) {
#pragma unroll
            for (int d = 0; d < D; ++d) atomicAdd(sm + b[d], v);
        } else if (MODE == 3 || MODE == 8) {
#pragma unroll
            for (int d = 0; d < D; ++d) { atomicAdd(scnt + b[d], 1u); atomicAdd(sm + b[d], v); }
        } else if (MODE == 4) {
#pragma unroll
            for (int d = 0; d < D; ++d) atomicAdd(scnt + b[d], 1u);
#pragma unroll
            for (int d = 0; d < D; d += 4) cas4(sum_sa + 8 * b[d], sum_sa + 8 * b[d + 1], sum_sa + 8 * b[d + 2], sum_sa + 8 * b[d + 3], v, sum_sa + 12 * D * W + 8 * (threadIdx.x & 31));
            __syncwarp();
        } else if (MODE == 5) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const uint32_t sa = sum_sa + 16 * b[d];
                unsigned long long o0, o1;
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(o0), "=l"(o1) : "r"(sa) : "memory");
                for (;;) {
                    unsigned long long n0 = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)o0) + v);
                    if (cas128(sa, o0, o1, n0, o1 + 1)) break;
                }
            }
            __syncwarp();
        } else if (MODE == 6) {
#pragma unroll
            for (int d = 0; d < D; ++d) { sm[b[d]] += v; scnt[b[d]] += 1u; }
        } else if (MODE == 10 || MODE == 11) {
            // staggered axes: at step s lane group g works on axis (s + g) & 7, so only 4 (MODE 10: 8 groups of 4 lanes)
            // or 16 (MODE 11: even / odd lanes in opposite order) lanes of a warp update the same axis' window at once
            // -- fewer same-bin collisions, fewer CAS retries; costs a select tree on the register-held bins
            const int g = MODE == 10 ? ((threadIdx.x >> 2) & 7) : ((threadIdx.x & 1) * 7);
#pragma unroll
            for (int st = 0; st < D; ++st) {
                const int a = MODE == 10 ? ((st + g) & 7) : (st ^ g);
                const int t0 = (a & 1) ? b[1] : b[0], t1 = (a & 1) ? b[3] : b[2], t2 = (a & 1) ? b[5] : b[4], t3 = (a & 1) ? b[7] : b[6];
                const int u0 = (a & 2) ? t1 : t0, u1 = (a & 2) ? t3 : t2;
                const int ba = (a & 4) ? u1 : u0;
                atomicAdd(scnt + ba, 1u);
                atomicAdd(sm + ba, v);
            }
            __syncwarp();
        } else if (MODE == 7) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const unsigned m = __match_any_sync(0xffffffffu, b[d]);
                const int leader = __ffs(m) - 1;
                const int cntm = __popc(m);
                if ((threadIdx.x & 31) == leader) { atomicAdd(scnt + b[d], (unsigned)cntm); atomicAdd(sm + b[d], v * cntm); }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    double t = acc;
    for (int i = threadIdx.x; i < D * W; i += blockDim.x) t += ssum[i] + (MODE == 5 ? 0.0 : (double)scnt[i]);
    if (t == 1.2345) out[0] = t;
}

// MODE 9 prototype: no fp64 atomics at all.  The last warp of the CTA OWNS the sums: sum word i belongs to
// its lane i % 32, and every update of it is executed by that lane with plain LDS / DADD / STS.  The
// other warps push {slot, value} into the owner lane's ring (slot reserved with a native u32 atomic,
// value stored first, then -- after one __threadfence_block per sample -- the slot word, which doubles
// as the "entry is valid" flag); counts stay native u32 reds.
template <int QCAP>
__global__ void __launch_bounds__(256) k_hist_owner(double* out, int W, int iters)
{
    double* ssum = sm;                                   // [D][W]
    unsigned* scnt = (unsigned*)(sm + D * W);            // [D][W]
    __shared__ uint4 q[32][QCAP];                        // {value lo, value hi, slot, sequence number}: one 16-byte store per push
    __shared__ unsigned qtail[32];
    __shared__ unsigned done_s;
    for (int i = threadIdx.x; i < D * W; i += blockDim.x) { ssum[i] = 0; scnt[i] = 0; }
    for (int i = threadIdx.x; i < 32 * QCAP; i += blockDim.x) (&q[0][0])[i] = make_uint4(0, 0, 0, (unsigned)(i % QCAP));   // cell sequence numbers (bounded MPSC queue after Vyukov)
    if (threadIdx.x < 32) qtail[threadIdx.x] = 0u;
    if (threadIdx.x == 0) done_s = 0u;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    double acc = 0;
    if (warp < nwarp - 1) {
        uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
        // the 7 producer warps do the work of 8: same samples per CTA as the other modes
        const int my_iters = (iters * nwarp + (nwarp - 2 - warp)) / (nwarp - 1);
        for (int it = 0; it < my_iters; ++it) {
            int b[D];
#pragma unroll
            for (int d = 0; d < D; ++d) b[d] = d * W + (int)(((unsigned long long)rng(s) * (unsigned)W) >> 32);
            const double v = 1.0 + (double)(s & 1023) * 1e-6;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                atomicAdd(scnt + b[d], 1u);
                const int o = b[d] & 31;
                const unsigned tk = atomicAdd(&qtail[o], 1u);             // ticket: cell tk % QCAP, in lap tk / QCAP
                const unsigned pos = tk % QCAP;
                const unsigned ea = (unsigned)__cvta_generic_to_shared(&q[o][pos]);
                // wait-and-store as ONE PTX loop with the store predicated inside it: a lane writes as soon as its
                // own slot is free.  (Written in C++ the compiler moves the store behind the loop's reconvergence
                // point, so the lanes of a warp hold their entries back until the slowest lane's slot frees -- and
                // the owner, who consumes in order, can be waiting for exactly those: deadlock, observed.)
                asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 z;\n"
                             "QPUSH:\n\tld.volatile.shared.u32 z, [%0+12];\n\tsetp.eq.u32 p, z, %5;\n\t"
                             "@p st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};\n\t@!p bra QPUSH;\n\t}"
                             :: "r"(ea), "r"((unsigned)__double2loint(v)), "r"((unsigned)__double2hiint(v)), "r"((unsigned)b[d]), "r"(tk + 1u), "r"(tk)
                             : "memory");
            }
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) atomicAdd(&done_s, 1u);
    } else {
        unsigned head = 0;
        bool last = false;
        for (;;) {
            const bool fin = *(volatile unsigned*)&done_s == (unsigned)(nwarp - 1);
            for (;;) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(&q[lane][head % QCAP]);
                uint4 ent;
                asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ent.x), "=r"(ent.y), "=r"(ent.z), "=r"(ent.w) : "r"(sa) : "memory");
                if (ent.w != head + 1u) break;                             // the cell's ticket has not been filled yet
                asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(sa + 12u), "r"(head + (unsigned)QCAP) : "memory");   // free it for the next lap
                ssum[ent.z] += __hiloint2double((int)ent.y, (int)ent.x);
                ++head;
            }
            if (__all_sync(0xffffffffu, last)) break;
            last = fin;
        }
    }
    __syncthreads();
    double t = acc;
    for (int i = threadIdx.x; i < D * W; i += blockDim.x) t += ssum[i] + (double)scnt[i];
    double cs = 0;                                       // every add is ~1: sums and counts must agree
    for (int i = threadIdx.x; i < D * W; i += blockDim.x) cs += ssum[i] - (double)scnt[i];
    if (t == 1.2345) out[0] = t;
    if (fabs(cs) > 0.01 * D * W) atomicAdd(out + 1, 1.0);    // updates were lost
}

template <int QCAP>
static int run_owner(int sms, int W, int nt, int bps, double mhz, double* out)
{
    auto kern = k_hist_owner<QCAP>;
    const int iters = 2000;
    size_t smem = (size_t)D * W * 12 + 512;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, smem));
    if (occ < bps) { printf("%-28s W=%4d nt=%3d x%d: does not fit (occ %d)\n", "owner-warp queues", W, nt, bps, occ); return 0; }
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    CK(cudaMemset(out, 0, 16));
    kern<<<sms * bps, nt, smem>>>(out, W, iters / 10);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    kern<<<sms * bps, nt, smem>>>(out, W, iters);
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double bad[2];
    CK(cudaMemcpy(bad, out, 16, cudaMemcpyDeviceToHost));
    const double nsamp = (double)sms * bps * nt * iters;
    printf("%-28s W=%4d nt=%3d x%d: %8.3f ms  %7.2f ps/sample  %6.2f cyc/sample/SM   (threads flagging lost updates: %.0f)\n",
           QCAP == 16 ? "owner-warp queues cap 16" : (QCAP == 32 ? "owner-warp queues cap 32" : "owner-warp queues cap 64"), W, nt, bps, ms, ms * 1e9 / nsamp, ms * 1e-3 * mhz * 1e6 * sms / nsamp, bad[1]);
    return 0;
}

template <int MODE>
static int run(const char* name, int sms, int W, int nt, int bps, double mhz, double* out)
{
    const int iters = 2000;
    size_t smem = (size_t)D * W * (MODE == 5 ? 16 : 12) + 512;
    CK(cudaFuncSetAttribute(k_hist<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_hist<MODE>, nt, smem));
    if (occ < bps) { printf("%-28s W=%4d nt=%3d x%d: does not fit (occ %d)\n", name, W, nt, bps, occ); return 0; }
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_hist<MODE><<<sms * bps, nt, smem>>>(out, W, iters / 10);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    k_hist<MODE><<<sms * bps, nt, smem>>>(out, W, iters);
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double nsamp = (double)sms * bps * nt * iters;
    printf("%-28s W=%4d nt=%3d x%d: %8.3f ms  %7.2f ps/sample  %6.2f cyc/sample/SM\n", name, W, nt, bps, ms,
           ms * 1e9 / nsamp, ms * 1e-3 * mhz * 1e6 * sms / nsamp);
    return 0;
}

int main(int argc, char** argv)
{
    setvbuf(stdout, nullptr, _IONBF, 0);
    const double mhz = argc > 1 ? atof(argv[1]) : 1965.0;
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    printf("device %s  SMs %d  (cycles at %.0f MHz)\n", pr.name, sms, mhz);
    double* out;
    CK(cudaMalloc(&out, 16));
    const bool only_owner = argc > 2;
    for (int W : {125, 1000}) {
        for (int cfg = 0; cfg < 3; ++cfg) {
            const int nt = cfg == 2 ? 128 : 256, bps = cfg == 0 ? 2 : (cfg == 1 ? 3 : 4);
            if (W == 1000 && cfg != 0) continue;
            if (nt == 256 && (run_owner<16>(sms, W, nt, bps, mhz, out) || run_owner<32>(sms, W, nt, bps, mhz, out) || run_owner<64>(sms, W, nt, bps, mhz, out))) return 1;
            if (only_owner) continue;
            if (run<0>("rng only", sms, W, nt, bps, mhz, out)) return 1;
            if (run<1>("u32 red", sms, W, nt, bps, mhz, out)) return 1;
            if (run<2>("f64 atomicAdd (CAST.SPIN)", sms, W, nt, bps, mhz, out)) return 1;
            if (run<3>("f64 atomicAdd + u32 red", sms, W, nt, bps, mhz, out)) return 1;
            if (run<4>("lock-step-4 CAS + u32 red", sms, W, nt, bps, mhz, out)) return 1;
            if (run<5>("CAS.128 {sum,count}", sms, W, nt, bps, mhz, out)) return 1;
            if (run<6>("plain RMW (not atomic)", sms, W, nt, bps, mhz, out)) return 1;
            if (run<7>("match_any leader + atomics", sms, W, nt, bps, mhz, out)) return 1;
            if (run<10>("staggered axes, 8 x 4 lanes", sms, W, nt, bps, mhz, out)) return 1;
            if (run<11>("staggered axes, 2 x 16 lanes", sms, W, nt, bps, mhz, out)) return 1;
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
