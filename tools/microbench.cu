// microbench.cu -- design probes for the histogram / RNG parts of the engine (not product code).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../vegas_b200/csrc/common.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// D global f64 REDs + D u64 REDs per sample on [D][1000] bins; window = bins spread per axis
__global__ void k_hist_global(double* sum, unsigned long long* cnt, int D, int nbin, int window, long long nsamp, int with_cnt)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x + 1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nsamp; i += (long long)gridDim.x * blockDim.x) {
        int base = (int)((i / 4096) % (nbin / window)) * window;
        for (int d = 0; d < D; ++d) {
            int b = (d < 2 ? lcg(s) % nbin : base + lcg(s) % window);
            atomicAdd(sum + d * nbin + b, 1.0000001);
            if (with_cnt) atomicAdd(cnt + d * nbin + b, 1ull);
        }
    }
}

// same, but into shared-memory bins (f64 CAS + u32 native), flushed at the end
__global__ void k_hist_smem(double* sum, unsigned long long* cnt, int D, int nbin, int window, long long nsamp, int with_cnt)
{
    extern __shared__ double sm[];
    double* ssum = sm;
    unsigned* scnt = (unsigned*)(sm + D * nbin);
    for (int i = threadIdx.x; i < D * nbin; i += blockDim.x) { ssum[i] = 0; scnt[i] = 0; }
    __syncthreads();
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x + 1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nsamp; i += (long long)gridDim.x * blockDim.x) {
        int base = (int)((i / 4096) % (nbin / window)) * window;
        for (int d = 0; d < D; ++d) {
            int b = (d < 2 ? lcg(s) % nbin : base + lcg(s) % window);
            atomicAdd(ssum + d * nbin + b, 1.0000001);
            if (with_cnt) atomicAdd(scnt + d * nbin + b, 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D * nbin; i += blockDim.x) {
        if (scnt[i] || ssum[i] != 0) { atomicAdd(sum + i, ssum[i]); atomicAdd(cnt + i, (unsigned long long)scnt[i]); }
    }
}

// warp-private smem bins without atomics: software conflict resolution through a tag array
__global__ void k_hist_tag(double* sum, unsigned long long* cnt, int D, int nbin, int window, long long nsamp)
{
    extern __shared__ double sm[];
    double* ssum = sm;
    unsigned* tag = (unsigned*)(sm + D * nbin);
    for (int i = threadIdx.x; i < D * nbin; i += blockDim.x) { ssum[i] = 0; tag[i] = 0; }
    __syncthreads();
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const unsigned me = threadIdx.x + 1;
    for (long long i0 = blockIdx.x * (long long)blockDim.x; i0 < nsamp; i0 += (long long)gridDim.x * blockDim.x) {
        long long i = i0 + threadIdx.x;
        int base = (int)((i / 4096) % (nbin / window)) * window;
        for (int d = 0; d < D; ++d) {
            int b = d * nbin + (d < 2 ? lcg(s) % nbin : base + lcg(s) % window);
            bool done = false;
            // CTA-wide software lock-free: claim with atomicCAS on a u32 tag (native), add, release
            while (!done) {
                if (atomicCAS(tag + b, 0u, me) == 0u) {
                    ssum[b] += 1.0000001;
                    __threadfence_block();
                    atomicExch(tag + b, 0u);
                    done = true;
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D * nbin; i += blockDim.x) if (ssum[i] != 0) atomicAdd(sum + i, ssum[i]);
}

__global__ void k_philox(double* out, PhiloxKey K, long long nsamp, int npair)
{
    double acc = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nsamp; i += (long long)gridDim.x * blockDim.x)
        for (int p = 0; p < npair; ++p) { double a, b; philox_pair(K, 3, i >> 3, (uint32_t)(i & 7), p, a, b); acc += a + b; }
    if (acc == 1.2345) out[0] = acc;
}

template <class F> float timeit(F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main()
{
    const int D = 8, nbin = 1000;
    const long long N = 1LL << 26;
    double* sum; unsigned long long* cnt; double* out;
    CK(cudaMalloc(&sum, D * nbin * 8)); CK(cudaMalloc(&cnt, D * nbin * 8)); CK(cudaMalloc(&out, 8));
    CK(cudaMemset(sum, 0, D * nbin * 8)); CK(cudaMemset(cnt, 0, D * nbin * 8));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int sms = pr.multiProcessorCount;
    printf("device %s  SMs %d\n", pr.name, sms);
    for (int window : {125, 1000}) {
        for (int wc : {0, 1}) {
            float ms = timeit([&] { k_hist_global<<<sms * 8, 256>>>(sum, cnt, D, nbin, window, N, wc); });
            printf("hist_global window=%4d cnt=%d : %8.3f ms  %.3e samples/s  (%.3e atomics/s)\n", window, wc, ms, N / (ms * 1e-3), N * (double)D * (1 + wc) / (ms * 1e-3));
        }
        size_t smem = (size_t)D * nbin * 12;
        CK(cudaFuncSetAttribute(k_hist_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int nt : {256, 512, 1024}) for (int wc : {0, 1}) {
            float ms = timeit([&] { k_hist_smem<<<sms * 2, nt, smem>>>(sum, cnt, D, nbin, window, N, wc); });
            printf("hist_smem   window=%4d cnt=%d nt=%4d: %8.3f ms  %.3e samples/s\n", window, wc, nt, ms, N / (ms * 1e-3));
        }
        CK(cudaFuncSetAttribute(k_hist_tag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            float ms = timeit([&] { k_hist_tag<<<sms * 2, 512, smem>>>(sum, cnt, D, nbin, window, N); });
            printf("hist_tag    window=%4d nt=512      : %8.3f ms  %.3e samples/s\n", window, ms, N / (ms * 1e-3));
        }
    }
    PhiloxKey K; philox_make_key(12345, K);
    {
        float ms = timeit([&] { k_philox<<<sms * 8, 256>>>(out, K, N, 4); });
        printf("philox 4 calls/sample: %8.3f ms  %.3e samples/s\n", ms, N / (ms * 1e-3));
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
