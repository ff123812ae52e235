// fp64_ilp.cu -- DFMA throughput vs (independent chains per thread) x (warps per SM): what does
// it take to saturate the B200 FP64 pipe?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double* out, int iters, double a, double b)
{
    double v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) v[j] = threadIdx.x * 1e-9 + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) v[j] = fma(v[j], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += v[j];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
void run(int sms, double* out)
{
    for (int wps : {4, 8, 12, 16, 24, 32, 64}) {          // warps per SM
        int nt = wps >= 8 ? 256 : wps * 32;
        int bps = wps * 32 / nt;
        int iters = 200000 / ILP;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<ILP><<<sms * bps, nt>>>(out, iters / 10, 0.999999, 1e-9);
        cudaEventRecord(e0);
        k<ILP><<<sms * bps, nt>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 16 * ILP * (double)iters * sms * bps * nt;
        printf("ILP=%d warps/SM=%2d : %6.2f TFLOP/s\n", ILP, wps, fl / (ms * 1e-3) / 1e12);
    }
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    double* out; cudaMalloc(&out, 8);
    run<1>(pr.multiProcessorCount, out);
    run<2>(pr.multiProcessorCount, out);
    run<4>(pr.multiProcessorCount, out);
    run<8>(pr.multiProcessorCount, out);
    return 0;
}
