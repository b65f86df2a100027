// FP64 / FP32 FMA latency and throughput on this GPU (context for the pair-kernel roofline).
// nvcc -arch=sm_100a -O3 tools/fp64_pipe.cu -o /tmp/fp64pipe && /tmp/fp64pipe
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int ILP>
__global__ void k(T *out, int iters, long long *cycles)
{
    T a[ILP];
    for (int k2 = 0; k2 < ILP; k2++) a[k2] = (T)(threadIdx.x + k2) * (T)1e-3;
    const T b = (T)1.0000001, c = (T)1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k2 = 0; k2 < ILP; k2++) a[k2] = a[k2] * b + c;
    }
    long long t1 = clock64();
    T s = 0;
    for (int k2 = 0; k2 < ILP; k2++) s += a[k2];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <typename T, int ILP>
void run(const char *name, int warps_per_sm)
{
    int dev = 0, sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int threads = 32 * warps_per_sm > 1024 ? 1024 : 32 * warps_per_sm;
    int blocks_per_sm = (32 * warps_per_sm + threads - 1) / threads;
    T *out; long long *cyc; cudaMalloc(&out, sizeof(T) * threads * blocks_per_sm * sms); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    k<T, ILP><<<blocks_per_sm * sms, threads>>>(out, iters, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<T, ILP><<<blocks_per_sm * sms, threads>>>(out, iters, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double fmas = (double)iters * ILP * threads * blocks_per_sm * sms;
    printf("%-5s ILP %d warps/SM %2d : %6.2f cycles per dependent FMA step, %7.2f TFLOP/s (%.3f FMA/clk/SM)\n", name, ILP,
           warps_per_sm, (double)c / iters, 2 * fmas / (ms * 1e-3) / 1e12, fmas / ((double)c * sms));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<double, 1>("fp64", 1); run<double, 1>("fp64", 4); run<double, 1>("fp64", 16); run<double, 2>("fp64", 16);
    run<double, 4>("fp64", 16); run<double, 1>("fp64", 32); run<double, 4>("fp64", 32); run<double, 8>("fp64", 64);
    run<float, 1>("fp32", 1); run<float, 1>("fp32", 16); run<float, 4>("fp32", 24); run<float, 8>("fp32", 64);
    return 0;
}
