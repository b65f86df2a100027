// ffma2_pipe.cu -- throughput and dependent-issue latency of the packed FP32 instructions of sm_100 (FFMA2) next to FFMA:
// does a packed instruction cost one issue slot or two?   nvcc -arch=sm_100a tools/ffma2_pipe.cu -o /tmp/ffma2 && /tmp/ffma2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
template <int ILP, bool PACKED>
__global__ void k(float *out, int iters, long long *cycles)
{
    float a[ILP]; unsigned long long p[ILP];
    for (int k = 0; k < ILP; k++) { a[k] = threadIdx.x * 1e-3f + k; p[k] = (unsigned long long)__float_as_uint(a[k]) * 0x100000001ull; }
    const float m = 1.0000001f, c = 1e-7f;
    const unsigned long long M = (unsigned long long)__float_as_uint(m) * 0x100000001ull, C = (unsigned long long)__float_as_uint(c) * 0x100000001ull;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            if (PACKED) p[k] = fma2(p[k], M, C);
            else asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(a[k]) : "f"(a[k]), "f"(m), "f"(c));
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int k = 0; k < ILP; k++) s += PACKED ? __uint_as_float((unsigned)p[k]) + __uint_as_float((unsigned)(p[k] >> 32)) : a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP, bool PACKED>
void run(const char *name, int warps_per_sm)
{
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<ILP, PACKED><<<148, warps_per_sm * 32>>>(out, iters, cyc);
    k<ILP, PACKED><<<148, warps_per_sm * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_smsp = (double)iters * ILP * (warps_per_sm / 4.0);       // instructions issued per scheduler
    printf("%-6s ILP %d, %2d warps/SM: %.2f cycles per instruction per scheduler (%.2f per warp-instruction chain step)\n", name, ILP,
           warps_per_sm, (double)h / per_smsp, (double)h / iters / ILP);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<1, false>("FFMA", 4);  run<1, true>("FFMA2", 4);        // one warp per scheduler, dependent chain: latency
    run<8, false>("FFMA", 4);  run<8, true>("FFMA2", 4);        // one warp per scheduler, 8 independent chains
    run<8, false>("FFMA", 16); run<8, true>("FFMA2", 16);       // four warps per scheduler: throughput
    run<4, false>("FFMA", 32); run<4, true>("FFMA2", 32);
    return 0;
}
