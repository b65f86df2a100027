// cdp_tail_check.cu -- does a device-side tail launch (CUDA dynamic parallelism 2) order itself between two host launches
// of one stream on this driver, and what does it cost?   nvcc -arch=sm_100a -rdc=true tools/cdp_tail_check.cu -lcudadevrt
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_child(int *data, int n, int add)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) data[i] += add;
}
__global__ void k_parent(int *data, int n, const int *flag, int nchild)
{
    if (*flag)
        for (int c = 0; c < nchild; c++) k_child<<<(n + 255) / 256, 256, 0, cudaStreamTailLaunch>>>(data, n, 1);
}
__global__ void k_empty(const int *flag) { if (*flag == 12345) printf("never\n"); }
__global__ void k_check(const int *data, int n, int expect, int *bad)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && data[i] != expect) atomicAdd(bad, 1);
}

int main()
{
    const int n = 1 << 20;
    int *data, *flag, *bad;
    cudaMalloc(&data, n * 4); cudaMalloc(&flag, 4); cudaMalloc(&bad, 4);
    cudaMemset(data, 0, n * 4); cudaMemset(bad, 0, 4);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int expect = 0;
    for (int on = 0; on < 2; on++) {
        cudaMemcpy(flag, &on, 4, cudaMemcpyHostToDevice);
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0, st);
            for (int it = 0; it < 100; it++) {
                k_parent<<<1, 1, 0, st>>>(data, n, flag, 10);
                expect += on ? 10 : 0;
                k_check<<<(n + 255) / 256, 256, 0, st>>>(data, n, expect, bad);      // must see all ten children of THIS parent
            }
            cudaEventRecord(e1, st); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("flag %d: parent (+10 tail children) + check: %.2f us per iteration\n", on, ms * 10.0f);
        }
    }
    // ten empty host launches, for comparison with the early-exit alternative
    cudaEventRecord(e0, st);
    for (int it = 0; it < 1000; it++) k_empty<<<493, 256, 0, st>>>(flag);
    cudaEventRecord(e1, st); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("early-exit host launch of 493 CTAs: %.2f us each\n", ms);
    int hb; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("ordering violations: %d  (%s)\n", hb, cudaGetErrorString(cudaGetLastError()));
    return hb != 0;
}
