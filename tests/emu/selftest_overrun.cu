// selftest_overrun.cu -- TEST INFRASTRUCTURE: kernels that deliberately step past their shared memory.  Built through the
// same preprocess.py + emu.cpp as the product kernels (tests/test_emu_guard.py); each mode must die with SIGSEGV, the
// in-bounds mode must exit 0.  This is the property round 1 lacked: the emulator absorbed a shared-memory overrun of the
// pair kernel that the hardware turned into CUDA error 700.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

__global__ void k_dyn(int *out, int n_read)
{
    extern __shared__ __align__(16) unsigned char raw[];
    int *buf = reinterpret_cast<int *>(raw);
    buf[threadIdx.x] = (int)threadIdx.x;
    __syncthreads();
    int s = 0;
    for (int i = 0; i < n_read; i++) s += buf[i];          // n_read > blockDim.x reads past the launch's allocation
    out[threadIdx.x] = s;
}

__global__ void k_static(int *out, int n_read)
{
    __shared__ int tab[64];
    __shared__ double two[2][8];
    tab[threadIdx.x] = 1;
    if (threadIdx.x < 8) { two[0][threadIdx.x] = 1.0; two[1][threadIdx.x] = 2.0; }
    __syncthreads();
    int s = (int)two[1][threadIdx.x & 7];
    for (int i = 0; i < n_read; i++) s += tab[i];
    out[threadIdx.x] = s;
}

int main(int argc, char **argv)
{
    const char *mode = argc > 1 ? argv[1] : "ok";
    int *out = nullptr;
    cudaMalloc(&out, 64 * sizeof(int));
    if (!strcmp(mode, "ok")) {
        k_dyn<<<2, 64, 64 * sizeof(int)>>>(out, 64);
        k_static<<<2, 64>>>(out, 64);
    } else if (!strcmp(mode, "dyn")) {
        k_dyn<<<1, 64, 64 * sizeof(int)>>>(out, 64 + 8);      // 32 bytes past the end
    } else if (!strcmp(mode, "static")) {
        k_static<<<1, 64>>>(out, 64 + 8);
    }
    cudaDeviceSynchronize();
    printf("survived %s\n", mode);
    return 0;
}
