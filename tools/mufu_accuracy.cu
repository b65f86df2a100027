// Accuracy of the double-precision MUFU seeds (rcp/rsqrt.approx.ftz.f64) and of the Newton-refined values
// used by the pair kernel (csrc/sph_math.cuh).  nvcc -arch=sm_100a tools/mufu_accuracy.cu -o /tmp/mufu && /tmp/mufu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__global__ void k(int n, double *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = 1.0 + (double)i / n * 7.0 + 1e-7 * i;      // [1, 8): covers mantissa space and an exponent parity flip
    double y0, s0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(s0) : "d"(x));
    double y = y0, s = s0, hx = 0.5 * x, e;
    double err[10];
    err[0] = fabs(y0 * x - 1.0);
    err[4] = fabs(s0 * s0 * x - 1.0) * 0.5;
    for (int it = 1; it <= 3; it++) {
        e = fma(-x, y, 1.0); y = fma(y, e, y);
        e = fma(-hx * s, s, 0.5); s = fma(s, e, s);
        err[it] = fabs(y - 1.0 / x) * x;
        err[4 + it] = fabs(s - 1.0 / sqrt(x)) * sqrt(x);
    }
    {   // the third-order single step the pair kernel uses now (sph_math.cuh, OSPH_NEWTON_STEPS == 0)
        double e3 = fma(-x, y0, 1.0), t3 = fma(e3, e3, e3), y3 = fma(y0, t3, y0);
        double f3 = fma(-(x * s0), s0, 1.0), u3 = f3 * fma(f3, 0.375, 0.5), s3 = fma(s0, u3, s0);
        err[8] = fabs(y3 - 1.0 / x) * x;
        err[9] = fabs(s3 - 1.0 / sqrt(x)) * sqrt(x);
    }
    for (int k2 = 0; k2 < 10; k2++) out[k2 * n + i] = err[k2];
}

int main()
{
    const int n = 1 << 20;
    double *d, *h = new double[10 * n];
    cudaMalloc(&d, sizeof(double) * 10 * n);
    k<<<n / 256, 256>>>(n, d);
    cudaMemcpy(h, d, sizeof(double) * 10 * n, cudaMemcpyDeviceToHost);
    const char *names[10] = {"rcp seed", "rcp 1 NR", "rcp 2 NR", "rcp 3 NR", "rsqrt seed", "rsqrt 1 NR", "rsqrt 2 NR", "rsqrt 3 NR",
                             "rcp cubic", "rsqrt cubic"};
    for (int k2 = 0; k2 < 10; k2++) {
        double m = 0;
        for (int i = 0; i < n; i++) m = fmax(m, h[k2 * n + i]);
        printf("%-12s max rel err %.3e  (2^%.1f)\n", names[k2], m, m > 0 ? log2(m) : -99.0);
    }
    return 0;
}
