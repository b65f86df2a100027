// leaf.cu -- stand-alone device versions of the reference's leaf functions, for callers that use the
// plugin objects outside Solver.run() (examples/IceBreak.py's pressure probe, the reference's unit tests
// pointed at this package): kernel.evaluate / kernel.gradient on arrays, the Tait EOS ufuncs, computeH.
// Each call copies its small host arrays to the device, runs one elementwise kernel and copies back.
#include <vector>

#include "common.cuh"
#include "sph_math.cuh"

static thread_local std::string g_leaf_error;
extern "C" const char *osph_leaf_last_error(void) { return g_leaf_error.c_str(); }

#define LEAF_CUDA(call)                                                         \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) {                                               \
            g_leaf_error = std::string(#call) + ": " + cudaGetErrorString(e__); \
            cudaGetLastError();                                                 \
            for (void *p__ : bufs) cudaFree(p__);                               \
            return e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ? OSPH_E_NO_DEVICE : OSPH_E_CUDA; \
        }                                                                       \
    } while (0)

// evaluate (what == 0): reference CubicSpline.py:10-36, Wendland.py:9-33, Gaussian.py:16-25
// gradient (what == 1): CubicSpline.py:38-70, Wendland.py:35-64, Gaussian.py:38-59
template <int KID>
__global__ void k_leaf_kernel(int what, long long n, const double *__restrict__ x, const double *__restrict__ r,
                              const double *__restrict__ h, double *__restrict__ out)
{
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double rr = r[j], hh = h[j];
    double inv_h = 1.0 / hh;
    double q = rr * inv_h;
    double inv_r = rr < 1e-10 ? 0.0 : 1.0 / rr;
    if (KID == OSPH_KERNEL_GAUSSIAN) inv_r = (rr * hh > 1e-12) ? 1.0 / rr : 0.0;
    double w, g;
    sph_kernel<double, KID>(q, inv_h, inv_r, w, g);
    if (KID == OSPH_KERNEL_GAUSSIAN && !(q <= 3.0)) { w = 0.0; g = 0.0; }
    out[j] = what == 0 ? w : g * x[j];
}

__global__ void k_leaf_tait_pressure(long long n, const double *__restrict__ rho, const signed char *__restrict__ label,
                                     double gamma, double B, double rho0, double Pb, double *__restrict__ out)
{
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double p = 0.0;
    if (label[j] == OSPH_FLUID) p = (pow(rho[j] / rho0, gamma) - 1.0) * B;      // TaitEOS.py:6-31
    out[j] = p + Pb;                                                            // WCSPH.py:143-149
}

__global__ void k_leaf_tait_height(long long n, const double *__restrict__ y, double rho0, double H, double B,
                                   double gamma, double *__restrict__ out)
{
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double frac = rho0 * 9.81 * (H - y[j]) / B;                                 // TaitEOS.py:46-65
    out[j] = rho0 * pow(1.0 + frac, 1.0 / gamma);
}

__global__ void k_leaf_compute_h(long long n, double sigma, const double *__restrict__ m,
                                 const double *__restrict__ rho, double *__restrict__ out)
{
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    out[j] = rho[j] > 1e-12 ? sigma * sqrt(m[j] / rho[j]) : 0.0;                // SolverTools.py:106-118
}

static int grid_of(long long n) { return (int)((n + 255) / 256); }

extern "C" int osph_leaf_kernel(int device, int kernel, int what, int64_t n, const double *x, const double *r,
                                const double *h, double *out)
{
    std::vector<void *> bufs;
    if (n <= 0) return 0;
    if (!r || !h || !out || (what == 1 && !x) || kernel < 0 || kernel > 2) { g_leaf_error = "bad argument"; return OSPH_E_INVALID; }
    LEAF_CUDA(cudaSetDevice(device));
    double *d = nullptr;
    LEAF_CUDA(cudaMalloc(&d, sizeof(double) * 4 * n)); bufs.push_back(d);
    double *dx = d, *dr = d + n, *dh = d + 2 * n, *dout = d + 3 * n;
    if (what == 1) LEAF_CUDA(cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice));
    LEAF_CUDA(cudaMemcpy(dr, r, sizeof(double) * n, cudaMemcpyHostToDevice));
    LEAF_CUDA(cudaMemcpy(dh, h, sizeof(double) * n, cudaMemcpyHostToDevice));
    if (kernel == OSPH_KERNEL_CUBIC) k_leaf_kernel<OSPH_KERNEL_CUBIC><<<grid_of(n), 256>>>(what, n, dx, dr, dh, dout);
    else if (kernel == OSPH_KERNEL_WENDLAND) k_leaf_kernel<OSPH_KERNEL_WENDLAND><<<grid_of(n), 256>>>(what, n, dx, dr, dh, dout);
    else k_leaf_kernel<OSPH_KERNEL_GAUSSIAN><<<grid_of(n), 256>>>(what, n, dx, dr, dh, dout);
    LEAF_CUDA(cudaGetLastError());
    LEAF_CUDA(cudaMemcpy(out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}

extern "C" int osph_leaf_tait_pressure(int device, int64_t n, const double *rho, const int8_t *label, double gamma,
                                       double B, double rho0, double Pb, double *out)
{
    std::vector<void *> bufs;
    if (n <= 0) return 0;
    if (!rho || !label || !out) { g_leaf_error = "bad argument"; return OSPH_E_INVALID; }
    LEAF_CUDA(cudaSetDevice(device));
    double *d = nullptr; signed char *dl = nullptr;
    LEAF_CUDA(cudaMalloc(&d, sizeof(double) * 2 * n)); bufs.push_back(d);
    LEAF_CUDA(cudaMalloc(&dl, n)); bufs.push_back(dl);
    LEAF_CUDA(cudaMemcpy(d, rho, sizeof(double) * n, cudaMemcpyHostToDevice));
    LEAF_CUDA(cudaMemcpy(dl, label, n, cudaMemcpyHostToDevice));
    k_leaf_tait_pressure<<<grid_of(n), 256>>>(n, d, dl, gamma, B, rho0, Pb, d + n);
    LEAF_CUDA(cudaGetLastError());
    LEAF_CUDA(cudaMemcpy(out, d + n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaFree(dl);
    return 0;
}

extern "C" int osph_leaf_tait_height(int device, int64_t n, const double *y, double rho0, double H, double B,
                                     double gamma, double *out)
{
    std::vector<void *> bufs;
    if (n <= 0) return 0;
    if (!y || !out) { g_leaf_error = "bad argument"; return OSPH_E_INVALID; }
    LEAF_CUDA(cudaSetDevice(device));
    double *d = nullptr;
    LEAF_CUDA(cudaMalloc(&d, sizeof(double) * 2 * n)); bufs.push_back(d);
    LEAF_CUDA(cudaMemcpy(d, y, sizeof(double) * n, cudaMemcpyHostToDevice));
    k_leaf_tait_height<<<grid_of(n), 256>>>(n, d, rho0, H, B, gamma, d + n);
    LEAF_CUDA(cudaGetLastError());
    LEAF_CUDA(cudaMemcpy(out, d + n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}

extern "C" int osph_leaf_compute_h(int device, int64_t n, double sigma, const double *m, const double *rho, double *out)
{
    std::vector<void *> bufs;
    if (n <= 0) return 0;
    if (!m || !rho || !out) { g_leaf_error = "bad argument"; return OSPH_E_INVALID; }
    LEAF_CUDA(cudaSetDevice(device));
    double *d = nullptr;
    LEAF_CUDA(cudaMalloc(&d, sizeof(double) * 3 * n)); bufs.push_back(d);
    LEAF_CUDA(cudaMemcpy(d, m, sizeof(double) * n, cudaMemcpyHostToDevice));
    LEAF_CUDA(cudaMemcpy(d + n, rho, sizeof(double) * n, cudaMemcpyHostToDevice));
    k_leaf_compute_h<<<grid_of(n), 256>>>(n, sigma, d, d + n, d + 2 * n);
    LEAF_CUDA(cudaGetLastError());
    LEAF_CUDA(cudaMemcpy(out, d + 2 * n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}
