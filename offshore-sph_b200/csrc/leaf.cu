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

// ---------------------------------------------------------------------------------------------------------
// The four per-neighbour equations on one computed-neighbour table (include/osph.h: osph_leaf_equations).
// Grid-stride accumulation per thread, shared-memory tree per CTA, then one thread per output adds the CTA
// partials in index order: the result depends on J only, not on the schedule.
// ---------------------------------------------------------------------------------------------------------
#define LEAF_EQ_THREADS 256
#define LEAF_EQ_NV 7
#define LEAF_EQ_MAX_BLOCKS 1024

struct LeafEqParams { double self_p, self_rho, self_h, self_c, alpha, beta, epsilon, r0, D, p1, p2; };

__global__ void __launch_bounds__(LEAF_EQ_THREADS)
k_leaf_equations(long long J, const signed char *__restrict__ label, const double *__restrict__ cols, LeafEqParams P,
                 double *__restrict__ partial)
{
    __shared__ double sh[LEAF_EQ_NV][LEAF_EQ_THREADS];
    const int tid = threadIdx.x;
    double acc[LEAF_EQ_NV];
#pragma unroll
    for (int k = 0; k < LEAF_EQ_NV; k++) acc[k] = 0.0;
    const double slf = P.self_p / (P.self_rho * P.self_rho);                           // Momentum.py:27
    for (long long j = (long long)blockIdx.x * LEAF_EQ_THREADS + tid; j < J; j += (long long)gridDim.x * LEAF_EQ_THREADS) {
#define COL(k) cols[(long long)(k) * J + j]
        const bool fluid = label[j] == OSPH_FLUID;
        const double m = COL(OSPH_COMP_M), rho = COL(OSPH_COMP_RHO), r = COL(OSPH_COMP_R);
        const double x = COL(OSPH_COMP_X), y = COL(OSPH_COMP_Y), vx = COL(OSPH_COMP_VX), vy = COL(OSPH_COMP_VY);
        const double dwx = COL(OSPH_COMP_DWX), dwy = COL(OSPH_COMP_DWY);
        if (fluid) {
            acc[0] += m * (vx * dwx + vy * dwy);                                       // Continuity.py:12-16
            const double othr = COL(OSPH_COMP_P) / (rho * rho);                        // Momentum.py:35
            const double dot = vx * x + vy * y;
            double PI = 0.0;
            if (dot < 0) {                                                             // Momentum.py:40-49
                const double hij = 0.5 * (P.self_h + COL(OSPH_COMP_H));
                const double cij = 0.5 * (P.self_c + COL(OSPH_COMP_C));
                const double rhoij = 0.5 * (P.self_rho + rho);
                const double mu = hij * dot / (r * r + 0.01 * hij * hij);
                PI = mu * (P.beta * mu - P.alpha * cij) / rhoij;
            }
            const double factor = slf + othr + PI;
            acc[1] += -m * factor * dwx;
            acc[2] += -m * factor * dwy;
        } else if (!(r > P.r0) && r > 1e-12) {                                         // BoundaryForce.py:30-40
            const double frac = P.r0 / r;
            const double fac = P.D * (pow(frac, P.p1) - pow(frac, P.p2));
            acc[5] += fac * x / (r * r);
            acc[6] += fac * y / (r * r);
        }
        {                                                                              // XSPH.py:24-29, every label
            const double fac = -P.epsilon * m * COL(OSPH_COMP_W) / (0.5 * (P.self_rho + rho));
            acc[3] += fac * vx;
            acc[4] += fac * vy;
        }
#undef COL
    }
#pragma unroll
    for (int k = 0; k < LEAF_EQ_NV; k++) sh[k][tid] = acc[k];
    __syncthreads();
    for (int s = LEAF_EQ_THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) {
#pragma unroll
            for (int k = 0; k < LEAF_EQ_NV; k++) sh[k][tid] += sh[k][tid + s];
        }
        __syncthreads();
    }
    if (tid < LEAF_EQ_NV) partial[(long long)blockIdx.x * LEAF_EQ_NV + tid] = sh[tid][0];
}

__global__ void k_leaf_equations_sum(int nblocks, const double *__restrict__ partial, double *__restrict__ out)
{
    const int k = threadIdx.x;
    if (k >= LEAF_EQ_NV) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; b++) s += partial[(long long)b * LEAF_EQ_NV + k];
    out[k] = s;
}

// Courant.py:17-31: h_min starts at 10e10, c_max at 1e-10 (so an empty table gives alpha * 1e11 / 1e-10, as there)
__global__ void k_leaf_courant_init(unsigned long long *mm)
{
    mm[0] = enc_f64(10e10); mm[1] = enc_f64(1e-10);
}
__global__ void __launch_bounds__(256)
k_leaf_courant(long long J, const double *__restrict__ h, const double *__restrict__ c, unsigned long long *mm)
{
    double hm = INFINITY, cm = -INFINITY;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < J; j += (long long)gridDim.x * blockDim.x) {
        hm = fmin(hm, h[j]); cm = fmax(cm, c[j]);
    }
    hm = warp_min(hm); cm = warp_max(cm);                   // all 32 lanes arrive: no early exit above
    if ((threadIdx.x & 31) == 0) {
        if (hm < INFINITY) atomicMin(&mm[0], enc_f64(hm));
        if (cm > -INFINITY) atomicMax(&mm[1], enc_f64(cm));
    }
}
__global__ void k_leaf_courant_final(double alpha, const unsigned long long *mm, double *out)
{
    out[0] = alpha * dec_f64(mm[0]) / dec_f64(mm[1]);
}

// _assignProps, SolverTools.py:97-101: column k of `out` = self4[k] - column k of `nbr`, k = x, y, vx, vy
__global__ void k_leaf_differences(long long J, double sx, double sy, double svx, double svy, const double *__restrict__ nbr,
                                   double *__restrict__ out)
{
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= J) return;
    out[j] = sx - nbr[j];
    out[J + j] = sy - nbr[J + j];
    out[2 * J + j] = svx - nbr[2 * J + j];
    out[3 * J + j] = svy - nbr[3 * J + j];
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

extern "C" int osph_leaf_equations(int device, int64_t J, const int8_t *label, const double *cols, double self_p,
                                   double self_rho, double self_h, double self_c, double alpha, double beta, double epsilon,
                                   double r0, double D, double p1, double p2, double out[7])
{
    std::vector<void *> bufs;
    if (!out || J < 0 || (J > 0 && (!label || !cols))) { g_leaf_error = "bad argument"; return OSPH_E_INVALID; }
    for (int k = 0; k < LEAF_EQ_NV; k++) out[k] = 0.0;
    if (J == 0) return 0;
    LEAF_CUDA(cudaSetDevice(device));
    const long long want = (J + LEAF_EQ_THREADS - 1) / LEAF_EQ_THREADS;
    const int nblocks = (int)(want < LEAF_EQ_MAX_BLOCKS ? want : LEAF_EQ_MAX_BLOCKS);
    double *d = nullptr; signed char *dl = nullptr;
    const size_t ncol = (size_t)OSPH_COMP_NCOLS * (size_t)J;
    LEAF_CUDA(cudaMalloc(&d, sizeof(double) * (ncol + (size_t)nblocks * LEAF_EQ_NV + LEAF_EQ_NV))); bufs.push_back(d);
    LEAF_CUDA(cudaMalloc(&dl, (size_t)J)); bufs.push_back(dl);
    LEAF_CUDA(cudaMemcpy(d, cols, sizeof(double) * ncol, cudaMemcpyHostToDevice));
    LEAF_CUDA(cudaMemcpy(dl, label, (size_t)J, cudaMemcpyHostToDevice));
    double *partial = d + ncol, *dout = partial + (size_t)nblocks * LEAF_EQ_NV;
    LeafEqParams P = {self_p, self_rho, self_h, self_c, alpha, beta, epsilon, r0, D, p1, p2};
    k_leaf_equations<<<nblocks, LEAF_EQ_THREADS>>>(J, dl, d, P, partial);
    LEAF_CUDA(cudaGetLastError());
    k_leaf_equations_sum<<<1, 32>>>(nblocks, partial, dout);
    LEAF_CUDA(cudaGetLastError());
    LEAF_CUDA(cudaMemcpy(out, dout, sizeof(double) * LEAF_EQ_NV, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaFree(dl);
    return 0;
}

extern "C" int osph_leaf_courant(int device, double alpha, int64_t J, const double *h, const double *c, double *out)
{
    std::vector<void *> bufs;
    if (!out || J < 0 || (J > 0 && (!h || !c))) { g_leaf_error = "bad argument"; return OSPH_E_INVALID; }
    LEAF_CUDA(cudaSetDevice(device));
    double *d = nullptr;
    LEAF_CUDA(cudaMalloc(&d, sizeof(double) * (2 * (size_t)J + 4))); bufs.push_back(d);
    unsigned long long *mm = reinterpret_cast<unsigned long long *>(d + 2 * (size_t)J);
    double *dout = d + 2 * (size_t)J + 2;
    if (J > 0) {
        LEAF_CUDA(cudaMemcpy(d, h, sizeof(double) * (size_t)J, cudaMemcpyHostToDevice));
        LEAF_CUDA(cudaMemcpy(d + J, c, sizeof(double) * (size_t)J, cudaMemcpyHostToDevice));
    }
    k_leaf_courant_init<<<1, 1>>>(mm);
    LEAF_CUDA(cudaGetLastError());
    if (J > 0) {
        const long long want = (J + 255) / 256;
        k_leaf_courant<<<(int)(want < 1024 ? want : 1024), 256>>>(J, d, d + J, mm);
        LEAF_CUDA(cudaGetLastError());
    }
    k_leaf_courant_final<<<1, 1>>>(alpha, mm, dout);
    LEAF_CUDA(cudaGetLastError());
    LEAF_CUDA(cudaMemcpy(out, dout, sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}

extern "C" int osph_leaf_differences(int device, int64_t J, const double self4[4], const double *nbr, double *out)
{
    std::vector<void *> bufs;
    if (J <= 0) return 0;
    if (!self4 || !nbr || !out) { g_leaf_error = "bad argument"; return OSPH_E_INVALID; }
    LEAF_CUDA(cudaSetDevice(device));
    double *d = nullptr;
    LEAF_CUDA(cudaMalloc(&d, sizeof(double) * 8 * (size_t)J)); bufs.push_back(d);
    LEAF_CUDA(cudaMemcpy(d, nbr, sizeof(double) * 4 * (size_t)J, cudaMemcpyHostToDevice));
    k_leaf_differences<<<grid_of(J), 256>>>(J, self4[0], self4[1], self4[2], self4[3], d, d + 4 * (size_t)J);
    LEAF_CUDA(cudaGetLastError());
    LEAF_CUDA(cudaMemcpy(out, d + 4 * (size_t)J, sizeof(double) * 4 * (size_t)J, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}
