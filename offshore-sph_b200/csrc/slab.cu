// slab.cu -- 1-D slab decomposition along x across the GPUs of one box.
// One context per GPU owns the particles with x_lo <= x < x_hi.  Per step, after the predictor:
//   osph_slab_pack    classifies the owned particles, packs migrants (full records) and halo particles (light
//                     records) for both neighbours into caller-owned device buffers (torch tensors that NCCL
//                     sends straight from), and keeps light copies of the migrants as this rank's first ghosts;
//   [the caller exchanges counts + payloads over NCCL/NVLink; halos are received directly into the ghost buffer]
//   osph_slab_commit  removes the migrants from the owned set (hole filling), appends the received ones and
//                     installs the all-reduced grid scalars, so every rank forms the SAME reference grid.
// The reference has no counterpart: it is a single-threaded, single-process code.
#include <stdlib.h>

#include "common.cuh"
#include "step.cuh"

struct SlabPackArgs {
    int n;
    const signed char *label;
    const int *row;
    const double *f[OSPH_NUM_FIELDS];
    double x_lo, x_hi, width;
    double *mig_left, *mig_right, *halo_left, *halo_right, *ghost;
    int mig_cap, halo_cap, ghost_cap;
    int *counters;      // [0] mig_left [1] mig_right [2] halo_left [3] halo_right [4] ghosts [5] overflow flag
    int *mig_slots;     // slots of the migrants (unordered)
    int mode;           // bit 0: classify and pack migrants, bit 1: halos (slab cadence: two passes on a sorting step)
    int *halo_idx_l, *halo_idx_r;   // optional: storage slot of the particle behind every halo record (frozen halo lists)
};

__device__ __forceinline__ void write_light(double *dst, const SlabPackArgs &a, int i)
{
    dst[0] = a.f[OSPH_F_X][i]; dst[1] = a.f[OSPH_F_Y][i]; dst[2] = a.f[OSPH_F_VX][i]; dst[3] = a.f[OSPH_F_VY][i];
    dst[4] = a.f[OSPH_F_RHO][i]; dst[5] = a.f[OSPH_F_M][i]; dst[6] = a.f[OSPH_F_H][i]; dst[7] = (double)a.label[i];
}
__device__ __forceinline__ void write_full(double *dst, const SlabPackArgs &a, int i)
{
#pragma unroll
    for (int k = 0; k < OSPH_NUM_FIELDS; k++) dst[k] = a.f[k][i];
    dst[OSPH_NUM_FIELDS] = (double)a.label[i];
    dst[OSPH_NUM_FIELDS + 1] = (double)a.row[i];
}

// One slot per lane of a group, one atomic per warp: the lanes that take part (`take`) are counted with a ballot, the first
// of them reserves that many slots, every lane gets base + its rank.  Halo particles come in runs of a few dozen
// consecutive particles (the cells next to a face, row by row), so a warp usually reserves several slots at once
// instead of hitting one counter per particle.  All 32 lanes must call.
__device__ __forceinline__ int warp_reserve(int *counter, bool take)
{
    const unsigned int m = __ballot_sync(0xffffffffu, take);
    if (m == 0) return 0;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(256)
k_slab_pack(SlabPackArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < a.n;
    const double x = valid ? a.f[OSPH_F_X][i] : 0.0;
    int side = !valid ? -1 : (x < a.x_lo ? 0 : (x >= a.x_hi ? 1 : -1));
    const bool outside = side >= 0;
    if (!(a.mode & 1)) side = -1;                  // halo-only pass: nobody migrates (the migrants of this step have left already)
    // migrants: full record to the new owner, light copy kept here as a ghost, slot remembered for the hole filling
    const int kl = warp_reserve(&a.counters[0], side == 0), kr = warp_reserve(&a.counters[1], side == 1);
    const int g = warp_reserve(&a.counters[4], side >= 0), m = warp_reserve(&a.counters[6], side >= 0);
    if (side >= 0) {
        const int k = side ? kr : kl;
        if (k < a.mig_cap && g < a.ghost_cap) {
            write_full((side ? a.mig_right : a.mig_left) + (size_t)k * OSPH_WIRE_FULL, a, i);
            write_light(a.ghost + (size_t)g * OSPH_WIRE_HALO, a, i);
            a.mig_slots[m] = i;
        } else a.counters[5] = 1;
    }
    // halo: light records for the neighbour on that side (a particle of a narrow slab can be in both halos)
    const bool halo = valid && (a.mode & 2) && !(outside && (a.mode & 1));
    const bool hl = halo && x < a.x_lo + a.width, hr = halo && x >= a.x_hi - a.width;
    const int khl = warp_reserve(&a.counters[2], hl), khr = warp_reserve(&a.counters[3], hr);
    if (hl) {
        if (khl < a.halo_cap) { write_light(a.halo_left + (size_t)khl * OSPH_WIRE_HALO, a, i); if (a.halo_idx_l) a.halo_idx_l[khl] = i; }
        else a.counters[5] = 1;
    }
    if (hr) {
        if (khr < a.halo_cap) { write_light(a.halo_right + (size_t)khr * OSPH_WIRE_HALO, a, i); if (a.halo_idx_r) a.halo_idx_r[khr] = i; }
        else a.counters[5] = 1;
    }
}

__global__ void k_slab_meta(const int *counters, const StepScalars *sc, double *meta)
{
    meta[0] = counters[0]; meta[1] = counters[1]; meta[2] = counters[2]; meta[3] = counters[3];
    meta[4] = dec_f64(sc->xmin); meta[5] = -dec_f64(sc->xmax); meta[6] = dec_f64(sc->ymin); meta[7] = -dec_f64(sc->ymax);
    meta[8] = dec_f64(sc->hmin_all); meta[9] = -dec_f64(sc->hmax_all);
    meta[10] = counters[5];
    meta[11] = -dec_f64(sc->disp2max);        // (negated: the ranks' values are combined with min) largest displacement^2 since the last sort
}

__global__ void k_slab_set_bounds(StepScalars *sc, double xmin, double nxmax, double ymin, double nymax, double hmin,
                                  double nhmax)
{
    sc->xmin = enc_f64(xmin); sc->xmax = enc_f64(-nxmax); sc->ymin = enc_f64(ymin); sc->ymax = enc_f64(-nymax);
    sc->hmin_all = enc_f64(hmin); sc->hmax_all = enc_f64(-nhmax);
}

// Slab cadence, a step that reuses the binning: the SAME particles as at the last sort, in the same record slots, with their
// current values (the receiver's sorted order refers to these slots).  No atomics, no classification.
__global__ void __launch_bounds__(256)
k_slab_repack(SlabPackArgs a, const int *__restrict__ idx_l, int n_l, const int *__restrict__ idx_r, int n_r)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_l) write_light(a.halo_left + (size_t)k * OSPH_WIRE_HALO, a, idx_l[k]);
    else if (k < n_l + n_r) write_light(a.halo_right + (size_t)(k - n_l) * OSPH_WIRE_HALO, a, idx_r[k - n_l]);
}

// ... and the receiver's side of such a step: global bounds, h extrema and displacement from the all-gathered meta rows,
// on the device (on a sorting step the host does this on its way through the counts).
__global__ void k_slab_install(const double *__restrict__ all_meta, int world, StepScalars *sc)
{
    double b[6], d = all_meta[11];
    for (int k = 0; k < 6; k++) b[k] = all_meta[4 + k];
    for (int r = 1; r < world; r++) {
        for (int k = 0; k < 6; k++) b[k] = fmin(b[k], all_meta[12 * r + 4 + k]);
        d = fmin(d, all_meta[12 * r + 11]);
    }
    sc->xmin = enc_f64(b[0]); sc->xmax = enc_f64(-b[1]); sc->ymin = enc_f64(b[2]); sc->ymax = enc_f64(-b[3]);
    sc->hmin_all = enc_f64(b[4]); sc->hmax_all = enc_f64(-b[5]);
    sc->disp2max = enc_f64(-d);
}

// hole filling: the m migrants leave; the last m slots are vacated, their non-migrant occupants fill the holes
__global__ void k_slab_mark(const int *__restrict__ mig_slots, int m, int n_new, int *__restrict__ tail_flag,
                            int *__restrict__ holes, int *__restrict__ nh)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    int h = mig_slots[k];
    if (h >= n_new) tail_flag[h - n_new] = 1;
    else holes[atomicAdd(nh, 1)] = h;
}
__global__ void k_slab_fillers(const int *__restrict__ tail_flag, int m, int n_new, int *__restrict__ fillers,
                               int *__restrict__ nf)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    if (!tail_flag[t]) fillers[atomicAdd(nf, 1)] = n_new + t;
}
struct SlabMoveArgs { double *f[OSPH_NUM_FIELDS]; signed char *label; int *row; };
__global__ void k_slab_move(SlabMoveArgs a, const int *__restrict__ holes, const int *__restrict__ fillers,
                            const int *__restrict__ nh, int m)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int k = t / (OSPH_NUM_FIELDS + 2), c = t % (OSPH_NUM_FIELDS + 2);
    if (k >= m || k >= *nh) return;
    int dst = holes[k], src = fillers[k];
    if (c < OSPH_NUM_FIELDS) a.f[c][dst] = a.f[c][src];
    else if (c == OSPH_NUM_FIELDS) a.label[dst] = a.label[src];
    else a.row[dst] = a.row[src];
}
__global__ void k_slab_append(SlabMoveArgs a, const double *__restrict__ rec, int n_in, int base)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int k = t / (OSPH_NUM_FIELDS + 2), c = t % (OSPH_NUM_FIELDS + 2);
    if (k >= n_in) return;
    double v = rec[(size_t)k * OSPH_WIRE_FULL + c];
    if (c < OSPH_NUM_FIELDS) a.f[c][base + k] = v;
    else if (c == OSPH_NUM_FIELDS) a.label[base + k] = (signed char)v;
    else a.row[base + k] = (int)v;
}

// Small exchanges (the common case: a few dozen migrants per face and step) are committed by ONE CTA: mark, pick
// fillers, move, append and install the all-reduced bounds, instead of two memsets and five launches.
#define SLAB_SMALL 4096
__global__ void __launch_bounds__(1024)
k_slab_commit_small(SlabMoveArgs a, const int *__restrict__ mig_slots, int m, int n_new, int *__restrict__ tail_flag,
                    int *__restrict__ holes, int *__restrict__ fillers, const double *__restrict__ rec_l, int n_in_l,
                    const double *__restrict__ rec_r, int n_in_r, StepScalars *sc, int set_bounds, double xmin, double nxmax, double ymin, double nymax, double hmin,
                    double nhmax)
{
    __shared__ int nh, nf;
    const int t = threadIdx.x, nt = blockDim.x;
    if (t == 0) { nh = 0; nf = 0; }
    for (int k = t; k < m; k += nt) tail_flag[k] = 0;
    __syncthreads();
    for (int k = t; k < m; k += nt) {
        int h = mig_slots[k];
        if (h >= n_new) tail_flag[h - n_new] = 1; else holes[atomicAdd(&nh, 1)] = h;
    }
    __syncthreads();
    for (int k = t; k < m; k += nt) if (!tail_flag[k]) fillers[atomicAdd(&nf, 1)] = n_new + k;
    __syncthreads();
    const int W = OSPH_NUM_FIELDS + 2, pairs = nh;
    for (int u = t; u < pairs * W; u += nt) {
        int k = u / W, c = u % W, dst = holes[k], src = fillers[k];
        if (c < OSPH_NUM_FIELDS) a.f[c][dst] = a.f[c][src];
        else if (c == OSPH_NUM_FIELDS) a.label[dst] = a.label[src];
        else a.row[dst] = a.row[src];
    }
    __syncthreads();
    const int n_in = n_in_l + n_in_r;
    for (int u = t; u < n_in * W; u += nt) {
        int k = u / W, c = u % W;
        double v = k < n_in_l ? rec_l[(size_t)k * OSPH_WIRE_FULL + c] : rec_r[(size_t)(k - n_in_l) * OSPH_WIRE_FULL + c];
        if (c < OSPH_NUM_FIELDS) a.f[c][n_new + k] = v;
        else if (c == OSPH_NUM_FIELDS) a.label[n_new + k] = (signed char)v;
        else a.row[n_new + k] = (int)v;
    }
    if (t == 0 && set_bounds) {
        sc->xmin = enc_f64(xmin); sc->xmax = enc_f64(-nxmax); sc->ymin = enc_f64(ymin); sc->ymax = enc_f64(-nymax);
        sc->hmin_all = enc_f64(hmin); sc->hmax_all = enc_f64(-nhmax);
    }
}

// uniform_c > 0: the step before deferred its corrector, nothing reduced c_max; it is the uniform co (k_timestep, fused == 2)
__global__ void k_slab_dt_local(const StepScalars *sc, double *out, double uniform_c)
{
    out[0] = dec_f64(sc->hmin_fluid); out[1] = uniform_c > 0.0 ? -uniform_c : -dec_f64(sc->cmax_fluid);
    out[2] = -dec_f64(sc->a2max_fluid);
}
__global__ void k_slab_ids(const int *__restrict__ row, const signed char *__restrict__ label, int n,
                           int *__restrict__ ids, signed char *__restrict__ lab)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { ids[i] = row[i]; if (lab) lab[i] = label[i]; }
}

static SlabMoveArgs move_args(osph_ctx *ctx)
{
    SlabMoveArgs a;
    for (int k = 0; k < OSPH_NUM_FIELDS; k++) a.f[k] = ctx->f[k];
    a.label = ctx->label; a.row = ctx->d_row;
    return a;
}

#define CHECK_CTX()                                                        \
    if (!ctx) return OSPH_E_INVALID;                                       \
    OSPH_CUDA(cudaSetDevice(ctx->device))

extern "C" int osph_slab_configure(osph_ctx *ctx, double x_lo, double x_hi, void *d_ghost, int64_t ghost_capacity)
{
    CHECK_CTX();
    if (!(x_lo < x_hi) || !d_ghost || ghost_capacity <= 0) { ctx->err = "osph_slab_configure: bad arguments"; return OSPH_E_INVALID; }
    ctx->slab = true; ctx->x_lo = x_lo; ctx->x_hi = x_hi;
    ctx->slab_cadence_force = 0;             // a sequencer with a cadence sets it anew at every step
    ctx->d_ghost = (double *)d_ghost; ctx->ghost_cap = ghost_capacity; ctx->n_ghost = 0;
    if (!ctx->d_slab_counters) {
        OSPH_CUDA(cudaMalloc(&ctx->d_slab_counters, sizeof(int) * 16));
    }
    return 0;
}

static int slab_pack_impl(osph_ctx *ctx, double halo_width, int mode, void *d_mig_left, void *d_mig_right, int64_t mig_cap,
                          void *d_halo_left, void *d_halo_right, int64_t halo_cap, int *d_idx_l, int *d_idx_r, double *d_meta)
{
    if (!ctx->slab || ctx->n <= 0) { ctx->err = "osph_slab_pack: context is not in slab mode"; return OSPH_E_INVALID; }
    if (mig_cap > ctx->slab_list_cap) {
        cudaFree(ctx->d_mig_slots); cudaFree(ctx->d_tail_flag); cudaFree(ctx->d_holes); cudaFree(ctx->d_fillers);
        OSPH_CUDA(cudaMalloc(&ctx->d_mig_slots, sizeof(int) * mig_cap * 2));
        OSPH_CUDA(cudaMalloc(&ctx->d_tail_flag, sizeof(int) * mig_cap * 2));
        OSPH_CUDA(cudaMalloc(&ctx->d_holes, sizeof(int) * mig_cap * 2));
        OSPH_CUDA(cudaMalloc(&ctx->d_fillers, sizeof(int) * mig_cap * 2));
        ctx->slab_list_cap = mig_cap;
    }
    OSPH_CUDA(cudaMemsetAsync(ctx->d_slab_counters, 0, sizeof(int) * 16, ctx->stream));
    SlabPackArgs a;
    a.n = (int)ctx->n; a.label = ctx->label; a.row = ctx->d_row;
    for (int k = 0; k < OSPH_NUM_FIELDS; k++) a.f[k] = ctx->f[k];
    a.x_lo = ctx->x_lo; a.x_hi = ctx->x_hi;
    // summation density: the ghosts an owned particle can see need their own kernel support inside the halo (pair.cu)
    a.width = ctx->cfg.summation_density ? 2.0 * halo_width : halo_width;
    a.mig_left = (double *)d_mig_left; a.mig_right = (double *)d_mig_right;
    a.halo_left = (double *)d_halo_left; a.halo_right = (double *)d_halo_right; a.ghost = ctx->d_ghost;
    a.mig_cap = (int)mig_cap; a.halo_cap = (int)halo_cap; a.ghost_cap = (int)ctx->ghost_cap;
    a.counters = ctx->d_slab_counters; a.mig_slots = ctx->d_mig_slots;
    a.mode = mode; a.halo_idx_l = d_idx_l; a.halo_idx_r = d_idx_r;
    k_slab_pack<<<div_up(ctx->n, 256), 256, 0, ctx->stream>>>(a); OSPH_LAUNCH_CHECK();
    k_slab_meta<<<1, 1, 0, ctx->stream>>>(ctx->d_slab_counters, ctx->d_sc, d_meta); OSPH_LAUNCH_CHECK();
    return 0;
}

extern "C" int osph_slab_pack(osph_ctx *ctx, double halo_width, void *d_mig_left, void *d_mig_right, int64_t mig_cap,
                              void *d_halo_left, void *d_halo_right, int64_t halo_cap, double *d_meta)
{
    CHECK_CTX();
    return slab_pack_impl(ctx, halo_width, 3, d_mig_left, d_mig_right, mig_cap, d_halo_left, d_halo_right, halo_cap, nullptr, nullptr, d_meta);
}

// Slab cadence (slab_p2p.cu).  mode 1: migrants only; mode 2: halos only, the storage slot behind every record is kept in
// d_idx_l / d_idx_r (capacity halo_cap each).
int osph_slab_pack_mode(osph_ctx *ctx, double halo_width, int mode, void *d_mig_left, void *d_mig_right, int64_t mig_cap,
                        void *d_halo_left, void *d_halo_right, int64_t halo_cap, int *d_idx_l, int *d_idx_r, double *d_meta)
{
    return slab_pack_impl(ctx, halo_width, mode, d_mig_left, d_mig_right, mig_cap, d_halo_left, d_halo_right, halo_cap, d_idx_l, d_idx_r, d_meta);
}

// Slab cadence, reuse step: refresh the records of the frozen halo lists in the neighbours' regions, then the meta row.
// d_meta == nullptr: the caller's mailbox kernel forms the meta row itself
int osph_slab_repack(osph_ctx *ctx, const int *d_idx_l, int64_t n_l, void *d_halo_left, const int *d_idx_r, int64_t n_r,
                     void *d_halo_right, double *d_meta)
{
    SlabPackArgs a;
    a.n = (int)ctx->n; a.label = ctx->label; a.row = ctx->d_row;
    for (int k = 0; k < OSPH_NUM_FIELDS; k++) a.f[k] = ctx->f[k];
    a.halo_left = (double *)d_halo_left; a.halo_right = (double *)d_halo_right;
    a.mig_left = a.mig_right = a.ghost = nullptr; a.counters = nullptr; a.mig_slots = nullptr; a.halo_idx_l = a.halo_idx_r = nullptr;
    a.mig_cap = a.halo_cap = a.ghost_cap = 0; a.mode = 0; a.x_lo = a.x_hi = a.width = 0.0;
    if (n_l + n_r > 0) {
        k_slab_repack<<<div_up(n_l + n_r, 256), 256, 0, ctx->stream>>>(a, d_idx_l, (int)n_l, d_idx_r, (int)n_r); OSPH_LAUNCH_CHECK();
    }
    if (d_meta) {
        OSPH_CUDA(cudaMemsetAsync(ctx->d_slab_counters, 0, sizeof(int) * 16, ctx->stream));
        k_slab_meta<<<1, 1, 0, ctx->stream>>>(ctx->d_slab_counters, ctx->d_sc, d_meta); OSPH_LAUNCH_CHECK();
    }
    return 0;
}

int osph_slab_install(osph_ctx *ctx, const double *d_all_meta, int world)
{
    k_slab_install<<<1, 1, 0, ctx->stream>>>(d_all_meta, world, ctx->d_sc); OSPH_LAUNCH_CHECK();
    ctx->prepared = true;
    return 0;
}

int osph_slab_commit_impl(osph_ctx *ctx, int64_t n_mig_out, const double *d_in_l, int64_t n_in_l, const double *d_in_r,
                          int64_t n_in_r, GhostMap gmap, int64_t n_ghost, const double global_bounds[6])
{
    if (!ctx->slab) { ctx->err = "osph_slab_commit: context is not in slab mode"; return OSPH_E_INVALID; }
    const int64_t n_mig_in = n_in_l + n_in_r;
    int64_t n_new = ctx->n - n_mig_out;
    if (n_new < 0 || n_new + n_mig_in + n_ghost > ctx->cap) {
        ctx->err = "osph_slab_commit: particle capacity exceeded (osph_reserve a larger capacity)"; return OSPH_E_CAPACITY;
    }
    int m = (int)n_mig_out;
    if (m + n_mig_in <= SLAB_SMALL) {
        const double z[6] = {0, 0, 0, 0, 0, 0};
        const double *b = global_bounds ? global_bounds : z;
        k_slab_commit_small<<<1, 1024, 0, ctx->stream>>>(move_args(ctx), ctx->d_mig_slots, m, (int)n_new, ctx->d_tail_flag,
                                                         ctx->d_holes, ctx->d_fillers, d_in_l, (int)n_in_l, d_in_r, (int)n_in_r,
                                                         ctx->d_sc, global_bounds ? 1 : 0, b[0], b[1], b[2], b[3], b[4], b[5]);
        OSPH_LAUNCH_CHECK();
    } else {
        if (m > 0) {
            int *nh = ctx->d_slab_counters + 8, *nf = ctx->d_slab_counters + 9;
            OSPH_CUDA(cudaMemsetAsync(nh, 0, sizeof(int) * 2, ctx->stream));
            OSPH_CUDA(cudaMemsetAsync(ctx->d_tail_flag, 0, sizeof(int) * m, ctx->stream));
            k_slab_mark<<<div_up(m, 256), 256, 0, ctx->stream>>>(ctx->d_mig_slots, m, (int)n_new, ctx->d_tail_flag, ctx->d_holes, nh);
            OSPH_LAUNCH_CHECK();
            k_slab_fillers<<<div_up(m, 256), 256, 0, ctx->stream>>>(ctx->d_tail_flag, m, (int)n_new, ctx->d_fillers, nf);
            OSPH_LAUNCH_CHECK();
            k_slab_move<<<div_up((int64_t)m * (OSPH_NUM_FIELDS + 2), 256), 256, 0, ctx->stream>>>(move_args(ctx), ctx->d_holes,
                                                                                                ctx->d_fillers, nh, m);
            OSPH_LAUNCH_CHECK();
        }
        if (n_in_l > 0) {
            k_slab_append<<<div_up(n_in_l * (OSPH_NUM_FIELDS + 2), 256), 256, 0, ctx->stream>>>(move_args(ctx), d_in_l, (int)n_in_l, (int)n_new);
            OSPH_LAUNCH_CHECK();
        }
        if (n_in_r > 0) {
            k_slab_append<<<div_up(n_in_r * (OSPH_NUM_FIELDS + 2), 256), 256, 0, ctx->stream>>>(move_args(ctx), d_in_r, (int)n_in_r,
                                                                                              (int)(n_new + n_in_l));
            OSPH_LAUNCH_CHECK();
        }
        if (global_bounds) {
            k_slab_set_bounds<<<1, 1, 0, ctx->stream>>>(ctx->d_sc, global_bounds[0], global_bounds[1], global_bounds[2],
                                                        global_bounds[3], global_bounds[4], global_bounds[5]);
            OSPH_LAUNCH_CHECK();
        }
    }
    ctx->n = n_new + n_mig_in;
    ctx->n_ghost = n_ghost;
    ctx->gmap = gmap;
    ctx->prepared = true;
    return 0;
}

extern "C" int osph_slab_commit(osph_ctx *ctx, int64_t n_mig_out, const void *d_mig_in, int64_t n_mig_in,
                                int64_t n_ghost, const double global_bounds[6])
{
    CHECK_CTX();
    if (n_ghost > ctx->ghost_cap) { ctx->err = "osph_slab_commit: ghost capacity exceeded"; return OSPH_E_CAPACITY; }
    GhostMap gm = {(int)n_ghost, 0, 0, 0};                    // NCCL receives are packed back to back
    return osph_slab_commit_impl(ctx, n_mig_out, (const double *)d_mig_in, n_mig_in, nullptr, 0, gm, n_ghost, global_bounds);
}

// Several slab steps in one call of a sequencer: as in osph_step, the corrector of step k is applied by the predictor
// pass of step k+1 (k_prepare<PEC, true, FUSED>), min h is reduced by that pass and max |a|^2 by the pair kernel, so the
// separate corrector pass exists only at the end of the call.  Migrants carry x0..rho0 and the rates in their 21-double
// records, so a particle that changed owner in between is corrected by its new owner with the same operands.
// OSPH_SLAB_FUSED=0 in the environment keeps every step in the plain form.
extern "C" int osph_slab_step_plan(osph_ctx *ctx, int32_t step, int32_t nsteps)
{
    if (!ctx || step < 0 || step >= nsteps) return OSPH_E_INVALID;
    static const bool enabled = [] { const char *e = getenv("OSPH_SLAB_FUSED"); return !(e && e[0] == '0'); }();
    const bool fuse = enabled && nsteps > 1 && ctx->cfg.integrator == OSPH_INTEGRATOR_PEC && !ctx->cfg.summation_density;
    if (ctx->slab_defer && step == 0) { ctx->err = "osph_slab_step_plan: the previous call ended with a deferred corrector"; return OSPH_E_INVALID; }
    ctx->slab_fused = fuse ? (step == 0 ? 1 : 2) : 0;
    ctx->slab_last = step == nsteps - 1;
    return 0;
}

extern "C" int osph_slab_dt_local(osph_ctx *ctx, double *d_out3)
{
    CHECK_CTX();
    if (ctx->slab_defer != (ctx->slab_fused == 2)) { ctx->err = "osph_slab_dt_local: step plan and deferred corrector disagree"; return OSPH_E_INVALID; }
    if (!ctx->reductions_valid) {
        int rc = osph_launch_correct(ctx, false, 0.0, 0.0, false);
        if (rc) return rc;
        ctx->reductions_valid = true;
    }
    k_slab_dt_local<<<1, 1, 0, ctx->stream>>>(ctx->d_sc, d_out3, ctx->slab_defer ? ctx->cfg.co : 0.0); OSPH_LAUNCH_CHECK();
    return 0;
}

// The same two calls for a sequencer whose dt mailbox computes the local triple and runs the time-step kernel itself
// (slab_p2p.cu: k_mbox_allgather, MboxExtra): the host-side checks of osph_slab_dt_local without its kernel ...
int osph_slab_dt_prepare(osph_ctx *ctx, double *uniform_c)
{
    if (ctx->slab_defer != (ctx->slab_fused == 2)) { ctx->err = "osph_slab_dt_local: step plan and deferred corrector disagree"; return OSPH_E_INVALID; }
    if (!ctx->reductions_valid) {
        int rc = osph_launch_correct(ctx, false, 0.0, 0.0, false);
        if (rc) return rc;
        ctx->reductions_valid = true;
    }
    *uniform_c = ctx->slab_defer ? ctx->cfg.co : 0.0;
    return 0;
}
// ... and osph_slab_step_begin without its k_timestep launch
int osph_slab_step_begin_after_dt(osph_ctx *ctx, double damping)
{
    if (!ctx->slab) { ctx->err = "osph_slab_step_begin: context is not in slab mode"; return OSPH_E_INVALID; }
    int rc;
    if ((rc = osph_launch_prepare(ctx, true, 0.0, damping, true, true, ctx->slab_fused))) return rc;
    ctx->slab_defer = false;
    ctx->neighbours_valid = false; ctx->reductions_valid = false;
    return 0;
}

extern "C" int osph_slab_step_begin(osph_ctx *ctx, const double *d_dt_reduced3, double fixed_dt, double damping)
{
    CHECK_CTX();
    if (!ctx->slab) { ctx->err = "osph_slab_step_begin: context is not in slab mode"; return OSPH_E_INVALID; }
    int rc;
    const int fused = ctx->slab_fused;
    if ((rc = osph_launch_timestep(ctx, fixed_dt > 0 ? fixed_dt : -1.0, true, true, d_dt_reduced3, fused))) return rc;
    if ((rc = osph_launch_prepare(ctx, true, 0.0, damping, true, true, fused))) return rc;      // fused == 2: corrector of the step before first
    ctx->slab_defer = false;
    ctx->neighbours_valid = false; ctx->reductions_valid = false;
    return 0;
}

int osph_size_cell_table(osph_ctx *ctx);     // api.cu

// Inside a fused call (osph_slab_step_plan) every step but the last leaves its corrector to the next predictor pass.
extern "C" int osph_slab_step_end(osph_ctx *ctx, double damping)
{
    CHECK_CTX();
    if (!ctx->slab) { ctx->err = "osph_slab_step_end: context is not in slab mode"; return OSPH_E_INVALID; }
    int rc;
    const bool fuse = ctx->slab_fused != 0, defer = fuse && !ctx->slab_last;
    if ((rc = osph_size_cell_table(ctx))) return rc;
    if ((rc = osph_launch_build(ctx, !fuse))) return rc;
    ctx->pair_reduce_a2 = fuse;
    rc = osph_launch_pair(ctx);
    ctx->pair_reduce_a2 = false;
    if (rc) return rc;
    ctx->c_uniform = true;
    if (!defer && (rc = osph_launch_correct(ctx, true, 0.0, damping, true, true))) return rc;
    ctx->slab_defer = defer;
    ctx->slab_fused = 0; ctx->slab_last = true;
    ctx->prepared = false; ctx->neighbours_valid = false; ctx->reductions_valid = true;
    ctx->step_counter++;
    return 0;
}

extern "C" int osph_slab_export(osph_ctx *ctx, int32_t *d_ids, int8_t *d_label, int32_t nfields, const int32_t *fields,
                                double *const *d_cols)
{
    CHECK_CTX();
    if (ctx->n <= 0) return 0;
    k_slab_ids<<<div_up(ctx->n, 256), 256, 0, ctx->stream>>>(ctx->d_row, ctx->label, (int)ctx->n, d_ids, (signed char *)d_label);
    OSPH_LAUNCH_CHECK();
    for (int k = 0; k < nfields; k++) {
        if (fields[k] < 0 || fields[k] >= OSPH_NUM_FIELDS) { ctx->err = "osph_slab_export: bad field"; return OSPH_E_INVALID; }
        if (fields[k] == OSPH_F_C && ctx->c_uniform) {
            int rc = osph_launch_fill(ctx, d_cols[k], ctx->cfg.co);
            if (rc) return rc;
        } else {
            OSPH_CUDA(cudaMemcpyAsync(d_cols[k], ctx->f[fields[k]], sizeof(double) * ctx->n, cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int osph_set_row_ids(osph_ctx *ctx, const int32_t *ids, int64_t n)
{
    CHECK_CTX();
    if (n != ctx->n || !ids) { ctx->err = "osph_set_row_ids: need one id per active particle"; return OSPH_E_INVALID; }
    OSPH_CUDA(cudaMemcpyAsync(ctx->d_row, ids, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int osph_reserve(osph_ctx *ctx, int64_t particle_capacity)
{
    CHECK_CTX();
    if (particle_capacity < 0) return OSPH_E_INVALID;
    ctx->reserve = particle_capacity;
    return 0;
}
