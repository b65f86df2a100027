// slab_p2p.cu -- the slab exchange over NVLink peer memory, without NCCL in the data path.
//
// Every rank owns one "window" in HBM, exported to the other ranks of the box with CUDA IPC.  Per step:
//   * the pack kernel (slab.cu: k_slab_pack) writes this rank's migrants and halo particles DIRECTLY into the
//     neighbours' windows -- packing and sending are one kernel, the records cross NVLink as they are produced;
//   * two tiny "mailbox" all-gathers (k_mbox_allgather: every rank stores its payload into every peer's window,
//     publishes a sequence number behind a system-scope fence, then spins on its own window) replace
//     ncclAllReduce(dt) and ncclAllGather(counts + grid bounds).  The second one doubles as the "data has landed"
//     signal: it is enqueued after the pack kernel on the same stream.
// Received halos are used in place: k_keys / k_gather read ghost records from the window's regions (GhostMap).
// A region is single-buffered: a neighbour can only start writing step k+1 after the dt mailbox of step k+1, to
// which this rank contributes only after its step k has completed (stream order), so readers and writers of a
// region never overlap.
// NCCL (slab_nccl.cu) stays as the portable sequencer; this one needs all ranks on one NVLink/NVSwitch box.
#include <stdio.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "step.cuh"

extern "C" {
int osph_slab_configure(osph_ctx *ctx, double x_lo, double x_hi, void *d_ghost, int64_t ghost_capacity);
int osph_slab_dt_local(osph_ctx *ctx, double *d_out3);
int osph_slab_step_begin(osph_ctx *ctx, const double *d_dt_reduced3, double fixed_dt, double damping);
int osph_slab_pack(osph_ctx *ctx, double halo_width, void *d_mig_left, void *d_mig_right, int64_t mig_cap,
                   void *d_halo_left, void *d_halo_right, int64_t halo_cap, double *d_meta);
int osph_slab_step_end(osph_ctx *ctx, double damping);
}
// slab.cu, slab cadence
int osph_slab_pack_mode(osph_ctx *ctx, double halo_width, int mode, void *d_mig_left, void *d_mig_right, int64_t mig_cap,
                        void *d_halo_left, void *d_halo_right, int64_t halo_cap, int *d_idx_l, int *d_idx_r, double *d_meta);
int osph_slab_repack(osph_ctx *ctx, const int *d_idx_l, int64_t n_l, void *d_halo_left, const int *d_idx_r, int64_t n_r,
                     void *d_halo_right, double *d_meta);
int osph_slab_install(osph_ctx *ctx, const double *d_all_meta, int world);
int osph_slab_dt_prepare(osph_ctx *ctx, double *uniform_c);
int osph_slab_step_begin_after_dt(osph_ctx *ctx, double damping);

#define P2P_MAX_WORLD 16
#define MBOX_STRIDE 16            // doubles per (slot, sender) cell of a mailbox
#define P2P_HDR_ABORT 8           // window header word (as u64): a rank that leaves the protocol with an error sets it in every peer
#define P2P_ST_ABORT 1u           // status bits of k_mbox_allgather: a peer aborted / a peer did not answer within the spin limit
#define P2P_ST_TIMEOUT 2u

struct osph_slab_p2p {
    int rank = 0, world = 1;
    double x_lo = 0, x_hi = 0, r0 = 0, hmax = 0, hmin = 0;       // hmin: smallest h over ALL active particles (reference cell size)
    int kernel = 0;
    int64_t mig_cap = 0, halo_cap = 0;
    double *win = nullptr;                         // this rank's window
    double *peer[P2P_MAX_WORLD] = {nullptr};       // every rank's window as mapped here (peer[rank] == win)
    double **d_peer = nullptr;                     // the same table on the device
    // window layout, offsets in doubles
    size_t off_dt = 0, off_dt_flag = 0, off_meta = 0, off_meta_flag = 0;
    size_t off_ghost = 0, off_halo_from_left = 0, off_halo_from_right = 0, off_mig_from_left = 0, off_mig_from_right = 0;
    size_t win_doubles = 0;
    double *d_meta = nullptr, *d_all_meta = nullptr, *d_dt3 = nullptr, *d_all_dt = nullptr, *h_all_meta = nullptr;
    unsigned long long seq = 0;
    // OSPH_SLAB_PROFILE=1: CUDA events between the phases of a step (GPU timeline incl. idle gaps), summed per phase
    bool profile = false;
    cudaEvent_t pev[8] = {nullptr};
    double pms[7] = {0, 0, 0, 0, 0, 0, 0}, phost_ms = 0;
    long long psteps = 0;
    long long spin_limit = 0;                      // clock64 ticks a mailbox waits for one peer before it gives up
    bool aborted = false;
    // ---- slab cadence: the ranks sort TOGETHER every few steps; in between the halo is the same particles in the same
    // record slots (frozen lists), migration waits for the next sort, and the host waits for nothing ----
    bool merge = true;                 // OSPH_SLAB_MERGE: scalar kernels of the step folded into the mailbox kernels
    bool cadence = false;              // OSPH_SLAB_CADENCE
    bool lists_valid = false;          // frozen halo lists describe the resident particles
    int *d_idx_l = nullptr, *d_idx_r = nullptr;      // storage slots behind my halo records towards the left / right neighbour
    int64_t halo_n_l = 0, halo_n_r = 0;
    double skin = 0.1, gs = 0.0;       // skin (fraction of the pair radius) and cell size of the current binning
    double D_known = 0.0, dstep = -1.0;              // largest displacement since the sort (all ranks, lagged one step) and per step
    int steps_since_sort = 0;
    unsigned long long pending_seq = 0;              // meta mailbox of a reuse step whose host copy has not been read yet
    int64_t sorts = 0, reuses = 0;
    int64_t last_counts[8] = {0};
    int64_t steps = 0;
};

// One CTA, one warp per peer.  Warp r: store my payload into rank r's window, fence, publish `seq`; then wait until
// rank r's payload for `seq` is in MY window and copy it out.  reduce_min != 0 appends the element-wise minimum.
// The wait is bounded: it ends when the peer's sequence number arrives, when some rank has set the abort word of my
// window (it left the protocol with an error), or after `spin_limit` clock ticks; the last two set a bit in *status,
// which the host reads at its one synchronisation point per step -- a rank-local failure then raises on every rank
// instead of leaving the others inside this kernel for ever.
// What a mailbox kernel does besides the all-gather (the slab step is a chain of latency-bound launches: every scalar kernel
// folded into a mailbox is a launch less on that chain).
struct MboxExtra {
    // payload computed here instead of by a kernel before: 1 = the dt triple {h_min, -c_max, -a2_max} of this rank
    // (k_slab_dt_local), 2 = the meta row with zero counts (k_slab_meta of a reuse step)
    int payload_from_sc;
    double uniform_c;
    StepScalars *sc;
    // tail, run by thread 0 after the gather: 1 = k_timestep with the all-reduced triple, 2 = k_slab_install (global bounds,
    // h extrema and displacement of a reuse step)
    int tail;
    double gamma_c, gamma_f, fixed_dt, co;
    double *dt_log;
    long long dt_log_cap;
    int ts_fused;
};

__global__ void k_mbox_allgather(double *const *__restrict__ peers, int me, int world, size_t off_data, size_t off_flag,
                                 int slot, const double *__restrict__ payload, int n, unsigned long long seq,
                                 double *__restrict__ out_all, int reduce_min, long long spin_limit, unsigned int *status,
                                 double *host_all, volatile unsigned long long *host_flag, MboxExtra x)
{
    const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (r < world) {
        double *dst = peers[r] + off_data + (size_t)(slot * world + me) * MBOX_STRIDE;
        volatile unsigned long long *dflag = reinterpret_cast<volatile unsigned long long *>(peers[r] + off_flag) + slot * world + me;
        if (lane < n) {
            double v;
            if (x.payload_from_sc == 1) {
                const StepScalars *sc = x.sc;
                v = lane == 0 ? dec_f64(sc->hmin_fluid) : (lane == 1 ? (x.uniform_c > 0.0 ? -x.uniform_c : -dec_f64(sc->cmax_fluid)) : -dec_f64(sc->a2max_fluid));
            } else if (x.payload_from_sc == 2) {
                const StepScalars *sc = x.sc;
                v = 0.0;
                if (lane == 4) v = dec_f64(sc->xmin); else if (lane == 5) v = -dec_f64(sc->xmax);
                else if (lane == 6) v = dec_f64(sc->ymin); else if (lane == 7) v = -dec_f64(sc->ymax);
                else if (lane == 8) v = dec_f64(sc->hmin_all); else if (lane == 9) v = -dec_f64(sc->hmax_all);
                else if (lane == 11) v = -dec_f64(sc->disp2max);
            } else v = payload[lane];
            dst[lane] = v;
        }
        __threadfence_system();
        __syncwarp();
        if (lane == 0) *dflag = seq;
        const volatile double *src = peers[me] + off_data + (size_t)(slot * world + r) * MBOX_STRIDE;
        volatile unsigned long long *sflag = reinterpret_cast<volatile unsigned long long *>(peers[me] + off_flag) + slot * world + r;
        volatile unsigned long long *abort_word = reinterpret_cast<volatile unsigned long long *>(peers[me]) + P2P_HDR_ABORT;
        if (lane == 0) {
            const long long t0 = clock64();
            while (*sflag != seq) {
                if (*abort_word != 0ull) { atomicOr(status, P2P_ST_ABORT); break; }
                if (clock64() - t0 > spin_limit) { atomicOr(status, P2P_ST_TIMEOUT); break; }
                __nanosleep(64);
            }
        }
        __syncwarp();
        __threadfence_system();
        if (lane < n) out_all[r * n + lane] = src[lane];
    }
    if (reduce_min) {
        __syncthreads();
        if (threadIdx.x < n) {
            double m = out_all[threadIdx.x];
            for (int q = 1; q < world; q++) m = fmin(m, out_all[q * n + threadIdx.x]);
            out_all[world * n + threadIdx.x] = m;
        }
    }
    if (x.tail) {
        __syncthreads();
        if (threadIdx.x == 0) {
            if (x.tail == 1) {
                // k_timestep of the step, with the all-reduced triple (as osph_slab_step_begin would launch it)
                timestep_body(x.sc, x.gamma_c, x.gamma_f, x.fixed_dt, x.dt_log, x.dt_log_cap, 1, out_all + world * n, x.ts_fused, x.co);
            } else {
                // k_slab_install: global bounds, h extrema, displacement
                StepScalars *sc = x.sc;
                double b[6], d = out_all[11];
                for (int k = 0; k < 6; k++) b[k] = out_all[4 + k];
                for (int q = 1; q < world; q++) {
                    for (int k = 0; k < 6; k++) b[k] = fmin(b[k], out_all[n * q + 4 + k]);
                    d = fmin(d, out_all[n * q + 11]);
                }
                sc->xmin = enc_f64(b[0]); sc->xmax = enc_f64(-b[1]); sc->ymin = enc_f64(b[2]); sc->ymax = enc_f64(-b[3]);
                sc->hmin_all = enc_f64(b[4]); sc->hmax_all = enc_f64(-b[5]);
                sc->disp2max = enc_f64(-d);
            }
        }
    }
    if (host_all) {
        // The host needs this gather (message sizes of the step): it goes straight into pinned host memory, followed by the
        // sequence number the host spins on -- no cudaMemcpy + stream synchronisation on the critical path of the step.
        __syncthreads();
        for (int k = threadIdx.x; k < world * n; k += blockDim.x) host_all[k] = out_all[k];
        if (threadIdx.x == 0) reinterpret_cast<unsigned int *>(host_all + world * n)[0] = *status;
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) *host_flag = seq;
    }
}

// A rank that returns an error from the step loop tells its peers first: one 8-byte store into every window header.
static void p2p_abort_peers(osph_ctx *ctx, osph_slab_p2p *s)
{
    if (s->aborted) return;
    s->aborted = true;
    const unsigned long long one = 1ull;
    for (int r = 0; r < s->world; r++)
        if (s->peer[r]) cudaMemcpy(reinterpret_cast<unsigned long long *>(s->peer[r]) + P2P_HDR_ABORT, &one, sizeof(one), cudaMemcpyHostToDevice);
    (void)ctx;
}

extern "C" int osph_slab_p2p_create(osph_ctx *ctx, int rank, int world, double x_lo, double x_hi, double r0, double hmax,
                                    int64_t mig_cap, int64_t halo_cap, osph_slab_p2p **out, char handle_out[64])
{
    if (!ctx || !out || !handle_out || world < 1 || world > P2P_MAX_WORLD || rank < 0 || rank >= world) return OSPH_E_INVALID;
    OSPH_CUDA(cudaSetDevice(ctx->device));
    osph_slab_p2p *s = new osph_slab_p2p();
    s->rank = rank; s->world = world; s->x_lo = x_lo; s->x_hi = x_hi; s->r0 = r0; s->hmax = hmax;
    s->kernel = ctx->cfg.kernel; s->mig_cap = mig_cap; s->halo_cap = halo_cap;
    size_t o = 0;
    auto take = [&](size_t doubles) { size_t at = o; o += (doubles + 15) / 16 * 16; return at; };
    const size_t off_header = take(16);                                           // {world, mig_cap, halo_cap}: layout check
    s->off_dt = take((size_t)2 * world * MBOX_STRIDE); s->off_dt_flag = take((size_t)2 * world);
    s->off_meta = take((size_t)2 * world * MBOX_STRIDE); s->off_meta_flag = take((size_t)2 * world);
    s->off_ghost = take((size_t)2 * mig_cap * OSPH_WIRE_HALO);                   // this rank's own migrants, kept as ghosts
    s->off_halo_from_left = take((size_t)halo_cap * OSPH_WIRE_HALO);
    s->off_halo_from_right = take((size_t)halo_cap * OSPH_WIRE_HALO);
    s->off_mig_from_left = take((size_t)mig_cap * OSPH_WIRE_FULL);
    s->off_mig_from_right = take((size_t)mig_cap * OSPH_WIRE_FULL);
    s->win_doubles = o;
    OSPH_CUDA(cudaMalloc(&s->win, sizeof(double) * o));
    OSPH_CUDA(cudaMemset(s->win, 0, sizeof(double) * o));
    const double header[3] = {(double)world, (double)mig_cap, (double)halo_cap};
    OSPH_CUDA(cudaMemcpy(s->win + off_header, header, sizeof(header), cudaMemcpyHostToDevice));
    OSPH_CUDA(cudaMalloc(&s->d_peer, sizeof(double *) * world));
    OSPH_CUDA(cudaMalloc(&s->d_meta, sizeof(double) * MBOX_STRIDE));
    OSPH_CUDA(cudaMalloc(&s->d_all_meta, sizeof(double) * MBOX_STRIDE * (world + 2)));    // + min row + the status word
    OSPH_CUDA(cudaMemset(s->d_all_meta, 0, sizeof(double) * MBOX_STRIDE * (world + 2)));
    {
        // bounded mailbox waits: OSPH_P2P_SPIN_SECONDS (default 30 s; clock64 ticks at about 2 GHz)
        const char *e = getenv("OSPH_P2P_SPIN_SECONDS");
        double sec = e ? atof(e) : 30.0;
        if (!(sec > 0.0)) sec = 30.0;
        s->spin_limit = (long long)(sec * 2.0e9);
    }
    OSPH_CUDA(cudaMalloc(&s->d_dt3, sizeof(double) * 4));
    OSPH_CUDA(cudaMalloc(&s->d_all_dt, sizeof(double) * 4 * (world + 1)));
    OSPH_CUDA(cudaMallocHost(&s->h_all_meta, sizeof(double) * (12 * world + 2)));      // + status word + sequence flag (device-written)
    memset(s->h_all_meta, 0, sizeof(double) * (12 * world + 2));
    {
        // OSPH_SLAB_CADENCE=0 restores the exchange that sorts, migrates and re-packs the halo at every step
        const char *e = getenv("OSPH_SLAB_CADENCE");
        s->cadence = !(e && e[0] == '0') && !ctx->cfg.summation_density;
        const char *m = getenv("OSPH_SLAB_MERGE");
        s->merge = !(m && m[0] == '0');
        if (s->cadence) {
            OSPH_CUDA(cudaMalloc(&s->d_idx_l, sizeof(int) * (size_t)halo_cap));
            OSPH_CUDA(cudaMalloc(&s->d_idx_r, sizeof(int) * (size_t)halo_cap));
        }
    }
    {
        const char *e = getenv("OSPH_SLAB_PROFILE");
        s->profile = e && e[0] == '1';
        if (s->profile) for (int k = 0; k < 8; k++) OSPH_CUDA(cudaEventCreate(&s->pev[k]));
    }
    cudaIpcMemHandle_t h;
    OSPH_CUDA(cudaIpcGetMemHandle(&h, s->win));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle_out, &h, 64);
    *out = s;
    return 0;
}

// all_handles: world x 64 bytes, rank-major (all-gathered by the caller)
extern "C" int osph_slab_p2p_connect(osph_ctx *ctx, osph_slab_p2p *s, const char *all_handles)
{
    if (!ctx || !s || !all_handles) return OSPH_E_INVALID;
    OSPH_CUDA(cudaSetDevice(ctx->device));
    for (int r = 0; r < s->world; r++) {
        if (r == s->rank) { s->peer[r] = s->win; continue; }
        cudaIpcMemHandle_t h; memcpy(&h, all_handles + 64 * r, 64);
        void *p = nullptr;
        OSPH_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s->peer[r] = (double *)p;
        // every rank addresses its peers' windows with its own offsets: the layouts must be identical
        double header[3] = {0, 0, 0};
        OSPH_CUDA(cudaMemcpy(header, s->peer[r], sizeof(header), cudaMemcpyDeviceToHost));
        if (header[0] != (double)s->world || header[1] != (double)s->mig_cap || header[2] != (double)s->halo_cap) {
            ctx->err = "osph_slab_p2p_connect: ranks were created with different world size / region capacities";
            return OSPH_E_INVALID;
        }
    }
    OSPH_CUDA(cudaMemcpy(s->d_peer, s->peer, sizeof(double *) * s->world, cudaMemcpyHostToDevice));
    return osph_slab_configure(ctx, s->x_lo, s->x_hi, s->win + s->off_ghost, 2 * s->mig_cap);
}

extern "C" int osph_slab_p2p_destroy(osph_ctx *ctx, osph_slab_p2p *s)
{
    if (!s) return OSPH_E_INVALID;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); ctx->d_ghost = nullptr; ctx->n_ghost = 0; ctx->slab = false; }
    if (s->profile && s->psteps > 0) {
        static const char *names[7] = {"dt reduce + mailbox", "timestep + predictor", "pack + meta mailbox", "copy to host (wait)",
                                       "commit", "neighbour structure + pair", "corrector / tail"};
        fprintf(stderr, "[osph slab profile] rank %d, %lld steps, us per step:", s->rank, s->psteps);
        double tot = 0;
        for (int k = 0; k < 7; k++) { fprintf(stderr, " %s %.1f;", names[k], 1e3 * s->pms[k] / s->psteps); tot += s->pms[k]; }
        fprintf(stderr, " sum %.1f; host wall between sync and last enqueue %.1f\n", 1e3 * tot / s->psteps, 1e3 * s->phost_ms / s->psteps);
        for (int k = 0; k < 8; k++) cudaEventDestroy(s->pev[k]);
    }
    for (int r = 0; r < s->world; r++) if (r != s->rank && s->peer[r]) cudaIpcCloseMemHandle(s->peer[r]);
    cudaFree(s->d_idx_l); cudaFree(s->d_idx_r);
    if (ctx) ctx->slab_cadence_force = 0;
    cudaFree(s->win); cudaFree(s->d_peer); cudaFree(s->d_meta); cudaFree(s->d_all_meta); cudaFree(s->d_dt3); cudaFree(s->d_all_dt);
    cudaFreeHost(s->h_all_meta);
    delete s;
    return 0;
}

extern "C" int osph_slab_p2p_attach(osph_ctx *ctx, osph_slab_p2p *s)
{
    if (!ctx || !s) return OSPH_E_INVALID;
    s->lists_valid = false;                      // the particle set was replaced: the next step sorts
    return osph_slab_configure(ctx, s->x_lo, s->x_hi, s->win + s->off_ghost, 2 * s->mig_cap);
}

extern "C" int osph_slab_p2p_set_bounds(osph_ctx *ctx, osph_slab_p2p *s, double x_lo, double x_hi)
{
    if (!ctx || !s || !(x_lo < x_hi)) return OSPH_E_INVALID;
    s->x_lo = x_lo; s->x_hi = x_hi;
    s->lists_valid = false;                      // new cuts: the next step migrates and sorts
    return osph_slab_configure(ctx, x_lo, x_hi, s->win + s->off_ghost, 2 * s->mig_cap);
}

extern "C" int osph_slab_p2p_run(osph_ctx *ctx, osph_slab_p2p *s, int32_t nsteps, double fixed_dt, double damping)
{
    if (!ctx || !s) return OSPH_E_INVALID;
    OSPH_CUDA(cudaSetDevice(ctx->device));
    const int W = s->world, me = s->rank;
    const int left = me > 0 ? me - 1 : -1, right = me < W - 1 ? me + 1 : -1;
    const double q = s->kernel == OSPH_KERNEL_GAUSSIAN ? 3.0 : 2.0;
    // where my outgoing records land: in the LEFT neighbour's "from right" regions and vice versa; without a
    // neighbour nothing is ever written (the outer slabs are unbounded), any valid address will do
    double *mig_l = left >= 0 ? s->peer[left] + s->off_mig_from_right : s->win + s->off_mig_from_left;
    double *halo_l = left >= 0 ? s->peer[left] + s->off_halo_from_right : s->win + s->off_halo_from_left;
    double *mig_r = right >= 0 ? s->peer[right] + s->off_mig_from_left : s->win + s->off_mig_from_right;
    double *halo_r = right >= 0 ? s->peer[right] + s->off_halo_from_left : s->win + s->off_halo_from_right;
    int rc;
    unsigned int *d_status = reinterpret_cast<unsigned int *>(s->d_all_meta + MBOX_STRIDE * (W + 1));
    // every error return below first sets the abort word of all peers: they leave their mailbox waits and raise too
    auto fail = [&](int code) { p2p_abort_peers(ctx, s); return code; };
    if (s->aborted) { ctx->err = "slab exchange: this group was aborted by an earlier error"; return OSPH_E_INVALID; }
    volatile unsigned long long *const host_flag = reinterpret_cast<volatile unsigned long long *>(s->h_all_meta + 12 * W + 1);
    // wait until the meta mailbox with sequence number `want` has arrived in pinned host memory (a CUDA error or a dead
    // stream ends the wait through cudaStreamQuery), then check the status word it carries
    auto wait_meta = [&](unsigned long long want) -> int {
        const auto t0 = std::chrono::steady_clock::now();
        long long spins = 0;
        while (*host_flag != want) {
            if ((++spins & 0xfff) == 0) {
                const cudaError_t qe = cudaStreamQuery(ctx->stream);
                if (qe != cudaSuccess && qe != cudaErrorNotReady) { ctx->err = std::string("slab exchange: ") + cudaGetErrorString(qe); return OSPH_E_CUDA; }
                if (qe == cudaSuccess && *host_flag != want) { ctx->err = "slab exchange: the mailbox kernel ended without publishing its result"; return OSPH_E_CUDA; }
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 2.0 * (double)s->spin_limit / 2.0e9 + 5.0) {
                    ctx->err = "slab exchange: timed out waiting for the meta mailbox"; return OSPH_E_PEER;
                }
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        unsigned int st; memcpy(&st, s->h_all_meta + 12 * W, sizeof(st));
        if (st) {
            char msg[200];
            snprintf(msg, sizeof msg, "slab exchange on rank %d, step %lld: %s", me, (long long)s->steps,
                     (st & P2P_ST_ABORT) ? "a peer rank left the step loop with an error"
                                         : "a peer rank did not answer within OSPH_P2P_SPIN_SECONDS");
            ctx->err = msg;
            return OSPH_E_PEER;
        }
        return 0;
    };
    // all-gather of this rank's meta row (s->d_meta) through the second mailbox; the result also lands in pinned host memory
    MboxExtra plain; memset(&plain, 0, sizeof(plain));
    // reuse: 1 = a step that reuses the binning: the kernel forms the meta row itself (no counts) and installs the global
    // bounds and displacement afterwards (with OSPH_SLAB_MERGE)
    auto meta_mailbox = [&](bool reuse_merged) -> int {
        s->seq++;
        MboxExtra x = plain;
        if (reuse_merged) { x.payload_from_sc = 2; x.sc = ctx->d_sc; x.tail = 2; }
        k_mbox_allgather<<<1, 32 * W, 0, ctx->stream>>>(s->d_peer, me, W, s->off_meta, s->off_meta_flag, (int)(s->seq & 1ull), s->d_meta, 12,
                                                       s->seq, s->d_all_meta, 0, s->spin_limit, d_status, s->h_all_meta, host_flag, x);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) { ctx->err = "kernel launch: mailbox all-gather"; return OSPH_E_CUDA; }
        return 0;
    };
    auto check_overflow = [&](const double *M) -> int {
        for (int r = 0; r < W; r++)
            if (M[12 * r + 10] != 0.0) {
                char msg[320];
                snprintf(msg, sizeof msg, "slab exchange regions overflowed on rank %d at step %lld: migrants %.0f/%.0f (cap %lld), "
                         "halo %.0f/%.0f (cap %lld), flag %.3g, seq %llu", r, (long long)s->steps, M[12 * r], M[12 * r + 1],
                         (long long)s->mig_cap, M[12 * r + 2], M[12 * r + 3], (long long)s->halo_cap, M[12 * r + 10], s->seq);
                ctx->err = msg;
                return OSPH_E_CAPACITY;
            }
        return 0;
    };
    auto pair_radius = [&](double hmax) { return std::max(q * hmax, std::min(s->r0, 3.0 * hmax)); };
    const double *M = s->h_all_meta;
    GhostMap gm;
    gm.b1 = (long long)((s->off_halo_from_left - s->off_ghost) / OSPH_WIRE_HALO);
    gm.b2 = (long long)((s->off_halo_from_right - s->off_ghost) / OSPH_WIRE_HALO);

    for (int step = 0; step < nsteps; step++) {
        const bool prof = s->profile && step > 0 && step < nsteps - 1;       // steady-state steps of a multi-step call
#define P2P_MARK(k) do { if (prof) cudaEventRecord(s->pev[k], ctx->stream); } while (0)
        P2P_MARK(0);
        // ---- slab cadence: does this step sort (migrate, re-pack the halo) or reuse?  Decided from numbers every rank has
        // (the all-gathered meta rows of the step before), so all ranks decide alike. ----
        bool sort_step = true;
        if (s->cadence) {
            if (s->pending_seq) {                        // the previous step reused: its meta row has not been looked at yet
                if ((rc = wait_meta(s->pending_seq))) return fail(rc);
                s->pending_seq = 0;
                double d2 = M[11], nh = M[9];
                for (int r = 1; r < W; r++) { d2 = std::min(d2, M[12 * r + 11]); nh = std::min(nh, M[12 * r + 9]); }
                const double D = std::sqrt(std::max(-d2, 0.0));
                s->dstep = std::max(D - s->D_known, D / std::max(1, s->steps_since_sort));
                s->D_known = D; s->hmax = -nh;
                double hm = M[8];
                for (int r = 1; r < W; r++) hm = std::min(hm, M[12 * r + 8]);
                s->hmin = hm;
            }
            // the displacement is known up to the step before this one; this step adds at most about dstep (x 2: margin)
            const double R_now = pair_radius(s->hmax) * 1.02 * (1.0 + 1e-6);
            // (where the pair radius reaches the reference cell -- small N -- the acceleration grid IS the reference grid and
            // follows the bounds: k_grid_params sorts at every build there, and so does this loop)
            double cs = s->hmin * ctx->cfg.nn_scale;
            if (cs < 1e-6) cs = 1.0;
            sort_step = !s->lists_valid || !(s->dstep >= 0.0) || !std::isfinite(s->D_known) || R_now >= cs ||
                        R_now + 2.0 * (s->D_known + 2.0 * s->dstep) > s->gs;
        }
        if ((rc = osph_slab_step_plan(ctx, step, nsteps))) return fail(rc);    // corrector of step k fused into the predictor of k+1
        // ---- identical dt on every rank: mailbox all-gather + min (+ the local triple before and k_timestep after it, in
        // the same kernel) ----
        if (s->merge) {
            MboxExtra x = plain;
            if ((rc = osph_slab_dt_prepare(ctx, &x.uniform_c))) return fail(rc);
            x.payload_from_sc = 1; x.sc = ctx->d_sc; x.tail = 1;
            x.gamma_c = ctx->cfg.cfl_courant; x.gamma_f = ctx->cfg.cfl_force; x.fixed_dt = fixed_dt > 0 ? fixed_dt : -1.0; x.co = ctx->cfg.co;
            x.dt_log = ctx->d_dt_log; x.dt_log_cap = (long long)ctx->dt_log_cap; x.ts_fused = ctx->slab_fused;
            s->seq++;
            k_mbox_allgather<<<1, 32 * W, 0, ctx->stream>>>(s->d_peer, me, W, s->off_dt, s->off_dt_flag, (int)(s->seq & 1ull), s->d_dt3, 3, s->seq,
                                                           s->d_all_dt, 1, s->spin_limit, d_status, nullptr, nullptr, x);
            ctx->launches++;
            if (cudaGetLastError() != cudaSuccess) { ctx->err = "kernel launch: mailbox all-gather"; return fail(OSPH_E_CUDA); }
            P2P_MARK(1);
            if ((rc = osph_slab_step_begin_after_dt(ctx, damping))) return fail(rc);
        } else {
            if ((rc = osph_slab_dt_local(ctx, s->d_dt3))) return fail(rc);
            s->seq++;
            k_mbox_allgather<<<1, 32 * W, 0, ctx->stream>>>(s->d_peer, me, W, s->off_dt, s->off_dt_flag, (int)(s->seq & 1ull), s->d_dt3, 3, s->seq,
                                                           s->d_all_dt, 1, s->spin_limit, d_status, nullptr, nullptr, plain);
            ctx->launches++;
            if (cudaGetLastError() != cudaSuccess) { ctx->err = "kernel launch: mailbox all-gather"; return fail(OSPH_E_CUDA); }
            P2P_MARK(1);
            if ((rc = osph_slab_step_begin(ctx, s->d_all_dt + W * 3, fixed_dt, damping))) return fail(rc);
        }
        P2P_MARK(2);
        const auto host_t0 = std::chrono::steady_clock::now();
        // A flow too fast for the largest skin (not even one reusing step between two sorts) takes the one-pass exchange
        // below: one host wait instead of the two of a sorting step.  The builds keep writing the reference positions, so the
        // displacement per step stays known and the cadence resumes when the flow calms down.
        const bool fast_flow = s->cadence && sort_step && s->dstep >= 0.0 &&
                               s->dstep > OSPH_SKIN_MAX * pair_radius(s->hmax) / 4.0;
        if (!s->cadence || fast_flow) {
            // ---- every step: classify + pack straight into the neighbours' windows; counts and bounds through the mailbox ----
            const double width = pair_radius(s->hmax) * 1.1;
            if ((rc = osph_slab_pack(ctx, width, mig_l, mig_r, s->mig_cap, halo_l, halo_r, s->halo_cap, s->d_meta))) return fail(rc);
            if ((rc = meta_mailbox(false))) return fail(rc);
            P2P_MARK(3);
            if ((rc = wait_meta(s->seq))) return fail(rc);                // the one host wait of the step
            P2P_MARK(4);
            double bounds[6];
            for (int k = 0; k < 6; k++) { bounds[k] = M[4 + k]; for (int r = 1; r < W; r++) bounds[k] = std::min(bounds[k], M[12 * r + 4 + k]); }
            if ((rc = check_overflow(M))) return fail(rc);
            const double *mine = M + 12 * me;
            const int64_t out_l = (int64_t)mine[0], out_r = (int64_t)mine[1], halo_out_l = (int64_t)mine[2], halo_out_r = (int64_t)mine[3];
            const int64_t in_mig_l = left >= 0 ? (int64_t)M[12 * left + 1] : 0, in_halo_l = left >= 0 ? (int64_t)M[12 * left + 3] : 0;
            const int64_t in_mig_r = right >= 0 ? (int64_t)M[12 * right + 0] : 0, in_halo_r = right >= 0 ? (int64_t)M[12 * right + 2] : 0;
            s->hmax = -bounds[5];
            gm.c0 = (int)(out_l + out_r); gm.c1 = (int)in_halo_l;
            const int64_t n_ghost = gm.c0 + in_halo_l + in_halo_r;
            // ---- owned set update (migrants are already here), same grid everywhere ----
            if ((rc = osph_slab_commit_impl(ctx, out_l + out_r, s->win + s->off_mig_from_left, in_mig_l,
                                            s->win + s->off_mig_from_right, in_mig_r, gm, n_ghost, bounds))) return fail(rc);
            const int64_t c[8] = {out_l, out_r, halo_out_l, halo_out_r, in_mig_l, in_mig_r, in_halo_l, in_halo_r};
            memcpy(s->last_counts, c, sizeof(c));
            if (fast_flow) {
                double d2 = M[11];
                for (int r = 1; r < W; r++) d2 = std::min(d2, M[12 * r + 11]);
                const double D = std::sqrt(std::max(-d2, 0.0));
                if (std::isfinite(D) && s->steps_since_sort > 0) s->dstep = std::max(D - s->D_known, D / s->steps_since_sort);
                s->hmin = bounds[4];
                s->D_known = 0.0; s->steps_since_sort = 0; s->lists_valid = false; s->sorts++;
                ctx->slab_cadence_force = 1; ctx->slab_skin = 0.0;        // sort, no skin, reference positions refreshed
            }
        } else if (sort_step) {
            // ---- sorting step, pass A: the migrants change owner first, so that the halo lists of pass B are built on the
            // final owned set (and nobody keeps stale copies of particles it gave away) ----
            if ((rc = osph_slab_pack_mode(ctx, 0.0, 1, mig_l, mig_r, s->mig_cap, halo_l, halo_r, s->halo_cap, nullptr, nullptr, s->d_meta))) return fail(rc);
            if ((rc = meta_mailbox(false))) return fail(rc);
            if ((rc = wait_meta(s->seq))) return fail(rc);
            if ((rc = check_overflow(M))) return fail(rc);
            const int64_t out_l = (int64_t)M[12 * me + 0], out_r = (int64_t)M[12 * me + 1];
            const int64_t in_mig_l = left >= 0 ? (int64_t)M[12 * left + 1] : 0, in_mig_r = right >= 0 ? (int64_t)M[12 * right + 0] : 0;
            gm.c0 = 0; gm.c1 = 0;
            if ((rc = osph_slab_commit_impl(ctx, out_l + out_r, s->win + s->off_mig_from_left, in_mig_l,
                                            s->win + s->off_mig_from_right, in_mig_r, gm, 0, nullptr))) return fail(rc);
            // ---- pass B: halos, with a skin; the storage slot behind every record is remembered ----
            // skin: about ten steps' worth of the displacement per step observed since the last sort, 3 %..25 % of the radius
            const double R = pair_radius(s->hmax);
            // (capped where the pair kernel's staged candidate rows still fit its shared-memory records: k_grid_params)
            s->skin = s->dstep >= 0.0 ? std::min(std::max(2.0 * s->dstep * 10.0 * 1.1 / R, 0.03), (double)OSPH_SKIN_MAX) : 0.03;
            const double width = R * (1.0 + s->skin) * 1.1;
            if ((rc = osph_slab_pack_mode(ctx, width, 2, mig_l, mig_r, s->mig_cap, halo_l, halo_r, s->halo_cap, s->d_idx_l, s->d_idx_r, s->d_meta))) return fail(rc);
            if ((rc = meta_mailbox(false))) return fail(rc);
            P2P_MARK(3);
            if ((rc = wait_meta(s->seq))) return fail(rc);
            P2P_MARK(4);
            if ((rc = check_overflow(M))) return fail(rc);
            double bounds[6], d2 = M[11];
            for (int k = 0; k < 6; k++) { bounds[k] = M[4 + k]; for (int r = 1; r < W; r++) bounds[k] = std::min(bounds[k], M[12 * r + 4 + k]); }
            for (int r = 1; r < W; r++) d2 = std::min(d2, M[12 * r + 11]);
            const int64_t halo_out_l = (int64_t)M[12 * me + 2], halo_out_r = (int64_t)M[12 * me + 3];
            const int64_t in_halo_l = left >= 0 ? (int64_t)M[12 * left + 3] : 0, in_halo_r = right >= 0 ? (int64_t)M[12 * right + 2] : 0;
            // displacement per step, from the stretch that ends here (valid when the lists were: the reference positions were)
            const double D = std::sqrt(std::max(-d2, 0.0));
            if (s->lists_valid && std::isfinite(D) && s->steps_since_sort > 0)
                s->dstep = std::max(D - s->D_known, D / s->steps_since_sort);
            s->hmax = -bounds[5]; s->hmin = bounds[4];
            s->gs = pair_radius(s->hmax) * (1.0 + 1e-6) * (1.0 + s->skin);          // what k_grid_params forms from the same numbers
            s->halo_n_l = halo_out_l; s->halo_n_r = halo_out_r;
            s->D_known = 0.0; s->steps_since_sort = 0; s->lists_valid = true; s->sorts++;
            gm.c0 = 0; gm.c1 = (int)in_halo_l;
            if ((rc = osph_slab_commit_impl(ctx, 0, nullptr, 0, nullptr, 0, gm, in_halo_l + in_halo_r, bounds))) return fail(rc);
            ctx->slab_cadence_force = 1; ctx->slab_skin = s->skin;
            const int64_t c[8] = {out_l, out_r, halo_out_l, halo_out_r, in_mig_l, in_mig_r, in_halo_l, in_halo_r};
            memcpy(s->last_counts, c, sizeof(c));
        } else {
            // ---- reuse: same particles, same record slots, current values; nothing for the host to wait for.  The meta
            // mailbox still runs: it is the "records have landed" signal and carries bounds and displacement ----
            if ((rc = osph_slab_repack(ctx, s->d_idx_l, s->halo_n_l, halo_l, s->d_idx_r, s->halo_n_r, halo_r, s->merge ? nullptr : s->d_meta))) return fail(rc);
            if ((rc = meta_mailbox(s->merge))) return fail(rc);
            s->pending_seq = s->seq;
            P2P_MARK(3); P2P_MARK(4);
            if (s->merge) ctx->prepared = true;
            else if ((rc = osph_slab_install(ctx, s->d_all_meta, W))) return fail(rc);
            ctx->slab_cadence_force = 4; ctx->slab_skin = s->skin;
            s->reuses++;
        }
        s->steps_since_sort++;
        P2P_MARK(5);
        if ((rc = osph_slab_step_end(ctx, damping))) return fail(rc);
        P2P_MARK(6);
        if (prof) {
            s->phost_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
            cudaEventRecord(s->pev[7], ctx->stream);
            cudaEventSynchronize(s->pev[7]);
            for (int k = 0; k < 7; k++) { float ms = 0; cudaEventElapsedTime(&ms, s->pev[k], s->pev[k + 1]); s->pms[k] += ms; }
            s->psteps++;
        }
        s->steps++;
    }
    if (s->cadence && s->pending_seq) {
        // the call ends on a reuse step: pick up its meta row now (status, displacement), so that errors of the last step are
        // not carried into a later call and the host's view is complete when the caller looks at the results
        if ((rc = wait_meta(s->pending_seq))) return fail(rc);
        s->pending_seq = 0;
        double d2 = M[11], nh = M[9];
        for (int r = 1; r < W; r++) { d2 = std::min(d2, M[12 * r + 11]); nh = std::min(nh, M[12 * r + 9]); }
        const double D = std::sqrt(std::max(-d2, 0.0));
        s->dstep = std::max(D - s->D_known, D / std::max(1, s->steps_since_sort));
        s->D_known = D; s->hmax = -nh;
    }
    return 0;
}

extern "C" int osph_slab_p2p_last_counts(const osph_slab_p2p *s, int64_t out[8])
{
    if (!s) return OSPH_E_INVALID;
    memcpy(out, s->last_counts, sizeof(s->last_counts));
    return 0;
}

// steps of this group that sorted (migration + fresh halo lists) / that reused the binning and the frozen halo lists
extern "C" int osph_slab_p2p_stats(const osph_slab_p2p *s, int64_t out[2])
{
    if (!s || !out) return OSPH_E_INVALID;
    out[0] = s->cadence ? s->sorts : s->steps; out[1] = s->cadence ? s->reuses : 0;
    return 0;
}
