// common.cuh -- context, device-side scalars and small device helpers shared by all
// translation units of libosph_b200.so (sm_100a only; no other architecture is built).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/osph.h"

#ifndef OSPH_PAIR_THREADS
#define OSPH_PAIR_THREADS 256         // one CTA of the pair kernel owns this many consecutive sorted particles
#endif
#ifndef PAIR_UH
#define PAIR_UH 1                      // 1: Solver(h=value) runs the pair kernel's uniform-smoothing-length instantiation (pair.cu)
#endif
#define OSPH_MAX_CELL_BITS 28
#define OSPH_SKIN_MAX 0.04             // largest adaptive skin of the sort cadence, as a fraction of the pair radius (k_grid_params)
#define OSPH_WIRE_HALO 8               // doubles per ghost record: x y vx vy rho m h label
#define OSPH_WIRE_FULL 21              // doubles per migrant record: 19 columns, label, global id
#define OSPH_PAIR_EVENTS 512           // pair-kernel launches timed between two osph_pair_kernel_time calls

// ---------------------------------------------------------------------------------------------
// Device-resident scalars.  Everything the step needs between kernels lives here so that a whole
// step is enqueued without a host round trip (and can be captured in a CUDA graph).
// ---------------------------------------------------------------------------------------------
struct GridParams {
    // reference grid (NNLinkedList._init, reference src/Tools/NNLinkedList.py:86-127)
    double xmin, xmax, ymin, ymax, cell_size;
    double rcell;          // RN(1 / cell_size) for the constant-divisor division of cell_of (step.cu: div_den), 0: divide
    long long ncx, ncy, n_cells;
    // acceleration grid used on the device (regime A: identical to the reference grid;
    // regime B: cells of the pair-interaction radius)
    double gsize, ginv;
    int gnx, gny;
    int regime_a;          // 1: acceleration grid == reference grid (ids by the reference formula)
    int adj_always;        // 1: one-cell fallback grid (table too small for the reference grid): test reference-cell adjacency on every pair
    int reach_set;         // cells to walk per side to cover q <= 3 (neighbour-set emitter)
    double hmax;           // max h over active particles after the refresh
    double pair_r2;        // square of the radius inside which a pair can contribute
    // Sort cadence (regime B, single GPU): the acceleration grid is frozen at the last sort -- origin (gox, goy), cell size
    // gsize = pair radius + skin -- and the sorted order and the cell table are reused while
    //     pair radius now + 2 x (largest displacement of a particle since that sort) <= gsize,
    // i.e. while two particles within the pair radius are still guaranteed to sit in adjacent cells OF THE BINNING.
    // k_grid_params decides per build (do_sort); the sort kernels exit at once when it is 0.
    double gox, goy;       // origin of the acceleration grid (regime A: the reference origin)
    double disp;           // largest displacement since the last sort (0 on a sorting build)
    int do_sort;           // this build bins and sorts
    unsigned int sort_count;   // sorts so far (parity selects the tile-total buffer of the counting sort)
    int steps_since_sort;
    int reserved_i;
};

// Ghost records live in up to three segments of the ghost buffer (slab mode): the rank's own migrants, the halo
// received from the left neighbour and the one from the right.  NCCL receives pack them back to back (c0 = all);
// peer-to-peer writes land in fixed regions of the rank's window (b1, b2 = region starts, in records).
struct GhostMap {
    int c0, c1;
    long long b1, b2;
};

struct StepScalars {
    // order-preserving encodings (see enc_f64) so that atomicMin/atomicMax work on doubles
    unsigned long long xmin, xmax, ymin, ymax;   // bounds of active particles
    unsigned long long hmin_all;                 // min h over ALL active (cell size rule), before refresh
    unsigned long long hmax_all;                 // max h over active, after refresh
    unsigned long long hmin_fluid, cmax_fluid, a2max_fluid;   // TimeStep.computeVars
    unsigned int status;                         // OSPH_S_* bits
    unsigned int reserved;
    double dt[3];                                // {dt, dt_c, dt_f} of the current step
    double ke;
    long long dt_log_count;
    double dt_prev;                              // dt of the step before (fused corrector + predictor, step.cu)
    unsigned long long disp2max;                 // max |x - x_ref|^2 over the particles, x_ref = position at the last sort (k_prepare)
    long long builds, sorts;                     // neighbour-structure builds / those that sorted (osph_sort_stats)
    unsigned int prepare_ticket, pair_ticket;    // CTAs done: the last one of k_prepare / k_pair runs the fused scalar kernel
};

struct osph_export_ring;          // export.cu

struct osph_ctx {
    osph_config cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;

    int64_t n_total = 0;      // rows of the host array
    int64_t stride = 154;
    int64_t n = 0;            // active particles resident on the device
    int64_t n_fluid = 0;
    int64_t cap = 0;          // allocated particle capacity

    // raw record mirror + maps
    unsigned char *d_aos = nullptr;  // n_total * stride bytes
    int *d_row = nullptr;            // storage slot -> row of the host array
    int *d_act = nullptr;            // storage slot -> index in the compacted active array

    // state, SoA, storage order (OSPH_NUM_FIELDS columns of doubles + label)
    double *f[OSPH_NUM_FIELDS] = {nullptr};
    signed char *label = nullptr;
    double *scratch = nullptr;       // one spare column for the physical reorder
    double *d_stage = nullptr;       // one spare column for column transfers
    size_t aos_bytes = 0;
    bool c_uniform = false;          // c == co for every active particle (true after the first compute)

    // sort workspace
    unsigned int *key[2] = {nullptr, nullptr};
    unsigned int *idx[2] = {nullptr, nullptr};   // sorted position -> storage slot
    int sorted_buf = 0;
    unsigned int *hist = nullptr;    // [256][nblocks]
    unsigned int *digit_tot = nullptr;
    int sort_blocks = 0;
    int key_bits = 0;

    // cell table of the acceleration grid: (begin, end) per cell
    int2 *cell_range = nullptr;
    int64_t cell_cap = 0;
    // sort cadence: positions at the last sort (storage order) and whether they describe the resident particles
    double *xref = nullptr, *yref = nullptr;
    bool timestep_fusion = false;
    bool grid_fusion = false;          // osph_step: k_grid_params runs in the last CTA of the predictor pass, k_timestep in the pair kernel's
    bool skin_valid = false;          // false: the next build must sort (upload, external edits, slab mode, table re-sized)
    double skin_frac = -1.0;          // OSPH_SKIN: skin as a fraction of the pair radius; < 0 adaptive; 0 sorts every build
    // counting sort by cell (binsort.cu): cell histogram, tile states of its single-pass scan (+ the ticket counter)
    bool bin_sort = true;             // false (OSPH_SORT=radix): LSD radix sort of sort.cu
    unsigned int *bin_counts = nullptr;
    unsigned int *bin_tiles = nullptr;   // totals of the scan tiles (BIN_TILE cells each), two buffers used alternately
    int64_t bin_tile_cap = 0;

    // per-particle cell info in sorted order
    int4 *s_coarse = nullptr;        // reference grid: (bin_cx, bin_cy, query_cx, query_cy); bin_cx < 0: unbinned
    int2 *s_gcell = nullptr;         // acceleration-grid query cell
    // pair-kernel inputs in sorted order (Real = double or float)
    double2 *s_pos = nullptr;
    void *s_vel = nullptr, *s_rm = nullptr, *s_hp = nullptr;   // Real2 each
    int *s_info = nullptr;           // bit0: fluid
    // slab mode: ghost particles as light wire records (OSPH_WIRE_HALO doubles each), appended after the owned ones
    double *d_ghost = nullptr;
    int64_t n_ghost = 0, ghost_cap = 0;
    GhostMap gmap = {0, 0, 0, 0};
    bool ghost_external = false;
    unsigned int *scan_block = nullptr;   // block sums of the scan utility
    bool slab = false;
    int slab_fused = 0;               // osph_slab_step_plan: 0 plain step, 1 first / 2 later step of a fused multi-step call
    bool slab_defer = false;          // this step's corrector is applied by the next step's predictor pass
    bool slab_last = true;            // osph_slab_step_plan: last step of the call (ends with the plain corrector)
    double x_lo = 0, x_hi = 0;
    int slab_cadence_force = 0;       // slab cadence (slab_p2p.cu): 0 off; else what the next build does: 1 sort, 4 reuse
    double slab_skin = 0.0;           // ... with this skin (fraction of the pair radius), the same on every rank
    int *d_slab_counters = nullptr, *d_mig_slots = nullptr, *d_tail_flag = nullptr, *d_holes = nullptr, *d_fillers = nullptr;
    int64_t slab_list_cap = 0;
    int64_t reserve = 0;              // particle capacity requested by osph_reserve
    int64_t scan_cap = 0;

    GridParams *d_grid = nullptr;
    StepScalars *d_sc = nullptr;
    double *d_dt_log = nullptr;
    int64_t dt_log_cap = 0;
    double *d_partial = nullptr;     // block partials for deterministic sums

    bool neighbours_valid = false;
    bool prepared = false;           // bounds / h refresh done for the current positions
    bool sized = false;              // cell table sized for the current particle set
    bool reductions_valid = false;   // hmin_fluid/cmax/a2max describe the current state
    int64_t step_counter = 0;
    int64_t build_counter = 0;
    int64_t launches = 0;

    // timers
    std::vector<cudaEvent_t> phase_ev;   // (start, stop) pairs of the per-phase timers
    std::vector<int> phase_id;
    double timers_ms[6] = {0, 0, 0, 0, 0, 0};
    cudaEvent_t pair_ev[2 * OSPH_PAIR_EVENTS] = {nullptr};   // (start, stop) per pair-kernel launch
    int pair_ev_used = 0;
    bool time_pair = true;
    bool pair_next_timestep = false; // osph_step's fused loop: the pair kernel's last CTA computes the next step's dt (k_timestep fused)
    double pair_ts_fixed_dt = -1.0;
    bool pair_reduce_a2 = false;     // osph_step's fused loop: the pair kernel reduces max |a|^2 for the next dt

    // pinned staging for transfers
    unsigned char *h_pinned = nullptr;
    size_t h_pinned_bytes = 0;

    // PAIR_UH: loop constants of the uniform-smoothing-length pair kernel, valid for (uh_for_h, uh_for_kernel, uh_for_prec)
    bool uh_ready = false;
    double uh_for_h = 0.0;
    int uh_for_kernel = -1, uh_for_prec = -1;
    double uh_d[5] = {0, 0, 0, 0, 0};
    float uh_f[5] = {0, 0, 0, 0, 0};
    double *d_uh = nullptr;
    int64_t pair_launches = 0, pair_uh_launches = 0;    // osph_pair_kernel_info

    osph_export_ring *xring = nullptr;   // asynchronous column export (export.cu), created on first use
    unsigned char *d_rows_buf = nullptr; // workspace of the row transfers (export.cu)
    size_t rows_bytes = 0;
};

void osph_export_free(osph_ctx *ctx);

#define OSPH_CUDA(call)                                                                    \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                \
            return OSPH_E_CUDA;                                                            \
        }                                                                                  \
    } while (0)

#define OSPH_LAUNCH_CHECK()                                                                \
    do {                                                                                   \
        ctx->launches++;                                                                   \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess) {                                                          \
            ctx->err = std::string("kernel launch: ") + cudaGetErrorString(e__);           \
            return OSPH_E_CUDA;                                                            \
        }                                                                                  \
    } while (0)

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ const double *ghost_record(const double *ghost, const GhostMap &gm, int g)
{
    long long slot = g < gm.c0 ? g : (g < gm.c0 + gm.c1 ? gm.b1 + (g - gm.c0) : gm.b2 + (g - gm.c0 - gm.c1));
    return ghost + slot * OSPH_WIRE_HALO;
}

// Monotone map double -> uint64 so that unsigned atomicMin/atomicMax order doubles correctly.
__device__ __forceinline__ unsigned long long enc_f64(double v)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u & 0x8000000000000000ULL) ? ~u : (u | 0x8000000000000000ULL);
}
__host__ __device__ __forceinline__ double dec_f64(unsigned long long u)
{
    u = (u & 0x8000000000000000ULL) ? (u & 0x7fffffffffffffffULL) : ~u;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
#define ENC_POS_INF 0xfff0000000000000ULL   // enc(+inf)
#define ENC_NEG_INF 0x000fffffffffffffULL   // enc(-inf)

// Warp-wide max / min of 64-bit keys with the redux unit: two 32-bit reductions (high words, then the low words of the
// lanes that hold the winning high word) instead of five shuffle rounds with a 64-bit compare-and-select each.  Doubles go
// through enc_f64, so the result is the same value a chain of fmin / fmax would return.  All 32 lanes must call.
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long e)
{
    const unsigned int hi = (unsigned int)(e >> 32), lo = (unsigned int)e;
    const unsigned int mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned int mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    return ((unsigned long long)mhi << 32) | mlo;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long e) { return ~warp_max_u64(~e); }

// NaN is skipped, as fmin / fmax would: it maps to the identity of the reduction.
__device__ __forceinline__ unsigned long long enc_for_min(double v) { return v == v ? enc_f64(v) : ENC_POS_INF; }
__device__ __forceinline__ unsigned long long enc_for_max(double v) { return v == v ? enc_f64(v) : ENC_NEG_INF; }

__device__ __forceinline__ double warp_min(double v) { return dec_f64(warp_min_u64(enc_for_min(v))); }
__device__ __forceinline__ double warp_max(double v) { return dec_f64(warp_max_u64(enc_for_max(v))); }
// Counting sort by cell (binsort.cu): the canonical position of the particle in storage slot `mine`, which the atomics of
// k_bin_keys placed at sorted position s of a cell occupying r = [begin, end): begin + the number of cell mates with a
// smaller storage slot.  Cells fuller than BIN_DENSE_CELL keep their arrival order (the count would be quadratic).
#define BIN_DENSE_CELL 32768
#define BIN_TILE 4096                  // cells per tile of the cell-count scan
__device__ __forceinline__ int bin_canonical_slot(int2 r, int s, unsigned int mine, const unsigned int *__restrict__ idx_arrival,
                                                  StepScalars *sc)
{
    if (r.y - r.x > BIN_DENSE_CELL) {
        if (s == r.x) atomicOr(&sc->status, OSPH_S_DENSE_CELL);
        return s;
    }
    int before = 0;
    for (int t = r.x; t < r.y; t++) before += idx_arrival[t] < mine ? 1 : 0;
    return r.x + before;
}

// TimeStep.compute / courant / force (src/Equations/TimeStep.py:11-56), strict IEEE: one thread.  Kernel of its own
// (k_timestep, step.cu) or the tail of the pair kernel's last CTA (osph_step's loop).
// fused: 1 = the dt scalars are reset here, right after use (the predictor and the pair kernel of this step reduce the next
// ones); 2 = additionally c_max is the uniform co (no corrector pass reduced it).
// All reads of *sc are volatile: in a last-CTA tail the words were just written by the atomics of other CTAs.
__device__ __attribute__((noinline)) static void timestep_body(StepScalars *sc, double gamma_c, double gamma_f, double fixed_dt, double *dt_log,
                                                  long long dt_log_cap, int reset_prepare, const double *reduced3, int fused, double co)
{
    auto rd = [](const unsigned long long *p) { return *reinterpret_cast<const volatile unsigned long long *>(p); };
    if (reduced3) {           // slab mode: all-reduced {h_min, -c_max, -a2_max} replaces the local reduction
        sc->hmin_fluid = enc_f64(reduced3[0]); sc->cmax_fluid = enc_f64(-reduced3[1]); sc->a2max_fluid = enc_f64(-reduced3[2]);
    }
    if (reset_prepare) {      // folded k_reset_prepare_scalars: the predictor that follows reduces into these
        sc->xmin = ENC_POS_INF; sc->ymin = ENC_POS_INF; sc->xmax = ENC_NEG_INF; sc->ymax = ENC_NEG_INF;
        sc->hmin_all = ENC_POS_INF; sc->hmax_all = ENC_NEG_INF; sc->disp2max = ENC_NEG_INF;
    }
    double out0, c = 0.0, f = 0.0;
    if (fixed_dt > 0.0) { out0 = fixed_dt; }
    else {
        double hmin = dec_f64(rd(&sc->hmin_fluid)), cmax = fused == 2 ? co : dec_f64(rd(&sc->cmax_fluid)), a2 = dec_f64(rd(&sc->a2max_fluid));
        c = __ddiv_rn(__dmul_rn(gamma_c, hmin), cmax);
        f = (a2 < 1e-12) ? 1e10 : __dmul_rn(gamma_f, __dsqrt_rn(__ddiv_rn(hmin, a2)));
        out0 = c < f ? c : f;
        if (out0 < 1e-6) atomicOr(&sc->status, OSPH_S_SMALL_DT);
    }
    if (fused) { sc->hmin_fluid = ENC_POS_INF; sc->cmax_fluid = ENC_NEG_INF; sc->a2max_fluid = ENC_NEG_INF; }
    sc->dt_prev = *reinterpret_cast<const volatile double *>(&sc->dt[0]);
    sc->dt[0] = out0; sc->dt[1] = c; sc->dt[2] = f;
    if (dt_log) {
        long long k = *reinterpret_cast<const volatile long long *>(&sc->dt_log_count);
        if (k < dt_log_cap) { dt_log[3 * k] = out0; dt_log[3 * k + 1] = c; dt_log[3 * k + 2] = f; }
        sc->dt_log_count = k + 1;
    }
}

__device__ __forceinline__ int warp_min_i(int v) { return __reduce_min_sync(0xffffffffu, v); }
__device__ __forceinline__ int warp_max_i(int v) { return __reduce_max_sync(0xffffffffu, v); }

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------
// launchers implemented in the other translation units
// ---------------------------------------------------------------------------------------------
int osph_bin_alloc(osph_ctx *ctx, int64_t cell_cap);                              // binsort.cu
void osph_bin_free(osph_ctx *ctx);
int osph_bin_sort(osph_ctx *ctx, int64_t n_all, bool rank_now);

int osph_sort_pairs(osph_ctx *ctx, int64_t n, int bits, bool first_hist_done);   // sort.cu: key[sorted_buf], idx[sorted_buf]
int osph_sort_alloc(osph_ctx *ctx, int64_t cap);
void osph_sort_free(osph_ctx *ctx);
int osph_launch_pair(osph_ctx *ctx);                                      // pair.cu

#ifdef __CUDACC__
// The reference record is packed (154 bytes, src/Common.py:26-57): its doubles sit at 2-byte alignment only.
__device__ __forceinline__ double load_f64_unaligned(const unsigned char *p)
{
    const unsigned short *s = reinterpret_cast<const unsigned short *>(p);
    unsigned long long u = (unsigned long long)s[0] | ((unsigned long long)s[1] << 16) |
                           ((unsigned long long)s[2] << 32) | ((unsigned long long)s[3] << 48);
    return __longlong_as_double((long long)u);
}
__device__ __forceinline__ void store_f64_unaligned(unsigned char *p, double v)
{
    unsigned short *s = reinterpret_cast<unsigned short *>(p);
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    s[0] = (unsigned short)u; s[1] = (unsigned short)(u >> 16);
    s[2] = (unsigned short)(u >> 32); s[3] = (unsigned short)(u >> 48);
}
#endif
