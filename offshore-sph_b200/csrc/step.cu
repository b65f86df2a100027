// step.cu -- everything of the WCSPH step except the radix sort and the fused pair kernel:
// record (un)packing, predictor / corrector with their fused reductions, grid parameters, cell keys,
// cell table, gather of the pair-kernel inputs, neighbour-list emitter and point query.
//
// All arithmetic that the reference evaluates in strict IEEE double (neighbour search, time step:
// numba jitclass without fastmath) uses __dadd_rn/__dmul_rn/__ddiv_rn/__dsqrt_rn so that nvcc cannot
// contract it into FMAs; integrator updates do the same so they are bit-identical to the CPU oracle.
#include "common.cuh"
#include "step.cuh"
#include "sort.cuh"
#include "sph_math.cuh"

// ---------------------------------------------------------------------------------------------
// K10: packed 154-byte records <-> SoA columns.  Doubles sit at byte offset 2 + 8k (2-byte aligned
// only), so they are moved as four 16-bit words.
// ---------------------------------------------------------------------------------------------

struct Columns { double *f[OSPH_NUM_FIELDS]; };

// Active-row list on the device: flag = !deleted, exclusive scan (scan.cu), then compaction.  counters[0] = number
// of active rows, counters[1] = number of fluid rows among them.
__global__ void k_active_flags(const unsigned char *__restrict__ aos, long long stride, int n, unsigned int *__restrict__ flag)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) flag[r] = aos[(long long)r * stride] ? 0u : 1u;
}
__global__ void k_active_compact(const unsigned char *__restrict__ aos, long long stride, int n,
                                 const unsigned int *__restrict__ pos, int *__restrict__ row, int *__restrict__ act,
                                 int *__restrict__ counters)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    int fluid = 0;
    if (r < n) {
        const unsigned char *rec = aos + (long long)r * stride;
        bool alive = rec[0] == 0;
        if (alive) { unsigned int p = pos[r]; row[p] = r; act[p] = (int)p; fluid = rec[1] == OSPH_FLUID; }
        if (r == n - 1) counters[0] = (int)pos[r] + (alive ? 1 : 0);
    }
    unsigned int m = __ballot_sync(0xffffffffu, fluid);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters[1], __popc(m));
}

// osph_set_active: the `deleted` byte of every record from a mask of active rows
__global__ void k_set_deleted(unsigned char *__restrict__ aos, long long stride, int n, const unsigned char *__restrict__ active)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) aos[(long long)r * stride] = active[r] ? 0 : 1;
}

__global__ void k_unpack_aos(const unsigned char *__restrict__ aos, long long stride, const int *__restrict__ row,
                             int n, Columns c, signed char *__restrict__ label)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char *rec = aos + (long long)row[i] * stride;
    label[i] = (signed char)rec[1];
#pragma unroll
    for (int k = 0; k < OSPH_NUM_FIELDS; k++) c.f[k][i] = load_f64_unaligned(rec + 2 + 8 * k);
}

// Pack: a CTA transposes 128 particles through shared memory (coalesced column reads -> packed records),
// then each warp streams whole records to their rows with contiguous 16-bit stores.  `ids` (optional) receives
// the row ids; `identity` writes record i to row i (owned-set export) and also sets deleted = 0 / label.
#define PACK_ROWS 128
__global__ void __launch_bounds__(PACK_ROWS)
k_pack_aos(unsigned char *__restrict__ aos, long long stride, const int *__restrict__ row, int n, Columns c,
           const signed char *__restrict__ label, int c_uniform, double co, int identity, int *__restrict__ ids)
{
    __shared__ __align__(16) unsigned short sh[PACK_ROWS * 77];
    __shared__ int sh_row[PACK_ROWS];
    const int t = threadIdx.x, i0 = blockIdx.x * PACK_ROWS, i = i0 + t;
    if (i < n) {
        unsigned short *rec = sh + t * 77;
        rec[0] = identity ? (unsigned short)((unsigned char)label[i]) << 8 : 0;      // bytes 0,1 = deleted(0), label
#pragma unroll
        for (int k = 0; k < OSPH_NUM_FIELDS; k++) {
            double v = c.f[k][i];
            if (k == OSPH_F_C && c_uniform) v = co;
            unsigned long long u = (unsigned long long)__double_as_longlong(v);
            rec[1 + 4 * k] = (unsigned short)u; rec[2 + 4 * k] = (unsigned short)(u >> 16);
            rec[3 + 4 * k] = (unsigned short)(u >> 32); rec[4 + 4 * k] = (unsigned short)(u >> 48);
        }
        int r = row[i];
        sh_row[t] = identity ? i : r;
        if (ids) ids[i] = r;
    }
    __syncthreads();
    const int lane = t & 31, w = t >> 5, rows = min(PACK_ROWS, n - i0);
    for (int r = w; r < rows; r += PACK_ROWS / 32) {
        unsigned short *dst = reinterpret_cast<unsigned short *>(aos + (long long)sh_row[r] * stride);
        const unsigned short *src = sh + r * 77;
        for (int hw = lane + (identity ? 0 : 1); hw < 77; hw += 32) dst[hw] = src[hw];
    }
}

// column of the active particles in active order <-> storage order
__global__ void k_col_to_active(const double *__restrict__ col, const int *__restrict__ act, int n,
                                double *__restrict__ out, int fill, double fill_value)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[act[i]] = fill ? fill_value : col[i];
}
__global__ void k_col_from_active(double *__restrict__ col, const int *__restrict__ act, int n,
                                  const double *__restrict__ in)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) col[i] = in[act[i]];
}

__global__ void k_fill_c(double *__restrict__ c, int n, double co)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) c[i] = co;
}

// Solver.setup() on the fluid rows (reference src/Solver.py:184-196): smoothing length from the uploaded
// density, hydrostatic density (WCSPH.initialize, TaitEOS_height), then pressure and speed of sound.
__global__ void k_setup(int n, const signed char *__restrict__ label, const double *__restrict__ m,
                        const double *__restrict__ y, double *__restrict__ rho, double *__restrict__ h,
                        double *__restrict__ p, double *__restrict__ c, int dynamic_h, double fixed_h, double h_sigma,
                        double rho0, double H, double B, double gamma, double Pb, double co)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || label[i] != OSPH_FLUID) return;
    if (dynamic_h == OSPH_H_DYNAMIC) h[i] = rho[i] > 1e-12 ? h_sigma * sqrt(m[i] / rho[i]) : 0.0;
    else if (dynamic_h == OSPH_H_FIXED) h[i] = fixed_h;
    double frac = rho0 * 9.81 * (H - y[i]) / B;
    double r = rho0 * pow(1.0 + frac, 1.0 / gamma);
    rho[i] = r;
    p[i] = (pow(r / rho0, gamma) - 1.0) * B + Pb;
    c[i] = co;
}

// ---------------------------------------------------------------------------------------------
// scalars
// ---------------------------------------------------------------------------------------------
__global__ void k_reset_prepare_scalars(StepScalars *sc)
{
    sc->xmin = ENC_POS_INF; sc->ymin = ENC_POS_INF; sc->xmax = ENC_NEG_INF; sc->ymax = ENC_NEG_INF;
    sc->hmin_all = ENC_POS_INF; sc->hmax_all = ENC_NEG_INF; sc->disp2max = ENC_NEG_INF;
}
__global__ void k_reset_dt_scalars(StepScalars *sc)
{
    sc->hmin_fluid = ENC_POS_INF; sc->cmax_fluid = ENC_NEG_INF; sc->a2max_fluid = ENC_NEG_INF;
}
__global__ void k_init_scalars(StepScalars *sc)
{
    sc->status = 0; sc->dt[0] = sc->dt[1] = sc->dt[2] = 0.0; sc->ke = 0.0; sc->dt_log_count = 0;
    sc->builds = 0; sc->sorts = 0; sc->disp2max = ENC_NEG_INF; sc->prepare_ticket = 0u; sc->pair_ticket = 0u;
}

template <int NV>
__device__ __forceinline__ void block_minmax_atomic(double (&vmin)[NV], unsigned long long *const (&pmin)[NV],
                                                    int nmin)
{
    // vmin[k] for k < nmin are minima, the rest maxima.  Reduced as order-preserving 64-bit keys with the redux unit
    // (warp_min_u64 / warp_max_u64): the shuffle + fmin / fmax form cost ~60 instructions per value and thread and made
    // the predictor pass instruction-bound (ncu: 47 % issue utilisation at 3.3 TB/s, profiles/r01).
    __shared__ unsigned long long sh[NV][32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        const unsigned long long e = k < nmin ? warp_min_u64(enc_for_min(vmin[k])) : warp_max_u64(enc_for_max(vmin[k]));
        if (lane == 0) sh[k][w] = e;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            unsigned long long e = lane < nw ? sh[k][lane] : (k < nmin ? ENC_POS_INF : ENC_NEG_INF);
            e = k < nmin ? warp_min_u64(e) : warp_max_u64(e);
            if (lane == 0) {
                // look before the atomic: thousands of CTAs hit the same few words and almost none improves them
                const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(pmin[k]);
                if (k < nmin) { if (e < ENC_POS_INF && e < cur) atomicMin(pmin[k], e); }
                else { if (e > ENC_NEG_INF && e > cur) atomicMax(pmin[k], e); }
            }
        }
    }
}

// v / den for the loop-invariant den = 1 + damping / 2 (PEC.py:44-45, 72-73).  With rden = RN(1 / den), one
// correction step on q = RN(v * rden) -- r = v - q * den exactly (fma), q' = RN(q + r * rden) -- gives the correctly
// rounded quotient (Markstein), i.e. the bits of the division the reference performs, for 3 instructions instead of
// the division subroutine; checked against v / den on 6.6e8 random operands.  rden == 0 (set on
// the host when den's significand is all ones, the one case the theorem excludes) selects the plain division.
__device__ __forceinline__ double div_den(double v, double den, double rden)
{
    if (rden == 0.0) return __ddiv_rn(v, den);
    const double q = __dmul_rn(v, rden);
    const double r = __fma_rn(-q, den, v);
    return __fma_rn(r, rden, q);
}

static double rden_for(double damping)
{
    const double den = 1.0 + 0.5 * damping;
    unsigned long long bits;
    memcpy(&bits, &den, 8);
    if (!(den > 0.0) || !std::isfinite(den) || (bits & 0xfffffffffffffULL) == 0xfffffffffffffULL) return 0.0;
    const double r = 1.0 / den;
    return std::isfinite(r) ? r : 0.0;
}

// ---------------------------------------------------------------------------------------------
// K1 + K2: predictor (PEC: reference src/Integrators/PEC.py:32-60, Verlet.py:28-36, Euler: identity)
// fused with the reductions NNLinkedList._init needs (bounds of ALL active particles,
// NNLinkedList.py:89-93; min h before the refresh, :103-110) and with the smoothing-length refresh of
// Solver._compute (src/Solver.py:242-246, computeH SolverTools.py:106-118).
// dt is read from device memory so a multi-step loop needs no host round trip.
// ---------------------------------------------------------------------------------------------
// FUSED (PEC only, osph_step's multi-step loop): the corrector of the PREVIOUS step (same arithmetic as k_correct, with
// sc->dt_prev) is applied in registers first, so the corrected state is never written and re-read between two steps:
// one pass over the state per step boundary instead of two.
#ifndef PREP_ITEMS
#define PREP_ITEMS 8
#endif
__device__ void grid_params_body(StepScalars *sc, GridParams *g, double nn_scale, double pair_radius_q, double r0,
                                 long long cell_cap, int reset_dt, int force_sort, double skin_frac);
template <int INTEG, bool PREDICT, bool FUSED>
__global__ void __launch_bounds__(256)
k_prepare(PrepareArgs a)
{
    // A CTA walks PREP_ITEMS consecutive groups of 256 particles and reduces once: the eight-value block reduction and the
    // CTA launch are paid per 256 * PREP_ITEMS particles (the reduction was most of the kernel's instructions).
    double Tbx = INFINITY, Tby = INFINITY, Thmn = INFINITY, TBx = -INFINITY, TBy = -INFINITY, Thmx = -INFINITY, Thfl = INFINITY;
    double Td2 = -INFINITY;
#pragma unroll 1
    for (int it = 0; it < PREP_ITEMS; it++) {
    const int i = blockIdx.x * (256 * PREP_ITEMS) + it * 256 + threadIdx.x;
    double bx = INFINITY, by = INFINITY, hmn = INFINITY, Bx = -INFINITY, By = -INFINITY, hmx = -INFINITY, hfl = INFINITY;
    double d2 = -INFINITY;            // |position after this pass - position at the last sort|^2 (sort cadence, k_grid_params)
    if (i < a.n) {
        double xr = 0.0, yr = 0.0;
        if (a.xref) { xr = a.xref[i]; yr = a.yref[i]; }
        // Every input of the thread is requested before the first one is used.  The argument block carries no
        // __restrict__, so a load written after a store has to wait for it: the straightforward form (load, compute,
        // store, load the next rate, ...) walked through six dependent memory round trips per particle and re-read the
        // rates between corrector and predictor (SASS of the first version; ncu: long_scoreboard-bound at 40 % occupancy).
        const bool fluid = a.label[i] == OSPH_FLUID;
        double h = a.h[i];
        constexpr bool RATES = FUSED || (PREDICT && INTEG == OSPH_INTEGRATOR_PEC);     // ax, ay, drho, xsph are read
        const double dt_prev = FUSED ? a.sc->dt_prev : 0.0;
        const double dt = PREDICT ? (a.use_dev_dt ? a.sc->dt[0] : a.dt) : 0.0;
        double x = 0.0, y = 0.0, vx = 0.0, vy = 0.0, rho = 0.0;
        double x0 = 0.0, y0 = 0.0, vx0 = 0.0, vy0 = 0.0, rho0 = 0.0, cvx = 0.0, cvy = 0.0;
        double xsx = 0.0, xsy = 0.0, axv = 0.0, ayv = 0.0, drhov = 0.0, m = 0.0;
        // (the fluid rows' inputs are requested without waiting for the label: 99 % of the rows are fluid, the arrays
        // cover every row, and the label would otherwise be one more dependent round trip)
        if (FUSED) {
            x0 = a.x0[i]; y0 = a.y0[i]; vx0 = a.vx0[i]; vy0 = a.vy0[i]; rho0 = a.rho0[i];
            if (!a.integ_xsph) { cvx = a.vx[i]; cvy = a.vy[i]; }           // the corrector moves with the evaluated velocity
            if (!fluid) { x = a.x[i]; y = a.y[i]; }
        } else {
            x = a.x[i]; y = a.y[i]; rho = a.rho[i];
            if (PREDICT) { vx = a.vx[i]; vy = a.vy[i]; }
        }
        if (RATES) {
            if (a.integ_xsph) { xsx = a.xsphx[i]; xsy = a.xsphy[i]; }
            axv = a.ax[i]; ayv = a.ay[i]; drhov = a.drho[i];
        }
        if (a.dynamic_h == OSPH_H_DYNAMIC) m = a.m[i];
        hmn = h;
        if (fluid) {
            if (FUSED) {
                // corrector of the step before: PEC.py:62-88, identical operation order to k_correct
                const double hdtp = __dmul_rn(0.5, dt_prev);
                const double ux = a.integ_xsph ? xsx : cvx, uy = a.integ_xsph ? xsy : cvy;
                const double mx = __dadd_rn(x0, __dmul_rn(hdtp, ux));
                const double my = __dadd_rn(y0, __dmul_rn(hdtp, uy));
                const double den = __dadd_rn(1.0, __dmul_rn(0.5, a.damping));
                const double mvx = div_den(__dadd_rn(vx0, __dmul_rn(hdtp, axv)), den, a.rden);
                const double mvy = div_den(__dadd_rn(vy0, __dmul_rn(hdtp, ayv)), den, a.rden);
                x = __dadd_rn(__dmul_rn(2.0, mx), -x0);
                y = __dadd_rn(__dmul_rn(2.0, my), -y0);
                vx = __dadd_rn(__dmul_rn(2.0, mvx), -vx0);
                vy = __dadd_rn(__dmul_rn(2.0, mvy), -vy0);
                const double mrho = __dadd_rn(rho0, __dmul_rn(hdtp, drhov));
                rho = __dadd_rn(__dmul_rn(2.0, mrho), -rho0);
                if (a.strict && rho < 0.0) rho = 0.0;
                if (!(isfinite(x) && isfinite(y) && isfinite(vx) && isfinite(vy) && isfinite(rho)))
                    atomicOr(&a.sc->status, OSPH_S_NONFINITE);
            }
            if (PREDICT) {
                const double hdt = __dmul_rn(0.5, dt);
                if (INTEG == OSPH_INTEGRATOR_PEC) {
                    a.x0[i] = x; a.y0[i] = y; a.vx0[i] = vx; a.vy0[i] = vy;
                    const double ux = a.integ_xsph ? xsx : vx, uy = a.integ_xsph ? xsy : vy;
                    x = __dadd_rn(x, __dmul_rn(hdt, ux));
                    y = __dadd_rn(y, __dmul_rn(hdt, uy));
                    const double den = __dadd_rn(1.0, __dmul_rn(0.5, a.damping));
                    a.vx[i] = div_den(__dadd_rn(vx, __dmul_rn(hdt, axv)), den, a.rden);
                    a.vy[i] = div_den(__dadd_rn(vy, __dmul_rn(hdt, ayv)), den, a.rden);
                    a.rho0[i] = rho;
                    rho = __dadd_rn(rho, __dmul_rn(hdt, drhov));
                    if (a.strict && rho < 0.0) rho = 0.0;
                    a.rho[i] = rho;
                } else if (INTEG == OSPH_INTEGRATOR_VERLET) {
                    x = __dadd_rn(x, __dmul_rn(hdt, vx));
                    y = __dadd_rn(y, __dmul_rn(hdt, vy));
                }
                a.x[i] = x; a.y[i] = y;
            }
            // smoothing-length refresh
            if (a.dynamic_h == OSPH_H_DYNAMIC) {
                h = 0.0;
                if (rho > 1e-12) h = __dmul_rn(a.h_sigma, sqrt(__ddiv_rn(m, rho)));
                a.h[i] = h;
            } else if (a.dynamic_h == OSPH_H_FIXED) {
                h = a.fixed_h;
                a.h[i] = h;
            }
        }
        // a particle that has left the floating-point line must not define the grid (it is parked in the overflow cell)
        if (isfinite(x) && isfinite(y)) {
            bx = x; Bx = x; by = y; By = y;
            const double ddx = x - xr, ddy = y - yr;
            d2 = ddx * ddx + ddy * ddy;
            if (!(d2 == d2)) d2 = INFINITY;        // a reference that is not a number: sort
        }
        else atomicOr(&a.sc->status, OSPH_S_NONFINITE);
        hmx = h;
        if (fluid) hfl = h;
    }
    Tbx = fmin(Tbx, bx); Tby = fmin(Tby, by); Thmn = fmin(Thmn, hmn); Thfl = fmin(Thfl, hfl);
    TBx = fmax(TBx, Bx); TBy = fmax(TBy, By); Thmx = fmax(Thmx, hmx); Td2 = fmax(Td2, d2);
    }
    const double bx = Tbx, by = Tby, hmn = Thmn, hfl = Thfl, Bx = TBx, By = TBy, hmx = Thmx, d2 = Td2;
    if (a.reduce_hmin_fluid) {
        double v[8] = {bx, by, hmn, hfl, Bx, By, hmx, d2};
        unsigned long long *const p[8] = {&a.sc->xmin, &a.sc->ymin, &a.sc->hmin_all, &a.sc->hmin_fluid,
                                          &a.sc->xmax, &a.sc->ymax, &a.sc->hmax_all, &a.sc->disp2max};
        block_minmax_atomic<8>(v, p, 4);
    } else {
        double v[7] = {bx, by, hmn, Bx, By, hmx, d2};
        unsigned long long *const p[7] = {&a.sc->xmin, &a.sc->ymin, &a.sc->hmin_all,
                                          &a.sc->xmax, &a.sc->ymax, &a.sc->hmax_all, &a.sc->disp2max};
        block_minmax_atomic<7>(v, p, 3);
    }
    if (a.grid) {
        // Fused grid parameters (osph_step's loop): the CTA that finishes last has every reduction of the pass behind it and
        // forms the grid and the sort decision right here -- no k_grid_params launch between this pass and the sort kernels.
        // Thread 0 issued this CTA's atomics itself (lane 0 of warp 0), so its fence orders them before the ticket.
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int ticket = atomicAdd(&a.sc->prepare_ticket, 1u);
            if (ticket == gridDim.x - 1) {
                a.sc->prepare_ticket = 0u;
                __threadfence();
                grid_params_body(a.sc, a.grid, a.g_nn_scale, a.g_pair_radius_q, a.g_r0, a.g_cell_cap, a.g_reset_dt, a.g_force_sort,
                                 a.g_skin_frac);
            }
        }
    }
}

template __global__ void k_prepare<OSPH_INTEGRATOR_PEC, true, false>(PrepareArgs);
template __global__ void k_prepare<OSPH_INTEGRATOR_PEC, true, true>(PrepareArgs);
template __global__ void k_prepare<OSPH_INTEGRATOR_VERLET, true, false>(PrepareArgs);
template __global__ void k_prepare<OSPH_INTEGRATOR_PEC, false, false>(PrepareArgs);

// ---------------------------------------------------------------------------------------------
// Grid parameters: the reference grid (NNLinkedList.py:86-127) and the acceleration grid.
// ---------------------------------------------------------------------------------------------
__device__ void grid_params_body(StepScalars *sc, GridParams *g, double nn_scale, double pair_radius_q, double r0,
                                 long long cell_cap, int reset_dt, int force_sort, double skin_frac)
{
    if (reset_dt) {           // folded k_reset_dt_scalars: the corrector of this step reduces into these
        sc->hmin_fluid = ENC_POS_INF; sc->cmax_fluid = ENC_NEG_INF; sc->a2max_fluid = ENC_NEG_INF;
        }
    // (volatile: inside k_prepare's last CTA these words were just written by the atomics of other CTAs, and the SM may
    // hold an older copy of the line in L1 from its own plain loads of sc->dt)
    auto rd = [](const unsigned long long *p) { return *reinterpret_cast<const volatile unsigned long long *>(p); };
    double xmin = dec_f64(rd(&sc->xmin)), xmax = dec_f64(rd(&sc->xmax)), ymin = dec_f64(rd(&sc->ymin)), ymax = dec_f64(rd(&sc->ymax));
    double hmin = dec_f64(rd(&sc->hmin_all)), hmax = dec_f64(rd(&sc->hmax_all));
    g->xmin = xmin; g->xmax = xmax; g->ymin = ymin; g->ymax = ymax; g->hmax = hmax;
    double cs = __dmul_rn(hmin, nn_scale);           // :104-105
    if (cs < 1e-6) cs = 1.0;                          // :107-108
    g->cell_size = cs;
    double inv = __ddiv_rn(1.0, cs);                  // :113
    g->rcell = ((unsigned long long)__double_as_longlong(cs) & 0xfffffffffffffULL) == 0xfffffffffffffULL || !isfinite(inv) ? 0.0 : inv;
    long long ncx = (long long)ceil(__dmul_rn(inv, __dadd_rn(xmax, -xmin)));
    long long ncy = (long long)ceil(__dmul_rn(inv, __dadd_rn(ymax, -ymin)));
    if (ncx < 1) ncx = 1;
    if (ncy < 1) ncy = 1;
    g->ncx = ncx; g->ncy = ncy; g->n_cells = ncx * ncy;

    // Radius inside which a pair can contribute: kernel support (q < 2 cubic/Wendland, q <= 3 Gaussian)
    // and the Lennard-Jones range min(r0, 3 h_ij); h_ij <= hmax.
    double lj = fmin(r0, 3.0 * hmax);
    double R = fmax(pair_radius_q * hmax, lj) * (1.0 + 1e-6);
    if (!(R > 0.0)) R = cs;
    g->pair_r2 = R * R;
    const double rs = 3.0 * hmax * (1.0 + 1e-6);          // reach of the reference neighbour set (q <= 3)

    // ---- sort cadence: can this build reuse the binning of the last sort? ----
    // disp = largest distance of a particle from where it was binned (k_prepare).  The frozen cells are gsize wide; two
    // particles now within R were at most R + 2 disp apart when they were binned, hence in adjacent cells as long as that
    // does not exceed gsize.  Regime A (the acceleration grid IS the reference grid, which follows the bounds) and the
    // one-cell fallback always sort.
    // force_sort: 0 the device decides; 1 sort (a skin is still added, for the builds that follow); 2 sort without skin (slab
    // mode without cadence and the radix path sort every build); 3 = the host's sizing pass for the cell table: grid only, no
    // bookkeeping; 4 = reuse, decided by the host (slab cadence)
    const bool dry = force_sort == 3;
    if (!dry) sc->builds++;
    const double disp = sqrt(fmax(dec_f64(rd(&sc->disp2max)), 0.0));
    const bool valid = g->sort_count > 0 && !g->regime_a && !g->adj_always && !(R >= cs) &&
                       isfinite(disp) && R + 2.0 * disp * (1.0 + 1e-9) <= g->gsize;
    bool reuse = !force_sort && skin_frac != 0.0 && valid;
    if (force_sort == 4) {
        // slab cadence (slab_p2p.cu): the ranks decided TOGETHER, on the host, that this build reuses the binning -- the ghost
        // records arrived in the slots of the last sort, so sorting on one's own is not an option.  The host plans with a
        // margin; should the condition fail all the same, pairs may be missing: say so loudly.
        reuse = true;
        if (!valid) atomicOr(&sc->status, OSPH_S_SKIN_EXHAUSTED);
    }
    if (reuse) {
        g->do_sort = 0; g->disp = disp; g->steps_since_sort++;
        int reach = (int)ceil((rs + 2.0 * disp) / g->gsize);
        if (reach < 1) reach = 1;
        g->reach_set = reach;
        return;
    }
    // skin of the new binning.  Adaptive (skin_frac < 0): the particles needed steps_since_sort builds to use up the last
    // skin, so the largest displacement per build was about disp / steps; size the skin for about ten builds at that pace
    // (each percent of skin costs about two percent more candidates in the pair kernel's scan, a sort about 25 us).
    double skin = 0.0;
    if (!(R >= cs) && skin_frac != 0.0 && force_sort != 2 && !dry) {
        if (skin_frac > 0.0) skin = skin_frac * R;
        else {
            const int steps = g->sort_count > 0 && g->steps_since_sort > 0 ? g->steps_since_sort : 0;
            const double per_build = steps > 0 && isfinite(disp) ? disp / steps : 0.0;
            // Capped at 4 % of R: the pair kernel stages the three candidate rows of a CTA in shared memory (PAIR_CAP
            // records), typically 835 of 1024 at skin 0; from about 5 % on a growing share of the CTAs falls back to batched
            // staging (measured: 294 us at 3 %, 308 us at 5 %).  A flow too fast to get two builds out of that skin sorts at
            // every build instead (skin 0: no candidates paid for nothing).
            skin = fmin(fmax(2.0 * per_build * 10.0 * 1.1, 0.03 * R), OSPH_SKIN_MAX * R);
            if (per_build > 0.0 && skin < 2.0 * per_build * 2.0) skin = 0.0;
        }
    }
    int regime_a = (R >= cs) ? 1 : 0, adj_always = 0;
    double gs = regime_a ? cs : R + skin;
    long long gnx, gny;
    if (regime_a) { gnx = ncx; gny = ncy; }
    else {
        // cells of the pair radius (+ skin); if the table cannot hold them (the domain grew since it was sized) coarsen in one
        // shot to the cell size that fits, then nudge
        const double ex = fmax(xmax - xmin, 0.0), ey = fmax(ymax - ymin, 0.0);
        const double fit = sqrt((ex + gs) * (ey + gs) / (0.9 * (double)(cell_cap - 2)));
        if (gs < fit) { gs = fit; atomicOr(&sc->status, OSPH_S_GRID_COARSE); }
        for (int it = 0; it < 64; it++) {
            gnx = (long long)floor(ex / gs) + 1;
            gny = (long long)floor(ey / gs) + 1;
            if (gnx * gny + 1 <= cell_cap) break;
            gs *= 1.1;
        }
    }
    if (!(gnx >= 1 && gny >= 1 && gnx * gny + 1 <= cell_cap) || !isfinite(gs)) {
        // regime A cannot be coarsened (it IS the reference grid), or the bounds are unusable: fall back to ONE cell holding
        // everything.  Candidates are then all particles and membership is decided by the stored reference cell ids and the
        // distance alone (adj_always), so the step stays exact -- O(n^2), but regime A only exists for a few thousand
        // particles -- and the host re-sizes the table at its next status read (osph_sync).
        atomicOr(&sc->status, 0x80000000u);
        gnx = 1; gny = 1; gs = fmax(fmax(xmax - xmin, ymax - ymin), 0.0) + 1.0; regime_a = 0; adj_always = 1;
        if (!isfinite(gs)) gs = 1.0;
    }
    g->adj_always = adj_always;
    g->regime_a = regime_a; g->gsize = gs; g->ginv = 1.0 / gs; g->gnx = (int)gnx; g->gny = (int)gny;
    g->gox = xmin; g->goy = ymin;
    g->do_sort = 1; g->disp = 0.0; g->steps_since_sort = 0;
    if (!dry) { g->sort_count++; sc->sorts++; }
    int reach = regime_a ? 1 : (int)ceil(rs / gs);
    if (reach < 1) reach = 1;
    g->reach_set = reach;
}

__global__ void k_grid_params(StepScalars *sc, GridParams *g, double nn_scale, double pair_radius_q, double r0,
                              long long cell_cap, int reset_dt, int force_sort, double skin_frac)
{
    grid_params_body(sc, g, nn_scale, pair_radius_q, r0, cell_cap, reset_dt, force_sort, skin_frac);
}

// ---------------------------------------------------------------------------------------------
// K3: cell keys.  Reference cell id (NNLinkedList.py:129-141, 164-176, 200-208) in strict IEEE.
// The same function is re-evaluated by the gather kernel (deterministic, so the ids agree) instead of
// carrying 24 bytes of cell info per particle through the sort.
// ---------------------------------------------------------------------------------------------
struct CellInfo { int4 coarse; int2 gcell; unsigned int key; bool binned; };

__device__ __forceinline__ CellInfo cell_of(double x, double y, const GridParams &g)
{
    CellInfo c;
    double rx = __dadd_rn(x, -g.xmin), ry = __dadd_rn(y, -g.ymin);
    long long cx = (long long)floor(div_den(rx, g.cell_size, g.rcell));       // == rx / cell_size, correctly rounded
    long long cy = (long long)floor(div_den(ry, g.cell_size, g.rcell));
    long long flat = cx + g.ncx * cy;
    c.binned = flat >= 0 && flat < g.n_cells;
    c.coarse.x = c.binned ? (int)(flat % g.ncx) : -1000000;
    c.coarse.y = c.binned ? (int)(flat / g.ncx) : -1000000;
    c.coarse.z = (int)cx; c.coarse.w = (int)cy;
    if (!(isfinite(x) && isfinite(y))) {                                        // parked: never found, finds nothing
        c.binned = false; c.coarse = make_int4(-1000000, -1000000, -1000000, -1000000);
        c.gcell = make_int2(-1000000, -1000000);
        c.key = (unsigned int)g.gnx * (unsigned int)g.gny;                      // the extra cell behind the table
        return c;
    }
    if (g.regime_a) {
        c.gcell = make_int2((int)cx, (int)cy);                                  // query cell: raw reference ids
        c.key = c.binned ? (unsigned int)flat : (unsigned int)g.n_cells;        // bin cell: the reference's flat id
    } else {
        // (the acceleration grid keeps the origin of the last sort; a particle outside it is clamped into the edge cells,
        // which preserves "within one cell size => same or adjacent cells")
        int gx = (int)floor((x - g.gox) * g.ginv), gy = (int)floor((y - g.goy) * g.ginv);
        gx = min(max(gx, 0), g.gnx - 1); gy = min(max(gy, 0), g.gny - 1);
        c.gcell = make_int2(gx, gy);
        c.key = (unsigned int)gy * (unsigned int)g.gnx + (unsigned int)gx;
    }
    return c;
}

// ghosts (slab mode) are light wire records of OSPH_WIRE_HALO doubles: x y vx vy rho m h label.
// The kernel walks the radix sort's tiles and also emits the first pass's per-tile digit histogram.
template <int BITS>
__global__ void __launch_bounds__(SORT_THREADS, 4)     // <= 64 registers: the 493 tiles of 1 M particles stay one wave (4 x 148 slots)
k_keys(const double *__restrict__ x, const double *__restrict__ y, int n_owned, const double *__restrict__ ghost,
       GhostMap gmap, int n_all, const GridParams *__restrict__ gp, StepScalars *sc, unsigned int *__restrict__ key,
       unsigned int *__restrict__ idx, int nblocks, unsigned int *__restrict__ hist)
{
    constexpr int RADIX = 1 << BITS;
    __shared__ unsigned int cnt[RADIX];
    for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS) cnt[d] = 0;
    __syncthreads();
    const GridParams g = *gp;
    bool unbinned = false;
    // all positions of the thread first, then the cell arithmetic: with the load inside the loop below every item waited
    // for its own round trip behind the previous item's warp vote and shared-memory atomic (SASS of the first version)
    double px[SORT_ITEMS], py[SORT_ITEMS];
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        const int i = blockIdx.x * SORT_TILE + r * SORT_THREADS + threadIdx.x;
        px[r] = 0.0; py[r] = 0.0;
        if (i < n_owned) { px[r] = x[i]; py[r] = y[i]; }
        else if (i < n_all) { const double *rec = ghost_record(ghost, gmap, i - n_owned); px[r] = rec[0]; py[r] = rec[1]; }
    }
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        int i = blockIdx.x * SORT_TILE + r * SORT_THREADS + threadIdx.x;
        const bool ok = i < n_all;
        unsigned int digit = 0;
        if (ok) {
            CellInfo c = cell_of(px[r], py[r], g);
            unbinned |= !c.binned;
            key[i] = c.key; idx[i] = (unsigned int)i;
            digit = c.key & (RADIX - 1);
        }
        warp_hist_add(cnt, digit, ok);
    }
    if (unbinned) atomicOr(&sc->status, OSPH_S_UNBINNED);
    __syncthreads();
    for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS) hist[(size_t)d * nblocks + blockIdx.x] = cnt[d];
}

// K3 of the counting sort (binsort.cu): cell key + arrival rank inside the cell + cell histogram.  Neighbouring particles
// share their cell, so the warp first groups equal keys (match_any) and the group's leader takes ONE global atomic for all
// of them; the arrival order is made canonical later (k_bin_rank).
#define BINK_ITEMS 4
#define BINK_WINDOW 16            // scan tiles around the CTA's first key whose totals are collected in shared memory
__global__ void __launch_bounds__(256)
k_bin_keys(const double *__restrict__ x, const double *__restrict__ y, int n_owned, const double *__restrict__ ghost,
           GhostMap gmap, int n_all, const GridParams *__restrict__ gp, StepScalars *sc, unsigned int *__restrict__ key,
           unsigned int *__restrict__ arrival, unsigned int *__restrict__ counts, unsigned int *__restrict__ tile_sums_base,
           long long tile_cap, double *__restrict__ xref, double *__restrict__ yref)
{
    if (!gp->do_sort) return;                             // this build reuses the binning of the last sort (k_grid_params)
    unsigned int *__restrict__ tile_sums = tile_sums_base + (size_t)(gp->sort_count & 1u) * tile_cap;
    // The totals of the scan tiles (BIN_TILE cells) are accumulated here so that k_bin_scan needs no look-back.  A tile is
    // hit by thousands of particles: straight global atomics serialise on a few hundred addresses (52 us measured), so the
    // CTA collects them in shared memory -- storage order follows cell order closely, a CTA's keys sit in one or two tiles --
    // and flushes one atomic per tile it touched.  Keys outside the window go to global memory directly.
    __shared__ unsigned int sh_tile[BINK_WINDOW];
    __shared__ int sh_tile0;
    const GridParams g = *gp;
    const int lane = threadIdx.x & 31;
    const int base = blockIdx.x * (256 * BINK_ITEMS) + (threadIdx.x >> 5) * (32 * BINK_ITEMS) + lane;
    double px[BINK_ITEMS], py[BINK_ITEMS];
#pragma unroll
    for (int r = 0; r < BINK_ITEMS; r++) {                  // all loads of the thread first (one round trip)
        const int i = base + r * 32;
        px[r] = 0.0; py[r] = 0.0;
        if (i < n_owned) { px[r] = x[i]; py[r] = y[i]; }
        else if (i < n_all) { const double *rec = ghost_record(ghost, gmap, i - n_owned); px[r] = rec[0]; py[r] = rec[1]; }
    }
    if (xref) {                                             // where every particle is binned: the displacements count from here
#pragma unroll
        for (int r = 0; r < BINK_ITEMS; r++) {
            const int i = base + r * 32;
            if (i < n_owned) { xref[i] = px[r]; yref[i] = py[r]; }
        }
    }
    if (threadIdx.x < BINK_WINDOW) sh_tile[threadIdx.x] = 0u;
    if (threadIdx.x == 0) sh_tile0 = (int)(cell_of(px[0], py[0], g).key / BIN_TILE) - BINK_WINDOW / 2;
    __syncthreads();
    const int tile0 = sh_tile0;
    bool unbinned = false;
#pragma unroll
    for (int r = 0; r < BINK_ITEMS; r++) {
        const int i = base + r * 32;
        const bool ok = i < n_all;
        unsigned int k = 0xffffffffu;
        if (ok) {
            CellInfo c = cell_of(px[r], py[r], g);
            unbinned |= !c.binned;
            k = c.key;
        }
        const unsigned int peers = __match_any_sync(0xffffffffu, k);
        const int leader = __ffs(peers) - 1;
        unsigned int first = 0;
        if (ok && lane == leader) {
            const unsigned int cnt = (unsigned int)__popc(peers);
            first = atomicAdd(&counts[k], cnt);
            const int t = (int)(k / BIN_TILE) - tile0;
            if (t >= 0 && t < BINK_WINDOW) atomicAdd(&sh_tile[t], cnt);
            else atomicAdd(&tile_sums[k / BIN_TILE], cnt);
        }
        first = __shfl_sync(0xffffffffu, first, leader);
        if (ok) { key[i] = k; arrival[i] = first + __popc(peers & ((1u << lane) - 1u)); }
    }
    if (unbinned) atomicOr(&sc->status, OSPH_S_UNBINNED);
    __syncthreads();
    if (threadIdx.x < BINK_WINDOW && sh_tile[threadIdx.x]) atomicAdd(&tile_sums[tile0 + (int)threadIdx.x], sh_tile[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------
// K5: gather of the pair-kernel inputs into sorted order, fused with the Tait EOS
// (reference WCSPH.compute_pressure, src/Methods/WCSPH.py:131-149; TaitEOS.py:6-31).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double tait_ratio_pow(double ratio, double gamma)
{
    if (gamma == 7.0) { double r2 = ratio * ratio, r4 = r2 * r2; return r4 * r2 * ratio; }
    return pow(ratio, gamma);
}

template <typename Real2>
#ifndef GATHER_MINB
#define GATHER_MINB 5
#endif
__global__ void __launch_bounds__(256, GATHER_MINB)          // 5 CTAs / SM (<= 51 registers), as before the grid parameters moved into registers
k_gather(GatherArgs a, Real2 *__restrict__ s_vel, Real2 *__restrict__ s_rm, Real2 *__restrict__ s_hp)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_all) return;
    // key, neighbouring keys, slot and the grid parameters are requested together (one round trip); the gathers below are
    // the second.  Written in program order (key, table store, idx, gathers, stores, *gp) the compiler had to keep four.
    unsigned int k = 0, kp = 0, kn = 0;
    if (a.key) {                                      // radix path only: the counting sort has written the table already
        k = a.key[s];
        kp = s > 0 ? a.key[s - 1] : k; kn = s < a.n_all - 1 ? a.key[s + 1] : k;
    }
    // (both candidate slots and the cell range are requested before the grid parameters say which applies: no extra round trip)
    int2 rr = make_int2(0, 0);
    unsigned int i_arr = 0;
    if (a.rank_ranges) { rr = a.rank_ranges[s]; i_arr = a.idx_arrival[s]; }
    const unsigned int i_fin = a.idx[s];
    const GridParams g = *a.gp;
    const bool ranking = a.rank_ranges != nullptr && g.do_sort != 0;     // counting sort, on a build that sorts
    int i = (int)(ranking ? i_arr : i_fin);
    if (a.key) {   // K6 fused: cell table from the sorted keys (table zeroed before: empty cells have begin == end == 0)
        if (s == 0 || kp != k) a.cell_range[k].x = s;
        if (s == a.n_all - 1 || kn != k) a.cell_range[k].y = s + 1;
    }
    double x, y, vx, vy, rho, m, h;
    int lab, info;
    if (i < a.n_owned) {
        x = a.x[i]; y = a.y[i]; vx = a.vx[i]; vy = a.vy[i]; rho = a.rho[i]; m = a.m[i]; h = a.h[i];
        lab = a.label[i]; info = 2;                                   // bit1: owned = a target of the pair kernel
    } else {
        const double *r = ghost_record(a.ghost, a.gmap, i - a.n_owned);
        x = r[0]; y = r[1]; vx = r[2]; vy = r[3]; rho = r[4]; m = r[5]; h = r[6]; lab = (int)r[7]; info = 0;
    }
    if (ranking) {
        // counting sort: this thread holds the particle that ARRIVED at position s; its final position is decided by the
        // storage slots of its cell mates (the state loads above are in flight while they are counted)
        s = bin_canonical_slot(rr, s, (unsigned int)i, a.idx_arrival, a.sc);
        a.idx_out[s] = (unsigned int)i;
    }
    bool fluid = lab == OSPH_FLUID;
#if PAIR_UH
    // the uniform-h pair kernel takes h of every fluid particle from a constant: hold the state to it (k_prepare wrote
    // fixed_h into every owned fluid row, ghosts carry what their owner wrote), loudly
    if (a.uh_h != 0.0 && fluid && h != a.uh_h) atomicOr(&a.sc->status, OSPH_S_H_NOT_UNIFORM);
#endif
    double p = a.Pb;
    if (fluid) p = (tait_ratio_pow(rho / a.rho0, a.gamma) - 1.0) * a.B + a.Pb;
    if (i < a.n_owned) a.p[i] = p;
    double pr2 = fluid ? p / (rho * rho) : 0.0;
    a.s_pos[s] = make_double2(x, y);
    typedef decltype(Real2().x) Real;
    Real2 v; v.x = (Real)vx; v.y = (Real)vy; s_vel[s] = v;
    Real2 rm; rm.x = (Real)rho; rm.y = (Real)m; s_rm[s] = rm;
    Real2 hp; hp.x = (Real)h; hp.y = (Real)pr2; s_hp[s] = hp;
    CellInfo c = cell_of(x, y, g);
    // bit2: the reference bins this particle in a cell other than the one it queries from (wrapped last column,
    // unbinned, non-finite): cell adjacency then does not follow from distance and the pair kernel tests it
    const bool irregular = c.coarse.x != c.coarse.z || c.coarse.y != c.coarse.w || !c.binned;
    // bit3: not binned by the reference (never found).  Bits 4..17 / 18..31: the reference bin cell modulo 2^14 per axis, for
    // the float pair kernel's adjacency test (two particles within the pair radius are a handful of cells apart, so the
    // difference modulo 2^14, read as a signed 14-bit number, is in [-1, 1] exactly when the cells are adjacent)
    a.s_info[s] = info | (fluid ? 1 : 0) | (irregular ? 4 : 0) | (c.binned ? 0 : 8) |
                  ((c.coarse.x & 0x3fff) << 4) | (int)((unsigned int)(c.coarse.y & 0x3fff) << 18);
    a.s_coarse[s] = c.coarse;
    if (g.do_sort) a.s_gcell[s] = c.gcell;          // the cell the particle is BINNED in: unchanged while the binning is reused
}
template __global__ void k_gather<double2>(GatherArgs, double2 *, double2 *, double2 *);
template __global__ void k_gather<float2>(GatherArgs, float2 *, float2 *, float2 *);

// physical reorder of one column: dst[s] = src[idx[s]]
template <typename T>
__global__ void k_permute(const T *__restrict__ src, const unsigned int *__restrict__ idx, int n, T *__restrict__ dst)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) dst[s] = src[idx[s]];
}
template __global__ void k_permute<double>(const double *, const unsigned int *, int, double *);
template __global__ void k_permute<int>(const int *, const unsigned int *, int, int *);
template __global__ void k_permute<signed char>(const signed char *, const unsigned int *, int, signed char *);

__global__ void k_iota(unsigned int *__restrict__ idx, int n)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) idx[s] = (unsigned int)s;
}

// ---------------------------------------------------------------------------------------------
// K8: corrector (PEC.py:62-88, Euler.py:17-26, Verlet.py:38-55) fused with the reduction TimeStep
// needs for the NEXT step (TimeStep.computeVars, src/Equations/TimeStep.py:58-91) and the
// non-finite check.
// ---------------------------------------------------------------------------------------------
template <int INTEG, bool CORRECT>
__global__ void __launch_bounds__(256)
k_correct(CorrectArgs a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double hmn = INFINITY, cmx = -INFINITY, amx = -INFINITY;
    if (i < a.n && a.label[i] == OSPH_FLUID) {
        double ax = a.ax[i], ay = a.ay[i];
        // requested here, with the other inputs: read after the stores below they cost one more dependent round trip
        const double h_i = a.h[i], c_i = a.c_uniform ? a.co : a.c[i];
        if (CORRECT) {
            const double dt = a.use_dev_dt ? a.sc->dt[0] : a.dt;
            double drho = a.drho[i];
            double x, y, vx, vy, rho;
            if (INTEG == OSPH_INTEGRATOR_PEC) {
                const double hdt = __dmul_rn(0.5, dt);
                double x0 = a.x0[i], y0 = a.y0[i], vx0 = a.vx0[i], vy0 = a.vy0[i], rho0 = a.rho0[i];
                double ux, uy;
                if (a.integ_xsph) { ux = a.xsphx[i]; uy = a.xsphy[i]; } else { ux = a.vx[i]; uy = a.vy[i]; }
                double mx = __dadd_rn(x0, __dmul_rn(hdt, ux));
                double my = __dadd_rn(y0, __dmul_rn(hdt, uy));
                const double den = __dadd_rn(1.0, __dmul_rn(0.5, a.damping));
                double mvx = div_den(__dadd_rn(vx0, __dmul_rn(hdt, ax)), den, a.rden);
                double mvy = div_den(__dadd_rn(vy0, __dmul_rn(hdt, ay)), den, a.rden);
                x = __dadd_rn(__dmul_rn(2.0, mx), -x0);
                y = __dadd_rn(__dmul_rn(2.0, my), -y0);
                vx = __dadd_rn(__dmul_rn(2.0, mvx), -vx0);
                vy = __dadd_rn(__dmul_rn(2.0, mvy), -vy0);
                double mrho = __dadd_rn(rho0, __dmul_rn(hdt, drho));
                rho = __dadd_rn(__dmul_rn(2.0, mrho), -rho0);
                if (a.strict && rho < 0.0) rho = 0.0;
            } else if (INTEG == OSPH_INTEGRATOR_EULER) {
                x = a.x[i]; y = a.y[i]; vx = a.vx[i]; vy = a.vy[i]; rho = a.rho[i];
                const double hdt2 = __dmul_rn(__dmul_rn(0.5, dt), dt);
                x = __dadd_rn(__dadd_rn(x, __dmul_rn(dt, vx)), __dmul_rn(hdt2, ax));
                y = __dadd_rn(__dadd_rn(y, __dmul_rn(dt, vy)), __dmul_rn(hdt2, ay));
                vx = __dadd_rn(vx, __dmul_rn(dt, ax));
                vy = __dadd_rn(vy, __dmul_rn(dt, ay));
                rho = __dadd_rn(rho, __dmul_rn(dt, drho));
            } else {
                const double hdt = __dmul_rn(0.5, dt);
                x = a.x[i]; y = a.y[i]; vx = a.vx[i]; vy = a.vy[i]; rho = a.rho[i];
                vx = __dadd_rn(vx, __dmul_rn(dt, ax));
                vy = __dadd_rn(vy, __dmul_rn(dt, ay));
                double ux = vx, uy = vy;
                if (a.integ_xsph) { ux = a.xsphx[i]; uy = a.xsphy[i]; }
                x = __dadd_rn(x, __dmul_rn(hdt, ux));
                y = __dadd_rn(y, __dmul_rn(hdt, uy));
            }
            a.x[i] = x; a.y[i] = y; a.vx[i] = vx; a.vy[i] = vy; a.rho[i] = rho;
            if (!(isfinite(x) && isfinite(y) && isfinite(vx) && isfinite(vy) && isfinite(rho)))
                atomicOr(&a.sc->status, OSPH_S_NONFINITE);
        }
        hmn = h_i;
        cmx = c_i;
        amx = __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay));
        if (!(amx == amx)) amx = INFINITY;
    }
    double v[3] = {hmn, cmx, amx};
    unsigned long long *const p[3] = {&a.sc->hmin_fluid, &a.sc->cmax_fluid, &a.sc->a2max_fluid};
    block_minmax_atomic<3>(v, p, 1);
}
template __global__ void k_correct<OSPH_INTEGRATOR_PEC, true>(CorrectArgs);
template __global__ void k_correct<OSPH_INTEGRATOR_EULER, true>(CorrectArgs);
template __global__ void k_correct<OSPH_INTEGRATOR_VERLET, true>(CorrectArgs);
template __global__ void k_correct<OSPH_INTEGRATOR_PEC, false>(CorrectArgs);

// TimeStep.compute / courant / force (src/Equations/TimeStep.py:11-56), strict IEEE.
// fused (osph_step's multi-step loop): 1 = the dt scalars are reset here, right after use (the predictor and the pair
// kernel of this step reduce the next ones); 2 = additionally c_max is the uniform co (no corrector pass reduced it).
__global__ void k_timestep(StepScalars *sc, double gamma_c, double gamma_f, double fixed_dt,
                           double *dt_log, long long dt_log_cap, int reset_prepare, const double *reduced3, int fused, double co)
{
    timestep_body(sc, gamma_c, gamma_f, fixed_dt, dt_log, dt_log_cap, reset_prepare, reduced3, fused, co);
}

// KineticEnergy (src/Equations/KineticEnergy.py:6-12) over the fluid rows; deterministic two-stage sum.
__global__ void __launch_bounds__(256)
k_ke_partial(const double *__restrict__ m, const double *__restrict__ vx, const double *__restrict__ vy,
             const signed char *__restrict__ label, int n, double *__restrict__ partial)
{
    __shared__ double sh[256];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double k = 0.0;
    if (i < n && label[i] == OSPH_FLUID) k = 0.5 * m[i] * (vx[i] * vx[i] + vy[i] * vy[i]);
    sh[threadIdx.x] = k;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256)
k_ke_final(const double *__restrict__ partial, int nb, StepScalars *sc)
{
    __shared__ double sh[256];
    double k = 0.0;
    for (int b = threadIdx.x; b < nb; b += 256) k += partial[b];
    sh[threadIdx.x] = k;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) sc->ke = sh[0];
}

// ---------------------------------------------------------------------------------------------
// K11 / K12: neighbour sets with the reference predicate (NNLinkedList.py:50-71):
//   j is binned in one of the 3x3 valid reference cells around i's raw cell  AND  r / h_ij <= 3.0,
// all in strict IEEE double.  Used for parity tests and for the nearPos point query.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool ref_accept(double px, double py, double ph, int qcx, int qcy, double2 pj, double hj,
                                           int4 cj, double *r_out, double *q_out, double *h_out)
{
    if (abs(cj.x - qcx) > 1 || abs(cj.y - qcy) > 1) return false;
    double dx = __dadd_rn(px, -pj.x), dy = __dadd_rn(py, -pj.y);
    double r = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    double hij = __dmul_rn(0.5, __dadd_rn(ph, hj));
    double q = __ddiv_rn(r, hij);
    if (!(q <= 3.0)) return false;
    if (r_out) { *r_out = r; *q_out = q; *h_out = hij; }
    return true;
}

// Walks the acceleration grid around (gx, gy) with `reach` cells per side and calls f(s) per candidate.
template <typename F>
__device__ __forceinline__ void walk_cells(const GridParams &g, const int2 *__restrict__ cell_range, int gx, int gy,
                                           int reach, F f)
{
    for (int dy = -reach; dy <= reach; dy++) {
        int cy = gy + dy;
        if (cy < 0 || cy >= g.gny) continue;
        int x0 = max(gx - reach, 0), x1 = min(gx + reach, g.gnx - 1);
        for (int cx = x0; cx <= x1; cx++) {
            int2 r = cell_range[(long long)cy * g.gnx + cx];
            for (int s = r.x; s < r.y; s++) f(s);
        }
    }
}

// mode 0: count per active index; mode 1: fill (ascending active index within each list)
__global__ void __launch_bounds__(128)
k_neighbours(NeighbourArgs a, int mode)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n) return;
    int i = (int)a.idx[s];
    int ai = a.act[i];
    if (!(a.s_info[s] & 1)) { if (mode == 0) a.counts[ai] = 0; return; }
    const GridParams g = *a.gp;
    double2 pi = a.s_pos[s];
    double hi = a.h[i];
    int4 ci = a.s_coarse[s];
    int2 gc = a.s_gcell[s];
    long long base = mode ? a.offsets[ai] : 0;
    int cnt = 0;
    walk_cells(g, a.cell_range, gc.x, gc.y, g.reach_set, [&](int t) {
        if (ref_accept(pi.x, pi.y, hi, ci.z, ci.w, a.s_pos[t], a.h[a.idx[t]], a.s_coarse[t], nullptr, nullptr, nullptr)) {
            if (mode) {
                long long aj = a.act[a.idx[t]];
                long long k = cnt;                       // insertion sort keeps the list ascending
                while (k > 0 && a.out[base + k - 1] > aj) { a.out[base + k] = a.out[base + k - 1]; k--; }
                a.out[base + k] = aj;
            }
            cnt++;
        }
    });
    if (mode == 0) a.counts[ai] = cnt;
}

__global__ void k_near_pos(NeighbourArgs a, double px, double py, double ph, long long cap, long long *out_idx,
                           double *out_r, double *out_q, double *out_h, long long *out_count)
{
    const GridParams g = *a.gp;
    double rx = __dadd_rn(px, -g.xmin), ry = __dadd_rn(py, -g.ymin);
    int qcx = (int)floor(__ddiv_rn(rx, g.cell_size)), qcy = (int)floor(__ddiv_rn(ry, g.cell_size));
    int gx, gy, reach;
    if (g.regime_a) { gx = qcx; gy = qcy; reach = 1; }
    else {
        gx = (int)floor((px - g.gox) * g.ginv); gy = (int)floor((py - g.goy) * g.ginv);
        double rs = 1.5 * (ph + g.hmax) * (1.0 + 1e-6) + g.disp;      // the particles may have moved g.disp since they were binned
        reach = (int)ceil(rs / g.gsize); if (reach < 1) reach = 1;
    }
    long long cnt = 0;
    walk_cells(g, a.cell_range, gx, gy, reach, [&](int t) {
        double r, q, h;
        if (ref_accept(px, py, ph, qcx, qcy, a.s_pos[t], a.h[a.idx[t]], a.s_coarse[t], &r, &q, &h)) {
            long long aj = a.act[a.idx[t]];
            if (cnt < cap) {
                long long k = cnt;
                while (k > 0 && out_idx[k - 1] > aj) {
                    out_idx[k] = out_idx[k - 1]; out_r[k] = out_r[k - 1]; out_q[k] = out_q[k - 1]; out_h[k] = out_h[k - 1];
                    k--;
                }
                out_idx[k] = aj; out_r[k] = r; out_q[k] = q; out_h[k] = h;
            }
            cnt++;
        }
    });
    *out_count = cnt;
}

// Batched SPH pressure probe (reference examples/IceBreak.py:252-285, called there once per ice node and step):
// neighbours of the point by the reference predicate, W(r, h) with the PROBE's h, Shepard-normalised summation
// density over the fluid neighbours (src/Equations/Shepard.py:5-12, SummationDensity.py:6-13), Tait EOS.
template <int KID>
__global__ void __launch_bounds__(64)
k_probe_pressure(NeighbourArgs a, const double *__restrict__ m, const double *__restrict__ rho,
                 const signed char *__restrict__ label, int npts, const double *__restrict__ px,
                 const double *__restrict__ py, double ph, double rho0, double gamma, double B,
                 double *__restrict__ out_rho, double *__restrict__ out_p)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npts) return;
    const GridParams g = *a.gp;
    const double x = px[k], y = py[k];
    double rx = __dadd_rn(x, -g.xmin), ry = __dadd_rn(y, -g.ymin);
    int qcx = (int)floor(__ddiv_rn(rx, g.cell_size)), qcy = (int)floor(__ddiv_rn(ry, g.cell_size));
    int gx, gy, reach;
    if (g.regime_a) { gx = qcx; gy = qcy; reach = 1; }
    else {
        gx = (int)floor((x - g.gox) * g.ginv); gy = (int)floor((y - g.goy) * g.ginv);
        reach = (int)ceil((1.5 * (ph + g.hmax) * (1.0 + 1e-6) + g.disp) / g.gsize); if (reach < 1) reach = 1;
    }
    const double inv_h = 1.0 / ph;
    double norm = 0.0, sum = 0.0;
    walk_cells(g, a.cell_range, gx, gy, reach, [&](int t) {
        int j = (int)a.idx[t];
        double r, q, hij;
        if (!ref_accept(x, y, ph, qcx, qcy, a.s_pos[t], a.h[j], a.s_coarse[t], &r, &q, &hij)) return;
        if (label[j] != OSPH_FLUID) return;
        double w, gg, qq = r * inv_h;
        sph_kernel<double, KID>(qq, inv_h, 0.0, w, gg);
        if (KID == OSPH_KERNEL_GAUSSIAN && !(qq <= 3.0)) w = 0.0;
        if (rho[j] >= 1e-3) norm += w * m[j] / rho[j];
        sum += m[j] * w;
    });
    double r_ = sum / norm;                 // sum_j m_j (w_j / w_tilde)
    out_rho[k] = r_;
    out_p[k] = (pow(r_ / rho0, gamma) - 1.0) * B;
}

__global__ void k_cells_to_active(const int4 *__restrict__ s_coarse, const unsigned int *__restrict__ idx,
                                  const int *__restrict__ act, int n, const GridParams *__restrict__ gp,
                                  long long *__restrict__ out)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int4 c = s_coarse[s];
    out[act[idx[s]]] = (long long)c.z + gp->ncx * (long long)c.w;   // the unclamped value _bin computes
}

// =============================================================================================
// host launchers
// =============================================================================================
static Columns columns_of(osph_ctx *ctx)
{
    Columns c;
    for (int k = 0; k < OSPH_NUM_FIELDS; k++) c.f[k] = ctx->f[k];
    return c;
}

int osph_init_scalars(osph_ctx *ctx)
{
    k_init_scalars<<<1, 1, 0, ctx->stream>>>(ctx->d_sc); OSPH_LAUNCH_CHECK();
    k_reset_prepare_scalars<<<1, 1, 0, ctx->stream>>>(ctx->d_sc); OSPH_LAUNCH_CHECK();
    k_reset_dt_scalars<<<1, 1, 0, ctx->stream>>>(ctx->d_sc); OSPH_LAUNCH_CHECK();
    return 0;
}

int osph_launch_setup(osph_ctx *ctx)
{
    const osph_config &c = ctx->cfg;
    k_setup<<<div_up(ctx->n, 256), 256, 0, ctx->stream>>>((int)ctx->n, ctx->label, ctx->f[OSPH_F_M], ctx->f[OSPH_F_Y],
                                                          ctx->f[OSPH_F_RHO], ctx->f[OSPH_F_H], ctx->f[OSPH_F_P],
                                                          ctx->f[OSPH_F_C], c.dynamic_h, c.fixed_h, c.h_sigma, c.rho0,
                                                          c.height, c.B, c.gamma, c.Pb, c.co);
    OSPH_LAUNCH_CHECK();
    return 0;
}

int osph_launch_active_list(osph_ctx *ctx, int n_total, int *d_counters)
{
    if (n_total == 0) return 0;
    unsigned int *flag = ctx->key[0];
    k_active_flags<<<div_up(n_total, 256), 256, 0, ctx->stream>>>(ctx->d_aos, ctx->stride, n_total, flag); OSPH_LAUNCH_CHECK();
    int rc = osph_scan_exclusive(ctx, flag, n_total);
    if (rc) return rc;
    k_active_compact<<<div_up(n_total, 256), 256, 0, ctx->stream>>>(ctx->d_aos, ctx->stride, n_total, flag, ctx->d_row,
                                                                   ctx->d_act, d_counters);
    OSPH_LAUNCH_CHECK();
    return 0;
}

int osph_launch_set_deleted(osph_ctx *ctx, const unsigned char *d_active)
{
    k_set_deleted<<<div_up(ctx->n_total, 256), 256, 0, ctx->stream>>>(ctx->d_aos, ctx->stride, (int)ctx->n_total, d_active);
    OSPH_LAUNCH_CHECK();
    return 0;
}

int osph_launch_unpack(osph_ctx *ctx)
{
    if (ctx->n == 0) return 0;
    k_unpack_aos<<<div_up(ctx->n, 128), 128, 0, ctx->stream>>>(ctx->d_aos, ctx->stride, ctx->d_row, (int)ctx->n,
                                                                columns_of(ctx), ctx->label);
    OSPH_LAUNCH_CHECK();
    return 0;
}

int osph_launch_pack(osph_ctx *ctx)
{
    if (ctx->n == 0) return 0;
    k_pack_aos<<<div_up(ctx->n, PACK_ROWS), PACK_ROWS, 0, ctx->stream>>>(ctx->d_aos, ctx->stride, ctx->d_row, (int)ctx->n,
                                                                          columns_of(ctx), ctx->label, ctx->c_uniform ? 1 : 0,
                                                                          ctx->cfg.co, 0, nullptr);
    OSPH_LAUNCH_CHECK();
    return 0;
}

int osph_launch_pack_owned(osph_ctx *ctx, int *d_ids)
{
    if (ctx->n == 0) return 0;
    k_pack_aos<<<div_up(ctx->n, PACK_ROWS), PACK_ROWS, 0, ctx->stream>>>(ctx->d_aos, ctx->stride, ctx->d_row, (int)ctx->n,
                                                                          columns_of(ctx), ctx->label, ctx->c_uniform ? 1 : 0,
                                                                          ctx->cfg.co, 1, d_ids);
    OSPH_LAUNCH_CHECK();
    return 0;
}

static double pair_radius_q(const osph_ctx *ctx);

int osph_launch_prepare(osph_ctx *ctx, bool predict, double dt, double damping, bool use_dev_dt, bool skip_reset, int fused,
                        const BuildPlan *grid_plan, bool grid_reset_dt)
{
    PrepareArgs a;
    a.n = (int)ctx->n; a.label = ctx->label;
    a.x = ctx->f[OSPH_F_X]; a.y = ctx->f[OSPH_F_Y]; a.vx = ctx->f[OSPH_F_VX]; a.vy = ctx->f[OSPH_F_VY];
    a.rho = ctx->f[OSPH_F_RHO]; a.h = ctx->f[OSPH_F_H]; a.m = ctx->f[OSPH_F_M];
    a.ax = ctx->f[OSPH_F_AX]; a.ay = ctx->f[OSPH_F_AY]; a.drho = ctx->f[OSPH_F_DRHO];
    a.xsphx = ctx->f[OSPH_F_XSPHX]; a.xsphy = ctx->f[OSPH_F_XSPHY];
    a.x0 = ctx->f[OSPH_F_X0]; a.y0 = ctx->f[OSPH_F_Y0]; a.vx0 = ctx->f[OSPH_F_VX0]; a.vy0 = ctx->f[OSPH_F_VY0];
    a.rho0 = ctx->f[OSPH_F_RHO0];
    a.sc = ctx->d_sc; a.dt = dt; a.damping = damping; a.fixed_h = ctx->cfg.fixed_h; a.h_sigma = ctx->cfg.h_sigma;
    a.use_dev_dt = use_dev_dt ? 1 : 0; a.integ_xsph = ctx->cfg.integrator_xsph; a.strict = ctx->cfg.strict;
    a.dynamic_h = ctx->cfg.dynamic_h;
    a.reduce_hmin_fluid = fused ? 1 : 0;
    a.rden = rden_for(damping);
    a.xref = ctx->xref; a.yref = ctx->yref;
    a.grid = nullptr;
    if (grid_plan) {
        a.grid = ctx->d_grid; a.g_nn_scale = ctx->cfg.nn_scale; a.g_pair_radius_q = pair_radius_q(ctx); a.g_r0 = ctx->cfg.r0;
        a.g_skin_frac = ctx->skin_frac; a.g_cell_cap = (long long)ctx->cell_cap; a.g_reset_dt = grid_reset_dt ? 1 : 0;
        a.g_force_sort = grid_plan->force;
    } else {
        a.g_nn_scale = a.g_pair_radius_q = a.g_r0 = a.g_skin_frac = 0.0; a.g_cell_cap = 0; a.g_reset_dt = a.g_force_sort = 0;
    }
    if (!skip_reset) { k_reset_prepare_scalars<<<1, 1, 0, ctx->stream>>>(ctx->d_sc); OSPH_LAUNCH_CHECK(); }
    int grid = div_up(ctx->n, 256 * PREP_ITEMS);
    int integ = ctx->cfg.integrator;
    if (fused == 2) {
        if (!(predict && integ == OSPH_INTEGRATOR_PEC)) { ctx->err = "fused corrector + predictor exists for PEC only"; return OSPH_E_INVALID; }
        k_prepare<OSPH_INTEGRATOR_PEC, true, true><<<grid, 256, 0, ctx->stream>>>(a);
    }
    else if (predict && integ == OSPH_INTEGRATOR_PEC) k_prepare<OSPH_INTEGRATOR_PEC, true, false><<<grid, 256, 0, ctx->stream>>>(a);
    else if (predict && integ == OSPH_INTEGRATOR_VERLET) k_prepare<OSPH_INTEGRATOR_VERLET, true, false><<<grid, 256, 0, ctx->stream>>>(a);
    else k_prepare<OSPH_INTEGRATOR_PEC, false, false><<<grid, 256, 0, ctx->stream>>>(a);
    OSPH_LAUNCH_CHECK();
    return 0;
}

static double pair_radius_q(const osph_ctx *ctx) { return ctx->cfg.kernel == OSPH_KERNEL_GAUSSIAN ? 3.0 : 2.0; }

// Physical reorder of the owned state into sorted (cell) order, right after the sort.  `perm` maps new slot ->
// old slot.  Afterwards idx[] is rewritten so that it refers to the new slots.
__global__ void k_owned_flags(const unsigned int *__restrict__ idx, int n_all, int n_owned, unsigned int *__restrict__ flag)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_all) flag[s] = idx[s] < (unsigned int)n_owned ? 1u : 0u;
}
__global__ void k_owned_perm(unsigned int *__restrict__ idx, const unsigned int *__restrict__ pos, int n_all, int n_owned,
                             unsigned int *__restrict__ perm)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_all) return;
    unsigned int i = idx[s];
    if (i < (unsigned int)n_owned) { perm[pos[s]] = i; idx[s] = pos[s]; }
}

static int reorder_state(osph_ctx *ctx)
{
    unsigned int *idx = ctx->idx[ctx->sorted_buf];
    unsigned int *perm = ctx->idx[ctx->sorted_buf ^ 1];          // the other sort buffer is free after the sort
    int n = (int)ctx->n, n_all = (int)(ctx->n + ctx->n_ghost), grid = div_up(ctx->n, 256);
    if (ctx->n_ghost == 0) {
        OSPH_CUDA(cudaMemcpyAsync(perm, idx, sizeof(unsigned int) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        k_iota<<<grid, 256, 0, ctx->stream>>>(idx, n); OSPH_LAUNCH_CHECK();
    } else {
        unsigned int *flag = ctx->key[ctx->sorted_buf ^ 1];
        k_owned_flags<<<div_up(n_all, 256), 256, 0, ctx->stream>>>(idx, n_all, n, flag); OSPH_LAUNCH_CHECK();
        int rc = osph_scan_exclusive(ctx, flag, n_all);
        if (rc) return rc;
        k_owned_perm<<<div_up(n_all, 256), 256, 0, ctx->stream>>>(idx, flag, n_all, n, perm); OSPH_LAUNCH_CHECK();
    }
    for (int k = 0; k < OSPH_NUM_FIELDS; k++) {
        k_permute<double><<<grid, 256, 0, ctx->stream>>>(ctx->f[k], perm, n, ctx->scratch); OSPH_LAUNCH_CHECK();
        double *t = ctx->f[k]; ctx->f[k] = ctx->scratch; ctx->scratch = t;
    }
    // label / row / act go through the same spare column (it is 8 bytes wide)
    k_permute<signed char><<<grid, 256, 0, ctx->stream>>>(ctx->label, perm, n, (signed char *)ctx->scratch); OSPH_LAUNCH_CHECK();
    OSPH_CUDA(cudaMemcpyAsync(ctx->label, ctx->scratch, (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    k_permute<int><<<grid, 256, 0, ctx->stream>>>(ctx->d_row, perm, n, (int *)ctx->scratch); OSPH_LAUNCH_CHECK();
    OSPH_CUDA(cudaMemcpyAsync(ctx->d_row, ctx->scratch, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    k_permute<int><<<grid, 256, 0, ctx->stream>>>(ctx->d_act, perm, n, (int *)ctx->scratch); OSPH_LAUNCH_CHECK();
    OSPH_CUDA(cudaMemcpyAsync(ctx->d_act, ctx->scratch, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    // the positions the particles were binned at (sort cadence), in the new storage order: this build sorted, so they
    // are the current positions
    OSPH_CUDA(cudaMemcpyAsync(ctx->xref, ctx->f[OSPH_F_X], sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    OSPH_CUDA(cudaMemcpyAsync(ctx->yref, ctx->f[OSPH_F_Y], sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

int osph_launch_grid_params(osph_ctx *ctx, bool reset_dt, int force_sort)
{
    k_grid_params<<<1, 1, 0, ctx->stream>>>(ctx->d_sc, ctx->d_grid, ctx->cfg.nn_scale, pair_radius_q(ctx), ctx->cfg.r0,
                                            (long long)ctx->cell_cap, reset_dt ? 1 : 0, force_sort,
                                            (ctx->slab && ctx->slab_cadence_force) ? ctx->slab_skin : ctx->skin_frac);
    OSPH_LAUNCH_CHECK();
    return 0;
}

BuildPlan osph_plan_build(osph_ctx *ctx)
{
    BuildPlan p;
    int every = ctx->cfg.reorder_every > 0 ? ctx->cfg.reorder_every : 32;
    // physical re-sort on the 3rd build after an upload, then every `every` builds: a caller that uploads, steps once
    // and downloads (the host-buffer plugin call pattern) never pays for it, a resident run gets it early
    p.reorder_now = ctx->build_counter % every == (2 % every);
    // sort cadence (k_grid_params): the device may reuse the last binning unless the particle set was touched from outside
    // since (skin_valid), the state is about to be reordered physically, or the build always sorts (slab mode: the ghost
    // set changes every step; radix path)
    p.always = ctx->slab || ctx->n_ghost > 0 || !ctx->bin_sort || ctx->skin_frac == 0.0;
    p.force = p.always ? 2 : ((!ctx->skin_valid || p.reorder_now) ? 1 : 0);
    if (ctx->slab && ctx->slab_cadence_force && ctx->bin_sort) {
        // slab cadence: the sequencer tells every build whether it sorts (1) or reuses (4); no physical reorder (the frozen
        // halo lists hold storage slots)
        p.always = false; p.reorder_now = false; p.force = ctx->slab_cadence_force;
    }
    p.grid_done = false;
    return p;
}

int osph_launch_build(osph_ctx *ctx, bool reset_dt, const BuildPlan *given)
{
    int n = (int)ctx->n, n_all = (int)(ctx->n + ctx->n_ghost), grid = div_up(n_all, 256);
    const BuildPlan plan = given ? *given : osph_plan_build(ctx);
    const bool reorder_now = plan.reorder_now, always = plan.always;
    int rc = 0;
    if (!plan.grid_done && (rc = osph_launch_grid_params(ctx, reset_dt, plan.force))) return rc;
    ctx->skin_valid = !always;
    ctx->sorted_buf = 0;
    const double *px = ctx->f[OSPH_F_X], *py = ctx->f[OSPH_F_Y];
    bool rank_in_gather = false;
    if (ctx->bin_sort && plan.force == 4) {
        // host-planned reuse (slab cadence): nothing to launch, the gather kernel takes the stored permutation
    } else if (ctx->bin_sort) {
        // counting sort by cell: histogram + arrival ranks, then scan (= cell table), scatter, canonical order (binsort.cu)
        k_bin_keys<<<div_up(n_all, 256 * BINK_ITEMS), 256, 0, ctx->stream>>>(px, py, n, ctx->d_ghost, ctx->gmap, n_all, ctx->d_grid,
                                                                            ctx->d_sc, ctx->key[0], ctx->idx[0], ctx->bin_counts, ctx->bin_tiles,
                                                                            (long long)ctx->bin_tile_cap, always ? nullptr : ctx->xref,
                                                                            always ? nullptr : ctx->yref);
        OSPH_LAUNCH_CHECK();
        if ((rc = osph_bin_sort(ctx, n_all, reorder_now))) return rc;
        rank_in_gather = !reorder_now;
    } else {
        const int nblocks = div_up(n_all, SORT_TILE), db = osph_sort_digit_bits(ctx->key_bits);
        if (db == 8)
            k_keys<8><<<nblocks, SORT_THREADS, 0, ctx->stream>>>(px, py, n, ctx->d_ghost, ctx->gmap, n_all, ctx->d_grid, ctx->d_sc,
                                                                 ctx->key[0], ctx->idx[0], nblocks, ctx->hist);
        else if (db == 10)
            k_keys<10><<<nblocks, SORT_THREADS, 0, ctx->stream>>>(px, py, n, ctx->d_ghost, ctx->gmap, n_all, ctx->d_grid, ctx->d_sc,
                                                                  ctx->key[0], ctx->idx[0], nblocks, ctx->hist);
        else
            k_keys<11><<<nblocks, SORT_THREADS, 0, ctx->stream>>>(px, py, n, ctx->d_ghost, ctx->gmap, n_all, ctx->d_grid, ctx->d_sc,
                                                                  ctx->key[0], ctx->idx[0], nblocks, ctx->hist);
        OSPH_LAUNCH_CHECK();
        if ((rc = osph_sort_pairs(ctx, n_all, ctx->key_bits, true))) return rc;
        OSPH_CUDA(cudaMemsetAsync(ctx->cell_range, 0, sizeof(int2) * (size_t)ctx->cell_cap, ctx->stream));
    }
    if (reorder_now && (rc = reorder_state(ctx))) return rc;
    ctx->build_counter++;

    GatherArgs g;
    g.n_owned = n; g.n_all = n_all; g.key = ctx->bin_sort ? nullptr : ctx->key[ctx->sorted_buf]; g.cell_range = ctx->cell_range;
    g.idx = ctx->idx[ctx->sorted_buf]; g.idx_arrival = ctx->idx[1];
    g.rank_ranges = rank_in_gather ? reinterpret_cast<const int2 *>(ctx->scratch) : nullptr; g.idx_out = ctx->idx[0]; g.sc = ctx->d_sc;
    g.label = ctx->label; g.ghost = ctx->d_ghost; g.gmap = ctx->gmap;
    g.x = ctx->f[OSPH_F_X]; g.y = ctx->f[OSPH_F_Y]; g.vx = ctx->f[OSPH_F_VX]; g.vy = ctx->f[OSPH_F_VY];
    g.rho = ctx->f[OSPH_F_RHO]; g.m = ctx->f[OSPH_F_M]; g.h = ctx->f[OSPH_F_H]; g.p = ctx->f[OSPH_F_P];
    g.gp = ctx->d_grid;
    g.s_pos = ctx->s_pos; g.s_info = ctx->s_info; g.s_coarse = ctx->s_coarse; g.s_gcell = ctx->s_gcell;
    g.gamma = ctx->cfg.gamma; g.B = ctx->cfg.B; g.rho0 = ctx->cfg.rho0; g.Pb = ctx->cfg.Pb;
    g.uh_h = (ctx->cfg.dynamic_h == OSPH_H_FIXED && ctx->cfg.fixed_h > 0.0) ? ctx->cfg.fixed_h : 0.0;
    if (ctx->cfg.precision == OSPH_FP64)
        k_gather<double2><<<grid, 256, 0, ctx->stream>>>(g, (double2 *)ctx->s_vel, (double2 *)ctx->s_rm, (double2 *)ctx->s_hp);
    else
        k_gather<float2><<<grid, 256, 0, ctx->stream>>>(g, (float2 *)ctx->s_vel, (float2 *)ctx->s_rm, (float2 *)ctx->s_hp);
    OSPH_LAUNCH_CHECK();
    return 0;
}

int osph_launch_correct(osph_ctx *ctx, bool correct, double dt, double damping, bool use_dev_dt, bool skip_reset)
{
    CorrectArgs a;
    a.n = (int)ctx->n; a.label = ctx->label;
    a.x = ctx->f[OSPH_F_X]; a.y = ctx->f[OSPH_F_Y]; a.vx = ctx->f[OSPH_F_VX]; a.vy = ctx->f[OSPH_F_VY];
    a.rho = ctx->f[OSPH_F_RHO]; a.h = ctx->f[OSPH_F_H]; a.c = ctx->f[OSPH_F_C];
    a.ax = ctx->f[OSPH_F_AX]; a.ay = ctx->f[OSPH_F_AY]; a.drho = ctx->f[OSPH_F_DRHO];
    a.xsphx = ctx->f[OSPH_F_XSPHX]; a.xsphy = ctx->f[OSPH_F_XSPHY];
    a.x0 = ctx->f[OSPH_F_X0]; a.y0 = ctx->f[OSPH_F_Y0]; a.vx0 = ctx->f[OSPH_F_VX0]; a.vy0 = ctx->f[OSPH_F_VY0];
    a.rho0 = ctx->f[OSPH_F_RHO0];
    a.sc = ctx->d_sc; a.dt = dt; a.damping = damping; a.co = ctx->cfg.co;
    a.use_dev_dt = use_dev_dt ? 1 : 0; a.integ_xsph = ctx->cfg.integrator_xsph; a.strict = ctx->cfg.strict;
    a.c_uniform = ctx->c_uniform ? 1 : 0;
    a.rden = rden_for(damping);
    if (!skip_reset) { k_reset_dt_scalars<<<1, 1, 0, ctx->stream>>>(ctx->d_sc); OSPH_LAUNCH_CHECK(); }
    int grid = div_up(ctx->n, 256);
    if (!correct) k_correct<OSPH_INTEGRATOR_PEC, false><<<grid, 256, 0, ctx->stream>>>(a);
    else if (ctx->cfg.integrator == OSPH_INTEGRATOR_PEC) k_correct<OSPH_INTEGRATOR_PEC, true><<<grid, 256, 0, ctx->stream>>>(a);
    else if (ctx->cfg.integrator == OSPH_INTEGRATOR_EULER) k_correct<OSPH_INTEGRATOR_EULER, true><<<grid, 256, 0, ctx->stream>>>(a);
    else k_correct<OSPH_INTEGRATOR_VERLET, true><<<grid, 256, 0, ctx->stream>>>(a);
    OSPH_LAUNCH_CHECK();
    return 0;
}

int osph_launch_timestep(osph_ctx *ctx, double fixed_dt, bool log, bool reset_prepare, const double *d_reduced3, int fused)
{
    k_timestep<<<1, 1, 0, ctx->stream>>>(ctx->d_sc, ctx->cfg.cfl_courant, ctx->cfg.cfl_force, fixed_dt,
                                         log ? ctx->d_dt_log : nullptr, (long long)ctx->dt_log_cap, reset_prepare ? 1 : 0,
                                         d_reduced3, fused, ctx->cfg.co);
    OSPH_LAUNCH_CHECK();
    return 0;
}

int osph_launch_ke(osph_ctx *ctx)
{
    int nb = div_up(ctx->n, 256);
    k_ke_partial<<<nb, 256, 0, ctx->stream>>>(ctx->f[OSPH_F_M], ctx->f[OSPH_F_VX], ctx->f[OSPH_F_VY], ctx->label,
                                              (int)ctx->n, ctx->d_partial);
    OSPH_LAUNCH_CHECK();
    k_ke_final<<<1, 256, 0, ctx->stream>>>(ctx->d_partial, nb, ctx->d_sc);
    OSPH_LAUNCH_CHECK();
    return 0;
}

static NeighbourArgs neighbour_args(osph_ctx *ctx)
{
    NeighbourArgs a;
    a.n = (int)ctx->n; a.idx = ctx->idx[ctx->sorted_buf]; a.act = ctx->d_act; a.h = ctx->f[OSPH_F_H];   // single-GPU validation path: no ghosts
    a.s_pos = ctx->s_pos; a.s_info = ctx->s_info; a.s_coarse = ctx->s_coarse; a.s_gcell = ctx->s_gcell;
    a.cell_range = ctx->cell_range; a.gp = ctx->d_grid; a.counts = nullptr; a.offsets = nullptr; a.out = nullptr;
    return a;
}

int osph_launch_neighbours(osph_ctx *ctx, int mode, long long *d_counts, const long long *d_offsets, long long *d_out)
{
    NeighbourArgs a = neighbour_args(ctx);
    a.counts = d_counts; a.offsets = d_offsets; a.out = d_out;
    k_neighbours<<<div_up(ctx->n, 128), 128, 0, ctx->stream>>>(a, mode);
    OSPH_LAUNCH_CHECK();
    return 0;
}

int osph_launch_near_pos(osph_ctx *ctx, double x, double y, double h, long long cap, long long *d_idx, double *d_r,
                         double *d_q, double *d_h, long long *d_count)
{
    NeighbourArgs a = neighbour_args(ctx);
    k_near_pos<<<1, 1, 0, ctx->stream>>>(a, x, y, h, cap, d_idx, d_r, d_q, d_h, d_count);
    OSPH_LAUNCH_CHECK();
    return 0;
}

int osph_launch_probe(osph_ctx *ctx, int npts, const double *d_x, const double *d_y, double h, double *d_rho, double *d_p)
{
    NeighbourArgs a = neighbour_args(ctx);
    const osph_config &c = ctx->cfg;
    const double *m = ctx->f[OSPH_F_M], *rho = ctx->f[OSPH_F_RHO];
    int grid = div_up(npts, 64);
    if (c.kernel == OSPH_KERNEL_CUBIC)
        k_probe_pressure<OSPH_KERNEL_CUBIC><<<grid, 64, 0, ctx->stream>>>(a, m, rho, ctx->label, npts, d_x, d_y, h, c.rho0, c.gamma, c.B, d_rho, d_p);
    else if (c.kernel == OSPH_KERNEL_WENDLAND)
        k_probe_pressure<OSPH_KERNEL_WENDLAND><<<grid, 64, 0, ctx->stream>>>(a, m, rho, ctx->label, npts, d_x, d_y, h, c.rho0, c.gamma, c.B, d_rho, d_p);
    else
        k_probe_pressure<OSPH_KERNEL_GAUSSIAN><<<grid, 64, 0, ctx->stream>>>(a, m, rho, ctx->label, npts, d_x, d_y, h, c.rho0, c.gamma, c.B, d_rho, d_p);
    OSPH_LAUNCH_CHECK();
    return 0;
}

int osph_launch_cells(osph_ctx *ctx, long long *d_out)
{
    k_cells_to_active<<<div_up(ctx->n, 256), 256, 0, ctx->stream>>>(ctx->s_coarse, ctx->idx[ctx->sorted_buf], ctx->d_act,
                                                                    (int)ctx->n, ctx->d_grid, d_out);
    OSPH_LAUNCH_CHECK();
    return 0;
}

int osph_launch_col_to_active(osph_ctx *ctx, int field, double *d_out)
{
    int fill = (field == OSPH_F_C && ctx->c_uniform) ? 1 : 0;
    k_col_to_active<<<div_up(ctx->n, 256), 256, 0, ctx->stream>>>(ctx->f[field], ctx->d_act, (int)ctx->n, d_out, fill,
                                                                  ctx->cfg.co);
    OSPH_LAUNCH_CHECK();
    return 0;
}

int osph_launch_fill(osph_ctx *ctx, double *d_col, double value)
{
    k_fill_c<<<div_up(ctx->n, 256), 256, 0, ctx->stream>>>(d_col, (int)ctx->n, value);
    OSPH_LAUNCH_CHECK();
    return 0;
}

int osph_launch_col_from_active(osph_ctx *ctx, int field, const double *d_in)
{
    if (field == OSPH_F_C && ctx->c_uniform) {      // materialise c before a caller overwrites it
        k_fill_c<<<div_up(ctx->n, 256), 256, 0, ctx->stream>>>(ctx->f[OSPH_F_C], (int)ctx->n, ctx->cfg.co);
        OSPH_LAUNCH_CHECK();
        ctx->c_uniform = false;
    }
    k_col_from_active<<<div_up(ctx->n, 256), 256, 0, ctx->stream>>>(ctx->f[field], ctx->d_act, (int)ctx->n, d_in);
    OSPH_LAUNCH_CHECK();
    return 0;
}
