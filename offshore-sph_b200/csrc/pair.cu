// pair.cu -- K7, the fused pair-interaction kernel of the WCSPH step.
//
// One launch replaces the whole main loop of the reference's `_loop`
// (src/Tools/SolverTools.py:143-173): neighbour query (NNLinkedList.nearPos, NNLinkedList.py:41-80),
// computeProps with its three kernel passes (SolverTools.py:8-44), Continuity (Continuity.py:5-17),
// Momentum + artificial viscosity + gravity (Momentum.py:6-57, WCSPH.py:151-169), the Lennard-Jones
// BoundaryForce (BoundaryForce.py:7-42) and XSPH (XSPH.py:6-31, WCSPH.py:171-189).
//
// Mapping.  Particles are sorted by acceleration-grid cell (row-major), so the candidates of a particle
// are three contiguous runs of the sorted arrays (cells cx-1..cx+1 of rows cy-1, cy, cy+1).  A CTA owns
// 256 CONSECUTIVE sorted particles; per row offset the union of its threads' runs is again one contiguous
// interval, which the CTA stages through shared memory in batches of OSPH_CAP_STAGE candidates with
// coalesced 16-byte loads.  Each thread then walks only its own sub-interval of the staged batch; threads
// of one cell read the same shared-memory address (broadcast).  Outputs are scattered to the storage-order
// state columns.  No atomics, no neighbour list in memory, deterministic summation order.
//
// FP64 instantiation = validation mode: membership of a pair in the reference neighbour set is decided
// exactly (reference-cell adjacency on stored integer cell ids; r/h_ij <= 3.0 re-evaluated in strict IEEE
// when the cheap test is within 1e-13 of the threshold).  FP32 instantiation = performance mode: positions
// are staged relative to a per-CTA anchor (subtracted in double, then rounded), arithmetic in float.
#include "common.cuh"
#include "pair.cuh"

#include "sph_math.cuh"

template <typename Real, int KID, bool EXACT>
__global__ void __launch_bounds__(OSPH_PAIR_THREADS)
k_pair(PairArgs a)
{
    typedef typename R2<Real>::type Real2;
    constexpr int CAP = OSPH_CAP_STAGE;
    constexpr int NT = OSPH_PAIR_THREADS;
    __shared__ Real2 sh_pos[CAP];
    __shared__ Real2 sh_vel[CAP];
    __shared__ Real2 sh_rm[CAP];
    __shared__ Real2 sh_hp[CAP];
    __shared__ int sh_info[CAP];
    __shared__ int2 sh_cb[EXACT ? CAP : 1];
    __shared__ int sh_red[NT / 32][6];

    const Real2 *__restrict__ g_vel = reinterpret_cast<const Real2 *>(a.s_vel);
    const Real2 *__restrict__ g_rm = reinterpret_cast<const Real2 *>(a.s_rm);
    const Real2 *__restrict__ g_hp = reinterpret_cast<const Real2 *>(a.s_hp);

    const int tid = threadIdx.x;
    const int s0 = blockIdx.x * NT;
    const int s = s0 + tid;
    const bool valid = s < a.n;
    const GridParams *__restrict__ gp = a.gp;
    const int gnx = gp->gnx, gny = gp->gny;
    const Real pair_r2 = (Real)gp->pair_r2;

    // anchor for relative coordinates (FP32 mode): first particle of the CTA
    const double2 anchor = EXACT ? make_double2(0.0, 0.0) : a.s_pos[s0];

    Real xi = 0, yi = 0, vxi = 0, vyi = 0, rhoi = 1, hi = 0, slf = 0;
    int qcx = 0, qcy = 0;
    bool fluid_i = false;
    int ra[3], rb[3];
#pragma unroll
    for (int d = 0; d < 3; d++) { ra[d] = 0x7fffffff; rb[d] = 0; }
    if (valid) {
        double2 p = a.s_pos[s];
        xi = (Real)(p.x - anchor.x); yi = (Real)(p.y - anchor.y);
        Real2 v = g_vel[s]; vxi = v.x; vyi = v.y;
        Real2 rm = g_rm[s]; rhoi = rm.x;
        Real2 hp = g_hp[s]; hi = hp.x; slf = hp.y;
        fluid_i = (a.s_info[s] & 3) == 3;          // fluid AND owned (ghosts of a slab are sources only)
        if constexpr (EXACT) { int4 c = a.s_coarse[s]; qcx = c.z; qcy = c.w; }
        if (fluid_i) {
            int2 gc = a.s_gcell[s];
            int x0 = max(gc.x - 1, 0), x1 = min(gc.x + 1, gnx - 1);
#pragma unroll
            for (int d = 0; d < 3; d++) {
                int cy = gc.y + d - 1;
                if (cy < 0 || cy >= gny || x0 > x1) continue;
                const int2 *row = a.cell_range + (long long)cy * gnx;
                int lo = 0x7fffffff, hiE = 0;
                for (int cx = x0; cx <= x1; cx++) {
                    int2 r = row[cx];
                    if (r.y > r.x) { lo = min(lo, r.x); hiE = max(hiE, r.y); }
                }
                ra[d] = lo; rb[d] = hiE;
            }
        }
    }
    // CTA-wide union of the runs, per row offset
    {
        int lane = tid & 31, w = tid >> 5;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            int lo = warp_min_i(ra[d]), hiE = warp_max_i(rb[d]);
            if (lane == 0) { sh_red[w][d] = lo; sh_red[w][3 + d] = hiE; }
        }
        __syncthreads();
    }
    int ulo[3], uhi[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        int lo = 0x7fffffff, hiE = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) { lo = min(lo, sh_red[w][d]); hiE = max(hiE, sh_red[w][3 + d]); }
        ulo[d] = lo; uhi[d] = hiE;
    }

    const Real alpha_c = (Real)(a.alpha * a.c_half);      // alpha * 0.5 * c_i  (comp.c is never filled)
    const Real beta = (Real)a.beta;
    const Real r0 = (Real)a.r0, r0sq = r0 * r0;
    Real drho = 0, ax = 0, ay = 0, bx = 0, by = 0, xs = 0, ys = 0;

#pragma unroll 1
    for (int d = 0; d < 3; d++) {
#pragma unroll 1
        for (int base = ulo[d]; base < uhi[d]; base += CAP) {
            const int cnt = min(CAP, uhi[d] - base);
            __syncthreads();
            for (int t = tid; t < cnt; t += NT) {
                int g = base + t;
                double2 p = a.s_pos[g];
                Real2 pr; pr.x = (Real)(p.x - anchor.x); pr.y = (Real)(p.y - anchor.y);
                sh_pos[t] = pr;
                sh_vel[t] = g_vel[g];
                sh_rm[t] = g_rm[g];
                sh_hp[t] = g_hp[g];
                sh_info[t] = a.s_info[g];
                if constexpr (EXACT) { int4 c = a.s_coarse[g]; sh_cb[t] = make_int2(c.x, c.y); }
            }
            __syncthreads();
            if (!fluid_i) continue;
            const int j0 = max(ra[d], base) - base, j1 = min(rb[d], base + cnt) - base;
#pragma unroll 1
            for (int j = j0; j < j1; j++) {
                const Real2 pj = sh_pos[j];
                const Real dx = xi - pj.x, dy = yi - pj.y;
                const Real r2 = dx * dx + dy * dy;
                if (!(r2 <= pair_r2)) continue;
                const Real2 hpj = sh_hp[j];
                const Real hij = Real(0.5) * (hi + hpj.x);
                const bool fluid_j = (sh_info[j] & 1) != 0;
                const Real sup = (KID == OSPH_KERNEL_GAUSSIAN ? Real(3) : Real(2)) * hij;
                // Gaussian: the cut IS the set boundary, keep the band for the exact test below
                const bool kern = r2 <= sup * sup * (KID == OSPH_KERNEL_GAUSSIAN ? Real(1.0 + 1e-6) : Real(1));
                const bool lj = !fluid_j && r2 <= r0sq;
                if (!(kern || lj)) continue;
                // membership in the reference neighbour set: q <= 3 (matters for LJ and the Gaussian cut)
                {
                    const Real t9 = Real(9) * hij * hij;
                    if constexpr (EXACT) {
                        if (r2 > t9 * (1.0 - 1e-13)) {
                            if (r2 > t9 * (1.0 + 1e-13)) continue;
                            double rr = __dsqrt_rn(__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)));
                            if (!(__ddiv_rn(rr, (double)hij) <= 3.0)) continue;
                        }
                        const int2 cb = sh_cb[j];
                        if (abs(cb.x - qcx) > 1 || abs(cb.y - qcy) > 1) continue;
                    } else {
                        if (r2 > t9) continue;
                    }
                }
                const Real inv_rt = r2 > Real(1e-24) ? rsqrt_fast(r2) : Real(0);   // LJ guard: r > 1e-12
                const Real inv_r = r2 > Real(1e-20) ? inv_rt : Real(0);            // gradient guard: r >= 1e-10
                const Real r = r2 * inv_rt;
                const Real inv_h = rcp_fast(hij);
                const Real q = r * inv_h;
                Real w, g;
                sph_kernel<Real, KID>(q, inv_h, inv_r, w, g);
                const Real dwx = g * dx, dwy = g * dy;
                const Real2 vj = sh_vel[j];
                const Real dvx = vxi - vj.x, dvy = vyi - vj.y;
                const Real2 rmj = sh_rm[j];
                const Real mj = rmj.y;
                Real inv_rbar = 0;
                if (fluid_j) {
                    drho += mj * (dvx * dwx + dvy * dwy);
                    const Real dot = dvx * dx + dvy * dy;
                    Real PIij = 0;
                    if (dot < Real(0)) {
                        inv_rbar = rcp_fast(Real(0.5) * (rhoi + rmj.x));
                        const Real hbar = Real(0.5) * (hi + hij);          // h averaged twice (Momentum.py:43)
                        const Real mu = hbar * dot * rcp_fast(r2 + Real(0.01) * hbar * hbar);
                        PIij = mu * (beta * mu - alpha_c) * inv_rbar;
                    }
                    const Real fac = mj * (slf + hpj.y + PIij);
                    ax -= fac * dwx; ay -= fac * dwy;
                } else if (lj && r2 > Real(1e-24)) {
                    const Real frac = r0 * inv_rt;
                    Real tmp;
                    if (a.lj_42) { const Real f2 = frac * frac; tmp = f2 * f2 - f2; }
                    else tmp = pow_gen(frac, (Real)a.p1) - pow_gen(frac, (Real)a.p2);
                    const Real fac = (Real)a.D * tmp * inv_rt * inv_rt;
                    bx += fac * dx; by += fac * dy;
                }
                if (a.method_xsph && mj != Real(0)) {
                    if (inv_rbar == Real(0)) inv_rbar = rcp_fast(Real(0.5) * (rhoi + rmj.x));
                    const Real fac = -(Real)a.eps * mj * w * inv_rbar;
                    xs += fac * dvx; ys += fac * dvy;
                }
            }
        }
    }

    if (fluid_i) {
        const int slot = (int)a.idx[s];
        a.drho[slot] = a.summation_density ? 0.0 : (double)drho;
        a.ax[slot] = (double)ax + (double)bx;
        a.ay[slot] = ((double)ay - a.gravity) + (double)by;
        if (a.method_xsph) {
            a.xsphx[slot] = a.vx[slot] + (double)xs;
            a.xsphy[slot] = a.vy[slot] + (double)ys;
        } else {
            a.xsphx[slot] = 0.0; a.xsphy[slot] = 0.0;
        }
    }
}

template <typename Real, bool EXACT>
static int launch_kid(osph_ctx *ctx, const PairArgs &a, int grid)
{
    switch (ctx->cfg.kernel) {
    case OSPH_KERNEL_CUBIC:
        k_pair<Real, OSPH_KERNEL_CUBIC, EXACT><<<grid, OSPH_PAIR_THREADS, 0, ctx->stream>>>(a); break;
    case OSPH_KERNEL_WENDLAND:
        k_pair<Real, OSPH_KERNEL_WENDLAND, EXACT><<<grid, OSPH_PAIR_THREADS, 0, ctx->stream>>>(a); break;
    default:
        k_pair<Real, OSPH_KERNEL_GAUSSIAN, EXACT><<<grid, OSPH_PAIR_THREADS, 0, ctx->stream>>>(a); break;
    }
    return 0;
}

int osph_launch_pair(osph_ctx *ctx)
{
    PairArgs a;
    a.n = (int)(ctx->n + ctx->n_ghost);
    a.idx = ctx->idx[ctx->sorted_buf];
    a.s_pos = ctx->s_pos; a.s_vel = ctx->s_vel; a.s_rm = ctx->s_rm; a.s_hp = ctx->s_hp;
    a.s_info = ctx->s_info; a.s_coarse = ctx->s_coarse; a.s_gcell = ctx->s_gcell;
    a.cell_range = ctx->cell_range; a.gp = ctx->d_grid;
    a.vx = ctx->f[OSPH_F_VX]; a.vy = ctx->f[OSPH_F_VY];
    a.drho = ctx->f[OSPH_F_DRHO]; a.ax = ctx->f[OSPH_F_AX]; a.ay = ctx->f[OSPH_F_AY];
    a.xsphx = ctx->f[OSPH_F_XSPHX]; a.xsphy = ctx->f[OSPH_F_XSPHY];
    const osph_config &c = ctx->cfg;
    a.alpha = c.alpha; a.beta = c.beta; a.c_half = 0.5 * c.co; a.eps = c.epsilon;
    a.r0 = c.r0; a.D = c.D; a.p1 = c.p1; a.p2 = c.p2; a.gravity = c.gravity;
    a.lj_42 = (c.p1 == 4.0 && c.p2 == 2.0) ? 1 : 0;
    a.method_xsph = c.method_xsph; a.summation_density = c.summation_density;
    int grid = div_up(ctx->n + ctx->n_ghost, OSPH_PAIR_THREADS);
    const bool timed = ctx->time_pair && ctx->pair_ev_used < OSPH_PAIR_EVENTS;
    if (timed) cudaEventRecord(ctx->pair_ev[2 * ctx->pair_ev_used], ctx->stream);
    if (c.precision == OSPH_FP64) launch_kid<double, true>(ctx, a, grid);
    else launch_kid<float, false>(ctx, a, grid);
    OSPH_LAUNCH_CHECK();
    if (timed) { cudaEventRecord(ctx->pair_ev[2 * ctx->pair_ev_used + 1], ctx->stream); ctx->pair_ev_used++; }
    return 0;
}
