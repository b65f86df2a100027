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
// interval.  The three intervals are staged ONCE into shared memory as array-of-structure records (PAIR_CAP of
// them; longer runs fall back to batches) with coalesced 16-byte loads.  Each thread then walks only its own
// sub-intervals in two phases, warp-synchronously: a cheap scan that appends accepted record indices to a
// per-thread list, and a flush that evaluates the listed pairs with (almost) all lanes active.  Threads of one
// cell read the same shared-memory address (broadcast).  Outputs are scattered to the storage-order state
// columns.  No atomics on the sums, no neighbour list in memory, deterministic summation order.
//
// FP64 instantiation = validation mode: membership of a pair in the reference neighbour set is decided
// exactly (reference-cell adjacency on stored integer cell ids wherever distance does not already imply it;
// r/h_ij <= 3.0 re-evaluated in strict IEEE when the cheap test is within 1e-13 of the threshold).  FP32
// instantiation = performance mode: positions are staged relative to a per-CTA anchor (subtracted in double,
// then rounded), arithmetic in float.
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "pair.cuh"

#include "sph_math.cuh"

#ifndef PAIR_LIST64
#define PAIR_LIST64 48        // per-thread list of accepted candidates (16-bit shared-memory indices), double
#endif
#ifndef PAIR_LIST32
#define PAIR_LIST32 48        // same, float instantiation
#endif
#ifndef PAIR_MINB64
#define PAIR_MINB64 2         // __launch_bounds__ min CTAs per SM, double
#endif
#ifndef PAIR_MINB32
#define PAIR_MINB32 3
#endif
#ifndef PAIR_SCAN
#define PAIR_SCAN 6           // candidates tested between two warp votes (4: 380 us, 6: 373 us, 8: 375 us at 1 M, FP64)
#endif
#ifndef PAIR_SCAN_ILP
#define PAIR_SCAN_ILP 1       // 1: the PAIR_SCAN loads and distance tests of one round are independent of each other
#endif
#ifndef PAIR_SCAN_F32
#define PAIR_SCAN_F32 1       // double instantiation: scan on float copies of the positions with a conservative radius
#endif
#ifndef PAIR_NO_FMAX
#define PAIR_NO_FMAX 1        // no clamp in front of the rsqrt: the r = 0 lanes are removed by the selects that follow
#endif
#ifndef PAIR_PREFETCH_IDX
#define PAIR_PREFETCH_IDX 1   // heavy loop: fetch the next list entry while the current pair is evaluated
#endif
#ifndef PAIR_SCAN2
#define PAIR_SCAN2 0          // 1: scan on PAIRS of candidates with the packed FP32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2): the
#endif                        // float copies of the positions are staged as (x0, x1, y0, y1) per two records, one 16-byte load and
                              // four packed instructions test two candidates (both precisions scan on these copies then)
#ifndef PAIR_LISTPTR
#define PAIR_LISTPTR 0        // 1 (with PAIR_SCAN2): the candidate list is appended through a running shared-memory pointer
#endif
#ifndef PAIR_CAP
#define PAIR_CAP 1024         // candidate records resident in shared memory at once
#endif
#ifndef PAIR_LEAN
#define PAIR_LEAN 1           // 1: the scan drops the self pair (it contributes exactly 0), and the r -> 0 guards of the reference and
#endif                        // the (2 - q)+ clamp leave the common path of the heavy body: a second, guarded copy of the body serves the
                              // pairs that came through the rarely taken branch (coincident particles, wall pairs outside the kernel
                              // support, irregularly binned particles, the Gaussian).  Common path per listed pair, FP64 cubic, by SASS:
                              // 131 -> 116 instructions, 80 -> 68 on the FP64 pipe, no spills.  0: one guarded body for every pair (the
                              // build all GPU measurements of round 1 were made with; tools/build_round2_variants.sh builds it as `lean0`)

#ifndef PAIR_ISIGN
#define PAIR_ISIGN 1          // 1: the two clamps of the common path (inner spline term, approaching pairs) are decided by the sign bit on
#endif                        // the ALU pipe instead of a DSETP on the FP64 pipe, and the r > 1e-12 guard of the wall force -- implied by
                              // the `tiny` test for every pair on the common path -- leaves it: three FP64-pipe instructions per pair less
// PAIR_UH (common.cuh; 1: built in): Solver(h=value) gives every fluid particle the same smoothing length (k_prepare writes
// it before every build), so for a fluid neighbour h_ij, 1/h_ij, the support test, h~ of the viscosity and the kernel
// normalisation are loop constants.  k_pair<.., UH = true> takes them from the constant bank (PairArgs::uh_*; k_uh_constants
// forms them with the operations of the general body, so the results are the same bits) and sends every non-fluid neighbour
// through the general bodies: 68 -> 54 FP64-pipe instructions and one MUFU less per common pair (FP64 cubic).

// Software-pipelined flush loop (PAIR_UH_PIPE: the uniform-h instantiation, PAIR_GEN_PIPE: the general one; bit 0 double, bit 1
// float).  Stage A of entry k + 1 -- record position, separation, r^2, the "common pair" test, rsqrt seed and first residual --
// is issued with stage B of entry k (everything else), so the front of one pair's dependency chain (shared-memory load, three
// FP64 operations, MUFU, two more: about a third of it) runs under the previous pair's tail.  At four warps per scheduler the
// kernel was bound by that chain, not by instruction count: the uniform-h body removed 18 % of the instructions and 3 % of
// the time, this loop 13 % more.  Pairs that need the bodies of `interact` (wall and gate neighbours, the rare branch) are
// written back into the list and evaluated after the common ones of the same flush: same pairs, same arithmetic per pair,
// the summation order of the rare ones changed.  Measured on the B200, pair kernel, 1 M particles, FP64 / FP32 (profiles/r02):
//   uniform-h instantiation   not pipelined 286.7 / 164.9 us, unrolled by 1: 250.5 / 170.8, by 2: 248.0 / 161.2, by 3: 268.9 / 165.2
//   general instantiation     not pipelined 292.5 / 171.4 us, unrolled by 1: 264.6 / 177.9, by 2: 273.3 / 169.8
// hence: uniform-h both precisions unrolled by two (two entries' chains share one basic block), general double only, by one.
#ifndef PAIR_UH_PIPE
#define PAIR_UH_PIPE 3
#endif
#ifndef PAIR_GEN_PIPE
#define PAIR_GEN_PIPE 1
#endif
#ifndef PAIR_PIPE_UNROLL
#define PAIR_PIPE_UNROLL 2
#endif
#ifndef PAIR_GEN_PIPE_UNROLL
#define PAIR_GEN_PIPE_UNROLL 1
#endif

// One staged candidate.  Array-of-structures in shared memory: a single address computation per candidate,
// every field at a compile-time offset.  40 B (float) / 80 B (double) keeps 8 / 16-byte vector alignment.
template <typename Real, bool EXACT> struct Rec;
template <> struct __align__(8) Rec<float, false> { float2 pos, vel, rm, hp; int info; int pad; };    // info carries the bin cell mod 2^14
template <> struct __align__(16) Rec<double, true> { double2 pos, vel, rm, hp; int info; int cbx, cby; int pad; };

template <typename Real, bool EXACT>
constexpr size_t pair_smem_bytes()
{
    return sizeof(Rec<Real, EXACT>) * PAIR_CAP + sizeof(unsigned short) * (sizeof(Real) == 8 ? PAIR_LIST64 : PAIR_LIST32) * OSPH_PAIR_THREADS +
           sizeof(int) * (OSPH_PAIR_THREADS / 32) * 6 + (EXACT ? 0 : sizeof(int2) * OSPH_PAIR_THREADS) +
           (PAIR_SCAN2 ? sizeof(float4) * ((PAIR_CAP + PAIR_SCAN + 2) / 2) : ((EXACT && PAIR_SCAN_F32) ? sizeof(float2) * (PAIR_CAP + PAIR_SCAN) : 0));
}

template <typename Real, int KID, bool EXACT, bool UH>
__global__ void __launch_bounds__(OSPH_PAIR_THREADS, sizeof(Real) == 8 ? PAIR_MINB64 : PAIR_MINB32)
k_pair(PairArgs a)
{
    typedef typename R2<Real>::type Real2;
    typedef Rec<Real, EXACT> RecT;
    constexpr int PAIR_LIST = sizeof(Real) == 8 ? PAIR_LIST64 : PAIR_LIST32;
    constexpr int CAP = PAIR_CAP;
    constexpr int NT = OSPH_PAIR_THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RecT *sh_rec = reinterpret_cast<RecT *>(smem_raw);
    unsigned short *sh_list = reinterpret_cast<unsigned short *>(smem_raw + sizeof(RecT) * CAP);
    int(*sh_red)[6] = reinterpret_cast<int(*)[6]>(smem_raw + sizeof(RecT) * CAP + sizeof(unsigned short) * PAIR_LIST * NT);
    constexpr bool SCAN2 = PAIR_SCAN2 != 0;
    constexpr bool SCANF = SCAN2 || (EXACT && PAIR_SCAN_F32);
    float2 *sh_pf = reinterpret_cast<float2 *>(smem_raw + sizeof(RecT) * CAP + sizeof(unsigned short) * PAIR_LIST * NT +
                                               sizeof(int) * (NT / 32) * 6);
    float4 *sh_pf2 = reinterpret_cast<float4 *>(sh_pf);          // SCAN2: (x0, x1, y0, y1) of records 2m and 2m + 1
    // float instantiation: the thread's reference query cell, read back only on the rarely taken set-membership branch (a
    // register pair, or a global load there, spills the 80-register build)
    int2 *sh_qcell = reinterpret_cast<int2 *>(reinterpret_cast<unsigned char *>(sh_pf) +
                                             (PAIR_SCAN2 ? sizeof(float4) * ((PAIR_CAP + PAIR_SCAN + 2) / 2) : 0));

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
    // float scan of the double instantiation: positions relative to the CTA's first particle; every coordinate is
    // off by at most 2^-24 of the domain extent, so a radius enlarged by four such errors cannot lose a pair (a
    // candidate accepted in excess is rejected by the exact tests of the heavy body)
    const double2 anchor_f = SCANF ? (EXACT ? a.s_pos[s0] : anchor) : make_double2(0.0, 0.0);
    float xf = 0.f, yf = 0.f, thr_f = 0.f;
    if constexpr (SCANF) {
        // (float instantiation: the heavy body works on the same float coordinates; the margin only has to cover the
        // different rounding order of the packed instructions, and pair_r2 itself exceeds the largest support by 2e-6)
        const double ext = EXACT ? fmax(gp->xmax - gp->xmin, gp->ymax - gp->ymin) : 0.0;
        const double rad = sqrt(gp->pair_r2) + 4.0 * 5.97e-8 * ext;
        thr_f = __double2float_ru(rad * rad * (1.0 + 1e-6));
    }

    Real xi = 0, yi = 0, vxi = 0, vyi = 0, rhoi = 1, hi = 0, slf = 0, hi_half = 0, rhoi_half = Real(0.5);
    int qcx = 0, qcy = 0, info_i = 0;
    bool fluid_i = false;
    const bool need_adj = gp->regime_a != 0 || gp->adj_always != 0;
    bool adj_i = need_adj;                      // the reference-cell test is due for every pair of this particle
    int ra0 = 0x7fffffff, ra1 = 0x7fffffff, ra2 = 0x7fffffff, rb0 = 0, rb1 = 0, rb2 = 0;
    if (valid) {
        // every load of the prologue is issued before the first use: two global round trips (state + cell id, then
        // the nine cell ranges) instead of the four a test-then-load order costs
        const double2 p = a.s_pos[s];
        const Real2 v = g_vel[s], rm = g_rm[s], hp = g_hp[s];
        info_i = a.s_info[s];
        const int2 gc = a.s_gcell[s];
        // (the reference cells are only compared where adjacency does not follow from distance: regime A, the one-cell
        // fallback, irregularly binned particles; the float instantiation loads them only then)
        if constexpr (EXACT) { int4 c = a.s_coarse[s]; qcx = c.z; qcy = c.w; }
        else if (need_adj || (info_i & 4)) { int4 c = a.s_coarse[s]; sh_qcell[tid] = make_int2(c.z, c.w); }    // own slot: no barrier needed
        xi = (Real)(p.x - anchor.x); yi = (Real)(p.y - anchor.y);
        if constexpr (SCANF) { xf = (float)(p.x - anchor_f.x); yf = (float)(p.y - anchor_f.y); }
        vxi = v.x; vyi = v.y; rhoi = rm.x; hi = hp.x; slf = hp.y;
        hi_half = Real(0.5) * hi; rhoi_half = Real(0.5) * rhoi;
        fluid_i = (info_i & 3) == 3;               // fluid AND owned (ghosts of a slab are sources only)
        adj_i = need_adj || (info_i & 4);
        const int x0 = max(gc.x - 1, 0), x1 = min(gc.x + 1, gnx - 1);
        int2 cr[9];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const int cyc = min(max(gc.y + d - 1, 0), gny - 1);
            const int2 *row = a.cell_range + (long long)cyc * gnx;
#pragma unroll
            for (int e = 0; e < 3; e++) cr[3 * d + e] = row[min(max(x0 + e, 0), gnx - 1)];
        }
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const int cy = gc.y + d - 1;
            const bool rowok = fluid_i && cy >= 0 && cy < gny;
            int lo = 0x7fffffff, hiE = 0;
#pragma unroll
            for (int e = 0; e < 3; e++) {
                const int2 r = cr[3 * d + e];
                if (rowok && x0 + e <= x1 && r.y > r.x) { lo = min(lo, r.x); hiE = max(hiE, r.y); }
            }
            if (d == 0) { ra0 = lo; rb0 = hiE; } else if (d == 1) { ra1 = lo; rb1 = hiE; } else { ra2 = lo; rb2 = hiE; }
        }
    }
    // CTA-wide union of the runs, per row offset
    {
        int lane = tid & 31, w = tid >> 5;
        int lo0 = warp_min_i(ra0), lo1 = warp_min_i(ra1), lo2 = warp_min_i(ra2);
        int hi0 = warp_max_i(rb0), hi1 = warp_max_i(rb1), hi2 = warp_max_i(rb2);
        if (lane == 0) {
            sh_red[w][0] = lo0; sh_red[w][1] = lo1; sh_red[w][2] = lo2;
            sh_red[w][3] = hi0; sh_red[w][4] = hi1; sh_red[w][5] = hi2;
        }
        __syncthreads();
    }
    int ulo0 = 0x7fffffff, ulo1 = 0x7fffffff, ulo2 = 0x7fffffff, uhi0 = 0, uhi1 = 0, uhi2 = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) {
        ulo0 = min(ulo0, sh_red[w][0]); ulo1 = min(ulo1, sh_red[w][1]); ulo2 = min(ulo2, sh_red[w][2]);
        uhi0 = max(uhi0, sh_red[w][3]); uhi1 = max(uhi1, sh_red[w][4]); uhi2 = max(uhi2, sh_red[w][5]);
    }
    const int len0 = max(uhi0 - ulo0, 0), len1 = max(uhi1 - ulo1, 0), len2 = max(uhi2 - ulo2, 0);

    // alpha_c = alpha * 0.5 * c_i (comp.c is never filled); all of these come straight from the constant bank
#define PC(name) (sizeof(Real) == 8 ? (Real)a.name##_d : (Real)a.name##_f)
#define UHC(name) (sizeof(Real) == 8 ? (Real)a.uh_##name##_d : (Real)a.uh_##name##_f)
    const bool use_xsph = a.method_xsph != 0;
    Real drho = 0, ax = 0, ay = 0, xs = 0, ys = 0;          // the wall force is accumulated into (ax, ay) as well

    // stage sorted particles [g0, g0 + cnt) into records [dst, dst + cnt)
    int stage_adj = 0;                   // this thread staged a record of an irregularly binned particle
    auto stage_one = [&](const int g, const int dst) {
        RecT rec;
        double2 p = a.s_pos[g];
        rec.pos.x = (Real)(p.x - anchor.x); rec.pos.y = (Real)(p.y - anchor.y);
        if constexpr (SCAN2) {
            float *q = reinterpret_cast<float *>(sh_pf2 + (dst >> 1)) + (dst & 1);
            q[0] = (float)(p.x - anchor_f.x); q[2] = (float)(p.y - anchor_f.y);
        } else if constexpr (SCANF) sh_pf[dst] = make_float2((float)(p.x - anchor_f.x), (float)(p.y - anchor_f.y));
        rec.vel = g_vel[g]; rec.rm = g_rm[g]; rec.hp = g_hp[g]; rec.info = a.s_info[g];
        if constexpr (!EXACT) stage_adj |= rec.info & 4;
        rec.rm.x *= Real(0.5); rec.hp.x *= Real(0.5);       // staged as rho_j / 2 and h_j / 2: the pair means are one add
        if constexpr (EXACT) { int4 c = a.s_coarse[g]; rec.cbx = c.x; rec.cby = c.y; }
        rec.pad = 0;
        sh_rec[dst] = rec;
    };
    auto stage = [&](int g0, int cnt, int dst) {
        for (int t = tid; t < cnt; t += NT) stage_one(g0 + t, dst + t);
    };

    // One listed candidate: everything the reference evaluates per neighbour, fused.  Written as one straight
    // line (all shared-memory loads first, one early exit, the four reciprocal / rsqrt Newton chains independent of
    // each other, fluid-only terms masked instead of branched) so the scheduler can overlap the FP64 latencies;
    // only the rare paths (exact-predicate band, Lennard-Jones wall force) branch.
    // adj_tag: compile-time switch of the reference-cell test.  It is only due in CTAs that hold a particle the reference bins
    // irregularly, in regime A and on the one-cell fallback (cta_adj below); everywhere else the per-pair flag logic is
    // compiled out of the loop.
    auto interact = [&](auto adj_tag, const int j) {
        constexpr bool ADJ = decltype(adj_tag)::value;
        const RecT *__restrict__ rj = sh_rec + j;
        const Real2 pj = rj->pos;
        const Real2 hpj = rj->hp;
        const Real2 vj = rj->vel;
        const Real2 rmj = rj->rm;
        const int info_j = rj->info;
        int cbx = 0, cby = 0;
        if constexpr (EXACT) { cbx = rj->cbx; cby = rj->cby; }        // (float: read on the rare path)
        const Real dx = xi - pj.x, dy = yi - pj.y;
        const Real r2 = dx * dx + dy * dy;
        const bool fluid_j = (info_j & 1) != 0;
        // PAIR_LEAN: a pair closer than about 1.2e-10 (high word of r2 at or below that of 1e-20: a superset of the pairs
        // the two guards can bind for) joins the rare branch, an integer compare on the ALU pipe
#define PAIR_TINY() (PAIR_LEAN && (sizeof(Real) == 8 ? __double2hiint((double)r2) <= 0x3BC79CA1 : r2 <= Real(1.0001e-20)))
        auto body = [&](auto guarded_tag, auto uh_tag, const Real hij, const bool lj) {
        constexpr bool GUARDED = decltype(guarded_tag)::value;
        constexpr bool UHB = decltype(uh_tag)::value;              // fluid neighbour of the uniform-h kernel: loop constants
        constexpr bool ISG = PAIR_ISIGN && !GUARDED;
#if PAIR_NO_FMAX
        const Real rs = rsqrt_fast(r2);                            // r2 == 0: inf / NaN, discarded by the two selects below
#else
        const Real rs = rsqrt_fast(fmax(r2, Real(1e-30)));
#endif
        const Real inv_h = UHB ? UHC(inv_h) : rcp_fast(hij);
        const Real rbar = rhoi_half + rmj.x;                       // == 0.5 * (rho_i + rho_j)
        const Real inv_rbar = rcp_fast(rbar);
        const Real hbar = UHB ? UHC(h) : fma(Real(0.5), hij, hi_half);            // h averaged twice (Momentum.py:43)
        const Real inv_den = rcp_fast(UHB ? r2 + UHC(c01) * hbar : r2 + Real(0.01) * hbar * hbar);
        const Real inv_rt = !GUARDED || r2 > Real(1e-24) ? rs : Real(0);       // LJ guard: r > 1e-12
        const Real inv_r = !GUARDED || r2 > Real(1e-20) ? rs : Real(0);        // gradient guard: r >= 1e-10
        const Real r = r2 * inv_rt;
        const Real q = r * inv_h;
        Real w, g;
        if constexpr (UHB) {
            if constexpr (KID == OSPH_KERNEL_CUBIC) cubic_pair_a<Real, GUARDED, ISG>(q, UHC(alpha), inv_h, inv_r, w, g);
            else sph_kernel_a<Real, KID, GUARDED>(q, UHC(alpha), inv_h, inv_r, w, g);
        } else {
            if constexpr (KID == OSPH_KERNEL_CUBIC) cubic_pair<Real, GUARDED, ISG>(q, inv_h, inv_r, w, g);
            else sph_kernel<Real, KID, GUARDED>(q, inv_h, inv_r, w, g);
        }
        const Real dwx = g * dx, dwy = g * dy;
        const Real dvx = vxi - vj.x, dvy = vyi - vj.y;
        const Real mj = rmj.y;
        const Real mjf = (UHB || fluid_j) ? mj : Real(0);          // continuity and momentum: fluid neighbours only
        drho += mjf * (dvx * dwx + dvy * dwy);
        // artificial viscosity only for approaching pairs: min(dot, 0) makes it branch-free
        const Real dot = ISG ? neg_part_s(dvx * dx + dvy * dy) : neg_part(dvx * dx + dvy * dy);
        const Real mu = hbar * dot * inv_den;
        const Real PIij = mu * (PC(beta) * mu - PC(alpha_c)) * inv_rbar;
        const Real fac = mjf * (slf + hpj.y + PIij);
        ax -= fac * dwx; ay -= fac * dwy;
        if (use_xsph) {
            const Real fx = mj * w * inv_rbar;                     // -epsilon is applied once, to the sums
            xs += fx * dvx; ys += fx * dvy;
        }
        if constexpr (!UHB) {
        if (lj && (ISG || r2 > Real(1e-24))) {                     // wall / coupled particle inside r0 (ISG: not `tiny`, so r > 1e-10)
            const Real frac = PC(r0) * inv_rt;
            Real tmp;
            if (a.lj_42) { const Real f2 = frac * frac; tmp = f2 * f2 - f2; }
            else tmp = pow_gen(frac, (Real)a.p1) - pow_gen(frac, (Real)a.p2);
            const Real fl = PC(D) * tmp * inv_rt * inv_rt;
            ax += fl * dx; ay += fl * dy;
        }
        }
        };      // body
        if constexpr (UH && PAIR_LEAN && KID != OSPH_KERNEL_GAUSSIAN) {
            // uniform smoothing length: a fluid neighbour inside the (constant) support, no cell test due, not `tiny` -- the
            // common pair -- needs none of the per-pair h terms
            const bool adjq_u = ADJ && (adj_i || (info_j & 4));
            if (fluid_j && r2 <= UHC(h2c) && !PAIR_TINY() && !adjq_u) { body(std::false_type(), std::true_type(), Real(0), false); return; }
        }
        const Real hij = hi_half + hpj.x;                          // == 0.5 * (h_i + h_j) bit for bit
        const Real h2 = hij * hij;
        // kernel support (q <= 2, or the q <= 3 cut of the Gaussian, which IS the set boundary: keep a band for the
        // exact test) and Lennard-Jones range
        const bool kern = r2 <= h2 * (KID == OSPH_KERNEL_GAUSSIAN ? Real(9.0 * (1.0 + 1e-6)) : Real(4));
        const bool lj = !fluid_j && r2 <= PC(r0sq);
        // Membership in the reference neighbour set = adjacent reference cells AND q <= 3.
        //  * cells: where the acceleration grid is finer than the reference grid (regime B) two particles within the
        //    pair radius sit in adjacent reference cells by construction, unless the reference bins one of them
        //    irregularly (info bit 2); only then, and in regime A, the stored cell ids are compared;
        //  * q <= 3 can only bind for wall pairs outside the kernel support and at the cut of the Gaussian.
        // Everything that can reject a listed pair sits behind ONE rarely taken branch: inside the kernel support with
        // no cell test due, the pair is a member and the common path pays one compare and one predicate for it.
        bool via_rare = false;
        const bool tiny = PAIR_TINY();
        if constexpr (EXACT) {
            const bool adjq = ADJ && (adj_i || (info_j & 4));
            if (KID == OSPH_KERNEL_GAUSSIAN || adjq || !kern || tiny) {
                via_rare = true;
                bool ok = kern || lj;
                if (adjq) ok = ok && abs(cbx - qcx) <= 1 && abs(cby - qcy) <= 1;
                if (KID == OSPH_KERNEL_GAUSSIAN || !kern) {
                    ok = ok && r2 <= h2 * (9.0 * (1.0 + 1e-13));
                    if (ok && r2 > h2 * (9.0 * (1.0 - 1e-13))) {      // within 1e-13 of the threshold: decide in strict IEEE
                        double rr = __dsqrt_rn(__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)));
                        ok = __ddiv_rn(rr, (double)hij) <= 3.0;
                    }
                }
                if (!ok) return;
            }
        } else {
            // float instantiation: the same set rule (adjacent reference cells where distance does not imply it, q <= 3)
            // in float arithmetic -- without the cell test FP32 and FP64 mode would differ in WHICH pairs exist in regime A
            const bool adjq = ADJ && (adj_i || (info_j & 4));
            if (KID == OSPH_KERNEL_GAUSSIAN || adjq || !kern || tiny) {
                via_rare = true;
                bool ok = (kern || lj) && r2 <= h2 * Real(9);
                if (adjq && ok) {
                    const int2 qc = sh_qcell[tid];
                    // bin cell of j modulo 2^14 from its info word (k_gather); differences as signed 14-bit numbers
                    const int ddx = (((info_j >> 4) - qc.x) << 18) >> 18, ddy = (((int)((unsigned int)info_j >> 18) - qc.y) << 18) >> 18;
                    ok = abs(ddx) <= 1 && abs(ddy) <= 1 && !(info_j & 8);        // bit 3: the reference does not bin j at all
                }
                if (!ok) return;
            }
        }
        // GUARDED body: the r -> 0 guards of the reference and the clamp of the outer spline term are evaluated.  The default
        // build always is; with PAIR_LEAN only the pairs that came through the rare branch above are (coincident particles,
        // wall pairs outside the kernel support, the Gaussian) and the common path uses rs as it stands.
        if (PAIR_LEAN && !via_rare) body(std::false_type(), std::false_type(), hij, lj); else body(std::true_type(), std::false_type(), hij, lj);
    };

    // Two-phase walk, warp-synchronous.  (1) scan: cheap distance test over the thread's own sub-interval of
    // the staged records, accepted candidates appended to a per-thread list.  (2) flush: when any lane's list
    // is nearly full every lane evaluates its list.  The heavy body then runs with most lanes active instead
    // of the ~35-45% a fused test-and-evaluate loop achieves (profiles/r01).
    // The list is addressed by its write index li = (entries so far) * NT + tid: an append is one store and one add
    // (a separate entry count costs a multiply-add per append), the count is li / NT.
    int li = tid, slot = -1;
#if PAIR_LISTPTR && PAIR_SCAN2
    // the list is appended through a running 32-bit shared-window address: one store and one add per entry (written as
    // sh_list[li] the compiler forms base + 2 * li for every store; through a generic pointer it adds in 64 bits)
    const unsigned int la0 = (unsigned int)__cvta_generic_to_shared(sh_list + tid);
    unsigned int la = la0;
#define LIST_COUNT() ((int)((la - la0) / (2u * NT)))
#define LIST_RESET() (la = la0)
#define LIST_APPEND(v) do { sts_u16(la, (unsigned short)(v)); la += 2u * NT; } while (0)
#define LIST_NEARLY_FULL() (la >= la0 + 2u * (PAIR_LIST - PAIR_SCAN + 1) * NT)
#else
#define LIST_COUNT() (li / NT)
#define LIST_RESET() (li = tid)
#define LIST_APPEND(v) do { sh_list[li] = (unsigned short)(v); li += NT; } while (0)
#define LIST_NEARLY_FULL() (li >= (PAIR_LIST - PAIR_SCAN + 1) * NT)
#endif
    double vx_st = 0.0, vy_st = 0.0;
    bool have_v = false;
    auto flush = [&](auto adj_tag) {
#if (PAIR_UH_PIPE || PAIR_GEN_PIPE) && PAIR_LEAN
        constexpr int PIPE_BIT = sizeof(Real) == 8 ? 1 : 2;
        if constexpr (KID != OSPH_KERNEL_GAUSSIAN && (((UH ? PAIR_UH_PIPE : PAIR_GEN_PIPE) & PIPE_BIT) != 0)) {
            constexpr bool ADJ = decltype(adj_tag)::value;
            const int nl = LIST_COUNT();
            // One basic block per entry, so that the instruction scheduler interleaves stage A of entry k + 1 with stage B of
            // entry k (a warp issues in order: a stage A behind a branch of its own would simply run first).  Hence no
            // branches: stage A always runs (past the end it re-reads the last entry), stage B always runs, on operands made
            // harmless for an entry that is not a common pair -- r^2 := 1, h_ij := 1, m_j := 0, so every contribution is an
            // exact zero -- and such an entry goes back into the list, behind the read position, for the bodies of `interact`.
            // A common pair: fluid neighbour inside the kernel support, not `tiny`, no reference-cell test due.
            int wr = tid;
            int jn = 0;
            Real dxn = 0, dyn = 0, r2n = 1, yn = 0, en = 0, hijn = 1;
            bool qn = false;
            auto stage_a = [&](const int j) {
                const RecT *__restrict__ rj = sh_rec + j;
                const Real2 pj = rj->pos;
                const int info_j = rj->info;
                dxn = xi - pj.x; dyn = yi - pj.y;
                const Real r2 = dxn * dxn + dyn * dyn;
                Real thr = UHC(h2c), hij = 1;
                if constexpr (!UH) { hij = hi_half + rj->hp.x; thr = (hij * hij) * Real(4); }
                const bool adjq_u = ADJ && (adj_i || (info_j & 4));
                qn = (info_j & 1) != 0 && r2 <= thr && !PAIR_TINY() && !adjq_u;
                // (the seed is taken from r^2 as it stands and made harmless afterwards: the selects stay off the MUFU chain)
                Real y, e;
                rsqrt_begin(r2, y, e);
                r2n = qn ? r2 : Real(1); yn = qn ? y : Real(1); en = qn ? e : Real(0);
                if constexpr (!UH) hijn = qn ? hij : Real(1);
            };
            if (nl > 0) { jn = (int)sh_list[tid]; stage_a(jn); }
            const int last = (nl - 1) * NT + tid;
            int rd = tid;
            constexpr int PIPE_UNROLL = UH ? PAIR_PIPE_UNROLL : PAIR_GEN_PIPE_UNROLL;
#pragma unroll PIPE_UNROLL
            for (int k = 0; k < nl; k++) {
                const int j = jn;
                const Real dx = dxn, dy = dyn, r2 = r2n, y0 = yn, e0 = en, hij = hijn;
                const bool quick = qn;
                rd = min(rd + NT, last);
                jn = (int)sh_list[rd];
                stage_a(jn);
                // stage B: the common body -- the expressions of `body` (not GUARDED; UHB in the uniform-h instantiation), with rs
                // from its two halves: the same bits for the same pair
                const RecT *__restrict__ rj = sh_rec + j;
                const Real2 hpj = rj->hp;
                const Real2 vj = rj->vel;
                const Real2 rmj = rj->rm;
                const Real rs = rsqrt_end(y0, e0);
                const Real inv_h = UH ? UHC(inv_h) : rcp_fast(hij);
                const Real rbar = rhoi_half + rmj.x;
                const Real inv_rbar = rcp_fast(rbar);
                const Real hbar = UH ? UHC(h) : fma(Real(0.5), hij, hi_half);
                const Real inv_den = rcp_fast(UH ? r2 + UHC(c01) * hbar : r2 + Real(0.01) * hbar * hbar);
                const Real r = r2 * rs;
                const Real q = r * inv_h;
                Real w, g;
                if constexpr (UH) {
                    if constexpr (KID == OSPH_KERNEL_CUBIC) cubic_pair_a<Real, false, PAIR_ISIGN != 0>(q, UHC(alpha), inv_h, rs, w, g);
                    else sph_kernel_a<Real, KID, false>(q, UHC(alpha), inv_h, rs, w, g);
                } else {
                    if constexpr (KID == OSPH_KERNEL_CUBIC) cubic_pair<Real, false, PAIR_ISIGN != 0>(q, inv_h, rs, w, g);
                    else sph_kernel<Real, KID, false>(q, inv_h, rs, w, g);
                }
                const Real dwx = g * dx, dwy = g * dy;
                const Real dvx = vxi - vj.x, dvy = vyi - vj.y;
                const Real mj = quick ? rmj.y : Real(0);
                drho += mj * (dvx * dwx + dvy * dwy);
                const Real dot = PAIR_ISIGN ? neg_part_s(dvx * dx + dvy * dy) : neg_part(dvx * dx + dvy * dy);
                const Real mu = hbar * dot * inv_den;
                const Real PIij = mu * (PC(beta) * mu - PC(alpha_c)) * inv_rbar;
                const Real fac = mj * (slf + hpj.y + PIij);
                ax -= fac * dwx; ay -= fac * dwy;
                if (use_xsph) {
                    const Real fx = mj * w * inv_rbar;
                    xs += fx * dvx; ys += fx * dvy;
                }
                if (!quick) { sh_list[wr] = (unsigned short)j; wr += NT; }
            }
#pragma unroll 1
            for (int t = tid; t < wr; t += NT) interact(adj_tag, (int)sh_list[t]);
            LIST_RESET();
            return;
        }
#endif
#if PAIR_PREFETCH_IDX
        const int nl = LIST_COUNT();
        int jn = nl > 0 ? (int)sh_list[tid] : 0;
#pragma unroll 1
        for (int k = 0; k < nl;) {
            const int j = jn;
            k++;
            if (k < nl) jn = (int)sh_list[k * NT + tid];      // next index while this pair is evaluated
            interact(adj_tag, j);
        }
#else
        const int nl = LIST_COUNT();
#pragma unroll 1
        for (int k = 0; k < nl; k++) interact(adj_tag, (int)sh_list[k * NT + tid]);
#endif
        LIST_RESET();
    };
#if PAIR_SCAN2
    // Packed scan.  Candidates are tested two at a time: record pair m = (2m, 2m + 1) is one float4 (x0, x1, y0, y1), and
    //   d = (xi, xi) - (x0, x1);  e = (yi, yi) - (y0, y1);  r2 = d * d;  r2 = e * e + r2
    // are four FADD2 / FMUL2 / FFMA2 instructions.  A run [j0, j1) may start and end inside a pair: membership of a record in
    // the run is one unsigned compare (idx - j0 < j1 - j0).  Three pairs per round.
    const unsigned long long XF2 = pack_f32x2(xf, xf), YF2 = pack_f32x2(yf, yf);
    auto scan = [&](auto adj_tag, auto skip_tag, int j0, int j1, const int self) {          // all 32 lanes of a warp call this together
        constexpr bool SKIP = PAIR_LEAN && decltype(skip_tag)::value;
        constexpr int NP = PAIR_SCAN / 2;
        int m = j0 >> 1;
        unsigned int len = (unsigned int)(j1 - j0);
        bool warp_more = __any_sync(0xffffffffu, j0 < j1);
#pragma unroll 1
        while (warp_more) {
            unsigned long long d2[NP];
#pragma unroll
            for (int u = 0; u < NP; u++) {
                const float4 pq = sh_pf2[m + u];
                const unsigned long long dx = sub_f32x2(XF2, pack_f32x2(pq.x, pq.y)), dy = sub_f32x2(YF2, pack_f32x2(pq.z, pq.w));
                d2[u] = fma_f32x2(dy, dy, mul_f32x2(dx, dx));
            }
            const int rel = 2 * m - j0;
#pragma unroll
            for (int u = 0; u < NP; u++) {
                float r0v, r1v;
                unpack_f32x2(d2[u], r0v, r1v);
                const int i0 = 2 * (m + u), i1 = i0 + 1;
                if ((unsigned int)(rel + 2 * u) < len && r0v <= thr_f && (!SKIP || i0 != self)) LIST_APPEND(i0);
                if ((unsigned int)(rel + 2 * u + 1) < len && r1v <= thr_f && (!SKIP || i1 != self)) LIST_APPEND(i1);
            }
            m += NP;
            // a lane whose run is exhausted parks at pair 0 with an empty run (see the scalar scan below)
            if (2 * m >= j1) { m = 0; j0 = 0; j1 = 0; len = 0; }
            if (__any_sync(0xffffffffu, LIST_NEARLY_FULL())) flush(adj_tag);
            warp_more = __any_sync(0xffffffffu, 2 * m < j1);
        }
    };
#else
    // skip_tag (PAIR_LEAN only): the run holds the thread's own record at index `self`; it is not listed
    auto scan = [&](auto adj_tag, auto skip_tag, int j, int j1, const int self) {          // all 32 lanes of a warp call this together
        constexpr bool SKIP = PAIR_LEAN && decltype(skip_tag)::value;
        bool warp_more = __any_sync(0xffffffffu, j < j1);
#pragma unroll 1
        while (warp_more) {
#if PAIR_SCAN_ILP
            // PAIR_SCAN candidates per round: all loads first (a candidate past the end of the run is read and discarded:
            // the float copies are padded by PAIR_SCAN entries, the records are followed by the candidate lists, so the
            // address needs no clamp and the loads of a round share one base register), then the independent distance
            // tests, then the appends.  The
            // serial form (load, test, append, next) left the warp waiting on one shared-memory load and one
            // dependent FP chain at a time: 48 % of the kernel's stall samples on 25 % of its instructions.
            if constexpr (SCANF) {
                float d2[PAIR_SCAN];
#pragma unroll
                for (int u = 0; u < PAIR_SCAN; u++) {
                    const float2 pj = sh_pf[j + u];
                    const float dx = xf - pj.x, dy = yf - pj.y;
                    d2[u] = dx * dx + dy * dy;
                }
#pragma unroll
                for (int u = 0; u < PAIR_SCAN; u++)
                    if (j + u < j1 && d2[u] <= thr_f && (!SKIP || j + u != self)) { sh_list[li] = (unsigned short)(j + u); li += NT; }
            } else {
                Real d2[PAIR_SCAN];
#pragma unroll
                for (int u = 0; u < PAIR_SCAN; u++) {
                    const Real2 pj = sh_rec[j + u].pos;
                    const Real dx = xi - pj.x, dy = yi - pj.y;
                    d2[u] = dx * dx + dy * dy;
                }
#pragma unroll
                for (int u = 0; u < PAIR_SCAN; u++)
                    if (j + u < j1 && d2[u] <= pair_r2 && (!SKIP || j + u != self)) { sh_list[li] = (unsigned short)(j + u); li += NT; }
            }
            j += PAIR_SCAN;
            // a lane whose run is exhausted parks at record 0 (its tests are discarded by j + u < j1): the loads of the
            // rounds the other lanes still need stay inside the staged records whatever the run lengths are
            if (j >= j1) { j = 0; j1 = 0; }
#else
#pragma unroll
            for (int u = 0; u < PAIR_SCAN; u++) {
                if (j < j1) {
                    const Real2 pj = sh_rec[j].pos;
                    const Real dx = xi - pj.x, dy = yi - pj.y;
                    if (dx * dx + dy * dy <= pair_r2 && (!SKIP || j != self)) { sh_list[li] = (unsigned short)j; li += NT; }
                    j++;
                }
            }
#endif
            if (__any_sync(0xffffffffu, li >= (PAIR_LIST - PAIR_SCAN + 1) * NT)) flush(adj_tag);      // more than PAIR_LIST - PAIR_SCAN entries
            warp_more = __any_sync(0xffffffffu, j < j1);
        }
    };

#endif

    if (len0 + len1 + len2 <= CAP) {
        // common case: the three runs of the CTA fit in shared memory together; one staging pass, the
        // candidate list persists across the rows and is flushed only when full and once at the end
        const int o1 = len0, o2 = len0 + len1, total = o2 + len2;
        // one loop over the three runs: all of a thread's loads are in flight together
#pragma unroll 4
        for (int t = tid; t < total; t += NT)
            stage_one(t < o1 ? ulo0 + t : (t < o2 ? ulo1 + (t - o1) : ulo2 + (t - o2)), t);
        // the barrier that publishes the records also tells whether any pair of this CTA needs the reference-cell test
        // (float instantiation only: it is bound by instruction issue; the double one keeps a single copy of its loops, a
        // second one makes it spill)
        bool cta_adj = true;
        if constexpr (EXACT) __syncthreads();
        else cta_adj = __syncthreads_or((need_adj || stage_adj || (info_i & 4)) ? 1 : 0) != 0;
        // (an empty run is ra = INT_MAX, rb = 0: test it before doing index arithmetic on it)
        const bool h0 = fluid_i && rb0 > ra0, h1 = fluid_i && rb1 > ra1, h2 = fluid_i && rb2 > ra2;
#define PAIR_ROWS(adj_tag)                                                                                                             \
        do {                                                                                                                           \
            scan(adj_tag, std::false_type(), h0 ? ra0 - ulo0 : 0, h0 ? rb0 - ulo0 : 0, -1);                                            \
            scan(adj_tag, std::true_type(), h1 ? ra1 - ulo1 + o1 : 0, h1 ? rb1 - ulo1 + o1 : 0, s - ulo1 + o1);   /* holds s itself */   \
            if (fluid_i) slot = (int)a.idx[s];                       /* the epilogue's loads travel under the remaining work */         \
            scan(adj_tag, std::false_type(), h2 ? ra2 - ulo2 + o2 : 0, h2 ? rb2 - ulo2 + o2 : 0, -1);                                  \
            if constexpr (!EXACT) { if (fluid_i && a.method_xsph) { vx_st = a.vx[slot]; vy_st = a.vy[slot]; have_v = true; } }         \
            flush(adj_tag);                                                                                                            \
        } while (0)
        if constexpr (EXACT) PAIR_ROWS(std::true_type());
        else { if (cta_adj) PAIR_ROWS(std::true_type()); else PAIR_ROWS(std::false_type()); }
#undef PAIR_ROWS
    } else {
        // rare: a run longer than the buffer (very dense cells or sparse rows spanning the domain): batches
#pragma unroll 1
        for (int d = 0; d < 3; d++) {
            const int ulo = d == 0 ? ulo0 : (d == 1 ? ulo1 : ulo2), uhi = d == 0 ? uhi0 : (d == 1 ? uhi1 : uhi2);
            const int ra = d == 0 ? ra0 : (d == 1 ? ra1 : ra2), rb = d == 0 ? rb0 : (d == 1 ? rb1 : rb2);
#pragma unroll 1
            for (int base = ulo; base < uhi; base += CAP) {
                const int cnt = min(CAP, uhi - base);
                __syncthreads();
                stage(base, cnt, 0);
                __syncthreads();
                // the part of the thread's run that lies in THIS batch (empty for a run in an earlier or later batch)
                const int jb = max(ra, base), je = min(rb, base + cnt);
                const bool has = fluid_i && rb > ra && je > jb;
                scan(std::true_type(), std::true_type(), has ? jb - base : 0, has ? je - base : 0, d == 1 ? s - base : -1);
                flush(std::true_type());
            }
        }
    }

    if (a.a2max) {
        // fused step loop: the force criterion of the NEXT time step is reduced here, from the values stored below
        // (same operations as k_correct), so that no pass over the state is needed between two steps
        double a2 = -INFINITY;
        if (fluid_i) {
            const double axd = (double)ax, ayd = (double)ay - a.gravity;
            a2 = __dadd_rn(__dmul_rn(axd, axd), __dmul_rn(ayd, ayd));
            if (!(a2 == a2)) a2 = INFINITY;
        }
        a2 = warp_max(a2);
        if ((tid & 31) == 0 && a2 > -INFINITY) {
            const unsigned long long e = enc_f64(a2);
            if (e > *reinterpret_cast<volatile unsigned long long *>(a.a2max)) atomicMax(a.a2max, e);
        }
    }
    if (fluid_i) {
        if (slot < 0) slot = (int)a.idx[s];
        a.drho[slot] = a.summation_density ? 0.0 : (double)drho;
        a.ax[slot] = (double)ax;
        a.ay[slot] = (double)ay - a.gravity;
        if (a.method_xsph) {
            // xsph = v + correction (WCSPH.py:171-189).  Double instantiation: the sorted copy of v IS the state's v
            if constexpr (EXACT) { vx_st = (double)vxi; vy_st = (double)vyi; }
            else if (!have_v) { vx_st = a.vx[slot]; vy_st = a.vy[slot]; }
            a.xsphx[slot] = vx_st + (double)(PC(neg_eps) * xs);
            a.xsphy[slot] = vy_st + (double)(PC(neg_eps) * ys);
        } else {
            a.xsphx[slot] = 0.0; a.xsphy[slot] = 0.0;
        }
    }
    if (a.ts_sc) {
        // Fused time step of the next step: every CTA has reduced its max |a|^2 (the warps' atomics above) and min h came
        // from this step's predictor pass, so the CTA that finishes last computes dt -- no k_timestep launch between two
        // steps.  (The outputs stored above are not read by it.)
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned int ticket = atomicAdd(&a.ts_sc->pair_ticket, 1u);
            if (ticket == gridDim.x - 1) {
                a.ts_sc->pair_ticket = 0u;
                __threadfence();
                timestep_body(a.ts_sc, a.ts_gamma_c, a.ts_gamma_f, a.ts_fixed_dt, a.ts_dt_log, a.ts_dt_log_cap, 1, nullptr, a.ts_fused, a.ts_co);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Optional summation-density pass (reference src/Tools/SolverTools.py:129-140, SummationDensity.py:6-13):
//   rho_i = sum over FLUID neighbours j of m_j W(r_ij, h_ij),  neighbours = the reference set (q <= 3).
// The sum uses only m, W and h of the neighbours -- not their density -- so the reference's in-place, index-ordered
// loop is order-independent and this parallel pass reproduces it exactly.  p keeps the value formed from the old
// density (the reference evaluates the EOS before this pass); p / rho^2 is refreshed with the new density.
// Off in every shipped example: a plain one-thread-per-particle walk, always in double.
// ---------------------------------------------------------------------------------------------------------
// Slab mode: ghosts are sources of the pair kernel, and the pair kernel reads THEIR density too, so the pass also runs for
// the ghost fluid particles (their h, m and old density come from the wire records).  osph_slab_pack doubles the halo
// width when the option is on: a ghost within one pair radius of the face -- the only ghosts an owned particle can
// interact with -- then has its whole kernel support inside the halo and receives the same sum its owner computes.
struct SumDensArgs {
    int n_owned;
    const double *ghost;
    GhostMap gmap;
    const double *h64, *m64, *p_state;
    double *rho_state;
    double gamma, B, rho0, Pb;
};

template <int KID, typename Real2>
__global__ void __launch_bounds__(128)
k_summation_density(PairArgs a, SumDensArgs d, Real2 *__restrict__ s_rm, Real2 *__restrict__ s_hp)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n || !(a.s_info[s] & 1)) return;
    const GridParams g = *a.gp;
    const int slot = (int)a.idx[s];
    const bool owned = slot < d.n_owned;
    const double *rec_i = owned ? nullptr : ghost_record(d.ghost, d.gmap, slot - d.n_owned);
    const double2 pi = a.s_pos[s];
    const double hi = owned ? d.h64[slot] : rec_i[6];
    const int4 ci = a.s_coarse[s];
    const int2 gc = a.s_gcell[s];
    double rho = 0.0;
    const int reach = g.reach_set;
    for (int dy = -reach; dy <= reach; dy++) {
        int cy = gc.y + dy;
        if (cy < 0 || cy >= g.gny) continue;
        int x0 = max(gc.x - reach, 0), x1 = min(gc.x + reach, g.gnx - 1);
        for (int cx = x0; cx <= x1; cx++) {
            int2 r = a.cell_range[(long long)cy * g.gnx + cx];
            for (int t = r.x; t < r.y; t++) {
                if (!(a.s_info[t] & 1)) continue;
                const int4 cj = a.s_coarse[t];
                if (abs(cj.x - ci.z) > 1 || abs(cj.y - ci.w) > 1) continue;
                const int tj = (int)a.idx[t];
                double hj, mj;
                if (tj < d.n_owned) { hj = d.h64[tj]; mj = d.m64[tj]; }
                else { const double *rj = ghost_record(d.ghost, d.gmap, tj - d.n_owned); hj = rj[6]; mj = rj[5]; }
                const double2 pj = a.s_pos[t];
                const double dx = __dadd_rn(pi.x, -pj.x), dyy = __dadd_rn(pi.y, -pj.y);
                const double rr = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dyy, dyy)));
                const double hij = __dmul_rn(0.5, __dadd_rn(hi, hj));
                if (!(__ddiv_rn(rr, hij) <= 3.0)) continue;
                const double inv_h = 1.0 / hij;
                double w, gg;
                sph_kernel<double, KID>(rr * inv_h, inv_h, 0.0, w, gg);
                rho += mj * w;
            }
        }
    }
    double p_old;
    if (owned) { d.rho_state[slot] = rho; p_old = d.p_state[slot]; }
    else {
        // the owner's EOS value for the density the ghost arrived with (k_gather evaluates the same expression)
        const double ratio = rec_i[4] / d.rho0;
        double rp;
        if (d.gamma == 7.0) { const double r2 = ratio * ratio, r4 = r2 * r2; rp = r4 * r2 * ratio; } else rp = pow(ratio, d.gamma);
        p_old = (rp - 1.0) * d.B + d.Pb;
    }
    typedef decltype(Real2().x) Real;
    Real2 rm = s_rm[s]; rm.x = (Real)rho; s_rm[s] = rm;
    Real2 hp = s_hp[s]; hp.y = (Real)(p_old / (rho * rho)); s_hp[s] = hp;
}

template <typename Real2>
static void launch_summation(osph_ctx *ctx, const PairArgs &a)
{
    int grid = div_up(a.n, 128);
    Real2 *rm = (Real2 *)ctx->s_rm, *hp = (Real2 *)ctx->s_hp;
    SumDensArgs d;
    d.n_owned = (int)ctx->n; d.ghost = ctx->d_ghost; d.gmap = ctx->gmap;
    d.h64 = ctx->f[OSPH_F_H]; d.m64 = ctx->f[OSPH_F_M]; d.p_state = ctx->f[OSPH_F_P]; d.rho_state = ctx->f[OSPH_F_RHO];
    d.gamma = ctx->cfg.gamma; d.B = ctx->cfg.B; d.rho0 = ctx->cfg.rho0; d.Pb = ctx->cfg.Pb;
    switch (ctx->cfg.kernel) {
    case OSPH_KERNEL_CUBIC: k_summation_density<OSPH_KERNEL_CUBIC, Real2><<<grid, 128, 0, ctx->stream>>>(a, d, rm, hp); break;
    case OSPH_KERNEL_WENDLAND: k_summation_density<OSPH_KERNEL_WENDLAND, Real2><<<grid, 128, 0, ctx->stream>>>(a, d, rm, hp); break;
    default: k_summation_density<OSPH_KERNEL_GAUSSIAN, Real2><<<grid, 128, 0, ctx->stream>>>(a, d, rm, hp); break;
    }
}

template <typename Real, int KID, bool EXACT, bool UH>
static cudaError_t launch_one(const PairArgs &a, int grid, cudaStream_t stream, int device)
{
    static unsigned long long configured = 0;           // one bit per device ordinal (the attribute is per device)
    constexpr size_t smem = pair_smem_bytes<Real, EXACT>();
    if (!(configured >> (device & 63) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(k_pair<Real, KID, EXACT, UH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured |= 1ull << (device & 63);
    }
    k_pair<Real, KID, EXACT, UH><<<grid, OSPH_PAIR_THREADS, smem, stream>>>(a);
    return cudaSuccess;
}

template <typename Real, bool EXACT>
static int launch_kid(osph_ctx *ctx, const PairArgs &a, int grid, bool uh)
{
    cudaError_t e;
    switch (ctx->cfg.kernel) {
#if PAIR_UH
    case OSPH_KERNEL_CUBIC:
        e = uh ? launch_one<Real, OSPH_KERNEL_CUBIC, EXACT, true>(a, grid, ctx->stream, ctx->device)
               : launch_one<Real, OSPH_KERNEL_CUBIC, EXACT, false>(a, grid, ctx->stream, ctx->device);
        break;
    case OSPH_KERNEL_WENDLAND:
        e = uh ? launch_one<Real, OSPH_KERNEL_WENDLAND, EXACT, true>(a, grid, ctx->stream, ctx->device)
               : launch_one<Real, OSPH_KERNEL_WENDLAND, EXACT, false>(a, grid, ctx->stream, ctx->device);
        break;
#else
    case OSPH_KERNEL_CUBIC: e = launch_one<Real, OSPH_KERNEL_CUBIC, EXACT, false>(a, grid, ctx->stream, ctx->device); break;
    case OSPH_KERNEL_WENDLAND: e = launch_one<Real, OSPH_KERNEL_WENDLAND, EXACT, false>(a, grid, ctx->stream, ctx->device); break;
#endif
    default: e = launch_one<Real, OSPH_KERNEL_GAUSSIAN, EXACT, false>(a, grid, ctx->stream, ctx->device); break;
    }
    if (e != cudaSuccess) { ctx->err = std::string("pair kernel configuration: ") + cudaGetErrorString(e); return OSPH_E_CUDA; }
    return 0;
}

#if PAIR_UH
// Loop constants of the uniform-h instantiation, formed ON THE DEVICE with the operations the general body applies to every
// pair (the staged halves h_i / 2 + h_j / 2, rcp_fast, the kernel's own normalisation), so that a pair evaluated with them
// gives the bits the general body gives.  out: h_ij, 1 / h_ij, support threshold, 0.01 h~, alpha.
template <typename Real, int KID>
__global__ void k_uh_constants(double h, Real *out)
{
    const Real hh = Real(0.5) * (Real)h;                // k_pair: hi_half; staging: rec.hp.x *= 0.5
    const Real hij = hh + hh;
    const Real inv_h = rcp_fast(hij);
    const Real hbar = fma(Real(0.5), hij, hh);
    out[0] = hbar;                                      // == hij == h (exact halves): h~ of the viscosity
    out[1] = inv_h;
    out[2] = (hij * hij) * Real(4);
    out[3] = Real(0.01) * hbar;
    out[4] = kernel_alpha_pair<Real, KID>(inv_h);
}

// 0: use the general kernel; 1: a.uh_* are filled
static int pair_uh_constants(osph_ctx *ctx, PairArgs &a)
{
    const char *env = getenv("OSPH_UH");                 // read per launch: tests switch it inside one process
    const bool off = env && env[0] == '0';
    const osph_config &c = ctx->cfg;
    if (off || c.dynamic_h != OSPH_H_FIXED || !(c.fixed_h > 0.0) || !std::isfinite(c.fixed_h) || c.kernel == OSPH_KERNEL_GAUSSIAN) return 0;
    // The pipelined loop evaluates stage B for EVERY list entry and makes the operands of a non-common entry harmless with
    // r^2 := 1 and m_j := 0; in this instantiation q is then 1 / h, and the Wendland polynomial (degree 8 in q) leaves the float
    // range for h below about 1.5e-5: such a set-up runs the general instantiation, whose stand-in is q = 1 whatever h is.
    if (c.precision == OSPH_FP32 && c.kernel == OSPH_KERNEL_WENDLAND && c.fixed_h < 1e-3) return 0;
    if (!(ctx->uh_ready && ctx->uh_for_h == c.fixed_h && ctx->uh_for_kernel == c.kernel && ctx->uh_for_prec == c.precision)) {
        if (!ctx->d_uh && cudaMalloc(&ctx->d_uh, 8 * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return 0; }
        const bool cubic = c.kernel == OSPH_KERNEL_CUBIC;
        if (c.precision == OSPH_FP64) {
            if (cubic) k_uh_constants<double, OSPH_KERNEL_CUBIC><<<1, 1, 0, ctx->stream>>>(c.fixed_h, ctx->d_uh);
            else k_uh_constants<double, OSPH_KERNEL_WENDLAND><<<1, 1, 0, ctx->stream>>>(c.fixed_h, ctx->d_uh);
            if (cudaMemcpyAsync(ctx->uh_d, ctx->d_uh, 5 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return 0;
        } else {
            if (cubic) k_uh_constants<float, OSPH_KERNEL_CUBIC><<<1, 1, 0, ctx->stream>>>(c.fixed_h, (float *)ctx->d_uh);
            else k_uh_constants<float, OSPH_KERNEL_WENDLAND><<<1, 1, 0, ctx->stream>>>(c.fixed_h, (float *)ctx->d_uh);
            if (cudaMemcpyAsync(ctx->uh_f, ctx->d_uh, 5 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return 0;
        }
        ctx->launches++;
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return 0;       // once per context and smoothing length
        ctx->uh_ready = true; ctx->uh_for_h = c.fixed_h; ctx->uh_for_kernel = c.kernel; ctx->uh_for_prec = c.precision;
    }
    a.uh_h_d = ctx->uh_d[0]; a.uh_inv_h_d = ctx->uh_d[1]; a.uh_h2c_d = ctx->uh_d[2]; a.uh_c01_d = ctx->uh_d[3]; a.uh_alpha_d = ctx->uh_d[4];
    a.uh_h_f = ctx->uh_f[0]; a.uh_inv_h_f = ctx->uh_f[1]; a.uh_h2c_f = ctx->uh_f[2]; a.uh_c01_f = ctx->uh_f[3]; a.uh_alpha_f = ctx->uh_f[4];
    return 1;
}
#endif

int osph_pair_build_flags()
{
    return (PAIR_UH ? 1 : 0) | (PAIR_ISIGN ? 2 : 0) | (PAIR_UH && PAIR_UH_PIPE && PAIR_LEAN ? 4 : 0) | (PAIR_GEN_PIPE && PAIR_LEAN ? 8 : 0);
}

int osph_launch_pair(osph_ctx *ctx)
{
    PairArgs a;
    a.n = (int)(ctx->n + ctx->n_ghost);
    a.idx = ctx->idx[ctx->sorted_buf];
    a.s_pos = ctx->s_pos; a.s_vel = ctx->s_vel; a.s_rm = ctx->s_rm; a.s_hp = ctx->s_hp;
    a.s_info = ctx->s_info; a.s_coarse = ctx->s_coarse; a.s_gcell = ctx->s_gcell;
    a.cell_range = ctx->cell_range; a.gp = ctx->d_grid;
    a.a2max = ctx->pair_reduce_a2 ? &ctx->d_sc->a2max_fluid : nullptr;
    a.vx = ctx->f[OSPH_F_VX]; a.vy = ctx->f[OSPH_F_VY];
    a.drho = ctx->f[OSPH_F_DRHO]; a.ax = ctx->f[OSPH_F_AX]; a.ay = ctx->f[OSPH_F_AY];
    a.xsphx = ctx->f[OSPH_F_XSPHX]; a.xsphy = ctx->f[OSPH_F_XSPHY];
    const osph_config &c = ctx->cfg;
    a.p1 = c.p1; a.p2 = c.p2; a.gravity = c.gravity;
    a.lj_42 = (c.p1 == 4.0 && c.p2 == 2.0) ? 1 : 0;
    a.alpha_c_d = c.alpha * 0.5 * c.co; a.beta_d = c.beta; a.r0_d = c.r0; a.r0sq_d = c.r0 * c.r0; a.neg_eps_d = -c.epsilon; a.D_d = c.D;
    a.alpha_c_f = (float)a.alpha_c_d; a.beta_f = (float)c.beta; a.r0_f = (float)c.r0; a.r0sq_f = a.r0_f * a.r0_f;
    a.neg_eps_f = (float)a.neg_eps_d; a.D_f = (float)c.D;
    a.method_xsph = c.method_xsph; a.summation_density = c.summation_density;
    a.ts_sc = nullptr; a.ts_gamma_c = a.ts_gamma_f = a.ts_fixed_dt = a.ts_co = 0.0; a.ts_dt_log = nullptr; a.ts_dt_log_cap = 0; a.ts_fused = 0;
    if (ctx->pair_next_timestep) {
        a.ts_sc = ctx->d_sc; a.ts_gamma_c = c.cfl_courant; a.ts_gamma_f = c.cfl_force; a.ts_fixed_dt = ctx->pair_ts_fixed_dt;
        a.ts_co = c.co; a.ts_dt_log = ctx->d_dt_log; a.ts_dt_log_cap = (long long)ctx->dt_log_cap; a.ts_fused = 2;
    }
    int grid = div_up(ctx->n + ctx->n_ghost, OSPH_PAIR_THREADS);
    if (c.summation_density) {
        if (c.precision == OSPH_FP64) launch_summation<double2>(ctx, a); else launch_summation<float2>(ctx, a);
        OSPH_LAUNCH_CHECK();
    }
    bool uh = false;
    a.uh_h_d = a.uh_inv_h_d = a.uh_h2c_d = a.uh_c01_d = a.uh_alpha_d = 0.0;
    a.uh_h_f = a.uh_inv_h_f = a.uh_h2c_f = a.uh_c01_f = a.uh_alpha_f = 0.f;
#if PAIR_UH
    uh = pair_uh_constants(ctx, a) != 0;
#endif
    const bool timed = ctx->time_pair && ctx->pair_ev_used < OSPH_PAIR_EVENTS;
    if (timed) cudaEventRecord(ctx->pair_ev[2 * ctx->pair_ev_used], ctx->stream);
    int rc = c.precision == OSPH_FP64 ? launch_kid<double, true>(ctx, a, grid, uh) : launch_kid<float, false>(ctx, a, grid, uh);
    if (rc) return rc;
    OSPH_LAUNCH_CHECK();
    if (timed) { cudaEventRecord(ctx->pair_ev[2 * ctx->pair_ev_used + 1], ctx->stream); ctx->pair_ev_used++; }
    ctx->pair_launches++;
    if (uh) ctx->pair_uh_launches++;
    return 0;
}
