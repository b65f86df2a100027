/*
 * wcsph_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement (single-threaded like the reference; oracle_loop_mt runs the
 * same pair loop on several host threads for the bench's CPU legs) of the WCSPH time step of
 * KoningJasper/Offshore-SPH (the reference is pure Python + numba).  It is the
 * checker the CUDA path is compared against; nothing under offshore-sph_b200/
 * may link, import or call it.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every entry point
 * against golden vectors frozen from the unmodified reference running under
 * numba in the build container (tests/golden/make_golden.py), and against the
 * closed-form known answers of the reference's own unit tests.
 *
 * The neighbour search and the time-step reduction are strict IEEE double in
 * the reference (numba jitclass, no fastmath); they are restated here with the
 * same operation order so cell ids, neighbour sets, neighbour ORDER and dt are
 * bit-exact.  The field math of the reference is fastmath=True; it is restated
 * in plain IEEE (build with -O2 -ffp-contract=off, no -ffast-math) and agrees
 * with the reference to ~1e-13 relative.
 *
 * All arrays are SoA over the ACTIVE (non-deleted) particles, in the order of
 * the reference's `pA[indexes]` compaction (src/Solver.py:238,254).
 * Citations are file:line into the reference tree.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LABEL_FLUID 0 /* src/Common.py:8-12 */

typedef struct {
    int64_t n;
    int8_t *label;
    double *m, *rho, *p, *c, *drho, *h, *x, *y, *vx, *vy, *ax, *ay;
    double *xsphx, *xsphy, *x0, *y0, *vx0, *vy0, *rho0;
} oracle_particles;

/* src/Methods/WCSPH.py:35-79 */
typedef struct {
    double height, r0, rho0, Pb;
    double gamma, co, B, alpha, beta, epsilon, D, p1, p2;
    int useXSPH, useSummationDensity;
} oracle_wcsph;

/* src/Tools/NNLinkedList.py:7-19 */
typedef struct {
    double xmin, xmax, ymin, ymax, cell_size, scale;
    int64_t ncx, ncy, n_cells, n;
    int64_t *heads, *nexts;
} oracle_grid;

enum { KERNEL_CUBIC = 0, KERNEL_WENDLAND = 1, KERNEL_GAUSSIAN = 2 };

/* ------------------------------------------------------------------ */
/* Tait EOS and WCSPH constants                                        */
/* ------------------------------------------------------------------ */

/* src/Equations/TaitEOS.py:40-44 */
double oracle_tait_co(double H) { return 10.0 * sqrt(2 * 9.81 * H); }

/* src/Equations/TaitEOS.py:33-38 */
double oracle_tait_B(double co, double rho0, double gamma) { return co * co * rho0 / gamma; }

/* src/Equations/TaitEOS.py:6-31 */
double oracle_tait_p(double gamma, double B, double rho0, double rho, int label)
{
    if (label != LABEL_FLUID) return 0.0;
    return (pow(rho / rho0, gamma) - 1.0) * B;
}

/* src/Equations/TaitEOS.py:46-65 */
double oracle_tait_height(double rho0, double H, double B, double gamma, double y)
{
    double frac = rho0 * 9.81 * (H - y) / B;
    return rho0 * pow(1 + frac, 1 / gamma);
}

/* src/Methods/WCSPH.py:35-79 */
void oracle_wcsph_init(oracle_wcsph *w, double height, double r0, double rho0,
                       int useXSPH, double Pb, int useSummationDensity)
{
    w->height = height; w->rho0 = rho0; w->useXSPH = useXSPH;
    w->useSummationDensity = useSummationDensity;
    w->epsilon = 0.5;
    w->gamma = 7.0;
    w->co = oracle_tait_co(height);
    w->B = oracle_tait_B(w->co, rho0, w->gamma);
    w->Pb = Pb;
    w->alpha = 0.01; w->beta = 0.0;
    w->r0 = r0; w->D = 5 * 9.81 * height; w->p1 = 4; w->p2 = 2;
}

/* src/Methods/WCSPH.py:82-108 (hydrostatic initial density of the rows passed in) */
void oracle_wcsph_initialize(const oracle_wcsph *w, int64_t n, const double *y, double *rho)
{
    for (int64_t j = 0; j < n; j++)
        rho[j] = oracle_tait_height(w->rho0, w->height, w->B, w->gamma, y[j]);
}

/* src/Tools/SolverTools.py:106-118 */
void oracle_compute_h(double sigma, int64_t J, const double *m, const double *rho, double *h)
{
    for (int64_t j = 0; j < J; j++) {
        h[j] = 0.0;
        if (rho[j] > 1e-12) h[j] = sigma * pow(m[j] / rho[j], 0.5);
    }
}

/* ------------------------------------------------------------------ */
/* Smoothing kernels (array form, as the reference's evaluate/gradient) */
/* ------------------------------------------------------------------ */

/* src/Kernels/CubicSpline.py:10-36, Wendland.py:9-33, Gaussian.py:16-25 */
static double kernel_w(int kid, double r, double h)
{
    double q = r / h;
    if (kid == KERNEL_CUBIC) {
        double alpha = (10 / (7 * M_PI)) / (h * h);
        if (q > 2.0) return 0.0;
        if (q > 1.0) return alpha * 0.25 * pow(2 - q, 3);
        return alpha * (1 - 1.5 * pow(q, 2) * (1 - 0.5 * q));
    } else if (kid == KERNEL_WENDLAND) {
        double alpha = (9.0 / (4.0 * M_PI)) / (h * h);
        if (q >= 2.0) return 0.0;
        double inner = 1.0 - 0.5 * q;
        return alpha * (pow(inner, 6) * (35.0 / 12.0 * q * q + 3.0 * q + 1.0));
    } else {
        double alpha = 1 / M_PI;
        if (q <= 3) return (alpha / (h * h)) * exp(-q * q);
        return 0.0;
    }
}

/* src/Kernels/CubicSpline.py:38-70, Wendland.py:35-64, Gaussian.py:38-59 */
static double kernel_dw(int kid, double x, double r, double h)
{
    double q = r / h;
    if (kid == KERNEL_CUBIC) {
        double alpha = (10 / (7 * M_PI)) / (h * h);
        double grad;
        if (q > 2.0 || r < 1e-10) return 0.0;
        if (q > 1.0) grad = -0.75 * pow(2 - q, 2);
        else grad = -3 * q * (1 - 0.75 * q);
        return (alpha * grad / (h * r)) * x;
    } else if (kid == KERNEL_WENDLAND) {
        double alpha = (9.0 / (4.0 * M_PI)) / (h * h);
        if (q >= 2.0 || r < 1e-10) return 0.0;
        double inner = 1.0 - 0.5 * q;
        double grad = pow(inner, 5) * (-14.0 / 3.0) * q * (1 + 2.5 * q);
        return (alpha * grad / (h * r)) * x;
    } else {
        double alpha = 1 / M_PI;
        double tmp = r * h;
        if (!(tmp > 1e-12)) return 0.0;
        if (!(q <= 3)) return 0.0;
        double dwdq = -2 * q * (alpha / (h * h)) * exp(-q * q);
        return dwdq / tmp * x;
    }
}

void oracle_kernel_evaluate(int kid, int64_t J, const double *r, const double *h, double *out)
{
    for (int64_t j = 0; j < J; j++) out[j] = kernel_w(kid, r[j], h[j]);
}

void oracle_kernel_gradient(int kid, int64_t J, const double *x, const double *r,
                            const double *h, double *out)
{
    for (int64_t j = 0; j < J; j++) out[j] = kernel_dw(kid, x[j], r[j], h[j]);
}

/* ------------------------------------------------------------------ */
/* Cell-linked-list neighbour search                                   */
/* ------------------------------------------------------------------ */

static double arr_min(const double *a, int64_t n)
{ double v = a[0]; for (int64_t i = 1; i < n; i++) if (a[i] < v) v = a[i]; return v; }
static double arr_max(const double *a, int64_t n)
{ double v = a[0]; for (int64_t i = 1; i < n; i++) if (a[i] > v) v = a[i]; return v; }

void oracle_grid_free(oracle_grid *g)
{
    free(g->heads); free(g->nexts); g->heads = g->nexts = NULL;
}

/*
 * src/Tools/NNLinkedList.py:37-39 update = _init (:86-101) + _bin (:129-141).
 * Returns 0, or -1 when a particle would be binned past the end of `heads`
 * (the reference writes out of bounds there; the oracle leaves it unbinned).
 */
int oracle_nn_update(oracle_grid *g, double scale, int64_t n, const double *x,
                     const double *y, const double *h)
{
    int rc = 0;
    g->scale = scale; g->n = n;
    g->xmin = arr_min(x, n); g->xmax = arr_max(x, n);
    g->ymin = arr_min(y, n); g->ymax = arr_max(y, n);
    /* :103-110 */
    double cs = arr_min(h, n) * scale;
    if (cs < 1e-6) cs = 1.0;
    g->cell_size = cs;
    /* :112-127 */
    double inv = 1. / cs;
    int64_t ncx = (int64_t)ceil(inv * (g->xmax - g->xmin));
    int64_t ncy = (int64_t)ceil(inv * (g->ymax - g->ymin));
    if (ncx < 1) ncx = 1;
    if (ncy < 1) ncy = 1;
    g->ncx = ncx; g->ncy = ncy; g->n_cells = ncx * ncy;
    g->heads = (int64_t *)malloc(sizeof(int64_t) * (size_t)g->n_cells);
    g->nexts = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t c = 0; c < g->n_cells; c++) g->heads[c] = -1;
    for (int64_t i = 0; i < n; i++) g->nexts[i] = -1;
    /* :129-141, cell id :164-176, flatten :200-208; head insertion */
    for (int64_t i = 0; i < n; i++) {
        int64_t cx = (int64_t)floor((x[i] - g->xmin) / cs);
        int64_t cy = (int64_t)floor((y[i] - g->ymin) / cs);
        int64_t cid = cx + ncx * cy;
        if (cid < 0 || cid >= g->n_cells) { rc = -1; continue; }
        g->nexts[i] = g->heads[cid];
        g->heads[cid] = i;
    }
    return rc;
}

/* Reference flat cell id of every particle (the value `_bin` computes, unclamped). */
void oracle_cell_ids(const oracle_grid *g, int64_t n, const double *x, const double *y,
                     int64_t *cid)
{
    for (int64_t i = 0; i < n; i++) {
        int64_t cx = (int64_t)floor((x[i] - g->xmin) / g->cell_size);
        int64_t cy = (int64_t)floor((y[i] - g->ymin) / g->cell_size);
        cid[i] = cx + g->ncx * cy;
    }
}

/*
 * src/Tools/NNLinkedList.py:41-80 nearPos.  Writes up to `cap` accepted
 * neighbours in reference order and returns the number accepted (which may
 * exceed cap; the caller re-calls with a larger buffer).
 */
int64_t oracle_near_pos(const oracle_grid *g, double px, double py, double ph,
                        const double *x, const double *y, const double *h,
                        int64_t cap, int64_t *idx, double *r_out, double *q_out, double *h_out)
{
    int64_t cx0 = (int64_t)floor((px - g->xmin) / g->cell_size);
    int64_t cy0 = (int64_t)floor((py - g->ymin) / g->cell_size);
    int64_t cnt = 0;
    for (int ix = -1; ix <= 1; ix++) {
        for (int iy = -1; iy <= 1; iy++) {
            int64_t cx = cx0 + ix, cy = cy0 + iy;
            /* :178-198 validity */
            if (!(cx > -1 && cx < g->ncx && cy > -1 && cy < g->ncy)) continue;
            int64_t cell = cx + g->ncx * cy;
            if (!(cell > -1 && cell < g->n_cells)) continue;
            for (int64_t j = g->heads[cell]; j != -1; j = g->nexts[j]) {
                double dx = px - x[j], dy = py - y[j];
                double r = sqrt(dx * dx + dy * dy);
                double hij = 0.5 * (ph + h[j]);
                double q = r / hij;
                if (q <= 3.0) {
                    if (cnt < cap) {
                        if (idx) idx[cnt] = j;
                        if (r_out) r_out[cnt] = r;
                        if (q_out) q_out[cnt] = q;
                        if (h_out) h_out[cnt] = hij;
                    }
                    cnt++;
                }
            }
        }
    }
    return cnt;
}

/*
 * CSR neighbour lists of every FLUID particle (the only rows `_loop` queries,
 * src/Tools/SolverTools.py:143-148); non-fluid rows are empty.  offsets has
 * n+1 entries.  Returns the total count; idx may be NULL to size the buffer.
 */
int64_t oracle_neighbours_csr(const oracle_grid *g, int64_t n, const int8_t *label,
                              const double *x, const double *y, const double *h,
                              int64_t *offsets, int64_t cap, int64_t *idx)
{
    int64_t total = 0;
    for (int64_t i = 0; i < n; i++) {
        offsets[i] = total;
        if (label[i] != LABEL_FLUID) continue;
        int64_t room = (idx && cap > total) ? cap - total : 0;
        total += oracle_near_pos(g, x[i], y[i], h[i], x, y, h, room,
                                 room ? idx + total : NULL, NULL, NULL, NULL);
    }
    offsets[n] = total;
    return total;
}

/* ------------------------------------------------------------------ */
/* _loop: EOS + pair interactions                                       */
/* ------------------------------------------------------------------ */

/* main loop of `_loop` (:143-173) over the particles [i0, i1); every particle writes only its own rates
 * (drho, ax, ay, xsphx, xsphy) and reads fields no iteration writes, so ranges are independent of each other. */
static int64_t loop_range(oracle_particles *P, const oracle_wcsph *w, const oracle_grid *g,
                          int kid, int64_t stride, int64_t phase, int64_t i0, int64_t i1)
{
    int64_t pairs = 0;
    int64_t cap = 256;
    int64_t *nb = (int64_t *)malloc(sizeof(int64_t) * cap);
    double *rr = (double *)malloc(sizeof(double) * cap);
    double *qq = (double *)malloc(sizeof(double) * cap);
    double *hh = (double *)malloc(sizeof(double) * cap);
    for (int64_t i = i0; i < i1; i++) {
        if (P->label[i] != LABEL_FLUID) continue;
        if (stride > 1 && (i % stride) != phase) continue;
        int64_t J = oracle_near_pos(g, P->x[i], P->y[i], P->h[i], P->x, P->y, P->h,
                                    cap, nb, rr, qq, hh);
        if (J > cap) {
            cap = J * 2;
            nb = realloc(nb, sizeof(int64_t) * cap); rr = realloc(rr, sizeof(double) * cap);
            qq = realloc(qq, sizeof(double) * cap); hh = realloc(hh, sizeof(double) * cap);
            J = oracle_near_pos(g, P->x[i], P->y[i], P->h[i], P->x, P->y, P->h,
                                cap, nb, rr, qq, hh);
        }
        if (J == 0) continue;
        pairs += J;

        double xi = P->x[i], yi = P->y[i], vxi = P->vx[i], vyi = P->vy[i];
        double rhoi = P->rho[i], hi = P->h[i], ci = P->c[i];
        double slf = P->p[i] / (rhoi * rhoi);
        double drho = 0.0, ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0, xs = 0.0, ys = 0.0;
        for (int64_t k = 0; k < J; k++) {
            int64_t j = nb[k];
            /* _assignProps :74-104: differences are i - j; comp.h is h_ij; comp.c stays 0 */
            double dx = xi - P->x[j], dy = yi - P->y[j];
            double dvx = vxi - P->vx[j], dvy = vyi - P->vy[j];
            double r = rr[k], hij = hh[k];
            double wk = kernel_w(kid, r, hij);
            double dwx = kernel_dw(kid, dx, r, hij);
            double dwy = kernel_dw(kid, dy, r, hij);
            int fluid_j = (P->label[j] == LABEL_FLUID);
            if (fluid_j) {
                /* Continuity */
                drho += P->m[j] * (dvx * dwx + dvy * dwy);
                /* Momentum */
                double othr = P->p[j] / (P->rho[j] * P->rho[j]);
                double dot = dvx * dx + dvy * dy;
                double PI = 0.0;
                if (dot < 0) {
                    double hbar = 0.5 * (hi + hij);      /* h averaged twice, Momentum.py:43 */
                    double cbar = 0.5 * (ci + 0.0);      /* comp.c is never filled, SolverTools.py:90 */
                    double rbar = 0.5 * (rhoi + P->rho[j]);
                    double mu = hbar * dot / (r * r + 0.01 * hbar * hbar);
                    PI = mu * (w->beta * mu - w->alpha * cbar) / rbar;
                }
                double factor = slf + othr + PI;
                ax += -P->m[j] * factor * dwx;
                ay += -P->m[j] * factor * dwy;
            } else if (!(r > w->r0)) {
                /* BoundaryForce: non-fluid, r <= r0, r > 1e-12 */
                if (r > 1e-12) {
                    double frac = w->r0 / r;
                    double fac = w->D * (pow(frac, w->p1) - pow(frac, w->p2));
                    bx += fac * dx / pow(r, 2);
                    by += fac * dy / pow(r, 2);
                }
            }
            /* XSPH: every label */
            {
                double rbar = 0.5 * (rhoi + P->rho[j]);
                double fac = -w->epsilon * P->m[j] * wk / rbar;
                xs += fac * dvx; ys += fac * dvy;
            }
        }
        P->drho[i] = w->useSummationDensity ? 0.0 : drho;  /* WCSPH.py:191-203 */
        P->ax[i] = ax + bx;
        P->ay[i] = (ay - 9.81) + by;                        /* WCSPH.py:168-169, SolverTools.py:169-170 */
        if (w->useXSPH) { P->xsphx[i] = vxi + xs; P->xsphy[i] = vyi + ys; }
        else { P->xsphx[i] = 0.0; P->xsphy[i] = 0.0; }
    }
    free(nb); free(rr); free(qq); free(hh);
    return pairs;
}

/* :121-140: EOS for every active particle, then the optional summation density (in place, index order) */
static void loop_head(oracle_particles *P, const oracle_wcsph *w, const oracle_grid *g, int kid)
{
    int64_t n = P->n;
    int64_t cap = 256;
    int64_t *nb = (int64_t *)malloc(sizeof(int64_t) * cap);
    double *rr = (double *)malloc(sizeof(double) * cap);
    double *qq = (double *)malloc(sizeof(double) * cap);
    double *hh = (double *)malloc(sizeof(double) * cap);
    /* :121-127 pressure and speed of sound for every active particle */
    for (int64_t i = 0; i < n; i++) {
        P->p[i] = oracle_tait_p(w->gamma, w->B, w->rho0, P->rho[i], P->label[i]) + w->Pb;
        P->c[i] = w->co;
    }

    /* :129-140 optional summation density, applied in place in index order */
    if (w->useSummationDensity) {
        for (int64_t i = 0; i < n; i++) {
            if (P->label[i] != LABEL_FLUID) continue;
            int64_t J = oracle_near_pos(g, P->x[i], P->y[i], P->h[i], P->x, P->y, P->h,
                                        cap, nb, rr, qq, hh);
            if (J > cap) {
                cap = J * 2;
                nb = realloc(nb, sizeof(int64_t) * cap); rr = realloc(rr, sizeof(double) * cap);
                qq = realloc(qq, sizeof(double) * cap); hh = realloc(hh, sizeof(double) * cap);
                J = oracle_near_pos(g, P->x[i], P->y[i], P->h[i], P->x, P->y, P->h,
                                    cap, nb, rr, qq, hh);
            }
            double rho = 0.0; /* SummationDensity.py:6-13 */
            for (int64_t k = 0; k < J; k++)
                if (P->label[nb[k]] == LABEL_FLUID) rho += P->m[nb[k]] * kernel_w(kid, rr[k], hh[k]);
            P->rho[i] = rho;
        }
    }

    free(nb); free(rr); free(qq); free(hh);
}

/*
 * src/Tools/SolverTools.py:120-174 with the equations it calls:
 * Continuity.py:5-17, Momentum.py:6-57 (+gravity WCSPH.py:151-169),
 * BoundaryForce.py:7-42, XSPH.py:6-31 (+WCSPH.py:171-189).
 * `stride`/`phase` restrict the main loop to fluid particles with
 * (i % stride) == phase (bounded CPU-baseline samples); use 1, 0 for all.
 * Returns the number of accepted pairs evaluated.
 */
int64_t oracle_loop(oracle_particles *P, const oracle_wcsph *w, const oracle_grid *g,
                    int kid, int64_t stride, int64_t phase)
{
    loop_head(P, w, g, kid);
    return loop_range(P, w, g, kid, stride, phase, 0, P->n);
}

/*
 * The same loop on `nthreads` host threads (bench.py's cpu_baseline / --impl reference legs: "all the host threads it
 * can use").  The reference itself is single-threaded (numba without parallel=True); per-particle arithmetic and its
 * order are those of oracle_loop, so the results are bit-identical to it (tests/test_oracle_golden.py).  Blocks of
 * OR_MT_BLOCK particles are dealt round-robin to the threads.  The EOS pass and the summation-density pass (in place,
 * order-dependent) stay serial.
 */
#define OR_MT_BLOCK 2048
typedef struct {
    oracle_particles *P; const oracle_wcsph *w; const oracle_grid *g;
    int kid, tid, nthreads; int64_t stride, phase, pairs;
} loop_job;

static void *loop_worker(void *arg)
{
    loop_job *j = (loop_job *)arg;
    int64_t n = j->P->n, pairs = 0;
    for (int64_t b = (int64_t)j->tid * OR_MT_BLOCK; b < n; b += (int64_t)j->nthreads * OR_MT_BLOCK) {
        int64_t e = b + OR_MT_BLOCK < n ? b + OR_MT_BLOCK : n;
        pairs += loop_range(j->P, j->w, j->g, j->kid, j->stride, j->phase, b, e);
    }
    j->pairs = pairs;
    return NULL;
}

int64_t oracle_loop_mt(oracle_particles *P, const oracle_wcsph *w, const oracle_grid *g,
                       int kid, int64_t stride, int64_t phase, int nthreads)
{
    if (nthreads < 2) return oracle_loop(P, w, g, kid, stride, phase);
    if (nthreads > 256) nthreads = 256;
    loop_head(P, w, g, kid);
    pthread_t th[256];
    loop_job job[256];
    int started = 0;
    for (int t = 0; t < nthreads; t++) {
        job[t].P = P; job[t].w = w; job[t].g = g; job[t].kid = kid; job[t].tid = t; job[t].nthreads = nthreads;
        job[t].stride = stride; job[t].phase = phase; job[t].pairs = 0;
    }
    for (int t = 1; t < nthreads; t++) {
        if (pthread_create(&th[t], NULL, loop_worker, &job[t]) != 0) break;
        started = t;
    }
    /* threads that could not be started: their blocks are walked here, after the caller's own */
    loop_worker(&job[0]);
    for (int t = started + 1; t < nthreads; t++) loop_worker(&job[t]);
    int64_t pairs = 0;
    for (int t = 1; t <= started; t++) pthread_join(th[t], NULL);
    for (int t = 0; t < nthreads; t++) pairs += job[t].pairs;
    return pairs;
}

/* ------------------------------------------------------------------ */
/* The equations as stand-alone leaves on a table of J computed        */
/* neighbours (computed_dtype, src/Common.py:59-79: differences i - j,  */
/* comp.h = h_ij).  Same operation order as the inlined forms in        */
/* oracle_loop; these are what the reference's own equation tests       */
/* (test/test_numba_momentum.py, test_numba_continuity.py,              */
/* test_eq_boundary.py, test_eq_courant.py) call, so they pin the       */
/* per-pair formulas one by one (tests/test_oracle_golden.py).          */
/* ------------------------------------------------------------------ */

/* src/Equations/Continuity.py:5-17 */
double oracle_eq_continuity(int64_t J, const int8_t *label, const double *m, const double *vx,
                            const double *vy, const double *dwx, const double *dwy)
{
    double arho = 0.0;
    for (int64_t j = 0; j < J; j++) {
        if (label[j] != LABEL_FLUID) continue;
        arho += m[j] * (vx[j] * dwx[j] + vy[j] * dwy[j]);
    }
    return arho;
}

/* src/Equations/Momentum.py:6-57; self = (p, rho, h, c) of particle i */
void oracle_eq_momentum(double alpha, double beta, double self_p, double self_rho, double self_h,
                        double self_c, int64_t J, const int8_t *label, const double *p,
                        const double *rho, const double *x, const double *y, const double *vx,
                        const double *vy, const double *r, const double *h, const double *c,
                        const double *m, const double *dwx, const double *dwy, double out[2])
{
    double slf = self_p / (self_rho * self_rho);
    double ax = 0.0, ay = 0.0;
    for (int64_t j = 0; j < J; j++) {
        if (label[j] != LABEL_FLUID) continue;
        double othr = p[j] / (rho[j] * rho[j]);
        double dot = vx[j] * x[j] + vy[j] * y[j];
        double PI = 0.0;
        if (dot < 0) {
            double hij = 0.5 * (self_h + h[j]);
            double cij = 0.5 * (self_c + c[j]);
            double rhoij = 0.5 * (self_rho + rho[j]);
            double mu = hij * dot / (r[j] * r[j] + 0.01 * hij * hij);
            PI = mu * (beta * mu - alpha * cij) / rhoij;
        }
        double factor = slf + othr + PI;
        ax += -m[j] * factor * dwx[j];
        ay += -m[j] * factor * dwy[j];
    }
    out[0] = ax; out[1] = ay;
}

/* src/Equations/XSPH.py:6-31 (every label) */
void oracle_eq_xsph(double epsilon, double self_rho, int64_t J, const double *rho, const double *m,
                    const double *w, const double *vx, const double *vy, double out[2])
{
    double xs = 0.0, ys = 0.0;
    for (int64_t j = 0; j < J; j++) {
        double rho_ij = 0.5 * (self_rho + rho[j]);
        double fac = -epsilon * m[j] * w[j] / rho_ij;
        xs += fac * vx[j]; ys += fac * vy[j];
    }
    out[0] = xs; out[1] = ys;
}

/* src/Equations/BoundaryForce.py:7-42 */
void oracle_eq_boundary_force(double r0, double D, double p1, double p2, int64_t J,
                              const int8_t *label, const double *r, const double *x,
                              const double *y, double out[2])
{
    double fx = 0.0, fy = 0.0;
    for (int64_t j = 0; j < J; j++) {
        if (label[j] == LABEL_FLUID || r[j] > r0) continue;
        if (r[j] > 1e-12) {
            double frac = r0 / r[j];
            double fac = D * (pow(frac, p1) - pow(frac, p2));
            fx += fac * x[j] / pow(r[j], 2);
            fy += fac * y[j] / pow(r[j], 2);
        }
    }
    out[0] = fx; out[1] = fy;
}

/* src/Equations/Courant.py:4-31 (unused by the Solver, which calls TimeStep; kept because
 * test/test_eq_courant.py pins it) */
double oracle_eq_courant(double alpha, int64_t J, const double *h, const double *c)
{
    double h_min = 10e10, c_max = 1e-10;
    for (int64_t j = 0; j < J; j++) {
        if (h[j] < h_min) h_min = h[j];
        if (c[j] > c_max) c_max = c[j];
    }
    return alpha * h_min / c_max;
}

/* ------------------------------------------------------------------ */
/* Integrators (act on the fluid rows only, src/Solver.py:380,396)      */
/* ------------------------------------------------------------------ */

/* src/Integrators/PEC.py:32-60 */
void oracle_pec_predict(oracle_particles *P, const uint8_t *mask, double dt, double damping,
                        int useXSPH, int strict)
{
    for (int64_t j = 0; j < P->n; j++) {
        if (mask && !mask[j]) continue;
        P->x0[j] = P->x[j]; P->y0[j] = P->y[j];
        P->vx0[j] = P->vx[j]; P->vy0[j] = P->vy[j];
        if (useXSPH && P->label[j] == LABEL_FLUID) {
            P->x[j] = P->x[j] + 0.5 * dt * P->xsphx[j];
            P->y[j] = P->y[j] + 0.5 * dt * P->xsphy[j];
        } else {
            P->x[j] = P->x[j] + 0.5 * dt * P->vx[j];
            P->y[j] = P->y[j] + 0.5 * dt * P->vy[j];
        }
        P->vx[j] = (P->vx[j] + 0.5 * dt * P->ax[j]) / (1 + 0.5 * damping);
        P->vy[j] = (P->vy[j] + 0.5 * dt * P->ay[j]) / (1 + 0.5 * damping);
        P->rho0[j] = P->rho[j];
        P->rho[j] = P->rho[j] + 0.5 * dt * P->drho[j];
        if (strict && P->rho[j] < 0.0) P->rho[j] = 0.0;
    }
}

/* src/Integrators/PEC.py:62-88 */
void oracle_pec_correct(oracle_particles *P, const uint8_t *mask, double dt, double damping,
                        int useXSPH, int strict)
{
    for (int64_t j = 0; j < P->n; j++) {
        if (mask && !mask[j]) continue;
        double mx, my;
        if (useXSPH && P->label[j] == LABEL_FLUID) {
            mx = P->x0[j] + 0.5 * dt * P->xsphx[j];
            my = P->y0[j] + 0.5 * dt * P->xsphy[j];
        } else {
            mx = P->x0[j] + 0.5 * dt * P->vx[j];
            my = P->y0[j] + 0.5 * dt * P->vy[j];
        }
        double mvx = (P->vx0[j] + 0.5 * dt * P->ax[j]) / (1 + 0.5 * damping);
        double mvy = (P->vy0[j] + 0.5 * dt * P->ay[j]) / (1 + 0.5 * damping);
        P->x[j] = 2 * mx - P->x0[j];
        P->y[j] = 2 * my - P->y0[j];
        P->vx[j] = 2 * mvx - P->vx0[j];
        P->vy[j] = 2 * mvy - P->vy0[j];
        double mrho = P->rho0[j] + 0.5 * dt * P->drho[j];
        P->rho[j] = 2 * mrho - P->rho0[j];
        if (strict && P->rho[j] < 0.0) P->rho[j] = 0.0;
    }
}

/* src/Integrators/Euler.py:17-26 (predict is the identity, :13-15) */
void oracle_euler_correct(oracle_particles *P, const uint8_t *mask, double dt)
{
    for (int64_t j = 0; j < P->n; j++) {
        if (mask && !mask[j]) continue;
        P->x[j] = P->x[j] + dt * P->vx[j] + 0.5 * dt * dt * P->ax[j];
        P->y[j] = P->y[j] + dt * P->vy[j] + 0.5 * dt * dt * P->ay[j];
        P->vx[j] = P->vx[j] + dt * P->ax[j];
        P->vy[j] = P->vy[j] + dt * P->ay[j];
        P->rho[j] = P->rho[j] + dt * P->drho[j];
    }
}

/* src/Integrators/Verlet.py:28-36 */
void oracle_verlet_predict(oracle_particles *P, const uint8_t *mask, double dt)
{
    for (int64_t j = 0; j < P->n; j++) {
        if (mask && !mask[j]) continue;
        P->x[j] += 0.5 * dt * P->vx[j];
        P->y[j] += 0.5 * dt * P->vy[j];
    }
}

/* src/Integrators/Verlet.py:38-55 */
void oracle_verlet_correct(oracle_particles *P, const uint8_t *mask, double dt, int useXSPH)
{
    for (int64_t j = 0; j < P->n; j++) {
        if (mask && !mask[j]) continue;
        P->vx[j] += dt * P->ax[j];
        P->vy[j] += dt * P->ay[j];
        if (useXSPH) {
            P->x[j] += 0.5 * dt * P->xsphx[j];
            P->y[j] += 0.5 * dt * P->xsphy[j];
        } else {
            P->x[j] += 0.5 * dt * P->vx[j];
            P->y[j] += 0.5 * dt * P->vy[j];
        }
    }
}

/* ------------------------------------------------------------------ */
/* Time step, kinetic energy                                            */
/* ------------------------------------------------------------------ */

/*
 * src/Equations/TimeStep.py:11-91 over the fluid rows (src/Solver.py:217).
 * out = {min(dt_c, dt_f), dt_c, dt_f}.  Returns -1 when there is no fluid
 * particle (the reference raises from np.min of an empty array).
 */
int oracle_timestep(const oracle_particles *P, const uint8_t *mask, double gamma_c,
                    double gamma_f, double out[3])
{
    int any = 0;
    double min_h = 0, max_c = 0, max_a2 = 0;
    for (int64_t j = 0; j < P->n; j++) {
        if (mask && !mask[j]) continue;
        if (P->label[j] != LABEL_FLUID) continue;
        double a2 = P->ax[j] * P->ax[j] + P->ay[j] * P->ay[j];
        if (!any) { min_h = P->h[j]; max_c = P->c[j]; max_a2 = a2; any = 1; }
        else {
            if (P->h[j] < min_h) min_h = P->h[j];
            if (P->c[j] > max_c) max_c = P->c[j];
            if (a2 > max_a2) max_a2 = a2;
        }
    }
    if (!any) return -1;
    double c = gamma_c * min_h / max_c;                       /* :47-49 */
    double f = (max_a2 < 1e-12) ? 1e10 : gamma_f * sqrt(min_h / max_a2); /* :51-56 */
    out[0] = c < f ? c : f; out[1] = c; out[2] = f;
    return 0;
}

/* src/Equations/KineticEnergy.py:6-12 over the rows selected by mask */
double oracle_kinetic_energy(const oracle_particles *P, const uint8_t *mask)
{
    double k = 0.0;
    for (int64_t j = 0; j < P->n; j++) {
        if (mask && !mask[j]) continue;
        double v2 = pow(P->vx[j], 2) + pow(P->vy[j], 2);
        k += 0.5 * P->m[j] * v2;
    }
    return k;
}

/* ------------------------------------------------------------------ */
/* One whole step in the order of src/Solver.py:366-399                 */
/* ------------------------------------------------------------------ */

/*
 * integrator: 0 PEC, 1 Euler, 2 Verlet.  fixed_h < 0 selects the dynamic
 * smoothing length (computeH(1.3, ...), src/Solver.py:242-246).  fixed_dt <= 0
 * selects the dynamic time step.  `fluid` marks the rows the integrator and
 * TimeStep act on.  dt_out receives {dt, dt_c, dt_f}.
 */
int oracle_step(oracle_particles *P, const oracle_wcsph *w, int kid, int integrator,
                int integ_xsph, int strict, double damping, double fixed_h, double fixed_dt,
                const uint8_t *fluid, int64_t stride, double dt_out[3], int64_t *pairs_out)
{
    double dt3[3] = {fixed_dt, 0, 0};
    if (!(fixed_dt > 0)) {
        if (oracle_timestep(P, fluid, 0.25, 0.25, dt3) != 0) return -1;
    }
    double dt = dt3[0];
    if (integrator == 0) oracle_pec_predict(P, fluid, dt, damping, integ_xsph, strict);
    else if (integrator == 2) oracle_verlet_predict(P, fluid, dt);

    oracle_grid g; memset(&g, 0, sizeof g);
    oracle_nn_update(&g, 2.0, P->n, P->x, P->y, P->h);          /* Solver.py:109,238 */
    for (int64_t j = 0; j < P->n; j++) {                        /* Solver.py:242-246 */
        if (!fluid[j]) continue;
        if (fixed_h >= 0) P->h[j] = fixed_h;
        else {
            P->h[j] = 0.0;
            if (P->rho[j] > 1e-12) P->h[j] = 1.3 * pow(P->m[j] / P->rho[j], 0.5);
        }
    }
    int64_t pairs = oracle_loop(P, w, &g, kid, stride, 0);
    oracle_grid_free(&g);

    if (integrator == 0) oracle_pec_correct(P, fluid, dt, damping, integ_xsph, strict);
    else if (integrator == 1) oracle_euler_correct(P, fluid, dt);
    else oracle_verlet_correct(P, fluid, dt, integ_xsph);
    if (dt_out) { dt_out[0] = dt3[0]; dt_out[1] = dt3[1]; dt_out[2] = dt3[2]; }
    if (pairs_out) *pairs_out = pairs;
    return 0;
}
