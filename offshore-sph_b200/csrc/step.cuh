// step.cuh -- argument blocks and host launchers of step.cu
#pragma once
#include "common.cuh"

struct PrepareArgs {
    int n;
    const signed char *label;
    double *x, *y, *vx, *vy, *rho, *h;
    const double *m, *ax, *ay, *drho, *xsphx, *xsphy;
    double *x0, *y0, *vx0, *vy0, *rho0;
    StepScalars *sc;
    double dt, damping, fixed_h, h_sigma;
    int use_dev_dt, integ_xsph, strict, dynamic_h;
    const double *xref, *yref;        // positions at the last sort (nullptr: no displacement reduction)
    // fused k_grid_params (last CTA): nullptr = separate launch
    GridParams *grid;
    double g_nn_scale, g_pair_radius_q, g_r0, g_skin_frac;
    long long g_cell_cap;
    int g_reset_dt, g_force_sort;
    int reduce_hmin_fluid;            // fused loop: min h over the fluid rows (TimeStep of the next step) is taken here
    double rden;                      // RN(1 / (1 + damping / 2)) for div_den, 0: divide
};

struct CorrectArgs {
    int n;
    const signed char *label;
    double *x, *y, *vx, *vy, *rho;
    const double *h, *c, *ax, *ay, *drho, *xsphx, *xsphy, *x0, *y0, *vx0, *vy0, *rho0;
    StepScalars *sc;
    double dt, damping, co;
    double rden;                      // as PrepareArgs::rden
    int use_dev_dt, integ_xsph, strict, c_uniform;
};

struct GatherArgs {
    int n_owned, n_all;
    const unsigned int *key;      // radix sort only: sorted keys (the cell table is written from them), else nullptr
    int2 *cell_range;
    const unsigned int *idx;      // sorted position -> storage slot, final (radix sort; counting sort on a build that does
                                  // not sort, or whose order k_bin_rank made canonical)
    const unsigned int *idx_arrival;   // counting sort, sorting build: the same in arrival order inside each cell
    // counting sort: the kernel makes the order inside each cell canonical on the way (bin_canonical_slot) -- every record
    // is written to its final position and idx_out receives the final permutation
    const int2 *rank_ranges;      // per sorted position: (begin, end) of its cell; nullptr: idx is final on every build
    unsigned int *idx_out;
    StepScalars *sc;
    const signed char *label;
    const double *ghost;
    GhostMap gmap;
    const double *x, *y, *vx, *vy, *rho, *m, *h;
    double *p;
    const GridParams *gp;
    double2 *s_pos;
    int *s_info;
    int4 *s_coarse;
    int2 *s_gcell;
    double gamma, B, rho0, Pb;
    double uh_h;                  // PAIR_UH, OSPH_H_FIXED: the smoothing length every fluid particle must carry (0: not checked)
};

struct NeighbourArgs {
    int n;
    const unsigned int *idx;
    const int *act;
    const double *h;          // state column, storage order (exact FP64 h in both precisions)
    const double2 *s_pos;
    const int *s_info;
    const int4 *s_coarse;
    const int2 *s_gcell;
    const int2 *cell_range;
    const GridParams *gp;
    long long *counts;        // per active index
    const long long *offsets;
    long long *out;
};

int osph_launch_setup(osph_ctx *ctx);
int osph_launch_unpack(osph_ctx *ctx);
int osph_launch_set_deleted(osph_ctx *ctx, const unsigned char *d_active);
int osph_launch_active_list(osph_ctx *ctx, int n_total, int *d_counters);
int osph_launch_pack(osph_ctx *ctx);
int osph_launch_pack_owned(osph_ctx *ctx, int *d_ids);
// What the next neighbour-structure build will do, decided on the host before the predictor pass is launched (the pass can
// then run k_grid_params in its last CTA): force_sort as k_grid_params takes it, physical reorder of the state or not.
struct BuildPlan { int force; bool reorder_now; bool always; bool grid_done; };
BuildPlan osph_plan_build(osph_ctx *ctx);
int osph_launch_prepare(osph_ctx *ctx, bool predict, double dt, double damping, bool use_dev_dt, bool skip_reset = false,
                        int fused = 0, const BuildPlan *grid_plan = nullptr, bool grid_reset_dt = false);
                        // fused 1: reduce fluid h_min too; 2: also apply the previous step's corrector first
                        // grid_plan: run k_grid_params inside this pass (its last CTA) with that plan
int osph_launch_build(osph_ctx *ctx, bool reset_dt = false, const BuildPlan *plan = nullptr);   // grid params (unless plan->grid_done), keys, sort, (reorder), gather
int osph_launch_correct(osph_ctx *ctx, bool correct, double dt, double damping, bool use_dev_dt, bool skip_reset = false);
int osph_launch_timestep(osph_ctx *ctx, double fixed_dt, bool log, bool reset_prepare = false,
                         const double *d_reduced3 = nullptr, int fused = 0);   // fused 1: reset the dt scalars after use; 2: c_max = co too
int osph_launch_grid_params(osph_ctx *ctx, bool reset_dt = false, int force_sort = 1);   // force_sort: see k_grid_params
int osph_launch_ke(osph_ctx *ctx);
int osph_launch_neighbours(osph_ctx *ctx, int mode, long long *d_counts, const long long *d_offsets, long long *d_out);
int osph_launch_near_pos(osph_ctx *ctx, double x, double y, double h, long long cap, long long *d_idx, double *d_r,
                         double *d_q, double *d_h, long long *d_count);
int osph_launch_cells(osph_ctx *ctx, long long *d_out);
int osph_launch_probe(osph_ctx *ctx, int npts, const double *d_x, const double *d_y, double h, double *d_rho, double *d_p);
int osph_launch_col_to_active(osph_ctx *ctx, int field, double *d_out);
int osph_launch_col_from_active(osph_ctx *ctx, int field, const double *d_in);
int osph_init_scalars(osph_ctx *ctx);
// slab.cu: owned-set update after an exchange; migrants arrive in two buffers (from the left / right neighbour)
int osph_slab_commit_impl(osph_ctx *ctx, int64_t n_mig_out, const double *d_in_l, int64_t n_in_l, const double *d_in_r,
                          int64_t n_in_r, GhostMap gmap, int64_t n_ghost, const double global_bounds[6]);
int osph_launch_fill(osph_ctx *ctx, double *d_col, double value);
int osph_scan_exclusive(osph_ctx *ctx, unsigned int *d_data, int n);     // scan.cu, in place
