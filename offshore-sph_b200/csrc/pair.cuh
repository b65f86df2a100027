// pair.cuh -- argument block of the fused pair kernel (pair.cu)
#pragma once
#include "common.cuh"

struct PairArgs {
    int n;
    const unsigned int *idx;          // sorted position -> storage slot
    const double2 *s_pos;
    const void *s_vel, *s_rm, *s_hp;  // Real2 arrays: (vx,vy), (rho,m), (h, p/rho^2)
    const int *s_info;
    const int4 *s_coarse;
    const int2 *s_gcell;
    const int2 *cell_range;
    const GridParams *gp;
    unsigned long long *a2max;        // optional: order-preserving encoding of max |a|^2 over the fluid rows (TimeStep.py:58-91)
    const double *vx, *vy;            // state columns (storage order), for xsph = v + correction
    double *drho, *ax, *ay, *xsphx, *xsphy;
    double p1, p2, gravity;
    int lj_42, method_xsph, summation_density;
    // loop constants precomputed on the host in both precisions: as kernel parameters they are constant-bank operands
    // of the FP instructions and do not occupy registers in the pair loop
    double alpha_c_d, beta_d, r0_d, r0sq_d, neg_eps_d, D_d;
    float alpha_c_f, beta_f, r0_f, r0sq_f, neg_eps_f, D_f;
    // fused k_timestep of the NEXT step (osph_step's loop): run by the CTA that finishes last; ts_sc == nullptr: not fused
    StepScalars *ts_sc;
    double ts_gamma_c, ts_gamma_f, ts_fixed_dt, ts_co;
    double *ts_dt_log;
    long long ts_dt_log_cap;
    int ts_fused;
    // PAIR_UH (uniform smoothing length, Solver(h=value)): loop constants formed on the device once per context by
    // k_uh_constants with the very operations the general body applies per pair, then passed here (constant bank)
    double uh_h_d, uh_inv_h_d, uh_h2c_d, uh_c01_d, uh_alpha_d;
    float uh_h_f, uh_inv_h_f, uh_h2c_f, uh_c01_f, uh_alpha_f;
};
