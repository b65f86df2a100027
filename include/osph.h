/*
 * osph.h -- C ABI of libosph_b200.so: the B200-native (sm_100a) WCSPH time step that
 * replaces the numba hot path of KoningJasper/Offshore-SPH.
 *
 * The reference has no FFI: its "plugin API" is a set of duck-typed Python call sites inside
 * Solver.run() (reference src/Solver.py:366-399).  Every entry point below replaces one of those
 * call sites; the citation after "replaces:" is file:line in the reference tree.  Host Python
 * binds this header with ctypes (offshore-sph_b200/osph_b200/capi.py); INTEGRATION.md shows the
 * few lines a maintainer of the reference would add to src/Solver.py.
 *
 * Conventions
 *   - every call returns 0 on success, a negative OSPH_E_* code on failure; osph_last_error()
 *     gives the text.  There is no CPU fallback: without a CUDA device osph_create() fails.
 *   - the caller owns host buffers, the library owns device buffers.  Host arrays of particles
 *     are the reference's packed `particle_dtype` records (src/Common.py:26-57): 1 byte deleted,
 *     1 byte label, 19 doubles at byte offsets 2 + 8*k, record stride 154 unless stated.
 *   - "active" particles are the rows with deleted == 0, in row order; index spaces that the
 *     reference exposes (neighbour indices, cell ids, per-particle columns) are positions in that
 *     compacted active array, exactly as in the reference's `pA[indexes]` (src/Solver.py:238,254).
 *   - one host thread per context; all work is enqueued on one CUDA stream per context and a call
 *     blocks only when it returns data to the host.
 */
#ifndef OSPH_H
#define OSPH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSPH_VERSION 1

/* arithmetic of the pair kernel; integrator state is always double */
#define OSPH_FP64 0   /* validation mode: exact reference neighbour predicate, fields <= 1e-10 */
#define OSPH_FP32 1   /* performance mode: pair arithmetic in float on anchor-relative positions */

/* smoothing kernel -- replaces: src/Kernels/CubicSpline.py:10-70, Wendland.py:9-64, Gaussian.py:16-59 */
#define OSPH_KERNEL_CUBIC    0
#define OSPH_KERNEL_WENDLAND 1
#define OSPH_KERNEL_GAUSSIAN 2

/* integrator -- replaces: src/Integrators/PEC.py:12-88, Euler.py:7-26, Verlet.py:8-55 */
#define OSPH_INTEGRATOR_PEC    0
#define OSPH_INTEGRATOR_EULER  1
#define OSPH_INTEGRATOR_VERLET 2

/* smoothing-length refresh of Solver._compute (src/Solver.py:242-246) */
#define OSPH_H_FIXED   0   /* Solver(h=value): h[fluid] = fixed_h */
#define OSPH_H_DYNAMIC 1   /* Solver(h=None): h[fluid] = h_sigma * sqrt(m/rho), computeH (SolverTools.py:106-118) */
#define OSPH_H_KEEP    2   /* leave h as uploaded (stand-alone nn.update / _loop calls) */

/* particle labels, src/Common.py:8-12 */
#define OSPH_FLUID 0
#define OSPH_BOUNDARY 1
#define OSPH_TEMP_BOUNDARY 2
#define OSPH_COUPLED 3

/* real-valued columns of particle_dtype, in record order (byte offset = 2 + 8 * id) */
enum osph_field {
    OSPH_F_M = 0, OSPH_F_RHO, OSPH_F_P, OSPH_F_C, OSPH_F_DRHO, OSPH_F_H, OSPH_F_X, OSPH_F_Y,
    OSPH_F_VX, OSPH_F_VY, OSPH_F_AX, OSPH_F_AY, OSPH_F_XSPHX, OSPH_F_XSPHY, OSPH_F_X0, OSPH_F_Y0,
    OSPH_F_VX0, OSPH_F_VY0, OSPH_F_RHO0, OSPH_NUM_FIELDS
};

#define OSPH_E_INVALID   (-1)  /* bad argument / call order */
#define OSPH_E_CUDA      (-2)  /* CUDA runtime error (text in osph_last_error) */
#define OSPH_E_NO_DEVICE (-3)  /* no usable CUDA device: the library has no CPU path */
#define OSPH_E_CAPACITY  (-4)  /* a caller-provided buffer is too small */
#define OSPH_E_NO_FLUID  (-5)  /* time-step reduction over zero fluid particles */
#define OSPH_E_GRID      (-6)  /* the reference grid would need more cells than can be tabulated */
#define OSPH_E_PEER      (-7)  /* slab mode: another rank left the step loop with an error, or did not answer in time */

/* status bits reported by osph_sync() */
#define OSPH_S_NONFINITE   1u  /* a NaN/Inf acceleration or density was produced */
#define OSPH_S_SMALL_DT    2u  /* dt < 1e-6 (reference src/Solver.py:220-224 records this as ts_error) */
#define OSPH_S_UNBINNED    4u  /* a particle's reference cell id fell outside the table (reference: OOB write) */
#define OSPH_S_GRID_COARSE 8u  /* acceleration grid was coarsened to fit the allocated cell table */
#define OSPH_S_SKIN_EXHAUSTED 32u /* slab cadence: a build reused the binning although the particles had moved too far (pairs may be missing) */
#define OSPH_S_H_NOT_UNIFORM 64u /* OSPH_H_FIXED: a fluid particle reached the pair kernel's inputs with h != fixed_h (internal error) */
#define OSPH_S_DENSE_CELL  16u /* a cell holds more than 32768 particles: its summation order is not canonical (results exact) */

typedef struct osph_ctx osph_ctx;

/*
 * Everything the reference spreads over the constructors of WCSPH (src/Methods/WCSPH.py:35-79),
 * PEC/Verlet (src/Integrators/PEC.py:13-28), Solver (src/Solver.py:46,109,215) and the kernel choice.
 */
typedef struct osph_config {
    int32_t struct_size;          /* sizeof(osph_config), checked */
    int32_t device;               /* CUDA device ordinal */
    int32_t precision;            /* OSPH_FP64 | OSPH_FP32 */
    int32_t kernel;               /* OSPH_KERNEL_* */
    int32_t integrator;           /* OSPH_INTEGRATOR_* */
    int32_t method_xsph;          /* WCSPH.useXSPH */
    int32_t integrator_xsph;      /* PEC.useXSPH / Verlet.useXSPH */
    int32_t strict;               /* PEC.strict: clamp rho >= 0 */
    int32_t summation_density;    /* WCSPH.useSummationDensity (Jacobi pass on device) */
    int32_t dynamic_h;            /* OSPH_H_*: how the fluid smoothing length is refreshed before each force evaluation */
    int32_t reorder_every;        /* physical re-sort cadence of the device state, steps; 0 = default */
    int32_t reserved0;
    double  fixed_h;              /* Solver(h=...) when dynamic_h == 0 */
    double  h_sigma;              /* 1.3 */
    double  nn_scale;             /* NNLinkedList(scale=2.0), src/Solver.py:109 */
    double  gamma, B, rho0, Pb, co;       /* Tait EOS */
    double  alpha, beta;                  /* artificial viscosity */
    double  epsilon;                      /* XSPH */
    double  r0, D, p1, p2;                /* Lennard-Jones boundary force */
    double  gravity;                      /* 9.81, subtracted from ay (WCSPH.py:168-169) */
    double  cfl_courant, cfl_force;       /* 0.25, 0.25 (src/Solver.py:215) */
    double  height;                       /* WCSPH.height: fluid column height of the hydrostatic initialisation */
} osph_config;

/* Fill *cfg with the reference defaults for WCSPH(height, r0, rho0, ...). replaces: src/Methods/WCSPH.py:56-79 */
int osph_default_config(osph_config *cfg, double height, double r0, double rho0);

int osph_create(const osph_config *cfg, osph_ctx **out);
int osph_destroy(osph_ctx *ctx);
const char *osph_last_error(const osph_ctx *ctx);   /* ctx may be NULL: error of the last failed osph_create */
int osph_version(void);

/* ---- host <-> device state ------------------------------------------------------------------ */

/*
 * Upload the whole particle array (n rows, `stride` bytes apart).  Rows with deleted != 0 are kept
 * verbatim in a device mirror and never touched.  replaces: the `self.particleArray[self.indexes]`
 * fancy-index copies handed to nn.update/_loop (src/Solver.py:238,254).
 */
int osph_upload_aos(osph_ctx *ctx, const void *pA, int64_t n, int64_t stride);
/* Write the current device state back into the caller's array (same n/stride as the upload). */
int osph_download_aos(osph_ctx *ctx, void *pA, int64_t n, int64_t stride);
/* Same, but the caller supplies the device-side view: nothing crosses PCIe.  d_pA is a device pointer. */
int osph_import_device_aos(osph_ctx *ctx, const void *d_pA, int64_t n, int64_t stride);
int osph_export_device_aos(osph_ctx *ctx, void *d_pA, int64_t n, int64_t stride);
/*
 * Download `nfields` columns (enum osph_field) of the ACTIVE particles, in active order, into
 * cols[k][0..n_active).  replaces: the per-step `np.copy(self.particleArray[key])` of Solver._store
 * (src/Solver.py:477-482) without moving the other 150 bytes of every record.
 */
int osph_download_fields(osph_ctx *ctx, int32_t nfields, const int32_t *fields, double *const *cols);
/* Overwrite columns of the active particles from host arrays in active order (coupling write-back). */
int osph_upload_fields(osph_ctx *ctx, int32_t nfields, const int32_t *fields, const double *const *cols);
/*
 * Asynchronous, double-buffered form of osph_download_fields.  osph_export_begin snapshots the columns in stream
 * order (later steps may change the state) and starts their device->host copy on a separate copy stream into
 * pinned memory owned by the library; it returns at once with a ticket.  osph_export_end(ticket) waits for that
 * copy only and writes cols[k][0..n).  row_space == 0: active particles in active order, n = osph_num_active at the
 * time of the begin (as osph_download_fields).  row_space != 0: one value per ROW of the uploaded array in host row
 * order, n = rows uploaded; deleted rows carry the uploaded value -- exactly the array Solver._store appends, with no
 * host-side merge.  Two tickets may be in flight; a third begin returns OSPH_E_CAPACITY until the oldest is ended.
 * replaces: the blocking per-step copy of Solver._store (src/Solver.py:477-486) -- SURVEY section 8(f) rank 2.
 */
int osph_export_begin(osph_ctx *ctx, int32_t nfields, const int32_t *fields, int32_t row_space, int64_t *ticket);
int osph_export_end(osph_ctx *ctx, int64_t ticket, int32_t nfields, double *const *cols, int64_t n);
/*
 * Transfers of a FEW rows: the records of host rows rows[0..nrows) are written to / read from the caller's array
 * at pA + rows[k] * stride (same array convention as osph_download_aos); every row must be active.  Upload
 * overwrites the 19 doubles of those rows (label / deleted are ignored) and invalidates the neighbour structure.
 * replaces: the whole-array copies around the Coupled rows of a structure, `pA[c_indexes] =
 * couplingIntegrator.predict(...)` / `.correct(...)` and the coupling callback (src/Solver.py:381-398) --
 * SURVEY section 8(f) rank 1.
 */
int osph_download_rows(osph_ctx *ctx, int64_t nrows, const int64_t *rows, void *pA, int64_t stride);
int osph_upload_rows(osph_ctx *ctx, int64_t nrows, const int64_t *rows, const void *pA, int64_t stride);
/*
 * Switch rows off (or back on) on the device: active[r] != 0 keeps row r of the uploaded array, 0 marks it deleted.
 * The current state of every active row is first written back into the device-side record mirror, so a row that is
 * switched off keeps its last state for osph_download_aos, and a row switched on again resumes from its record.
 * n must equal the row count of the last upload.  Invalidates the neighbour structure; not available in slab mode.
 * replaces: the removal of the TempBoundary gate after settling, `pA['deleted'][inds] = True` followed by the
 * re-derivation of the index sets (src/Solver.py:428-442) -- n bytes cross PCIe instead of the particle array twice.
 */
int osph_set_active(osph_ctx *ctx, const uint8_t *active, int64_t n);
int64_t osph_num_active(const osph_ctx *ctx);
int64_t osph_num_fluid(const osph_ctx *ctx);

/*
 * Initialise the fluid rows as Solver.setup() does: h (fixed or computeH on the uploaded rho), hydrostatic
 * density rho0 (1 + rho0 g (H - y) / B)^(1/gamma), pressure and speed of sound.
 * replaces: src/Solver.py:184-196 (WCSPH.initialize src/Methods/WCSPH.py:82-108, TaitEOS.py:46-65).
 */
int osph_initialize(osph_ctx *ctx);

/* ---- the five per-step calls of Solver.run() ------------------------------------------------- */

/* out = {min(dt_c, dt_f), dt_c, dt_f}.  replaces: TimeStep().compute(...) src/Equations/TimeStep.py:11-91 via Solver._minTimeStep (src/Solver.py:210-230) */
int osph_timestep(osph_ctx *ctx, double out[3]);
/* replaces: integrator.predict(dt, pA[f_indexes], damping)  src/Solver.py:380 */
int osph_predict(osph_ctx *ctx, double dt, double damping);
/* replaces: nn.update(pA[indexes]) + the h refresh  src/Solver.py:238-246 (NNLinkedList.py:37-39,86-141) */
int osph_build_neighbours(osph_ctx *ctx);
/* replaces: _loop(pA[indexes], kernel.evaluate, kernel.gradient, method, nn)  src/Solver.py:254 (SolverTools.py:120-174) */
int osph_compute(osph_ctx *ctx);
/* replaces: integrator.correct(dt, pA[f_indexes], damping)  src/Solver.py:396 */
int osph_correct(osph_ctx *ctx, double dt, double damping);
/*
 * nsteps whole steps (timestep -> predict -> neighbours -> compute -> correct) without host round
 * trips; fixed_dt <= 0 selects the dynamic time step.  replaces: the body of the while loop
 * src/Solver.py:366-399 for runs without coupling.  The (dt, dt_c, dt_f) series is kept on the device
 * and fetched with osph_get_dt_log.
 */
int osph_step(osph_ctx *ctx, int32_t nsteps, double fixed_dt, double damping);
/* Up to cap triples of the series recorded by osph_step since the last call; returns the count via *count. */
int osph_get_dt_log(osph_ctx *ctx, double *out, int64_t cap, int64_t *count);
/* replaces: KineticEnergy(J, pA[f_indexes])  src/Equations/KineticEnergy.py:6-12 (sum over the fluid rows) */
int osph_kinetic_energy(osph_ctx *ctx, double *ke);
/*
 * Sort cadence: out[0] = neighbour-structure builds since osph_create, out[1] = those that binned and sorted the particles.
 * The others reused the sorted order and cell table of the last sort: cells are one pair radius plus a skin wide, and the
 * device re-sorts as soon as (pair radius + 2 x largest displacement since the sort) exceeds the cell size, so the 3 x 3
 * cell walk stays a complete candidate set; membership is always decided on the CURRENT positions.  Environment:
 * OSPH_SKIN = skin / pair radius ("auto" default, 0 = sort at every build).  Slab mode sorts at every build.
 */
int osph_sort_stats(osph_ctx *ctx, int64_t out[2]);
/* Block until all enqueued work is done; *status receives the OSPH_S_* bits accumulated since the last call. */
int osph_sync(osph_ctx *ctx, uint32_t *status);

/* ---- validation / query entry points --------------------------------------------------------- */

/*
 * Reference grid of the last osph_build_neighbours: grid = {xmin, xmax, ymin, ymax, cell_size, ncx, ncy}
 * and, per active particle, the flat cell id the reference's _bin computes (NNLinkedList.py:129-141).
 */
int osph_get_cells(osph_ctx *ctx, double grid[7], int64_t *cell_ids);
/*
 * CSR neighbour lists (reference predicate: 3x3 cells AND r/h_ij <= 3.0, NNLinkedList.py:50-71) of the
 * fluid rows, indices in active order, each list sorted ascending.  offsets has n_active+1 entries.
 * Pass idx == NULL to obtain only offsets/total.  replaces: nn.near(i, pA) for every i.
 */
int osph_get_neighbours_csr(osph_ctx *ctx, int64_t *offsets, int64_t *idx, int64_t cap, int64_t *total);
/*
 * Neighbours of an arbitrary point; same predicate, results sorted by index.  Returns the count in
 * *count (may exceed cap).  replaces: nn.nearPos(x, y, h, pA)  NNLinkedList.py:41-80 (IceBreak pressure probe).
 */
int osph_near_pos(osph_ctx *ctx, double x, double y, double h, int64_t cap,
                  int64_t *idx, double *r, double *q, double *hij, int64_t *count);
/*
 * SPH-interpolated density and Tait pressure at n arbitrary points (host arrays), all points in one launch:
 * neighbours by the reference predicate with query smoothing length h, W(r, h), Shepard normalisation, summation
 * density over the fluid neighbours.  replaces: pressure_SPH, examples/IceBreak.py:252-285 (nn.nearPos +
 * kernel.evaluate + Shepard + SummationDensity + Tait, once per ice node per step on the host).
 */
int osph_probe_pressure(osph_ctx *ctx, int64_t n, const double *x, const double *y, double h, double *rho_out,
                        double *p_out);
/* Device time of each phase since creation, ms: {timestep, predict, neighbours, compute, correct, transfer}. */
int osph_get_timers(osph_ctx *ctx, double out_ms[6]);
/* Kernel launches issued by this context since creation (bench.py reports them as gpu_launches). */
int64_t osph_launch_count(const osph_ctx *ctx);
/* CUDA stream of the context as a uintptr (so callers can record their own events on it). */
uint64_t osph_stream(const osph_ctx *ctx);
/* Average device time of the fused pair kernel over the launches since the last call, microseconds. */
int osph_pair_kernel_time(osph_ctx *ctx, double *avg_us, int64_t *launches);
/*
 * out[0] = launches of the fused pair kernel since osph_create, out[1] = those that ran its uniform-smoothing-length
 * instantiation: with OSPH_H_FIXED (Solver(h=value), the reference's DamBreak set-up) every fluid particle carries the same
 * h, so h_ij, 1/h_ij, the support test and the kernel normalisation of a fluid-fluid pair are loop constants.  Every pair is
 * evaluated with the operations of the general instantiation (tests/test_gpu_parity.py holds the two to the same bits, or,
 * where the software-pipelined flush loop defers the rare pairs of a flush, to summation order); OSPH_UH=0 in the environment
 * selects the general one.  out[2] = how the library was built: bit 0 uniform-h instantiation present (PAIR_UH), bit 1
 * sign-bit clamps (PAIR_ISIGN), bit 2 software-pipelined flush loop (PAIR_UH_PIPE).
 */
int osph_pair_kernel_info(osph_ctx *ctx, int64_t out[3]);

/* ---- 1-D slab decomposition: one context per GPU, exchange buffers owned by the caller -------------------
 *
 * The reference is single-process; this section has no counterpart there.  A rank owns
 * the particles with x_lo <= x < x_hi.  Buffers passed here are DEVICE pointers owned by the caller (torch
 * tensors that NCCL sends from / receives into); records are doubles:
 *   halo / ghost record  OSPH_WIRE_HALO = 8  : x y vx vy rho m h label
 *   migrant record       OSPH_WIRE_FULL = 21 : the 19 columns (enum osph_field order), label, global row id
 * Per step:  osph_slab_step_begin -> osph_slab_pack -> [exchange] -> osph_slab_commit -> osph_slab_step_end.
 */
#define OSPH_WIRE_HALO 8
#define OSPH_WIRE_FULL 21
/* Reserve room for `particle_capacity` owned + ghost particles at the next upload (slabs grow by migration). */
int osph_reserve(osph_ctx *ctx, int64_t particle_capacity);
/* Replace the row ids recorded at upload (0..n_active-1) by global ids, so migrants keep their identity. */
int osph_set_row_ids(osph_ctx *ctx, const int32_t *ids, int64_t n);
/* Enter slab mode; d_ghost has room for ghost_capacity halo records. */
int osph_slab_configure(osph_ctx *ctx, double x_lo, double x_hi, void *d_ghost, int64_t ghost_capacity);
/* Optional, before osph_slab_dt_local of step `step` of a sequencer call that runs `nsteps` steps: lets the library
 * fuse the corrector of every step but the last into the next step's predictor pass (PEC), exactly as osph_step does
 * on one GPU; results are bit-identical to unplanned steps.  The state is complete again after the last step.
 * OSPH_SLAB_FUSED=0 in the environment disables the fusion. */
int osph_slab_step_plan(osph_ctx *ctx, int32_t step, int32_t nsteps);
/* d_out3 (device) <- local {h_min, -c_max, -a2_max} over the owned fluid rows (all-reduce with MIN). */
int osph_slab_dt_local(osph_ctx *ctx, double *d_out3);
/* Install the all-reduced triple (device pointer, may be NULL to keep the local one), form dt on the device
 * (fixed_dt <= 0: dynamic) and run the predictor + local grid reductions. */
int osph_slab_step_begin(osph_ctx *ctx, const double *d_dt_reduced3, double fixed_dt, double damping);
/* Classify the owned particles: migrants (x outside the slab) are packed as full records and kept as this
 * rank's first ghosts; particles within halo_width of a face are packed as halo records.  d_meta (device, 12
 * doubles) <- {mig_left, mig_right, halo_left, halo_right, xmin, -xmax, ymin, -ymax, hmin_all, -hmax, overflow, 0}. */
int osph_slab_pack(osph_ctx *ctx, double halo_width, void *d_mig_left, void *d_mig_right, int64_t mig_cap,
                   void *d_halo_left, void *d_halo_right, int64_t halo_cap, double *d_meta);
/* Drop the n_mig_out migrants recorded by osph_slab_pack (hole filling), append the n_mig_in received ones,
 * declare n_ghost records present in the ghost buffer and install the all-reduced grid scalars
 * {xmin, -xmax, ymin, -ymax, hmin_all, -hmax} so that every rank forms the same reference grid. */
int osph_slab_commit(osph_ctx *ctx, int64_t n_mig_out, const void *d_mig_in, int64_t n_mig_in, int64_t n_ghost,
                     const double global_bounds[6]);
/* Neighbour structure over owned + ghost particles, fused pair kernel for the owned fluid rows, corrector. */
int osph_slab_step_end(osph_ctx *ctx, double damping);
/* Host copy of the owned particles as packed records in storage order (deleted = 0) and their global ids;
 * the pair (pA, ids) is valid input for osph_upload_aos + osph_set_row_ids. */
int osph_download_owned(osph_ctx *ctx, void *pA, int64_t cap_rows, int64_t stride, int32_t *ids, int64_t *n_out);
/* Copy ids, labels and columns of the owned particles (storage order) into caller-owned device buffers. */
int osph_slab_export(osph_ctx *ctx, int32_t *d_ids, int8_t *d_label, int32_t nfields, const int32_t *fields,
                     double *const *d_cols);

/*
 * The same protocol sequenced inside the library with direct NCCL calls (no interpreter between the collectives).
 * libnccl.so.2 is resolved at run time: the copy already loaded into the process (torch's) if any, else
 * `libnccl_path`, else the system one.  Rendezvous stays with the caller: rank 0 obtains a unique id with
 * osph_nccl_unique_id and distributes the 128 bytes (e.g. torch.distributed.broadcast).
 */
typedef struct osph_slab_comm osph_slab_comm;
int osph_nccl_unique_id(const char *libnccl_path, char out[128]);
const char *osph_nccl_last_error(void);
/* Communicator + exchange buffers for a context that already holds this rank's particles (x_lo <= x < x_hi).
 * hmax: largest smoothing length at start (halo width = 1.1 * pair radius, refreshed from the all-gathered h_max). */
int osph_slab_comm_create(osph_ctx *ctx, const char *libnccl_path, const char unique_id[128], int rank, int world,
                          double x_lo, double x_hi, double r0, double hmax, int64_t mig_cap, int64_t halo_cap,
                          osph_slab_comm **out);
int osph_slab_comm_destroy(osph_ctx *ctx, osph_slab_comm *comm);
/* Re-enter slab mode after osph_upload_aos replaced the particle set (host-buffer call pattern). */
int osph_slab_comm_attach(osph_ctx *ctx, osph_slab_comm *comm);
/* Move this rank's slab boundaries (load re-balancing; all ranks switch at the same step, the particles between the
 * old and the new cut reach their new owner through the next step's migration). */
int osph_slab_comm_set_bounds(osph_ctx *ctx, osph_slab_comm *comm, double x_lo, double x_hi);
/* nsteps slab-decomposed steps: all_reduce(dt) -> predict -> pack -> all_gather(counts, bounds) -> send/recv
 * (migrants + halos) -> commit -> neighbours + pair kernel + correct. */
int osph_slab_run(osph_ctx *ctx, osph_slab_comm *comm, int32_t nsteps, double fixed_dt, double damping);
/* {mig_out l,r, halo_out l,r, mig_in l,r, halo_in l,r} of the last step */
int osph_slab_last_counts(const osph_slab_comm *comm, int64_t out[8]);

/*
 * The same protocol over NVLink peer memory, no NCCL in the data path (csrc/slab_p2p.cu): every rank exports one
 * HBM window with CUDA IPC; the pack kernel writes migrants and halos straight into the neighbours' windows, two
 * mailbox all-gather kernels (store to every peer, system fence, sequence flag, spin on the own window) carry dt and
 * counts + grid bounds.  All ranks must live on one NVLink / NVSwitch box.  The caller all-gathers the 64-byte
 * handles (rank-major) between osph_slab_p2p_create and osph_slab_p2p_connect.
 */
typedef struct osph_slab_p2p osph_slab_p2p;
int osph_slab_p2p_create(osph_ctx *ctx, int rank, int world, double x_lo, double x_hi, double r0, double hmax,
                         int64_t mig_cap, int64_t halo_cap, osph_slab_p2p **out, char handle_out[64]);
int osph_slab_p2p_connect(osph_ctx *ctx, osph_slab_p2p *p2p, const char *all_handles);
int osph_slab_p2p_destroy(osph_ctx *ctx, osph_slab_p2p *p2p);
int osph_slab_p2p_attach(osph_ctx *ctx, osph_slab_p2p *p2p);
int osph_slab_p2p_set_bounds(osph_ctx *ctx, osph_slab_p2p *p2p, double x_lo, double x_hi);
int osph_slab_p2p_run(osph_ctx *ctx, osph_slab_p2p *p2p, int32_t nsteps, double fixed_dt, double damping);
int osph_slab_p2p_last_counts(const osph_slab_p2p *p2p, int64_t out[8]);
/* Slab cadence: out[0] = steps that sorted (migration, fresh halo lists), out[1] = steps that reused the binning: the same
 * halo particles re-sent into the same record slots, no migration, no host wait.  The ranks decide together from the
 * all-gathered displacement of the step before.  OSPH_SLAB_CADENCE=0: every step sorts. */
int osph_slab_p2p_stats(const osph_slab_p2p *p2p, int64_t out[2]);

/* ---- stand-alone leaf functions on host arrays (context-free; `device` is a CUDA ordinal) ---------- */

/* what == 0: kernel.evaluate(r, h); what == 1: kernel.gradient(x, r, h).
 * replaces: src/Kernels/CubicSpline.py:10-70, Wendland.py:9-64, Gaussian.py:16-59 */
int osph_leaf_kernel(int device, int kernel, int what, int64_t n, const double *x, const double *r,
                     const double *h, double *out);
/* replaces: WCSPH.compute_pressure src/Methods/WCSPH.py:131-149 (TaitEOS ufunc, TaitEOS.py:6-31) */
int osph_leaf_tait_pressure(int device, int64_t n, const double *rho, const int8_t *label, double gamma,
                            double B, double rho0, double Pb, double *out);
/* replaces: TaitEOS_height src/Equations/TaitEOS.py:46-65 */
int osph_leaf_tait_height(int device, int64_t n, const double *y, double rho0, double H, double B,
                          double gamma, double *out);
/* replaces: computeH src/Tools/SolverTools.py:106-118 */
int osph_leaf_compute_h(int device, int64_t n, double sigma, const double *m, const double *rho, double *out);
/*
 * The per-neighbour equations on ONE table of J computed neighbours (the reference's computed_dtype, src/Common.py:59-105:
 * differences are i - j, h is h_ij), for callers that use the equations outside the fused pair kernel -- the reference's
 * own equation tests (test/test_numba_momentum.py, test_numba_continuity.py, test_eq_boundary.py) and
 * WCSPH.compute_density_change / compute_acceleration / compute_velocity (src/Methods/WCSPH.py:151-203).
 * cols: OSPH_COMP_NCOLS columns of J doubles each, column k at cols + k * J, in the order OSPH_COMP_*;
 * self_*: p, rho, h, c of particle i.  One launch evaluates all four equations, deterministic summation:
 *   out[0]    Continuity     src/Equations/Continuity.py:5-17     sum over fluid j of m (v . dW)
 *   out[1,2]  Momentum       src/Equations/Momentum.py:6-57       (no gravity)
 *   out[3,4]  XSPH           src/Equations/XSPH.py:6-31           (the correction, every label)
 *   out[5,6]  BoundaryForce  src/Equations/BoundaryForce.py:7-42  (non-fluid j with 1e-12 < r <= r0)
 */
enum { OSPH_COMP_M = 0, OSPH_COMP_P, OSPH_COMP_RHO, OSPH_COMP_H, OSPH_COMP_C, OSPH_COMP_R, OSPH_COMP_W, OSPH_COMP_DWX,
       OSPH_COMP_DWY, OSPH_COMP_X, OSPH_COMP_Y, OSPH_COMP_VX, OSPH_COMP_VY, OSPH_COMP_NCOLS };
int osph_leaf_equations(int device, int64_t J, const int8_t *label, const double *cols, double self_p, double self_rho,
                        double self_h, double self_c, double alpha, double beta, double epsilon, double r0, double D,
                        double p1, double p2, double out[7]);
/* The positional columns of a computed-neighbour table: differences self - neighbour of x, y, vx, vy
 * (replaces the arithmetic of _assignProps, src/Tools/SolverTools.py:97-101; the other columns are copies).
 * self4 = {x, y, vx, vy} of particle i; nbr / out: 4 columns of J doubles each (x, y, vx, vy), column k at + k * J. */
int osph_leaf_differences(int device, int64_t J, const double self4[4], const double *nbr, double *out);
/* replaces: Courant src/Equations/Courant.py:4-31 (alpha * min h / max c, with the reference's 10e10 / 1e-10 seeds) */
int osph_leaf_courant(int device, double alpha, int64_t J, const double *h, const double *c, double *out);
const char *osph_leaf_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* OSPH_H */
