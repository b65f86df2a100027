// api.cu -- the extern "C" surface declared in include/osph.h: context life cycle, host<->device
// transfers of the reference's packed particle records, the five per-step calls of Solver.run()
// (reference src/Solver.py:366-399), the fused multi-step loop and the validation queries.
// There is deliberately no CPU implementation behind any entry point.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "step.cuh"

static std::string g_create_error;

enum { T_TIMESTEP = 0, T_PREDICT, T_NEIGHBOURS, T_COMPUTE, T_CORRECT, T_TRANSFER };

struct PhaseTimer {
    osph_ctx *ctx; int id; int slot;
    PhaseTimer(osph_ctx *c, int id_) : ctx(c), id(id_), slot(-1)
    {
        if ((int)ctx->phase_ev.size() / 2 >= 2048) drain(ctx);
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
        ctx->phase_ev.push_back(a); ctx->phase_ev.push_back(b); ctx->phase_id.push_back(id);
        slot = (int)ctx->phase_id.size() - 1;
        cudaEventRecord(a, ctx->stream);
    }
    ~PhaseTimer() { if (slot >= 0) cudaEventRecord(ctx->phase_ev[2 * slot + 1], ctx->stream); }
    static void drain(osph_ctx *ctx)
    {
        cudaStreamSynchronize(ctx->stream);
        for (size_t k = 0; k < ctx->phase_id.size(); k++) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ctx->phase_ev[2 * k], ctx->phase_ev[2 * k + 1]) == cudaSuccess)
                ctx->timers_ms[ctx->phase_id[k]] += ms;
            cudaEventDestroy(ctx->phase_ev[2 * k]); cudaEventDestroy(ctx->phase_ev[2 * k + 1]);
        }
        ctx->phase_ev.clear(); ctx->phase_id.clear();
    }
};

static void free_particles(osph_ctx *ctx)
{
    cudaFree(ctx->d_aos); cudaFree(ctx->d_row); cudaFree(ctx->d_act);
    for (int k = 0; k < OSPH_NUM_FIELDS; k++) { cudaFree(ctx->f[k]); ctx->f[k] = nullptr; }
    cudaFree(ctx->label); cudaFree(ctx->scratch); cudaFree(ctx->d_stage); cudaFree(ctx->xref); cudaFree(ctx->yref);
    ctx->xref = ctx->yref = nullptr;
    cudaFree(ctx->s_coarse); cudaFree(ctx->s_gcell); cudaFree(ctx->s_pos); cudaFree(ctx->s_vel);
    cudaFree(ctx->s_rm); cudaFree(ctx->s_hp); cudaFree(ctx->s_info);
    cudaFree(ctx->d_partial);
    osph_sort_free(ctx);
    ctx->d_aos = nullptr; ctx->d_row = ctx->d_act = nullptr; ctx->label = nullptr;
    ctx->scratch = ctx->d_stage = nullptr; ctx->s_coarse = nullptr; ctx->s_gcell = nullptr; ctx->s_pos = nullptr;
    ctx->s_vel = ctx->s_rm = ctx->s_hp = nullptr; ctx->s_info = nullptr;
    ctx->d_partial = nullptr;
    ctx->cap = 0; ctx->aos_bytes = 0;
}

static int alloc_particles(osph_ctx *ctx, int64_t n_active, int64_t n_total, int64_t stride)
{
    size_t aos_bytes = (size_t)std::max<int64_t>(n_total, ctx->reserve) * (size_t)stride;
    if (std::max<int64_t>(n_active, ctx->reserve) <= ctx->cap && aos_bytes <= ctx->aos_bytes) return 0;
    free_particles(ctx);
    int64_t cap = std::max<int64_t>(std::max<int64_t>(n_active, ctx->reserve), 1024);
    const size_t rs = ctx->cfg.precision == OSPH_FP64 ? sizeof(double2) : sizeof(float2);
    OSPH_CUDA(cudaMalloc(&ctx->d_aos, std::max<size_t>(aos_bytes, 16)));
    OSPH_CUDA(cudaMalloc(&ctx->d_row, sizeof(int) * cap));
    OSPH_CUDA(cudaMalloc(&ctx->d_act, sizeof(int) * cap));
    for (int k = 0; k < OSPH_NUM_FIELDS; k++) OSPH_CUDA(cudaMalloc(&ctx->f[k], sizeof(double) * cap));
    OSPH_CUDA(cudaMalloc(&ctx->label, cap));
    OSPH_CUDA(cudaMalloc(&ctx->scratch, sizeof(double) * cap));
    OSPH_CUDA(cudaMalloc(&ctx->d_stage, sizeof(double) * cap));
    OSPH_CUDA(cudaMalloc(&ctx->xref, sizeof(double) * cap));
    OSPH_CUDA(cudaMalloc(&ctx->yref, sizeof(double) * cap));
    ctx->skin_valid = false;
    OSPH_CUDA(cudaMalloc(&ctx->s_coarse, sizeof(int4) * cap));
    OSPH_CUDA(cudaMalloc(&ctx->s_gcell, sizeof(int2) * cap));
    OSPH_CUDA(cudaMalloc(&ctx->s_pos, sizeof(double2) * cap));
    OSPH_CUDA(cudaMalloc(&ctx->s_vel, rs * cap));
    OSPH_CUDA(cudaMalloc(&ctx->s_rm, rs * cap));
    OSPH_CUDA(cudaMalloc(&ctx->s_hp, rs * cap));
    OSPH_CUDA(cudaMalloc(&ctx->s_info, sizeof(int) * cap));
    OSPH_CUDA(cudaMalloc(&ctx->d_partial, sizeof(double) * (div_up(cap, 256) + 1)));
    int rc = osph_sort_alloc(ctx, cap);
    if (rc) return rc;
    ctx->cap = cap; ctx->aos_bytes = aos_bytes;
    return 0;
}

extern "C" int osph_version(void) { return OSPH_VERSION; }

extern "C" int osph_default_config(osph_config *cfg, double height, double r0, double rho0)
{
    if (!cfg) return OSPH_E_INVALID;
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = (int32_t)sizeof(osph_config);
    cfg->precision = OSPH_FP64; cfg->kernel = OSPH_KERNEL_CUBIC; cfg->integrator = OSPH_INTEGRATOR_PEC;
    cfg->method_xsph = 1; cfg->integrator_xsph = 1; cfg->strict = 1;
    cfg->dynamic_h = 1; cfg->fixed_h = 0.0; cfg->h_sigma = 1.3; cfg->nn_scale = 2.0;
    cfg->gamma = 7.0;
    cfg->co = 10.0 * sqrt(2 * 9.81 * height);          // TaitEOS_co, reference src/Equations/TaitEOS.py:40-44
    cfg->B = cfg->co * cfg->co * rho0 / cfg->gamma;    // TaitEOS_B, :33-38
    cfg->rho0 = rho0; cfg->Pb = 0.0;
    cfg->alpha = 0.01; cfg->beta = 0.0; cfg->epsilon = 0.5;
    cfg->r0 = r0; cfg->D = 5 * 9.81 * height; cfg->p1 = 4; cfg->p2 = 2;
    cfg->gravity = 9.81; cfg->cfl_courant = 0.25; cfg->cfl_force = 0.25; cfg->height = height;
    return 0;
}

extern "C" const char *osph_last_error(const osph_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int osph_create(const osph_config *cfg, osph_ctx **out)
{
    if (!cfg || !out || cfg->struct_size != (int32_t)sizeof(osph_config)) {
        g_create_error = "osph_create: null argument or osph_config.struct_size mismatch";
        return OSPH_E_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev) {
        g_create_error = std::string("osph_create: no usable CUDA device (") +
                         (e != cudaSuccess ? cudaGetErrorString(e) : "device ordinal out of range") +
                         "); libosph_b200 has no CPU path";
        cudaGetLastError();
        return OSPH_E_NO_DEVICE;
    }
    if (cfg->precision != OSPH_FP64 && cfg->precision != OSPH_FP32) { g_create_error = "bad precision"; return OSPH_E_INVALID; }
    if (cfg->kernel < 0 || cfg->kernel > 2 || cfg->integrator < 0 || cfg->integrator > 2) {
        g_create_error = "bad kernel / integrator id"; return OSPH_E_INVALID;
    }
    osph_ctx *ctx = new osph_ctx();
    ctx->cfg = *cfg; ctx->device = cfg->device;
    {   // OSPH_SORT=radix: the LSD radix sort of sort.cu instead of the counting sort by cell (same order, A/B runs)
        const char *e = getenv("OSPH_SORT");
        ctx->bin_sort = !(e && std::string(e) == "radix");
        // OSPH_SKIN: skin of the sort cadence as a fraction of the pair radius (e.g. 0.1); "auto" / unset = sized on the
        // device from the observed displacement per build; 0 = sort at every build (the behaviour before the cadence)
        // OSPH_FUSE_SCALARS (A/B runs, default 0): bit 0 = k_grid_params in the last CTA of the predictor pass, bit 1 =
        // k_timestep in the last CTA of the pair kernel.  Measured on the B200 (profiles/r02/README.md): neither pays -- the
        // single-thread tail is as long as the one-thread kernel it replaces (0.4027 vs 0.4014 ms per step), and in the pair
        // kernel the fence + barrier + ticket of every CTA costs 10 us (305 vs 295 us).
        const char *f = getenv("OSPH_FUSE_SCALARS");
        const int bits = f ? atoi(f) : 0;
        ctx->grid_fusion = (bits & 1) != 0; ctx->timestep_fusion = (bits & 2) != 0;
        const char *k = getenv("OSPH_SKIN");
        ctx->skin_frac = (k && *k && std::string(k) != "auto") ? atof(k) : -1.0;
        if (ctx->skin_frac > 1.0) ctx->skin_frac = 1.0;
    }
    auto fail = [&](const char *what, cudaError_t err) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        delete ctx; return OSPH_E_CUDA;
    };
    if ((e = cudaSetDevice(ctx->device)) != cudaSuccess) return fail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    for (int k = 0; k < 2 * OSPH_PAIR_EVENTS; k++)
        if ((e = cudaEventCreate(&ctx->pair_ev[k])) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaMalloc(&ctx->d_grid, sizeof(GridParams))) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMemset(ctx->d_grid, 0, sizeof(GridParams))) != cudaSuccess) return fail("cudaMemset", e);
    if ((e = cudaMalloc(&ctx->d_sc, sizeof(StepScalars))) != cudaSuccess) return fail("cudaMalloc", e);
    ctx->dt_log_cap = 1 << 16;
    if ((e = cudaMalloc(&ctx->d_dt_log, sizeof(double) * 3 * ctx->dt_log_cap)) != cudaSuccess) return fail("cudaMalloc", e);
    if (osph_init_scalars(ctx) != 0) { g_create_error = ctx->err; delete ctx; return OSPH_E_CUDA; }
    *out = ctx;
    return 0;
}

extern "C" int osph_destroy(osph_ctx *ctx)
{
    if (!ctx) return OSPH_E_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    PhaseTimer::drain(ctx);
    osph_export_free(ctx);
    free_particles(ctx);
    osph_bin_free(ctx);
    cudaFree(ctx->cell_range); cudaFree(ctx->d_grid); cudaFree(ctx->d_sc); cudaFree(ctx->d_dt_log);
    cudaFree(ctx->scan_block); cudaFree(ctx->d_slab_counters); cudaFree(ctx->d_mig_slots); cudaFree(ctx->d_tail_flag);
    cudaFree(ctx->d_holes); cudaFree(ctx->d_fillers); cudaFree(ctx->d_uh);
    for (int k = 0; k < 2 * OSPH_PAIR_EVENTS; k++) if (ctx->pair_ev[k]) cudaEventDestroy(ctx->pair_ev[k]);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

#define CHECK_CTX()                                                        \
    if (!ctx) return OSPH_E_INVALID;                                       \
    OSPH_CUDA(cudaSetDevice(ctx->device))
#define NEED_PARTICLES()                                                   \
    if (ctx->n <= 0) { ctx->err = "no particles uploaded"; return OSPH_E_INVALID; }

static void invalidate_state(osph_ctx *ctx)
{
    ctx->prepared = false; ctx->neighbours_valid = false; ctx->reductions_valid = false;
}

// ---- transfers ---------------------------------------------------------------------------------

// Rebuilds the active list from the device-side record mirror (flags, scan, compaction) and splits the active rows into
// the state columns.  Shared by the uploads and osph_set_active.
static int rebuild_active(osph_ctx *ctx)
{
    int rc;
    int *d_counters = (int *)ctx->d_partial;
    OSPH_CUDA(cudaMemsetAsync(d_counters, 0, sizeof(int) * 2, ctx->stream));
    if ((rc = osph_launch_active_list(ctx, (int)ctx->n_total, d_counters))) return rc;
    int counters[2] = {0, 0};
    OSPH_CUDA(cudaMemcpyAsync(counters, d_counters, sizeof(counters), cudaMemcpyDeviceToHost, ctx->stream));
    OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    if (counters[0] != ctx->n) ctx->sized = false;
    ctx->n = counters[0]; ctx->n_fluid = counters[1];
    if (ctx->n > 0 && (rc = osph_launch_unpack(ctx))) return rc;
    ctx->c_uniform = false; ctx->build_counter = 0; ctx->n_ghost = 0; ctx->slab = false;
    ctx->slab_fused = 0; ctx->slab_defer = false; ctx->slab_last = true;
    ctx->skin_valid = false;
    invalidate_state(ctx);
    return 0;
}

// Upload path: raw records -> device mirror; the list of active rows is built ON the device (flags, exclusive
// scan, compaction) so the host touches nothing but two counters.
static int ingest(osph_ctx *ctx, const void *src, bool src_on_device, int64_t n, int64_t stride)
{
    int rc = alloc_particles(ctx, n, n, stride);          // capacity for the case that every row is active
    if (rc) return rc;
    if (n != ctx->n_total) ctx->sized = false;
    ctx->n_total = n; ctx->stride = stride;
    OSPH_CUDA(cudaMemcpyAsync(ctx->d_aos, src, (size_t)n * stride,
                              src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    return rebuild_active(ctx);
}

// reference src/Solver.py:428-442 (removal of the TempBoundary gate after settling): rows are switched off (or on) ON the
// device.  The state goes back into the record mirror, the `deleted` byte of every row is set from the mask, and the active
// list is rebuilt -- n bytes cross PCIe instead of the whole particle array twice.
extern "C" int osph_set_active(osph_ctx *ctx, const uint8_t *active, int64_t n)
{
    CHECK_CTX();
    if (!active || n != ctx->n_total || n <= 0) { ctx->err = "osph_set_active: the mask must have one byte per uploaded row"; return OSPH_E_INVALID; }
    if (ctx->slab) { ctx->err = "osph_set_active: not available in slab mode"; return OSPH_E_INVALID; }
    PhaseTimer t(ctx, T_TRANSFER);
    int rc;
    if (ctx->n > 0 && (rc = osph_launch_pack(ctx))) return rc;          // current state -> records (also of rows about to go)
    unsigned char *d_mask = (unsigned char *)ctx->d_stage;             // one spare column: 8 bytes per particle capacity
    OSPH_CUDA(cudaMemcpyAsync(d_mask, active, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = osph_launch_set_deleted(ctx, d_mask))) return rc;
    return rebuild_active(ctx);
}

extern "C" int osph_upload_aos(osph_ctx *ctx, const void *pA, int64_t n, int64_t stride)
{
    CHECK_CTX();
    if (!pA || n < 0 || n > 0x7fffffff || stride < 2 + 8 * OSPH_NUM_FIELDS || (stride & 1)) { ctx->err = "osph_upload_aos: bad arguments (the record stride must be even: fields are moved as 16-bit words)"; return OSPH_E_INVALID; }
    PhaseTimer t(ctx, T_TRANSFER);
    return ingest(ctx, pA, false, n, stride);
}

extern "C" int osph_import_device_aos(osph_ctx *ctx, const void *d_pA, int64_t n, int64_t stride)
{
    CHECK_CTX();
    if (!d_pA || n < 0 || n > 0x7fffffff || stride < 2 + 8 * OSPH_NUM_FIELDS || (stride & 1)) { ctx->err = "osph_import_device_aos: bad arguments (the record stride must be even: fields are moved as 16-bit words)"; return OSPH_E_INVALID; }
    PhaseTimer t(ctx, T_TRANSFER);
    return ingest(ctx, d_pA, true, n, stride);
}

extern "C" int osph_download_aos(osph_ctx *ctx, void *pA, int64_t n, int64_t stride)
{
    CHECK_CTX();
    if (!pA || n != ctx->n_total || stride != ctx->stride) { ctx->err = "osph_download_aos: shape differs from the upload"; return OSPH_E_INVALID; }
    PhaseTimer t(ctx, T_TRANSFER);
    int rc = osph_launch_pack(ctx);
    if (rc) return rc;
    OSPH_CUDA(cudaMemcpyAsync(pA, ctx->d_aos, (size_t)n * stride, cudaMemcpyDeviceToHost, ctx->stream));
    OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int osph_download_owned(osph_ctx *ctx, void *pA, int64_t cap_rows, int64_t stride, int32_t *ids, int64_t *n_out)
{
    CHECK_CTX();
    if (!pA || !ids || stride != ctx->stride || cap_rows < ctx->n || (size_t)ctx->n * stride > ctx->aos_bytes) {
        ctx->err = "osph_download_owned: buffer too small or stride differs from the upload"; return OSPH_E_CAPACITY;
    }
    PhaseTimer t(ctx, T_TRANSFER);
    int *d_ids = (int *)ctx->d_stage;
    int rc = osph_launch_pack_owned(ctx, d_ids);
    if (rc) return rc;
    OSPH_CUDA(cudaMemcpyAsync(pA, ctx->d_aos, (size_t)ctx->n * stride, cudaMemcpyDeviceToHost, ctx->stream));
    OSPH_CUDA(cudaMemcpyAsync(ids, d_ids, sizeof(int) * ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
    OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n_out) *n_out = ctx->n;
    return 0;
}

extern "C" int osph_export_device_aos(osph_ctx *ctx, void *d_pA, int64_t n, int64_t stride)
{
    CHECK_CTX();
    if (!d_pA || n != ctx->n_total || stride != ctx->stride) { ctx->err = "osph_export_device_aos: shape differs from the import"; return OSPH_E_INVALID; }
    int rc = osph_launch_pack(ctx);
    if (rc) return rc;
    OSPH_CUDA(cudaMemcpyAsync(d_pA, ctx->d_aos, (size_t)n * stride, cudaMemcpyDeviceToDevice, ctx->stream));
    OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int osph_download_fields(osph_ctx *ctx, int32_t nfields, const int32_t *fields, double *const *cols)
{
    CHECK_CTX(); NEED_PARTICLES();
    PhaseTimer t(ctx, T_TRANSFER);
    for (int k = 0; k < nfields; k++) {
        if (fields[k] < 0 || fields[k] >= OSPH_NUM_FIELDS || !cols[k]) { ctx->err = "osph_download_fields: bad field"; return OSPH_E_INVALID; }
        int rc = osph_launch_col_to_active(ctx, fields[k], ctx->d_stage);
        if (rc) return rc;
        OSPH_CUDA(cudaMemcpyAsync(cols[k], ctx->d_stage, sizeof(double) * ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
        OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

extern "C" int osph_upload_fields(osph_ctx *ctx, int32_t nfields, const int32_t *fields, const double *const *cols)
{
    CHECK_CTX(); NEED_PARTICLES();
    PhaseTimer t(ctx, T_TRANSFER);
    for (int k = 0; k < nfields; k++) {
        if (fields[k] < 0 || fields[k] >= OSPH_NUM_FIELDS || !cols[k]) { ctx->err = "osph_upload_fields: bad field"; return OSPH_E_INVALID; }
        OSPH_CUDA(cudaMemcpyAsync(ctx->d_stage, cols[k], sizeof(double) * ctx->n, cudaMemcpyHostToDevice, ctx->stream));
        int rc = osph_launch_col_from_active(ctx, fields[k], ctx->d_stage);
        if (rc) return rc;
        OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    ctx->skin_valid = false;            // positions may have been replaced: the next build sorts
    invalidate_state(ctx);
    return 0;
}

extern "C" int64_t osph_num_active(const osph_ctx *ctx) { return ctx ? ctx->n : 0; }
extern "C" int64_t osph_num_fluid(const osph_ctx *ctx) { return ctx ? ctx->n_fluid : 0; }

extern "C" int osph_initialize(osph_ctx *ctx)
{
    CHECK_CTX(); NEED_PARTICLES();
    int rc = osph_launch_setup(ctx);
    if (rc) return rc;
    if (ctx->c_uniform) ctx->c_uniform = false;     // c now holds per-row values again (non-fluid rows keep theirs)
    ctx->skin_valid = false;
    invalidate_state(ctx);
    return 0;
}

// ---- the per-step calls --------------------------------------------------------------------------

// First build after an upload: size the cell table from the grid the device would like to use.
int osph_size_cell_table(osph_ctx *ctx)
{
    if (ctx->sized) return 0;
    int64_t saved = ctx->cell_cap;
    ctx->cell_cap = (int64_t)1 << OSPH_MAX_CELL_BITS;         // "unlimited" for the sizing pass
    ctx->skin_valid = false;                                  // the pass overwrites the grid parameters: the build that follows sorts
    int rc = osph_launch_grid_params(ctx, false, 3);
    ctx->cell_cap = saved;
    if (rc) return rc;
    GridParams g;
    OSPH_CUDA(cudaMemcpyAsync(&g, ctx->d_grid, sizeof(g), cudaMemcpyDeviceToHost, ctx->stream));
    OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    int64_t need = (int64_t)g.gnx * g.gny + 1;
    if (g.regime_a && need > ((int64_t)1 << OSPH_MAX_CELL_BITS)) {
        ctx->err = "reference grid needs more cells than can be tabulated"; return OSPH_E_GRID;
    }
    int64_t want = std::min<int64_t>((int64_t)1 << OSPH_MAX_CELL_BITS, need + need / 2 + 1024);
    if (want > ctx->cell_cap) {
        cudaFree(ctx->cell_range); ctx->cell_range = nullptr;
        OSPH_CUDA(cudaMalloc(&ctx->cell_range, sizeof(int2) * (size_t)want));
        ctx->cell_cap = want;
        int rc2 = osph_bin_alloc(ctx, want);
        if (rc2) return rc2;
    }
    int bits = 1;
    while (((int64_t)1 << bits) < ctx->cell_cap + 1) bits++;
    ctx->key_bits = bits;
    ctx->sized = true;
    return 0;
}

static int ensure_reductions(osph_ctx *ctx)
{
    if (ctx->reductions_valid) return 0;
    int rc = osph_launch_correct(ctx, false, 0.0, 0.0, false);
    if (rc) return rc;
    ctx->reductions_valid = true;
    return 0;
}

extern "C" int osph_timestep(osph_ctx *ctx, double out[3])
{
    CHECK_CTX(); NEED_PARTICLES();
    if (ctx->n_fluid == 0) { ctx->err = "osph_timestep: no fluid particles"; return OSPH_E_NO_FLUID; }
    int rc;
    {
        PhaseTimer t(ctx, T_TIMESTEP);
        if ((rc = ensure_reductions(ctx))) return rc;
        if ((rc = osph_launch_timestep(ctx, -1.0, false))) return rc;
    }
    StepScalars sc;
    OSPH_CUDA(cudaMemcpyAsync(&sc, ctx->d_sc, sizeof(sc), cudaMemcpyDeviceToHost, ctx->stream));
    OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    out[0] = sc.dt[0]; out[1] = sc.dt[1]; out[2] = sc.dt[2];
    return 0;
}

extern "C" int osph_predict(osph_ctx *ctx, double dt, double damping)
{
    CHECK_CTX(); NEED_PARTICLES();
    PhaseTimer t(ctx, T_PREDICT);
    int rc = osph_launch_prepare(ctx, true, dt, damping, false);
    if (rc) return rc;
    invalidate_state(ctx);
    ctx->prepared = true;
    return 0;
}

static int build_neighbours(osph_ctx *ctx)
{
    int rc;
    if (!ctx->prepared) {
        if ((rc = osph_launch_prepare(ctx, false, 0.0, 0.0, false))) return rc;
        ctx->prepared = true;
    }
    if ((rc = osph_size_cell_table(ctx))) return rc;
    if ((rc = osph_launch_build(ctx))) return rc;
    ctx->neighbours_valid = true;
    return 0;
}

extern "C" int osph_build_neighbours(osph_ctx *ctx)
{
    CHECK_CTX(); NEED_PARTICLES();
    PhaseTimer t(ctx, T_NEIGHBOURS);
    return build_neighbours(ctx);
}

extern "C" int osph_compute(osph_ctx *ctx)
{
    CHECK_CTX(); NEED_PARTICLES();
    int rc;
    if (!ctx->neighbours_valid) {
        PhaseTimer t(ctx, T_NEIGHBOURS);
        if ((rc = build_neighbours(ctx))) return rc;
    }
    PhaseTimer t(ctx, T_COMPUTE);
    if ((rc = osph_launch_pair(ctx))) return rc;
    ctx->c_uniform = true;
    ctx->reductions_valid = false;
    return 0;
}

extern "C" int osph_correct(osph_ctx *ctx, double dt, double damping)
{
    CHECK_CTX(); NEED_PARTICLES();
    PhaseTimer t(ctx, T_CORRECT);
    int rc = osph_launch_correct(ctx, true, dt, damping, false);
    if (rc) return rc;
    invalidate_state(ctx);
    ctx->reductions_valid = true;
    return 0;
}

extern "C" int osph_step(osph_ctx *ctx, int32_t nsteps, double fixed_dt, double damping)
{
    CHECK_CTX(); NEED_PARTICLES();
    if (ctx->n_fluid == 0 && !(fixed_dt > 0)) { ctx->err = "osph_step: no fluid particles"; return OSPH_E_NO_FLUID; }
    int rc;
    // PEC, several steps in one call: the corrector of step k and the predictor of step k+1 are one pass over the state
    // (k_prepare<.., FUSED>).  What TimeStep needs for step k+1 is reduced on the way: min h by the predictor of
    // step k (h is final after its refresh), max |a|^2 by the pair kernel of step k, c_max = co.  The last step of the
    // call ends with the plain corrector, so the state and the reductions are complete when the call returns.
    const bool fuse = nsteps > 1 && ctx->cfg.integrator == OSPH_INTEGRATOR_PEC && !ctx->slab && ctx->n_fluid > 0 &&
                      !ctx->cfg.summation_density;
    const bool ts_in_pair = fuse && ctx->timestep_fusion;
    for (int s = 0; s < nsteps; s++) {
        const bool fused_step = fuse && s > 0;
        if (!fused_step && (rc = ensure_reductions(ctx))) return rc;
        // the scalar resets ride along in k_timestep / k_grid_params
        // (from the second step of a fused call on, dt was computed by the last CTA of the previous step's pair kernel)
        const bool dt_done = fused_step && ts_in_pair;
        if (!dt_done && (rc = osph_launch_timestep(ctx, fixed_dt > 0 ? fixed_dt : -1.0, true, true, nullptr, fuse ? (fused_step ? 2 : 1) : 0)))
            return rc;
        // the predictor pass also forms the grid and the sort decision (last CTA) once the cell table has been sized
        BuildPlan plan = osph_plan_build(ctx);
        const bool grid_in_prepare = ctx->sized && ctx->grid_fusion;
        if ((rc = osph_launch_prepare(ctx, true, 0.0, damping, true, true, fuse ? (fused_step ? 2 : 1) : 0,
                                      grid_in_prepare ? &plan : nullptr, !fuse))) return rc;
        ctx->prepared = true;
        if ((rc = osph_size_cell_table(ctx))) return rc;
        plan.grid_done = grid_in_prepare;
        if ((rc = osph_launch_build(ctx, !fuse, &plan))) return rc;
        ctx->pair_reduce_a2 = fuse;
        ctx->pair_next_timestep = ts_in_pair && s < nsteps - 1;       // the next step of this call is a fused one: its dt here
        ctx->pair_ts_fixed_dt = fixed_dt > 0 ? fixed_dt : -1.0;
        rc = osph_launch_pair(ctx);
        ctx->pair_reduce_a2 = false; ctx->pair_next_timestep = false;
        if (rc) return rc;
        ctx->c_uniform = true;
        if (!fuse || s == nsteps - 1) {
            if ((rc = osph_launch_correct(ctx, true, 0.0, damping, true, true))) return rc;
        }
        invalidate_state(ctx);
        ctx->reductions_valid = true;
        ctx->step_counter++;
    }
    return 0;
}

extern "C" int osph_get_dt_log(osph_ctx *ctx, double *out, int64_t cap, int64_t *count)
{
    CHECK_CTX();
    StepScalars sc;
    OSPH_CUDA(cudaMemcpyAsync(&sc, ctx->d_sc, sizeof(sc), cudaMemcpyDeviceToHost, ctx->stream));
    OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    int64_t have = std::min<int64_t>(sc.dt_log_count, ctx->dt_log_cap);
    int64_t take = std::min<int64_t>(have, cap);
    if (take > 0 && out) OSPH_CUDA(cudaMemcpy(out, ctx->d_dt_log, sizeof(double) * 3 * take, cudaMemcpyDeviceToHost));
    if (count) *count = sc.dt_log_count;
    long long zero = 0;
    OSPH_CUDA(cudaMemcpy((char *)ctx->d_sc + offsetof(StepScalars, dt_log_count), &zero, sizeof(zero), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int osph_kinetic_energy(osph_ctx *ctx, double *ke)
{
    CHECK_CTX(); NEED_PARTICLES();
    int rc = osph_launch_ke(ctx);
    if (rc) return rc;
    StepScalars sc;
    OSPH_CUDA(cudaMemcpyAsync(&sc, ctx->d_sc, sizeof(sc), cudaMemcpyDeviceToHost, ctx->stream));
    OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    *ke = sc.ke;
    return 0;
}

extern "C" int osph_sort_stats(osph_ctx *ctx, int64_t out[2])
{
    CHECK_CTX();
    if (!out) return OSPH_E_INVALID;
    StepScalars sc;
    OSPH_CUDA(cudaMemcpyAsync(&sc, ctx->d_sc, sizeof(sc), cudaMemcpyDeviceToHost, ctx->stream));
    OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    out[0] = sc.builds; out[1] = sc.sorts;
    return 0;
}

extern "C" int osph_sync(osph_ctx *ctx, uint32_t *status)
{
    CHECK_CTX();
    StepScalars sc;
    OSPH_CUDA(cudaMemcpyAsync(&sc, ctx->d_sc, sizeof(sc), cudaMemcpyDeviceToHost, ctx->stream));
    OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    // bit 31: the reference grid outgrew the cell table and the steps since ran on the exact one-cell fallback
    // (k_grid_params); reported like a coarsened grid, and the next build re-sizes the table
    if (sc.status & 0x80000000u) { sc.status = (sc.status & 0x7fffffffu) | OSPH_S_GRID_COARSE; }
    if (status) *status = sc.status;
    unsigned int zero = 0;
    OSPH_CUDA(cudaMemcpy((char *)ctx->d_sc + offsetof(StepScalars, status), &zero, sizeof(zero), cudaMemcpyHostToDevice));
    if (sc.status & OSPH_S_GRID_COARSE) ctx->sized = false;    // re-size the cell table at the next build
    return 0;
}

// ---- validation / queries -----------------------------------------------------------------------

extern "C" int osph_get_cells(osph_ctx *ctx, double grid[7], int64_t *cell_ids)
{
    CHECK_CTX(); NEED_PARTICLES();
    int rc;
    if (!ctx->neighbours_valid && (rc = build_neighbours(ctx))) return rc;
    GridParams g;
    OSPH_CUDA(cudaMemcpyAsync(&g, ctx->d_grid, sizeof(g), cudaMemcpyDeviceToHost, ctx->stream));
    long long *d_out = nullptr;
    if (cell_ids) {
        OSPH_CUDA(cudaMalloc(&d_out, sizeof(long long) * ctx->n));
        if ((rc = osph_launch_cells(ctx, d_out))) { cudaFree(d_out); return rc; }
        cudaError_t e = cudaMemcpyAsync(cell_ids, d_out, sizeof(long long) * ctx->n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) { cudaFree(d_out); ctx->err = cudaGetErrorString(e); return OSPH_E_CUDA; }
    }
    {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        cudaFree(d_out);
        if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return OSPH_E_CUDA; }
    }
    if (grid) {
        grid[0] = g.xmin; grid[1] = g.xmax; grid[2] = g.ymin; grid[3] = g.ymax; grid[4] = g.cell_size;
        grid[5] = (double)g.ncx; grid[6] = (double)g.ncy;
    }
    return 0;
}

extern "C" int osph_get_neighbours_csr(osph_ctx *ctx, int64_t *offsets, int64_t *idx, int64_t cap, int64_t *total)
{
    CHECK_CTX(); NEED_PARTICLES();
    if (!offsets) { ctx->err = "osph_get_neighbours_csr: offsets is NULL"; return OSPH_E_INVALID; }
    int rc;
    if (!ctx->neighbours_valid && (rc = build_neighbours(ctx))) return rc;
    int64_t n = ctx->n;
    long long *d_counts = nullptr, *d_out = nullptr;
    OSPH_CUDA(cudaMalloc(&d_counts, sizeof(long long) * (n + 1)));
    if ((rc = osph_launch_neighbours(ctx, 0, d_counts, nullptr, nullptr))) { cudaFree(d_counts); return rc; }
    std::vector<long long> counts((size_t)n + 1);
    cudaError_t e = cudaMemcpyAsync(counts.data(), d_counts, sizeof(long long) * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(d_counts); ctx->err = cudaGetErrorString(e); return OSPH_E_CUDA; }
    long long run = 0;
    for (int64_t i = 0; i < n; i++) { long long c = counts[i]; offsets[i] = run; counts[i] = run; run += c; }
    offsets[n] = run; counts[n] = run;
    if (total) *total = run;
    if (idx) {
        if (cap < run) { cudaFree(d_counts); ctx->err = "osph_get_neighbours_csr: idx buffer too small"; return OSPH_E_CAPACITY; }
        e = cudaMalloc(&d_out, sizeof(long long) * std::max<long long>(run, 1));
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_counts, counts.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { cudaFree(d_counts); cudaFree(d_out); ctx->err = cudaGetErrorString(e); return OSPH_E_CUDA; }
        rc = osph_launch_neighbours(ctx, 1, nullptr, d_counts, d_out);
        if (!rc) {
            e = cudaMemcpyAsync(idx, d_out, sizeof(long long) * run, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = OSPH_E_CUDA; }
        }
    }
    cudaFree(d_counts); cudaFree(d_out);
    return rc;
}

extern "C" int osph_near_pos(osph_ctx *ctx, double x, double y, double h, int64_t cap, int64_t *idx, double *r,
                             double *q, double *hij, int64_t *count)
{
    CHECK_CTX(); NEED_PARTICLES();
    int rc;
    if (!ctx->neighbours_valid && (rc = build_neighbours(ctx))) return rc;
    int64_t c = std::max<int64_t>(cap, 1);
    long long *d_idx = nullptr, *d_count = nullptr; double *d_r = nullptr;
    if (cudaMalloc(&d_idx, sizeof(long long) * c) != cudaSuccess || cudaMalloc(&d_r, sizeof(double) * 3 * c) != cudaSuccess ||
        cudaMalloc(&d_count, sizeof(long long)) != cudaSuccess) {
        ctx->err = std::string("osph_near_pos: ") + cudaGetErrorString(cudaGetLastError());
        cudaFree(d_idx); cudaFree(d_r); cudaFree(d_count);
        return OSPH_E_CUDA;
    }
    rc = osph_launch_near_pos(ctx, x, y, h, cap, d_idx, d_r, d_r + c, d_r + 2 * c, d_count);
    long long cnt = 0;
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(&cnt, d_count, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        int64_t take = std::min<int64_t>(cnt, cap);
        if (e == cudaSuccess && take > 0) {
            if (idx) cudaMemcpy(idx, d_idx, sizeof(long long) * take, cudaMemcpyDeviceToHost);
            if (r) cudaMemcpy(r, d_r, sizeof(double) * take, cudaMemcpyDeviceToHost);
            if (q) cudaMemcpy(q, d_r + c, sizeof(double) * take, cudaMemcpyDeviceToHost);
            if (hij) cudaMemcpy(hij, d_r + 2 * c, sizeof(double) * take, cudaMemcpyDeviceToHost);
        }
        if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = OSPH_E_CUDA; }
    }
    if (count) *count = cnt;
    cudaFree(d_idx); cudaFree(d_r); cudaFree(d_count);
    return rc;
}

extern "C" int osph_probe_pressure(osph_ctx *ctx, int64_t n, const double *x, const double *y, double h, double *rho_out,
                                   double *p_out)
{
    CHECK_CTX(); NEED_PARTICLES();
    if (n <= 0) return 0;
    if (!x || !y || !rho_out || !p_out || !(h > 0) || ctx->n_ghost > 0) { ctx->err = "osph_probe_pressure: bad arguments"; return OSPH_E_INVALID; }
    int rc;
    if (!ctx->neighbours_valid && (rc = build_neighbours(ctx))) return rc;
    double *d = nullptr;
    OSPH_CUDA(cudaMalloc(&d, sizeof(double) * 4 * n));
    cudaError_t e = cudaMemcpyAsync(d, x, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, y, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream);
    rc = e == cudaSuccess ? osph_launch_probe(ctx, (int)n, d, d + n, h, d + 2 * n, d + 3 * n) : OSPH_E_CUDA;
    if (!rc) {
        e = cudaMemcpyAsync(rho_out, d + 2 * n, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(p_out, d + 3 * n, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = OSPH_E_CUDA; }
    cudaFree(d);
    return rc;
}

extern "C" int osph_get_timers(osph_ctx *ctx, double out_ms[6])
{
    CHECK_CTX();
    PhaseTimer::drain(ctx);
    for (int k = 0; k < 6; k++) out_ms[k] = ctx->timers_ms[k];
    return 0;
}

extern "C" int64_t osph_launch_count(const osph_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" uint64_t osph_stream(const osph_ctx *ctx) { return ctx ? (uint64_t)(uintptr_t)ctx->stream : 0; }

int osph_pair_build_flags();       // pair.cu

extern "C" int osph_pair_kernel_info(osph_ctx *ctx, int64_t out[3])
{
    CHECK_CTX();
    if (!out) return OSPH_E_INVALID;
    out[0] = ctx->pair_launches; out[1] = ctx->pair_uh_launches; out[2] = osph_pair_build_flags();
    return 0;
}

extern "C" int osph_pair_kernel_time(osph_ctx *ctx, double *avg_us, int64_t *launches)
{
    CHECK_CTX();
    OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    double sum = 0; int cnt = 0;
    for (int k = 0; k < ctx->pair_ev_used; k++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->pair_ev[2 * k], ctx->pair_ev[2 * k + 1]) == cudaSuccess) { sum += ms; cnt++; }
    }
    ctx->pair_ev_used = 0;
    if (avg_us) *avg_us = cnt ? sum * 1000.0 / cnt : 0.0;
    if (launches) *launches = cnt;
    return 0;
}
