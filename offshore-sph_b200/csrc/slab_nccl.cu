// slab_nccl.cu -- the slab-decomposed step sequenced in C++ with direct NCCL calls (NVLink 5 / NVSwitch).
//
// Same protocol as osph_b200/slabs.py (which stays as the readable, gloo-testable sequencer), but without a
// Python turnaround inside the step: after the single host sync of a step (message sizes), the sends/receives,
// the owned-set update and the force evaluation are enqueued within a few microseconds.
//   ncclAllReduce(MIN) 3 doubles -> dt;  ncclAllGather 12 doubles -> counts + grid bounds;
//   ncclGroup{Send,Recv} of migrants (21 doubles each) and halo particles (8 doubles each) with both x-neighbours,
//   halos received straight into the ghost buffer that k_keys / k_gather read.
// libnccl is resolved at run time (dlopen of the copy already loaded by torch, else by path), so the library has
// no link-time NCCL dependency and still loads on a box without it.
#include <dlfcn.h>
#include <stdlib.h>
#include <nccl.h>

#include <algorithm>
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
int osph_slab_commit(osph_ctx *ctx, int64_t n_mig_out, const void *d_mig_in, int64_t n_mig_in, int64_t n_ghost,
                     const double global_bounds[6]);
int osph_slab_step_end(osph_ctx *ctx, double damping);
}

namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok() const { return handle && GetUniqueId && CommInitRank && AllReduce && AllGather && Send && Recv && GroupStart && GroupEnd; }
};
NcclApi g_nccl;
std::string g_nccl_error;

bool load_nccl(const char *path)
{
    if (g_nccl.ok()) return true;
    // OSPH_NCCL_LIB=<path>: use exactly this NCCL build (symbols are taken from its handle, whatever else is loaded)
    const char *forced = getenv("OSPH_NCCL_LIB");
    void *h = forced && *forced ? dlopen(forced, RTLD_NOW | RTLD_LOCAL) : nullptr;
    if (forced && *forced && !h) { g_nccl_error = std::string("cannot load OSPH_NCCL_LIB: ") + dlerror(); return false; }
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);         // the copy torch already loaded
    if (!h && path && *path) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { g_nccl_error = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
    g_nccl.handle = h;
#define SYM(field, name) g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name))
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce"); SYM(AllGather, "ncclAllGather"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    if (!g_nccl.ok()) { g_nccl_error = "libnccl.so.2 lacks a required symbol"; return false; }
    return true;
}
}  // namespace

struct osph_slab_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    double x_lo = 0, x_hi = 0, r0 = 0, hmax = 0;
    int kernel = 0;
    int64_t mig_cap = 0, halo_cap = 0, ghost_cap = 0;
    double *d_ghost = nullptr, *d_mig_l = nullptr, *d_mig_r = nullptr, *d_mig_in = nullptr, *d_halo_l = nullptr,
           *d_halo_r = nullptr, *d_meta = nullptr, *d_all_meta = nullptr, *d_dt3 = nullptr;
    double *h_all_meta = nullptr;       // pinned
    int64_t last_counts[8] = {0};       // mig_out l/r, halo_out l/r, mig_in l/r, halo_in l/r
    int64_t steps = 0;
};

#define NCCL_CK(call)                                                                             \
    do {                                                                                          \
        ncclResult_t r__ = (call);                                                                \
        if (r__ != ncclSuccess) {                                                                 \
            ctx->err = std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error"); \
            return OSPH_E_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

extern "C" int osph_nccl_unique_id(const char *libnccl_path, char out[128])
{
    if (!load_nccl(libnccl_path)) return OSPH_E_INVALID;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_nccl_error = "ncclGetUniqueId failed"; return OSPH_E_CUDA; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out, &id, 128);
    return 0;
}

extern "C" const char *osph_nccl_last_error(void) { return g_nccl_error.c_str(); }

// Create the per-rank communicator and the exchange buffers; the context must already hold this rank's particles.
extern "C" int osph_slab_comm_create(osph_ctx *ctx, const char *libnccl_path, const char unique_id[128], int rank,
                                     int world, double x_lo, double x_hi, double r0, double hmax, int64_t mig_cap,
                                     int64_t halo_cap, osph_slab_comm **out)
{
    if (!ctx || !out || world < 1 || rank < 0 || rank >= world) return OSPH_E_INVALID;
    OSPH_CUDA(cudaSetDevice(ctx->device));
    if (!load_nccl(libnccl_path)) { ctx->err = g_nccl_error; return OSPH_E_INVALID; }
    osph_slab_comm *s = new osph_slab_comm();
    s->rank = rank; s->world = world; s->x_lo = x_lo; s->x_hi = x_hi; s->r0 = r0; s->hmax = hmax;
    s->kernel = ctx->cfg.kernel; s->mig_cap = mig_cap; s->halo_cap = halo_cap; s->ghost_cap = 2 * halo_cap + 2 * mig_cap;
    ncclUniqueId id; memcpy(&id, unique_id, 128);
    NCCL_CK(g_nccl.CommInitRank(&s->comm, world, id, rank));
    OSPH_CUDA(cudaMalloc(&s->d_ghost, sizeof(double) * OSPH_WIRE_HALO * s->ghost_cap));
    OSPH_CUDA(cudaMalloc(&s->d_mig_l, sizeof(double) * OSPH_WIRE_FULL * mig_cap));
    OSPH_CUDA(cudaMalloc(&s->d_mig_r, sizeof(double) * OSPH_WIRE_FULL * mig_cap));
    OSPH_CUDA(cudaMalloc(&s->d_mig_in, sizeof(double) * OSPH_WIRE_FULL * 2 * mig_cap));
    OSPH_CUDA(cudaMalloc(&s->d_halo_l, sizeof(double) * OSPH_WIRE_HALO * halo_cap));
    OSPH_CUDA(cudaMalloc(&s->d_halo_r, sizeof(double) * OSPH_WIRE_HALO * halo_cap));
    OSPH_CUDA(cudaMalloc(&s->d_meta, sizeof(double) * 12));
    OSPH_CUDA(cudaMalloc(&s->d_all_meta, sizeof(double) * 12 * world));
    OSPH_CUDA(cudaMalloc(&s->d_dt3, sizeof(double) * 3));
    OSPH_CUDA(cudaMallocHost(&s->h_all_meta, sizeof(double) * 12 * world));
    int rc = osph_slab_configure(ctx, x_lo, x_hi, s->d_ghost, s->ghost_cap);
    if (rc) return rc;
    *out = s;
    return 0;
}

extern "C" int osph_slab_comm_destroy(osph_ctx *ctx, osph_slab_comm *s)
{
    if (!s) return OSPH_E_INVALID;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); ctx->d_ghost = nullptr; ctx->n_ghost = 0; ctx->slab = false; }
    if (s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
    cudaFree(s->d_ghost); cudaFree(s->d_mig_l); cudaFree(s->d_mig_r); cudaFree(s->d_mig_in); cudaFree(s->d_halo_l);
    cudaFree(s->d_halo_r); cudaFree(s->d_meta); cudaFree(s->d_all_meta); cudaFree(s->d_dt3); cudaFreeHost(s->h_all_meta);
    delete s;
    return 0;
}

// Re-attach after an upload replaced the context's particle set (host-buffer call pattern).
extern "C" int osph_slab_comm_attach(osph_ctx *ctx, osph_slab_comm *s)
{
    if (!ctx || !s) return OSPH_E_INVALID;
    return osph_slab_configure(ctx, s->x_lo, s->x_hi, s->d_ghost, s->ghost_cap);
}

// New slab boundaries (load re-balancing); every rank must switch at the same step.
extern "C" int osph_slab_comm_set_bounds(osph_ctx *ctx, osph_slab_comm *s, double x_lo, double x_hi)
{
    if (!ctx || !s || !(x_lo < x_hi)) return OSPH_E_INVALID;
    s->x_lo = x_lo; s->x_hi = x_hi;
    return osph_slab_configure(ctx, x_lo, x_hi, s->d_ghost, s->ghost_cap);
}

extern "C" int osph_slab_run(osph_ctx *ctx, osph_slab_comm *s, int32_t nsteps, double fixed_dt, double damping)
{
    if (!ctx || !s) return OSPH_E_INVALID;
    OSPH_CUDA(cudaSetDevice(ctx->device));
    const int left = s->rank > 0 ? s->rank - 1 : -1, right = s->rank < s->world - 1 ? s->rank + 1 : -1;
    const double q = s->kernel == OSPH_KERNEL_GAUSSIAN ? 3.0 : 2.0;
    int rc;
    for (int step = 0; step < nsteps; step++) {
        if ((rc = osph_slab_step_plan(ctx, step, nsteps))) return rc;    // corrector of step k fused into the predictor of k+1
        // ---- identical dt on every rank ----
        if ((rc = osph_slab_dt_local(ctx, s->d_dt3))) return rc;
        NCCL_CK(g_nccl.AllReduce(s->d_dt3, s->d_dt3, 3, ncclDouble, ncclMin, s->comm, ctx->stream));
        if ((rc = osph_slab_step_begin(ctx, s->d_dt3, fixed_dt, damping))) return rc;
        // ---- classify + pack; counts and local grid bounds to everyone ----
        const double width = std::max(q * s->hmax, std::min(s->r0, 3.0 * s->hmax)) * 1.1;
        if ((rc = osph_slab_pack(ctx, width, s->d_mig_l, s->d_mig_r, s->mig_cap, s->d_halo_l, s->d_halo_r, s->halo_cap,
                                 s->d_meta))) return rc;
        NCCL_CK(g_nccl.AllGather(s->d_meta, s->d_all_meta, 12, ncclDouble, s->comm, ctx->stream));
        OSPH_CUDA(cudaMemcpyAsync(s->h_all_meta, s->d_all_meta, sizeof(double) * 12 * s->world, cudaMemcpyDeviceToHost, ctx->stream));
        OSPH_CUDA(cudaStreamSynchronize(ctx->stream));                  // the one host sync of the step
        const double *M = s->h_all_meta;
        double bounds[6];
        for (int k = 0; k < 6; k++) { bounds[k] = M[4 + k]; for (int r = 1; r < s->world; r++) bounds[k] = std::min(bounds[k], M[12 * r + 4 + k]); }
        for (int r = 0; r < s->world; r++)
            if (M[12 * r + 10] != 0.0) { ctx->err = "slab exchange buffers overflowed: raise the migrant / halo capacities"; return OSPH_E_CAPACITY; }
        const double *me = M + 12 * s->rank;
        const int64_t out_l = (int64_t)me[0], out_r = (int64_t)me[1], halo_l = (int64_t)me[2], halo_r = (int64_t)me[3];
        const int64_t in_mig_l = left >= 0 ? (int64_t)M[12 * left + 1] : 0, in_halo_l = left >= 0 ? (int64_t)M[12 * left + 3] : 0;
        const int64_t in_mig_r = right >= 0 ? (int64_t)M[12 * right + 0] : 0, in_halo_r = right >= 0 ? (int64_t)M[12 * right + 2] : 0;
        s->hmax = -bounds[5];
        const int64_t n_own_ghost = out_l + out_r, n_ghost = n_own_ghost + in_halo_l + in_halo_r;
        if (n_ghost > s->ghost_cap || in_mig_l + in_mig_r > 2 * s->mig_cap) { ctx->err = "slab receive buffers too small"; return OSPH_E_CAPACITY; }
        // ---- payloads: halos land directly behind this rank's own migrants in the ghost buffer ----
        double *g0 = s->d_ghost + n_own_ghost * OSPH_WIRE_HALO, *g1 = g0 + in_halo_l * OSPH_WIRE_HALO;
        double *m1 = s->d_mig_in + in_mig_l * OSPH_WIRE_FULL;
        NCCL_CK(g_nccl.GroupStart());
        if (left >= 0) {
            if (in_mig_l) NCCL_CK(g_nccl.Recv(s->d_mig_in, in_mig_l * OSPH_WIRE_FULL, ncclDouble, left, s->comm, ctx->stream));
            if (in_halo_l) NCCL_CK(g_nccl.Recv(g0, in_halo_l * OSPH_WIRE_HALO, ncclDouble, left, s->comm, ctx->stream));
            if (out_l) NCCL_CK(g_nccl.Send(s->d_mig_l, out_l * OSPH_WIRE_FULL, ncclDouble, left, s->comm, ctx->stream));
            if (halo_l) NCCL_CK(g_nccl.Send(s->d_halo_l, halo_l * OSPH_WIRE_HALO, ncclDouble, left, s->comm, ctx->stream));
        }
        if (right >= 0) {
            if (in_mig_r) NCCL_CK(g_nccl.Recv(m1, in_mig_r * OSPH_WIRE_FULL, ncclDouble, right, s->comm, ctx->stream));
            if (in_halo_r) NCCL_CK(g_nccl.Recv(g1, in_halo_r * OSPH_WIRE_HALO, ncclDouble, right, s->comm, ctx->stream));
            if (out_r) NCCL_CK(g_nccl.Send(s->d_mig_r, out_r * OSPH_WIRE_FULL, ncclDouble, right, s->comm, ctx->stream));
            if (halo_r) NCCL_CK(g_nccl.Send(s->d_halo_r, halo_r * OSPH_WIRE_HALO, ncclDouble, right, s->comm, ctx->stream));
        }
        NCCL_CK(g_nccl.GroupEnd());
        // ---- owned set update, same grid everywhere, force evaluation, corrector ----
        if ((rc = osph_slab_commit(ctx, out_l + out_r, s->d_mig_in, in_mig_l + in_mig_r, n_ghost, bounds))) return rc;
        if ((rc = osph_slab_step_end(ctx, damping))) return rc;
        const int64_t c[8] = {out_l, out_r, halo_l, halo_r, in_mig_l, in_mig_r, in_halo_l, in_halo_r};
        memcpy(s->last_counts, c, sizeof(c));
        s->steps++;
    }
    return 0;
}

extern "C" int osph_slab_last_counts(const osph_slab_comm *s, int64_t out[8])
{
    if (!s) return OSPH_E_INVALID;
    memcpy(out, s->last_counts, sizeof(s->last_counts));
    return 0;
}
