// export.cu -- asynchronous, double-buffered export of selected state columns.
//
// Replaces the per-step blocking copy of Solver._store (reference src/Solver.py:477-486: one full
// `np.copy(self.particleArray[key])` per exported property and step).  osph_export_begin gathers the requested
// columns of the active particles into a device slot ON THE COMPUTE STREAM (a few microseconds: the snapshot is
// taken in stream order, later steps may overwrite the state), then a dedicated copy stream moves the slot into
// library-owned pinned host memory while the compute stream runs the next step(s).  osph_export_end waits for
// that one copy only and hands the columns to the caller.  Two slots: the export of step k crosses PCIe while
// step k+1 is computed, and the host consumes step k-1.
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "step.cuh"

#define XSLOTS 2

// Row-space export: column `field` for EVERY row of the uploaded array, in host row order.  Inactive (deleted) rows
// keep the value of the device-side record mirror (what the host uploaded); active rows take the current state.
__global__ void k_rows_from_mirror(const unsigned char *__restrict__ aos, long long stride, int n_total, int field,
                                   double *__restrict__ out)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_total) out[r] = load_f64_unaligned(aos + (long long)r * stride + 2 + 8 * field);
}

__global__ void k_col_to_rows(const double *__restrict__ col, const int *__restrict__ row, int n, double *__restrict__ out,
                              int fill, double value)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[row[i]] = fill ? value : col[i];
}

struct osph_export_ring {
    cudaStream_t copy_stream = nullptr;
    struct Slot {
        double *d_buf = nullptr, *h_buf = nullptr;
        size_t doubles = 0;                  // allocated size of both buffers
        cudaEvent_t packed = nullptr, done = nullptr;
        int64_t ticket = -1;                 // ticket in flight / ready to be fetched, -1: free
        int64_t n = 0;
        int32_t nfields = 0;
    } slot[XSLOTS];
    int64_t next_ticket = 0;
};

static int ring_get(osph_ctx *ctx, osph_export_ring **out)
{
    if (!ctx->xring) {
        osph_export_ring *r = new osph_export_ring();
        ctx->xring = r;
        OSPH_CUDA(cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < XSLOTS; k++) {
            OSPH_CUDA(cudaEventCreateWithFlags(&r->slot[k].packed, cudaEventDisableTiming));
            OSPH_CUDA(cudaEventCreateWithFlags(&r->slot[k].done, cudaEventDisableTiming));
        }
    }
    *out = ctx->xring;
    return 0;
}

void osph_export_free(osph_ctx *ctx)
{
    osph_export_ring *r = ctx->xring;
    if (!r) return;
    if (r->copy_stream) cudaStreamSynchronize(r->copy_stream);
    for (int k = 0; k < XSLOTS; k++) {
        cudaFree(r->slot[k].d_buf);
        cudaFreeHost(r->slot[k].h_buf);
        if (r->slot[k].packed) cudaEventDestroy(r->slot[k].packed);
        if (r->slot[k].done) cudaEventDestroy(r->slot[k].done);
    }
    if (r->copy_stream) cudaStreamDestroy(r->copy_stream);
    delete r;
    ctx->xring = nullptr;
}

extern "C" int osph_export_begin(osph_ctx *ctx, int32_t nfields, const int32_t *fields, int32_t row_space, int64_t *ticket)
{
    if (!ctx) return OSPH_E_INVALID;
    OSPH_CUDA(cudaSetDevice(ctx->device));
    if (ctx->n <= 0) { ctx->err = "no particles uploaded"; return OSPH_E_INVALID; }
    if (nfields <= 0 || !fields || !ticket) { ctx->err = "osph_export_begin: bad argument"; return OSPH_E_INVALID; }
    for (int k = 0; k < nfields; k++)
        if (fields[k] < 0 || fields[k] >= OSPH_NUM_FIELDS) { ctx->err = "osph_export_begin: bad field"; return OSPH_E_INVALID; }
    if (row_space && ctx->slab) { ctx->err = "osph_export_begin: row-space export is not available in slab mode (rows are global ids)"; return OSPH_E_INVALID; }
    const int64_t len = row_space ? ctx->n_total : ctx->n;       // entries per exported column
    osph_export_ring *r;
    int rc = ring_get(ctx, &r);
    if (rc) return rc;
    osph_export_ring::Slot &s = r->slot[r->next_ticket % XSLOTS];
    if (s.ticket >= 0) {
        ctx->err = "osph_export_begin: both export slots are in flight; call osph_export_end on the oldest ticket first";
        return OSPH_E_CAPACITY;
    }
    const size_t need = (size_t)nfields * (size_t)len;
    if (need > s.doubles) {
        // the slot is free, so neither stream still touches its buffers
        cudaFree(s.d_buf); cudaFreeHost(s.h_buf);
        s.d_buf = s.h_buf = nullptr; s.doubles = 0;
        const size_t want = (size_t)nfields * (size_t)std::max(len, ctx->cap);
        OSPH_CUDA(cudaMalloc(&s.d_buf, sizeof(double) * want));
        OSPH_CUDA(cudaMallocHost(&s.h_buf, sizeof(double) * want));
        s.doubles = want;
    }
    for (int k = 0; k < nfields; k++) {
        double *dst = s.d_buf + (size_t)k * len;
        if (!row_space) {
            if ((rc = osph_launch_col_to_active(ctx, fields[k], dst))) return rc;
            continue;
        }
        if (ctx->n < ctx->n_total) {
            k_rows_from_mirror<<<div_up(ctx->n_total, 256), 256, 0, ctx->stream>>>(ctx->d_aos, ctx->stride, (int)ctx->n_total, fields[k], dst);
            OSPH_LAUNCH_CHECK();
        }
        k_col_to_rows<<<div_up(ctx->n, 256), 256, 0, ctx->stream>>>(ctx->f[fields[k]], ctx->d_row, (int)ctx->n, dst,
                                                                    (fields[k] == OSPH_F_C && ctx->c_uniform) ? 1 : 0, ctx->cfg.co);
        OSPH_LAUNCH_CHECK();
    }
    OSPH_CUDA(cudaEventRecord(s.packed, ctx->stream));
    OSPH_CUDA(cudaStreamWaitEvent(r->copy_stream, s.packed, 0));
    OSPH_CUDA(cudaMemcpyAsync(s.h_buf, s.d_buf, sizeof(double) * need, cudaMemcpyDeviceToHost, r->copy_stream));
    OSPH_CUDA(cudaEventRecord(s.done, r->copy_stream));
    s.ticket = r->next_ticket++;
    s.n = len; s.nfields = nfields;
    *ticket = s.ticket;
    return 0;
}

extern "C" int osph_export_end(osph_ctx *ctx, int64_t ticket, int32_t nfields, double *const *cols, int64_t n)
{
    if (!ctx) return OSPH_E_INVALID;
    OSPH_CUDA(cudaSetDevice(ctx->device));
    osph_export_ring *r = ctx->xring;
    if (!r || ticket < 0) { ctx->err = "osph_export_end: unknown ticket"; return OSPH_E_INVALID; }
    osph_export_ring::Slot &s = r->slot[ticket % XSLOTS];
    if (s.ticket != ticket) { ctx->err = "osph_export_end: unknown ticket"; return OSPH_E_INVALID; }
    if (nfields != s.nfields || n != s.n || !cols) {
        ctx->err = "osph_export_end: column count / length differ from the matching osph_export_begin";
        return OSPH_E_INVALID;
    }
    OSPH_CUDA(cudaEventSynchronize(s.done));
    for (int k = 0; k < nfields; k++) {
        if (!cols[k]) { ctx->err = "osph_export_end: null column"; return OSPH_E_INVALID; }
        memcpy(cols[k], s.h_buf + (size_t)k * s.n, sizeof(double) * (size_t)s.n);
    }
    s.ticket = -1;
    return 0;
}
