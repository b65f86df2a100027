// export.cu -- partial transfers: asynchronous, double-buffered export of selected state columns, and
// transfers of a few rows (coupled structures).
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
#include <vector>

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
    cudaFree(ctx->d_rows_buf); ctx->d_rows_buf = nullptr; ctx->rows_bytes = 0;
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

// ---------------------------------------------------------------------------------------------------------
// Row transfers: move the records of a FEW host rows (the Coupled rows of a structure, reference
// src/Solver.py:381-398 / examples/IceBreak.py:109-171) instead of the whole array.
// ---------------------------------------------------------------------------------------------------------
struct RowColumns { double *f[OSPH_NUM_FIELDS]; };

__global__ void k_inverse_rows(const int *__restrict__ row, int n, int *__restrict__ inv)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[row[i]] = i;
}

// pack != 0: staged[k] <- record of host row rows[k];  pack == 0: state of that row <- staged[k] (label / deleted ignored)
__global__ void k_rows_records(const long long *__restrict__ rows, int nrows, const int *__restrict__ inv, int n_total,
                               RowColumns c, const signed char *__restrict__ label, unsigned char *__restrict__ staged,
                               long long stride, int pack, int c_uniform, double co, int *__restrict__ bad)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nrows) return;
    long long r = rows[k];
    int slot = (r >= 0 && r < n_total) ? inv[r] : -1;
    if (slot < 0) { atomicAdd(bad, 1); return; }
    unsigned char *rec = staged + (long long)k * stride;
    if (pack) {
        rec[0] = 0; rec[1] = (unsigned char)label[slot];
#pragma unroll
        for (int f = 0; f < OSPH_NUM_FIELDS; f++)
            store_f64_unaligned(rec + 2 + 8 * f, (f == OSPH_F_C && c_uniform) ? co : c.f[f][slot]);
    } else {
#pragma unroll
        for (int f = 0; f < OSPH_NUM_FIELDS; f++) c.f[f][slot] = load_f64_unaligned(rec + 2 + 8 * f);
    }
}

static int rows_prepare(osph_ctx *ctx, int64_t nrows, const int64_t *rows, int64_t stride, unsigned char **d_staged,
                        long long **d_rows, int **d_inv, int **d_bad)
{
    if (ctx->n <= 0) { ctx->err = "no particles uploaded"; return OSPH_E_INVALID; }
    if (ctx->slab) { ctx->err = "row transfers are not available in slab mode (rows are global ids)"; return OSPH_E_INVALID; }
    if (nrows < 0 || (nrows > 0 && !rows) || stride < 154 || (stride & 1)) { ctx->err = "row transfer: bad argument"; return OSPH_E_INVALID; }
    // one device allocation: [staged records | row ids | inverse map | error counter]
    const size_t b_staged = ((size_t)nrows * stride + 15) / 16 * 16, b_rows = sizeof(long long) * (size_t)nrows;
    const size_t b_inv = sizeof(int) * (size_t)ctx->n_total, need = b_staged + b_rows + b_inv + 16;
    if (need > ctx->rows_bytes) {
        OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_rows_buf); ctx->d_rows_buf = nullptr; ctx->rows_bytes = 0;
        OSPH_CUDA(cudaMalloc(&ctx->d_rows_buf, need + need / 2));
        ctx->rows_bytes = need + need / 2;
    }
    *d_staged = ctx->d_rows_buf;
    *d_rows = reinterpret_cast<long long *>(ctx->d_rows_buf + b_staged);
    *d_inv = reinterpret_cast<int *>(ctx->d_rows_buf + b_staged + b_rows);
    *d_bad = reinterpret_cast<int *>(ctx->d_rows_buf + b_staged + b_rows + (b_inv + 7) / 8 * 8);
    OSPH_CUDA(cudaMemcpyAsync(*d_rows, rows, b_rows, cudaMemcpyHostToDevice, ctx->stream));
    OSPH_CUDA(cudaMemsetAsync(*d_inv, 0xff, b_inv, ctx->stream));                       // -1: row not active
    OSPH_CUDA(cudaMemsetAsync(*d_bad, 0, sizeof(int), ctx->stream));
    k_inverse_rows<<<div_up(ctx->n, 256), 256, 0, ctx->stream>>>(ctx->d_row, (int)ctx->n, *d_inv);
    OSPH_LAUNCH_CHECK();
    return 0;
}

static RowColumns row_columns(osph_ctx *ctx)
{
    RowColumns c;
    for (int f = 0; f < OSPH_NUM_FIELDS; f++) c.f[f] = ctx->f[f];
    return c;
}

extern "C" int osph_download_rows(osph_ctx *ctx, int64_t nrows, const int64_t *rows, void *pA, int64_t stride)
{
    if (!ctx) return OSPH_E_INVALID;
    OSPH_CUDA(cudaSetDevice(ctx->device));
    if (nrows == 0) return 0;
    if (!pA) { ctx->err = "osph_download_rows: null array"; return OSPH_E_INVALID; }
    unsigned char *d_staged; long long *d_rows; int *d_inv, *d_bad;
    int rc = rows_prepare(ctx, nrows, rows, stride, &d_staged, &d_rows, &d_inv, &d_bad);
    if (rc) return rc;
    k_rows_records<<<div_up(nrows, 128), 128, 0, ctx->stream>>>(d_rows, (int)nrows, d_inv, (int)ctx->n_total, row_columns(ctx),
                                                                ctx->label, d_staged, stride, 1, ctx->c_uniform ? 1 : 0,
                                                                ctx->cfg.co, d_bad);
    OSPH_LAUNCH_CHECK();
    std::vector<unsigned char> host((size_t)nrows * stride);
    int bad = 0;
    OSPH_CUDA(cudaMemcpyAsync(host.data(), d_staged, host.size(), cudaMemcpyDeviceToHost, ctx->stream));
    OSPH_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    OSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    if (bad) { ctx->err = "osph_download_rows: a requested row is out of range or not active (deleted)"; return OSPH_E_INVALID; }
    for (int64_t k = 0; k < nrows; k++)
        memcpy((unsigned char *)pA + rows[k] * stride, host.data() + k * stride, 154);
    return 0;
}

extern "C" int osph_upload_rows(osph_ctx *ctx, int64_t nrows, const int64_t *rows, const void *pA, int64_t stride)
{
    if (!ctx) return OSPH_E_INVALID;
    OSPH_CUDA(cudaSetDevice(ctx->device));
    if (nrows == 0) return 0;
    if (!pA) { ctx->err = "osph_upload_rows: null array"; return OSPH_E_INVALID; }
    unsigned char *d_staged; long long *d_rows; int *d_inv, *d_bad;
    int rc = rows_prepare(ctx, nrows, rows, stride, &d_staged, &d_rows, &d_inv, &d_bad);
    if (rc) return rc;
    std::vector<unsigned char> host((size_t)nrows * stride);
    for (int64_t k = 0; k < nrows; k++) {
        if (rows[k] < 0 || rows[k] >= ctx->n_total) { ctx->err = "osph_upload_rows: row out of range"; return OSPH_E_INVALID; }
        memcpy(host.data() + k * stride, (const unsigned char *)pA + rows[k] * stride, 154);
    }
    OSPH_CUDA(cudaMemcpyAsync(d_staged, host.data(), host.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->c_uniform) {            // c becomes per-row data again once a caller writes it
        if ((rc = osph_launch_fill(ctx, ctx->f[OSPH_F_C], ctx->cfg.co))) return rc;
        ctx->c_uniform = false;
    }
    k_rows_records<<<div_up(nrows, 128), 128, 0, ctx->stream>>>(d_rows, (int)nrows, d_inv, (int)ctx->n_total, row_columns(ctx),
                                                                ctx->label, d_staged, stride, 0, 0, ctx->cfg.co, d_bad);
    OSPH_LAUNCH_CHECK();
    int bad = 0;
    OSPH_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    OSPH_CUDA(cudaStreamSynchronize(ctx->stream));                 // `host` goes out of scope
    ctx->prepared = false; ctx->neighbours_valid = false; ctx->reductions_valid = false;
    ctx->skin_valid = false;            // rows were replaced from outside: the next build sorts
    if (bad) { ctx->err = "osph_upload_rows: a row is not active (deleted); the other rows were written"; return OSPH_E_INVALID; }
    return 0;
}
