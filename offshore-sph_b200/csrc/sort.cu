// sort.cu -- stable LSD radix sort of (cell key, storage slot) pairs, 8 bits per pass.
//
// Replaces the serial head-insertion binning of the reference (src/Tools/NNLinkedList.py:129-141):
// after the sort, the particles of one cell are contiguous and the cell table is two integers per cell.
// Three kernels per pass, all HBM-bound: per-tile digit histogram, per-digit scan along the tiles,
// stable scatter (warp-level match_any ranking, so equal keys keep their input order).
// Algorithmic traffic per pass and pair: read key twice + slot once, write both = 20 B.
#include "common.cuh"

#define SORT_THREADS 256
#define SORT_ITEMS 8
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)
#define RADIX 256

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist(const unsigned int *__restrict__ key, int n, int shift, int nblocks, unsigned int *__restrict__ hist)
{
    __shared__ unsigned int cnt[RADIX];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    int base = blockIdx.x * SORT_TILE;
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        int i = base + r * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&cnt[(key[i] >> shift) & 0xff], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = cnt[threadIdx.x];
}

// One block per digit: exclusive scan of hist[d][0..nblocks) in place, total to digit_tot[d].
__global__ void __launch_bounds__(256)
k_sort_rowscan(unsigned int *__restrict__ hist, int nblocks, unsigned int *__restrict__ digit_tot)
{
    __shared__ unsigned int wsum[8];
    __shared__ unsigned int carry_s;
    unsigned int *row = hist + (size_t)blockIdx.x * nblocks;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 256) {
        int i = base + threadIdx.x;
        unsigned int v = i < nblocks ? row[i] : 0u;
        unsigned int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) wsum[w] = s;
        __syncthreads();
        unsigned int woff = 0;
        for (int k = 0; k < w; k++) woff += wsum[k];
        unsigned int carry = carry_s;
        if (i < nblocks) row[i] = carry + woff + s - v;
        __syncthreads();
        if (threadIdx.x == 255) carry_s = carry + woff + s;
        __syncthreads();
    }
    if (threadIdx.x == 0) digit_tot[blockIdx.x] = carry_s;
}

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_scatter(const unsigned int *__restrict__ key_in, const unsigned int *__restrict__ val_in,
               unsigned int *__restrict__ key_out, unsigned int *__restrict__ val_out, int n, int shift,
               int nblocks, const unsigned int *__restrict__ hist, const unsigned int *__restrict__ digit_tot)
{
    __shared__ unsigned int cnt[SORT_THREADS / 32][RADIX];
    __shared__ unsigned int dbase[RADIX];
    __shared__ unsigned int wsum[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

    // exclusive scan of the 256 digit totals (every block repeats it; 1 KB from L2)
    {
        unsigned int v = digit_tot[threadIdx.x], s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) wsum[w] = s;
        for (int k = 0; k < SORT_THREADS / 32; k++) cnt[k][threadIdx.x] = 0;
        __syncthreads();
        unsigned int woff = 0;
        for (int k = 0; k < w; k++) woff += wsum[k];
        dbase[threadIdx.x] = woff + s - v + hist[threadIdx.x * nblocks + blockIdx.x];
    }
    __syncthreads();

    // warp w ranks keys [base + w*256, base + (w+1)*256) in input order
    const int base = blockIdx.x * SORT_TILE + w * (32 * SORT_ITEMS);
    unsigned int k[SORT_ITEMS], v[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        int i = base + r * 32 + lane;
        bool ok = i < n;
        k[r] = ok ? key_in[i] : 0u;
        v[r] = ok ? val_in[i] : 0u;
        unsigned int d = ok ? ((k[r] >> shift) & 0xff) : 0x100u;
        unsigned int peers = __match_any_sync(0xffffffffu, d);
        unsigned int before = __popc(peers & ((1u << lane) - 1u));
        int leader = __ffs(peers) - 1;
        unsigned int b = 0;
        if (ok && lane == leader) { b = cnt[w][d]; cnt[w][d] = b + __popc(peers); }
        b = __shfl_sync(0xffffffffu, b, leader);
        rank[r] = b + before;
        __syncwarp();
    }
    __syncthreads();
    {   // digit = threadIdx.x: turn the per-warp counts into per-warp offsets
        unsigned int run = dbase[threadIdx.x];
#pragma unroll
        for (int kk = 0; kk < SORT_THREADS / 32; kk++) {
            unsigned int t = cnt[kk][threadIdx.x];
            cnt[kk][threadIdx.x] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        int i = base + r * 32 + lane;
        if (i < n) {
            unsigned int pos = cnt[w][(k[r] >> shift) & 0xff] + rank[r];
            key_out[pos] = k[r];
            val_out[pos] = v[r];
        }
    }
}

int osph_sort_alloc(osph_ctx *ctx, int64_t cap)
{
    ctx->sort_blocks = div_up(cap, SORT_TILE);
    OSPH_CUDA(cudaMalloc(&ctx->hist, sizeof(unsigned int) * RADIX * (size_t)ctx->sort_blocks));
    OSPH_CUDA(cudaMalloc(&ctx->digit_tot, sizeof(unsigned int) * RADIX));
    for (int b = 0; b < 2; b++) {
        OSPH_CUDA(cudaMalloc(&ctx->key[b], sizeof(unsigned int) * (size_t)cap));
        OSPH_CUDA(cudaMalloc(&ctx->idx[b], sizeof(unsigned int) * (size_t)cap));
    }
    return 0;
}

void osph_sort_free(osph_ctx *ctx)
{
    cudaFree(ctx->hist); cudaFree(ctx->digit_tot);
    for (int b = 0; b < 2; b++) { cudaFree(ctx->key[b]); cudaFree(ctx->idx[b]); }
    ctx->hist = ctx->digit_tot = nullptr;
    ctx->key[0] = ctx->key[1] = ctx->idx[0] = ctx->idx[1] = nullptr;
}

// Sorts key[sorted_buf] / idx[sorted_buf] (n pairs, `bits` significant key bits); updates sorted_buf.
int osph_sort_pairs(osph_ctx *ctx, int64_t n, int bits)
{
    if (n <= 0) return 0;
    int nblocks = div_up(n, SORT_TILE);
    int passes = (bits + 7) / 8;
    if (passes < 1) passes = 1;
    int cur = ctx->sorted_buf;
    for (int p = 0; p < passes; p++) {
        int shift = 8 * p;
        k_sort_hist<<<nblocks, SORT_THREADS, 0, ctx->stream>>>(ctx->key[cur], (int)n, shift, nblocks, ctx->hist);
        OSPH_LAUNCH_CHECK();
        k_sort_rowscan<<<RADIX, 256, 0, ctx->stream>>>(ctx->hist, nblocks, ctx->digit_tot);
        OSPH_LAUNCH_CHECK();
        k_sort_scatter<<<nblocks, SORT_THREADS, 0, ctx->stream>>>(ctx->key[cur], ctx->idx[cur], ctx->key[cur ^ 1],
                                                                    ctx->idx[cur ^ 1], (int)n, shift, nblocks,
                                                                    ctx->hist, ctx->digit_tot);
        OSPH_LAUNCH_CHECK();
        cur ^= 1;
    }
    ctx->sorted_buf = cur;
    return 0;
}
