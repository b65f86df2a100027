// sort.cu -- stable LSD radix sort of (cell key, storage slot) pairs.
//
// Replaces the serial head-insertion binning of the reference (src/Tools/NNLinkedList.py:129-141):
// after the sort, the particles of one cell are contiguous and the cell table is two integers per cell.
// 8-bit digits, ceil(bits/8) passes (3 at 1 M .. 16 M particles).  Three kernels per pass, all
// HBM/latency-bound: per-tile digit histogram (the first one is fused into the key kernel by the caller),
// per-digit scan along the tiles, stable scatter (warp-level match_any ranking, so equal keys keep their
// input order and the neighbour summation order is deterministic).
// Algorithmic traffic per pass and pair: read key twice + slot once, write both = 20 B.
#include "common.cuh"
#include "sort.cuh"

template <int BITS>
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist(const unsigned int *__restrict__ key, int n, int shift, int nblocks, unsigned int *__restrict__ hist)
{
    constexpr int RADIX = 1 << BITS;
    __shared__ unsigned int cnt[RADIX];
    for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS) cnt[d] = 0;
    __syncthreads();
    int base = blockIdx.x * SORT_TILE;
    unsigned int kk[SORT_ITEMS];                      // all loads in flight before the first vote (see k_keys)
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        int i = base + r * SORT_THREADS + threadIdx.x;
        kk[r] = i < n ? key[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        int i = base + r * SORT_THREADS + threadIdx.x;
        const bool ok = i < n;
        warp_hist_add(cnt, ok ? ((kk[r] >> shift) & (RADIX - 1)) : 0u, ok);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS) hist[(size_t)d * nblocks + blockIdx.x] = cnt[d];
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

template <int BITS>
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_scatter(const unsigned int *__restrict__ key_in, const unsigned int *__restrict__ val_in,
               unsigned int *__restrict__ key_out, unsigned int *__restrict__ val_out, int n, int shift,
               int nblocks, const unsigned int *__restrict__ hist, const unsigned int *__restrict__ digit_tot)
{
    constexpr int RADIX = 1 << BITS;
    constexpr int NW = SORT_THREADS / 32;
    constexpr int PER = RADIX / SORT_THREADS;                 // digits per thread in the digit-total scan
    extern __shared__ unsigned int sort_smem[];
    unsigned int(*cnt)[RADIX] = reinterpret_cast<unsigned int(*)[RADIX]>(sort_smem);     // [NW][RADIX]
    unsigned int *dbase = sort_smem + NW * RADIX;                                        // [RADIX]
    __shared__ unsigned int wsum[NW];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

    // exclusive scan of the RADIX digit totals (every block repeats it; a few KB from L2)
    {
        unsigned int v[PER], t = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) { v[k] = digit_tot[threadIdx.x * PER + k]; t += v[k]; }
        unsigned int s = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        if (lane == 31) wsum[w] = s;
        for (int d = threadIdx.x; d < NW * RADIX; d += SORT_THREADS) sort_smem[d] = 0;
        __syncthreads();
        unsigned int run = s - t;
        for (int k = 0; k < w; k++) run += wsum[k];
#pragma unroll
        for (int k = 0; k < PER; k++) {
            int d = threadIdx.x * PER + k;
            dbase[d] = run + hist[(size_t)d * nblocks + blockIdx.x];
            run += v[k];
        }
    }
    __syncthreads();

    // warp w ranks keys [base + w*256, base + (w+1)*256) in input order
    const int base = blockIdx.x * SORT_TILE + w * (32 * SORT_ITEMS);
    unsigned int k[SORT_ITEMS], v[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {                    // all loads in flight before the first vote (see k_keys)
        int i = base + r * 32 + lane;
        bool ok = i < n;
        k[r] = ok ? key_in[i] : 0u;
        v[r] = ok ? val_in[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        int i = base + r * 32 + lane;
        bool ok = i < n;
        unsigned int d = ok ? ((k[r] >> shift) & (RADIX - 1)) : (unsigned int)RADIX;
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
    // per digit: turn the per-warp counts into per-warp offsets
    for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS) {
        unsigned int run = dbase[d];
#pragma unroll
        for (int kk = 0; kk < NW; kk++) {
            unsigned int t = cnt[kk][d];
            cnt[kk][d] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; r++) {
        int i = base + r * 32 + lane;
        if (i < n) {
            unsigned int pos = cnt[w][(k[r] >> shift) & (RADIX - 1)] + rank[r];
            key_out[pos] = k[r];
            val_out[pos] = v[r];
        }
    }
}

// Digit width.  Measured on B200 at 1 M pairs (profiles/r01/README.md): three 8-bit passes cost 3 x (6.8 + 4.5 + 9) us;
// two 11-bit passes cost 2 x (12 + 7 + 29) us because every 2048-key tile pays for 8 x 2048 per-warp counters.
// Wide digits only pay off with much larger tiles, so 8 bits it is; the templates keep the choice open.
int osph_sort_digit_bits(int bits)
{
    (void)bits;
    return 8;
}

int osph_sort_alloc(osph_ctx *ctx, int64_t cap)
{
    ctx->sort_blocks = div_up(cap, SORT_TILE);
    OSPH_CUDA(cudaMalloc(&ctx->hist, sizeof(unsigned int) * (size_t)(1 << SORT_MAX_BITS) * (size_t)ctx->sort_blocks));
    OSPH_CUDA(cudaMalloc(&ctx->digit_tot, sizeof(unsigned int) * (1 << SORT_MAX_BITS)));
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

template <int BITS>
static int sort_pass(osph_ctx *ctx, int n, int nblocks, int shift, int cur, bool have_hist)
{
    constexpr int RADIX = 1 << BITS;
    constexpr size_t smem = sizeof(unsigned int) * (SORT_THREADS / 32 + 1) * RADIX;
    static unsigned long long configured = 0;           // one bit per device ordinal
    if (!(configured >> (ctx->device & 63) & 1ull)) {
        OSPH_CUDA(cudaFuncSetAttribute(k_sort_scatter<BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured |= 1ull << (ctx->device & 63);
    }
    if (!have_hist) {
        k_sort_hist<BITS><<<nblocks, SORT_THREADS, 0, ctx->stream>>>(ctx->key[cur], n, shift, nblocks, ctx->hist);
        OSPH_LAUNCH_CHECK();
    }
    k_sort_rowscan<<<RADIX, 256, 0, ctx->stream>>>(ctx->hist, nblocks, ctx->digit_tot);
    OSPH_LAUNCH_CHECK();
    k_sort_scatter<BITS><<<nblocks, SORT_THREADS, smem, ctx->stream>>>(ctx->key[cur], ctx->idx[cur], ctx->key[cur ^ 1],
                                                                      ctx->idx[cur ^ 1], n, shift, nblocks, ctx->hist,
                                                                      ctx->digit_tot);
    OSPH_LAUNCH_CHECK();
    return 0;
}

// Sorts key[sorted_buf] / idx[sorted_buf] (n pairs, `bits` significant key bits); updates sorted_buf.
// first_hist_done: the caller already filled ctx->hist with the first pass's tile histograms (fused key kernel).
int osph_sort_pairs(osph_ctx *ctx, int64_t n, int bits, bool first_hist_done)
{
    if (n <= 0) return 0;
    int nblocks = div_up(n, SORT_TILE);
    int db = osph_sort_digit_bits(bits);
    int passes = (bits + db - 1) / db;
    if (passes < 1) passes = 1;
    int cur = ctx->sorted_buf;
    for (int p = 0; p < passes; p++) {
        int rc;
        bool have = p == 0 && first_hist_done;
        if (db == 8) rc = sort_pass<8>(ctx, (int)n, nblocks, db * p, cur, have);
        else if (db == 10) rc = sort_pass<10>(ctx, (int)n, nblocks, db * p, cur, have);
        else rc = sort_pass<11>(ctx, (int)n, nblocks, db * p, cur, have);
        if (rc) return rc;
        cur ^= 1;
    }
    ctx->sorted_buf = cur;
    return 0;
}
