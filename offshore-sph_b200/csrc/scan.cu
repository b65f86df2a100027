// scan.cu -- in-place exclusive prefix sum of 32-bit counters (three HBM-bound kernels: tile sums, scan of
// the tile sums by one CTA, tile rescan + offset).  Used to rank the owned particles inside the sorted order
// (slab mode) where owned and ghost particles are interleaved.
#include "common.cuh"
#include "step.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ unsigned int block_exclusive(unsigned int v, unsigned int *wsum, unsigned int &total)
{
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
    }
    if (lane == 31) wsum[w] = s;
    __syncthreads();
    unsigned int off = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < SCAN_THREADS / 32; k++) { unsigned int t = wsum[k]; if (k < w) off += t; tot += t; }
    total = tot;
    __syncthreads();
    return off + s - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_tile_sums(const unsigned int *__restrict__ d, int n, unsigned int *__restrict__ sums)
{
    __shared__ unsigned int wsum[SCAN_THREADS / 32];
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned int v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) v += d[base + k];
    unsigned int total;
    block_exclusive(v, wsum, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_sums(unsigned int *__restrict__ sums, int nb)
{
    __shared__ unsigned int wsum[SCAN_THREADS / 32];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += SCAN_THREADS) {
        int i = base + threadIdx.x;
        unsigned int v = i < nb ? sums[i] : 0u, total;
        unsigned int e = block_exclusive(v, wsum, total);
        unsigned int c = carry;
        if (i < nb) sums[i] = c + e;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_apply(unsigned int *__restrict__ d, int n, const unsigned int *__restrict__ sums)
{
    __shared__ unsigned int wsum[SCAN_THREADS / 32];
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned int v[SCAN_ITEMS], t = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = base + k < n ? d[base + k] : 0u; t += v[k]; }
    unsigned int total;
    unsigned int run = block_exclusive(t, wsum, total) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) d[base + k] = run; run += v[k]; }
}

int osph_scan_exclusive(osph_ctx *ctx, unsigned int *d, int n)
{
    if (n <= 0) return 0;
    int nb = div_up(n, SCAN_TILE);
    if (nb > ctx->scan_cap) {
        cudaFree(ctx->scan_block); ctx->scan_block = nullptr;
        OSPH_CUDA(cudaMalloc(&ctx->scan_block, sizeof(unsigned int) * (size_t)(nb + 1024)));
        ctx->scan_cap = nb + 1024;
    }
    k_scan_tile_sums<<<nb, SCAN_THREADS, 0, ctx->stream>>>(d, n, ctx->scan_block); OSPH_LAUNCH_CHECK();
    k_scan_sums<<<1, SCAN_THREADS, 0, ctx->stream>>>(ctx->scan_block, nb); OSPH_LAUNCH_CHECK();
    k_scan_apply<<<nb, SCAN_THREADS, 0, ctx->stream>>>(d, n, ctx->scan_block); OSPH_LAUNCH_CHECK();
    return 0;
}
