// binsort.cu -- counting sort of the particles by acceleration-grid cell, and the cell table that falls out of it.
//
// Replaces the serial head-insertion binning of the reference (src/Tools/NNLinkedList.py:129-141) like the radix sort of
// sort.cu did, with a third of the launches: the keys ARE cell numbers, so one histogram over the cells, one exclusive scan
// and one scatter sort the particles, and the scan is the cell table (begin, end).
//
//   k_bin_keys   (step.cu)  cell key of every particle + its arrival rank in the cell (one global atomic per distinct cell
//                           of a warp); counts[cell]++ and tile_sums[cell / BIN_TILE]++
//   k_bin_scan              one pass over the counts: exclusive offsets -> cell_range[cell] = (begin, end); the counts are
//                           zeroed for the next build on the way.  The tile totals are there already, so no CTA waits for
//                           another.
//   k_bin_scatter           slot = begin[cell] + arrival rank
//   ranking (k_gather)      arrival order inside a cell depends on the atomics; every particle counts the cell mates with a
//                           smaller storage slot and moves there (bin_canonical_slot, common.cuh).  The result is the order of
//                           a STABLE sort by cell (cell, then storage slot) -- bit-identical to the radix sort, so the summation
//                           order of the pair kernel stays deterministic.  Fused into the gather kernel, whose state loads travel
//                           under the ranking loop; a kernel of its own (k_bin_rank) only on builds that reorder the state.
//
// Algorithmic traffic per particle: keys 16 + 8, scatter 12 + 8, rank 8 + ~cell occupancy x 4 (cached) + 4; per cell 4 + 8.
#include "common.cuh"
#include "step.cuh"

#define BIN_THREADS 512
#define BIN_ITEMS 8
static_assert(BIN_TILE == BIN_THREADS * BIN_ITEMS, "scan tile = one CTA");

// Exclusive scan of the cell counts = the cell table.  The totals of the scan tiles (BIN_TILE cells each) were accumulated
// by k_bin_keys next to the counts, so a CTA sums the totals of the tiles before its own and scans its tile: no look-back,
// no spinning on other CTAs (a single-pass scan with tile tickets and a look-back spent 14-16 us waiting, measured).
// tile_sums is double-buffered by the parity of the sort count: this kernel clears the buffer of the NEXT sort.
__global__ void __launch_bounds__(BIN_THREADS)
k_bin_scan(unsigned int *__restrict__ counts, const GridParams *__restrict__ gp, int2 *__restrict__ cell_range,
           unsigned int *__restrict__ tile_sums_base, int n_tiles)
{
    if (!gp->do_sort) return;                             // this build reuses the binning of the last sort (k_grid_params)
    const unsigned int parity = gp->sort_count & 1u;
    const unsigned int *__restrict__ tile_sums = tile_sums_base + (size_t)parity * n_tiles;
    unsigned int *__restrict__ tile_sums_next = tile_sums_base + (size_t)(parity ^ 1u) * n_tiles;
    __shared__ unsigned int wsum[BIN_THREADS / 32];
    __shared__ unsigned int s_prefix;
    const unsigned int tile = blockIdx.x;
    const long long n_cells = (long long)gp->gnx * gp->gny + 1;          // + the cell behind the table (parked particles)
    const long long base = (long long)tile * BIN_TILE + (long long)threadIdx.x * BIN_ITEMS;
    // sums of the earlier tiles (at most a few hundred values)
    unsigned int part = 0;
    for (unsigned int p = threadIdx.x; p < tile; p += BIN_THREADS) part += tile_sums[p];
    static_assert(BIN_ITEMS == 8, "two uint4 per thread");
    unsigned int v[BIN_ITEMS], t = 0;
    if (base + BIN_ITEMS <= n_cells) {
        const uint4 a = *reinterpret_cast<const uint4 *>(counts + base), c = *reinterpret_cast<const uint4 *>(counts + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    } else {
#pragma unroll
        for (int k = 0; k < BIN_ITEMS; k++) v[k] = base + k < n_cells ? counts[base + k] : 0u;
    }
#pragma unroll
    for (int k = 0; k < BIN_ITEMS; k++) t += v[k];
    if (t) {                                                             // ready for the next build
        if (base + BIN_ITEMS <= n_cells) {
            *reinterpret_cast<uint4 *>(counts + base) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4 *>(counts + base + 4) = make_uint4(0, 0, 0, 0);
        } else {
#pragma unroll
            for (int k = 0; k < BIN_ITEMS; k++) if (base + k < n_cells) counts[base + k] = 0u;
        }
    }
    if (threadIdx.x == 0) { s_prefix = 0; if ((int)tile < n_tiles) tile_sums_next[tile] = 0u; }
    // block scan of the thread sums
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned int s = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int u = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 31) wsum[w] = s;
    __syncthreads();
    if (lane == 0 && part) atomicAdd(&s_prefix, part);
    unsigned int woff = 0;
#pragma unroll
    for (int k = 0; k < BIN_THREADS / 32; k++) { unsigned int u = wsum[k]; if (k < w) woff += u; }
    __syncthreads();
    unsigned int run = s_prefix + woff + s - t;
    if (base + BIN_ITEMS <= n_cells) {
        int4 o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            o[k].x = (int)run; run += v[2 * k]; o[k].y = (int)run;
            o[k].z = (int)run; run += v[2 * k + 1]; o[k].w = (int)run;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) reinterpret_cast<int4 *>(cell_range + base)[k] = o[k];
    } else {
#pragma unroll
        for (int k = 0; k < BIN_ITEMS; k++) {
            if (base + k < n_cells) cell_range[base + k] = make_int2((int)run, (int)(run + v[k]));
            run += v[k];
        }
    }
}

// slot = begin[cell] + arrival rank.  Besides the storage slot, every sorted position receives its cell's (begin, end): the
// ranking that follows (fused into the gather kernel, or k_bin_rank) then starts without a dependent table lookup.
#define BIN_SCATTER_ITEMS 4
__global__ void __launch_bounds__(256)
k_bin_scatter(const unsigned int *__restrict__ key, const unsigned int *__restrict__ arrival, int n,
              const int2 *__restrict__ cell_range, int2 *__restrict__ range_out, unsigned int *__restrict__ idx_out,
              const GridParams *__restrict__ gp)
{
    if (!gp->do_sort) return;                             // (a quarter of the CTAs of a one-item kernel: the early exit is cheaper)
    unsigned int k[BIN_SCATTER_ITEMS], a[BIN_SCATTER_ITEMS];
    const int base = blockIdx.x * (256 * BIN_SCATTER_ITEMS) + threadIdx.x;
#pragma unroll
    for (int r = 0; r < BIN_SCATTER_ITEMS; r++) {
        const int i = base + r * 256;
        k[r] = i < n ? key[i] : 0u; a[r] = i < n ? arrival[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < BIN_SCATTER_ITEMS; r++) {
        const int i = base + r * 256;
        if (i < n) {
            const int2 rg = cell_range[k[r]];
            const unsigned int pos = (unsigned int)rg.x + a[r];
            range_out[pos] = rg;
            idx_out[pos] = (unsigned int)i;
        }
    }
}

// Canonical order as a kernel of its own: only on the builds that physically reorder the state (it needs the final
// permutation before the gather); every other build ranks inside k_gather.
__global__ void __launch_bounds__(256)
k_bin_rank(const int2 *__restrict__ range_sorted, const unsigned int *__restrict__ idx_arrival, int n,
           unsigned int *__restrict__ idx_out, StepScalars *sc, const GridParams *__restrict__ gp)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || !gp->do_sort) return;
    const int2 r = range_sorted[s];
    const unsigned int mine = idx_arrival[s];
    idx_out[bin_canonical_slot(r, s, mine, idx_arrival, sc)] = mine;
}

int osph_bin_alloc(osph_ctx *ctx, int64_t cell_cap)
{
    cudaFree(ctx->bin_counts); cudaFree(ctx->bin_tiles); ctx->bin_counts = nullptr; ctx->bin_tiles = nullptr;
    const int64_t tiles = div_up(cell_cap + 1, BIN_TILE);
    OSPH_CUDA(cudaMalloc(&ctx->bin_counts, sizeof(unsigned int) * (size_t)(cell_cap + 1)));
    OSPH_CUDA(cudaMalloc(&ctx->bin_tiles, sizeof(unsigned int) * (size_t)(2 * tiles)));      // two buffers, by build parity
    OSPH_CUDA(cudaMemsetAsync(ctx->bin_counts, 0, sizeof(unsigned int) * (size_t)(cell_cap + 1), ctx->stream));
    OSPH_CUDA(cudaMemsetAsync(ctx->bin_tiles, 0, sizeof(unsigned int) * (size_t)(2 * tiles), ctx->stream));
    ctx->bin_tile_cap = tiles;
    return 0;
}

void osph_bin_free(osph_ctx *ctx)
{
    cudaFree(ctx->bin_counts); cudaFree(ctx->bin_tiles);
    ctx->bin_counts = nullptr; ctx->bin_tiles = nullptr;
}

// key[0] / idx[0] hold (cell key, arrival rank) per particle in storage order and counts[] the cell histogram (k_bin_keys).
// Afterwards: cell_range filled; idx[1] = sorted position -> storage slot in ARRIVAL order inside each cell and
// bin_ranges (the spare state column) = that position's cell range.  rank_now: also run k_bin_rank, idx[0] = the
// canonical permutation and sorted_buf = 0; otherwise the gather kernel ranks (and writes idx[0]) on the way.
int osph_bin_sort(osph_ctx *ctx, int64_t n_all, bool rank_now)
{
    if (n_all <= 0) return 0;
    const int tiles = (int)ctx->bin_tile_cap;
    k_bin_scan<<<tiles, BIN_THREADS, 0, ctx->stream>>>(ctx->bin_counts, ctx->d_grid, ctx->cell_range, ctx->bin_tiles, tiles);
    OSPH_LAUNCH_CHECK();
    const int grid = div_up(n_all, 256);
    int2 *ranges = reinterpret_cast<int2 *>(ctx->scratch);
    k_bin_scatter<<<div_up(n_all, 256 * BIN_SCATTER_ITEMS), 256, 0, ctx->stream>>>(ctx->key[0], ctx->idx[0], (int)n_all, ctx->cell_range, ranges,
                                                                                  ctx->idx[1], ctx->d_grid);
    OSPH_LAUNCH_CHECK();
    if (rank_now) {
        k_bin_rank<<<grid, 256, 0, ctx->stream>>>(ranges, ctx->idx[1], (int)n_all, ctx->idx[0], ctx->d_sc, ctx->d_grid);
        OSPH_LAUNCH_CHECK();
    }
    ctx->sorted_buf = 0;
    return 0;
}
