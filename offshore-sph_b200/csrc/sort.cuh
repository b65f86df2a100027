// sort.cuh -- tile geometry shared by the radix sort and the key kernel that feeds it
#pragma once
#include "common.cuh"

#define SORT_THREADS 256
#define SORT_ITEMS 8
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)
#define SORT_MAX_BITS 11              // widest digit (2048 buckets)

int osph_sort_digit_bits(int bits);   // digit width chosen for `bits` significant key bits: 8, 10 or 11

#ifdef __CUDACC__
// One shared-memory atomic per distinct digit of a warp: neighbouring particles share their cell, so a warp's keys
// repeat and plain per-thread atomics serialise on the same counter (ncu: the histogram kernel spent its time on a
// 4 MB read).  Every lane of the warp must call this.
__device__ __forceinline__ void warp_hist_add(unsigned int *cnt, unsigned int digit, bool ok)
{
    const unsigned int peers = __match_any_sync(0xffffffffu, ok ? digit : 0xffffffffu);
    if (ok && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cnt[digit], (unsigned int)__popc(peers));
}
#endif

