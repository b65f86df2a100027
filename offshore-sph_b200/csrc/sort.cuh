// sort.cuh -- tile geometry shared by the radix sort and the key kernel that feeds it
#pragma once
#include "common.cuh"

#define SORT_THREADS 256
#define SORT_ITEMS 8
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)
#define SORT_MAX_BITS 11              // widest digit (2048 buckets)

int osph_sort_digit_bits(int bits);   // digit width chosen for `bits` significant key bits: 8, 10 or 11
