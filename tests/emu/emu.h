// emu.h -- TEST INFRASTRUCTURE, never shipped and never loaded by the product.
//
// A small SIMT emulator that lets the UNMODIFIED kernel sources of offshore-sph_b200/csrc be compiled by g++ and
// executed on the CPU, so that `pytest -m "not gpu"` exercises the real kernel code (indexing, staging, warp-synchronous
// control flow, reductions, edge paths) against the oracle in a container without a GPU.  It says nothing about
// performance, about races between CTAs or about the memory model; the GPU tests remain the parity gate.
//
// Model: every CUDA thread of a CTA is a fiber (own stack, cooperative switch); CTAs run one after the other.
// __syncthreads() and the *_sync warp collectives are the only switch points: a fiber that reaches one deposits its
// operand and yields; when every live thread of the CTA / lane of the warp has arrived the scheduler computes the
// results and resumes them.  Exited threads count as arrived.  __shared__ variables become references to storage that ends
// in front of an inaccessible page (one CTA at a time), dynamic shared memory likewise, sized per launch: an access past
// the end of either faults at the instruction.  tests/emu/preprocess.py rewrites the three
// constructs g++ cannot parse: kernel<<<...>>>(...) launches, `extern __shared__` declarations and the two inline-PTX
// MUFU seeds (modelled with their documented 2^-20 accuracy: lower 32 mantissa bits zero).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>      // pulled in by nccl.h (slab_nccl.cu): parse them as host headers, before __CUDACC__ is defined
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <type_traits>

#undef __host__
#undef __device__
#undef __global__
#undef __shared__
#undef __forceinline__
#undef __launch_bounds__
#define __host__
#define __device__
#define __global__
#define __shared__ static
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#ifndef __CUDACC__
#define __CUDACC__ 1            // the product headers guard their device helpers with this
#endif
#define OSPH_EMU 1

namespace emu {

struct ThreadCtx {
    uint3 tid, bid;
    dim3 bdim, gdim;
    int linear, lane, warp;
};
extern ThreadCtx *cur;
extern unsigned char *dyn_smem_ptr;
inline void *dyn_smem() { return dyn_smem_ptr; }
inline size_t shared_offset(const void *p) { return (size_t)((const unsigned char *)p - dyn_smem_ptr); }   // __cvta_generic_to_shared
void *static_smem_alloc(size_t bytes, size_t align);
template <typename T> inline T *static_smem() { return (T *)static_smem_alloc(sizeof(T), alignof(T)); }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);

template <typename... P, typename... A>
void launch_k(dim3 grid, dim3 block, size_t smem, void (*kernel)(P...), A &&...args)
{
    std::tuple<typename std::decay<P>::type...> params(std::forward<A>(args)...);     // by value, like a launch
    launch(grid, block, smem, [&]() { std::apply(kernel, params); });
}

enum Op { OP_SHFL_IDX, OP_SHFL_XOR, OP_SHFL_UP, OP_SHFL_DOWN, OP_BALLOT, OP_MATCH, OP_SYNC };
void block_barrier();
uint64_t warp_collective(Op op, unsigned mask, uint64_t value, int param, int width);

// packed FP32 pairs (FADD2 / FMUL2 / FFMA2 of sm_100): element-wise IEEE single precision, lo element first
inline uint64_t f32x2_pack(float lo, float hi) { uint32_t a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4); return (uint64_t)a | ((uint64_t)b << 32); }
inline void f32x2_unpack(uint64_t v, float &lo, float &hi) { uint32_t a = (uint32_t)v, b = (uint32_t)(v >> 32); memcpy(&lo, &a, 4); memcpy(&hi, &b, 4); }
#define EMU_F32X2(name, expr_lo, expr_hi)                                                         \
    inline uint64_t name(uint64_t x, uint64_t y, uint64_t z = 0)                                  \
    {                                                                                             \
        float x0, x1, y0, y1, z0, z1; f32x2_unpack(x, x0, x1); f32x2_unpack(y, y0, y1); f32x2_unpack(z, z0, z1); \
        (void)z0; (void)z1;                                                                       \
        return f32x2_pack(expr_lo, expr_hi);                                                      \
    }
EMU_F32X2(f32x2_sub, x0 - y0, x1 - y1)
EMU_F32X2(f32x2_mul, x0 * y0, x1 * y1)
EMU_F32X2(f32x2_fma, std::fmaf(x0, y0, z0), std::fmaf(x1, y1, z1))
#undef EMU_F32X2
double rcp_approx_f64(double x);
double rsqrt_approx_f64(double x);

template <typename T> inline uint64_t to_bits(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> inline T from_bits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::cur->bid)
#define blockDim (emu::cur->bdim)
#define gridDim (emu::cur->gdim)
#define warpSize 32

// cuda_runtime.h only declares the typed overload under nvcc
template <typename T> inline cudaError_t cudaFuncSetAttribute(T *, cudaFuncAttribute, int) { return cudaSuccess; }

// ---- synchronisation and warp collectives ---------------------------------------------------------------------------
inline void __syncthreads() { emu::block_barrier(); }
// barrier + OR over the CTA: set between two barriers, read, and cleared by thread 0 behind a third one (the next call starts
// with a barrier, so the clearing cannot overtake a later contribution)
namespace emu { extern int block_or_flag; }
inline int __syncthreads_or(int pred)
{
    emu::block_barrier();
    if (pred) emu::block_or_flag = 1;
    emu::block_barrier();
    const int r = emu::block_or_flag;
    emu::block_barrier();
    if (emu::cur->linear == 0) emu::block_or_flag = 0;
    return r;
}
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_collective(emu::OP_SYNC, mask, 0, 0, 32); }
template <typename T> inline T __shfl_sync(unsigned m, T v, int src, int w = 32) { return emu::from_bits<T>(emu::warp_collective(emu::OP_SHFL_IDX, m, emu::to_bits(v), src, w)); }
template <typename T> inline T __shfl_xor_sync(unsigned m, T v, int x, int w = 32) { return emu::from_bits<T>(emu::warp_collective(emu::OP_SHFL_XOR, m, emu::to_bits(v), x, w)); }
template <typename T> inline T __shfl_up_sync(unsigned m, T v, unsigned d, int w = 32) { return emu::from_bits<T>(emu::warp_collective(emu::OP_SHFL_UP, m, emu::to_bits(v), (int)d, w)); }
template <typename T> inline T __shfl_down_sync(unsigned m, T v, unsigned d, int w = 32) { return emu::from_bits<T>(emu::warp_collective(emu::OP_SHFL_DOWN, m, emu::to_bits(v), (int)d, w)); }
inline unsigned __ballot_sync(unsigned m, int pred) { return (unsigned)emu::warp_collective(emu::OP_BALLOT, m, pred ? 1 : 0, 0, 32); }
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
template <typename T> inline unsigned __match_any_sync(unsigned m, T v) { return (unsigned)emu::warp_collective(emu::OP_MATCH, m, emu::to_bits(v), 0, 32); }
// redux.sync: same result as a butterfly of shuffles (full or partial masks whose lanes all arrive)
#define EMU_REDUX(name, T, expr)                                                                                        \
    inline T name(unsigned m, T v)                                                                                      \
    {                                                                                                                   \
        for (int o = 16; o > 0; o >>= 1) { T t = __shfl_xor_sync(m, v, o); v = (expr); }                                \
        return v;                                                                                                       \
    }
EMU_REDUX(__reduce_max_sync, unsigned, t > v ? t : v)
EMU_REDUX(__reduce_min_sync, unsigned, t < v ? t : v)
EMU_REDUX(__reduce_max_sync, int, t > v ? t : v)
EMU_REDUX(__reduce_min_sync, int, t < v ? t : v)
EMU_REDUX(__reduce_add_sync, unsigned, v + t)
EMU_REDUX(__reduce_add_sync, int, v + t)
EMU_REDUX(__reduce_or_sync, unsigned, v | t)
EMU_REDUX(__reduce_and_sync, unsigned, v & t)
#undef EMU_REDUX
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() {}
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }     // peers are other processes (shared memory)
void emu_yield_cpu();
inline void __nanosleep(unsigned) { emu_yield_cpu(); }
long long emu_clock64();                 // monotonic clock in half-nanoseconds (a 2 GHz SM clock)
inline long long clock64() { return emu_clock64(); }

// ---- integer / conversion intrinsics --------------------------------------------------------------------------------
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline long long __double_as_longlong(double v) { return emu::from_bits<long long>(emu::to_bits(v)); }
inline double __longlong_as_double(long long v) { return emu::from_bits<double>(emu::to_bits(v)); }
inline int __double2hiint(double v) { return (int)(emu::to_bits(v) >> 32); }
inline int __double2loint(double v) { return (int)(emu::to_bits(v) & 0xffffffffu); }
inline int __float_as_int(float v) { return emu::from_bits<int>(emu::to_bits(v)); }
inline float __int_as_float(int v) { return emu::from_bits<float>(emu::to_bits(v)); }

// ---- IEEE round-to-nearest arithmetic that the compiler must not contract (the TU is built with -ffp-contract=off) ----
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __double2float_ru(double v)
{
    float f = (float)v;
    if ((double)f < v) f = std::nextafter(f, INFINITY);
    return f;
}
inline float __double2float_rn(double v) { return (float)v; }
inline float __fdividef(float a, float b) { return a / b; }
inline float __expf(float a) { return std::exp(a); }
inline float __powf(float a, float b) { return std::pow(a, b); }
inline float __logf(float a) { return std::log(a); }
inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
inline double rsqrt(double a) { return 1.0 / std::sqrt(a); }

using std::isfinite;
using std::isinf;
using std::isnan;

template <typename A, typename B> inline typename std::common_type<A, B>::type min(A a, B b)
{
    typedef typename std::common_type<A, B>::type T;
    return (T)b < (T)a ? (T)b : (T)a;
}
template <typename A, typename B> inline typename std::common_type<A, B>::type max(A a, B b)
{
    typedef typename std::common_type<A, B>::type T;
    return (T)a < (T)b ? (T)b : (T)a;
}

// ---- atomics (CTAs run one after the other, but keep them real so that the engine may run CTAs on several threads) ----
template <typename T> inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline double atomicAdd(double *p, double v) { double o = *p; *p = o + v; return o; }
inline float atomicAdd(float *p, float v) { float o = *p; *p = o + v; return o; }
inline int atomicAdd(int *p, unsigned v) { return __atomic_fetch_add(p, (int)v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned *p, int v) { return __atomic_fetch_add(p, (unsigned)v, __ATOMIC_RELAXED); }
template <typename T> inline T atomicSub(T *p, T v) { return __atomic_fetch_sub(p, v, __ATOMIC_RELAXED); }
template <typename T> inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicOr(unsigned *p, int v) { return __atomic_fetch_or(p, (unsigned)v, __ATOMIC_RELAXED); }
template <typename T> inline T atomicAnd(T *p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
template <typename T> inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
template <typename T> inline T atomicCAS(T *p, T cmp, T v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return cmp; }
template <typename T> inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
