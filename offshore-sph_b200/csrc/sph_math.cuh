// sph_math.cuh -- device math shared by the fused pair kernel and the leaf entry points:
// fast double reciprocal / rsqrt (MUFU seed + Newton) and the three smoothing kernels.
#pragma once
#include "common.cuh"

#define PI_D 3.14159265358979323846
#ifndef OSPH_NEWTON_STEPS
#define OSPH_NEWTON_STEPS 0      // 0: one third-order step after the 2^-20 MUFU seed (error ~ seed^3 = 2^-60, below half an ulp);
#endif                           // 2 / 3: that many Newton steps (2^-40, then full double: tools/mufu_accuracy.cu)

template <typename Real> struct R2;
template <> struct R2<double> { typedef double2 type; };
template <> struct R2<float> { typedef float2 type; };

// ---- fast reciprocal / rsqrt in double: MUFU seed + refinement, ~1 ulp, no special-case branches ----------
// Third-order step: with y = (1/x)(1 + d), e = 1 - x y = -d exactly (fma), and 1/x = y / (1 - e) = y (1 + e + e^2 + e^3 ...):
// y (1 + e + e^2) is off by d^3.  Three dependent DFMAs instead of the four of two Newton steps.
__device__ __forceinline__ double rcp_fast(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if OSPH_NEWTON_STEPS == 0
    const double e = fma(-x, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
#else
    double e = fma(-x, y, 1.0); y = fma(y, e, y);
    e = fma(-x, y, 1.0); y = fma(y, e, y);
#if OSPH_NEWTON_STEPS > 2
    e = fma(-x, y, 1.0); y = fma(y, e, y);
#endif
    return y;
#endif
}
__device__ __forceinline__ float rcp_fast(float x) { return __fdividef(1.0f, x); }

// Same idea for 1/sqrt(x): e = 1 - x y^2, 1/sqrt(x) = y (1 - e)^(-1/2) = y (1 + e/2 + 3 e^2/8 + 5 e^3/16 ...); the step keeps the
// first two terms (error 5/16 e^3 with |e| <= 2^-19).  Five FP64 instructions instead of seven.
__device__ __forceinline__ double rsqrt_fast(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if OSPH_NEWTON_STEPS == 0
    const double e = fma(-(x * y), y, 1.0);
    const double t = e * fma(e, 0.375, 0.5);
    return fma(y, t, y);
#else
    // Newton for 1/sqrt(x): y <- y + y*(1 - x*y*y)/2
    double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5); y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5); y = fma(y, e, y);
#if OSPH_NEWTON_STEPS > 2
    e = fma(-hx * y, y, 0.5); y = fma(y, e, y);
#endif
    return y;
#endif
}
__device__ __forceinline__ float rsqrt_fast(float x) { return rsqrtf(x); }

// rsqrt_fast in two halves, for a software-pipelined loop (pair.cu, PAIR_UH_PIPE): the seed and the first residual of entry
// k + 1 are formed while entry k is evaluated.  Same operations as rsqrt_fast, hence the same bits.
__device__ __forceinline__ void rsqrt_begin(double x, double &y, double &e)
{
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if OSPH_NEWTON_STEPS == 0
    e = fma(-(x * y), y, 1.0);
#else
    e = x;
#endif
}
__device__ __forceinline__ double rsqrt_end(double y, double e)
{
#if OSPH_NEWTON_STEPS == 0
    const double t = e * fma(e, 0.375, 0.5);
    return fma(y, t, y);
#else
    const double hx = 0.5 * e;           // e carries x
    double r = fma(-hx * y, y, 0.5); y = fma(y, r, y);
    r = fma(-hx * y, y, 0.5); y = fma(y, r, y);
#if OSPH_NEWTON_STEPS > 2
    r = fma(-hx * y, y, 0.5); y = fma(y, r, y);
#endif
    return y;
#endif
}
__device__ __forceinline__ void rsqrt_begin(float x, float &y, float &e) { y = rsqrtf(x); e = 0.f; }
__device__ __forceinline__ float rsqrt_end(float y, float) { return y; }

__device__ __forceinline__ double exp_neg(double x) { return exp(-x); }
__device__ __forceinline__ float exp_neg(float x) { return __expf(-x); }

__device__ __forceinline__ double pow_gen(double a, double b) { return pow(a, b); }
__device__ __forceinline__ float pow_gen(float a, float b) { return __powf(a, b); }

// ---- packed single precision (sm_100: two FP32 operations per instruction, FADD2 / FMUL2 / FFMA2) ----------------
// A pair of floats travels as one 64-bit register pair (lo = first element).
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long sub_f32x2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul_f32x2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma_f32x2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// 16-bit store through a 32-bit shared-window address (__cvta_generic_to_shared)
__device__ __forceinline__ void sts_u16(unsigned int addr, unsigned short v)
{
    asm volatile("st.shared.u16 [%0], %1;" : : "r"(addr), "h"(v) : "memory");
}

// ---- smoothing kernels: value w and gradient factor g with  grad W = g * (dx, dy) ---------------------
// reference: CubicSpline.py:10-70, Wendland.py:9-64, Gaussian.py:16-59.  inv_r == 0 encodes r < 1e-10
// (the reference zeroes the gradient there).
// CUT = false (pair.cu, PAIR_LEAN): the caller knows r^2 <= 4 h^2, so q exceeds 2 by rounding only and the polynomials are
// used as they stand (Wendland: (1 - q/2)^5 is 1e-80 there).
// Normalisation of kernel KID, formed exactly as the pair kernel's evaluation forms it (the cubic spline of the fused kernel
// goes through cubic_pair, whose product is associated differently: see kernel_alpha_pair).
template <typename Real, int KID>
__device__ __forceinline__ Real kernel_alpha(Real inv_h)
{
    const Real ih2 = inv_h * inv_h;
    return Real(KID == OSPH_KERNEL_CUBIC ? 10.0 / (7.0 * PI_D) : (KID == OSPH_KERNEL_WENDLAND ? 9.0 / (4.0 * PI_D) : 1.0 / PI_D)) * ih2;
}
template <typename Real, int KID>
__device__ __forceinline__ Real kernel_alpha_pair(Real inv_h)
{
    if (KID == OSPH_KERNEL_CUBIC) return Real(10.0 / (7.0 * PI_D)) * inv_h * inv_h;
    return kernel_alpha<Real, KID>(inv_h);
}
template <typename Real, int KID, bool CUT = true>
__device__ __forceinline__ void sph_kernel_a(Real q, Real alpha, Real inv_h, Real inv_r, Real &w, Real &g);

template <typename Real, int KID, bool CUT = true>
__device__ __forceinline__ void sph_kernel(Real q, Real inv_h, Real inv_r, Real &w, Real &g)
{
    sph_kernel_a<Real, KID, CUT>(q, kernel_alpha<Real, KID>(inv_h), inv_h, inv_r, w, g);
}

template <typename Real, int KID, bool CUT>
__device__ __forceinline__ void sph_kernel_a(Real q, Real alpha, Real inv_h, Real inv_r, Real &w, Real &g)
{
    if (KID == OSPH_KERNEL_CUBIC) {
        Real wv, gv;
        if (q > Real(1)) { Real t = Real(2) - q; Real t2 = t * t; wv = Real(0.25) * t2 * t; gv = Real(-0.75) * t2; }
        else { wv = Real(1) - Real(1.5) * q * q * (Real(1) - Real(0.5) * q); gv = Real(-3) * q * (Real(1) - Real(0.75) * q); }
        if (q > Real(2)) { wv = Real(0); gv = Real(0); }
        w = alpha * wv;
        g = alpha * gv * inv_h * inv_r;
    } else if (KID == OSPH_KERNEL_WENDLAND) {
        Real in = Real(1) - Real(0.5) * q;
        Real in2 = in * in, in4 = in2 * in2, in5 = in4 * in;
        Real wv = in5 * in * (Real(35.0 / 12.0) * q * q + Real(3) * q + Real(1));
        Real gv = in5 * Real(-14.0 / 3.0) * q * (Real(1) + Real(2.5) * q);
        if (CUT && q >= Real(2)) { wv = Real(0); gv = Real(0); }
        w = alpha * wv;
        g = alpha * gv * inv_h * inv_r;
    } else {
        Real e = exp_neg(q * q);
        Real wv = alpha * e;                          // q <= 3 is decided by the caller (set membership)
        w = wv;
        g = Real(-2) * q * wv * inv_h * inv_r;      // dwdq / (r h) ; inv_r == 0 also covers r*h <= 1e-12
    }
}

// max(d, 0) and min(d, 0) for numbers.  fmax / fmin on doubles compile to DSETP.MAX plus six moves and selects for their
// NaN rules, and the compiler turns every C++ spelling of `d > 0 ? d : 0` (also on the bit pattern) back into max.f64;
// setp + selp written in PTX stay three instructions.  NaN maps to 0, as with fmax / fmin.
__device__ __forceinline__ double pos_part(double d)
{
    double r;
    asm("{ .reg .pred p; setp.gt.f64 p, %1, 0d0000000000000000; selp.f64 %0, %1, 0d0000000000000000, p; }" : "=d"(r) : "d"(d));
    return r;
}
__device__ __forceinline__ double neg_part(double d)
{
    double r;
    asm("{ .reg .pred p; setp.lt.f64 p, %1, 0d0000000000000000; selp.f64 %0, %1, 0d0000000000000000, p; }" : "=d"(r) : "d"(d));
    return r;
}
__device__ __forceinline__ float pos_part(float d) { return fmaxf(d, 0.0f); }
__device__ __forceinline__ float neg_part(float d) { return fminf(d, 0.0f); }
// The same two selections decided by the SIGN BIT, an integer compare on the ALU pipe instead of a DSETP on the FP64 pipe
// (PAIR_ISIGN, pair.cu: the heavy loop of the double instantiation keeps that pipe 74 % busy).  For finite d the result is
// the same number (d = -0 yields -0 from neg_part_s where neg_part yields +0: the sums that follow cannot tell); a NaN with
// a clear sign bit passes pos_part_s -- the callers only use it where d is finite.
__device__ __forceinline__ double pos_part_s(double d)
{
    double r;
    asm("{ .reg .pred p; .reg .b32 lo, hi; mov.b64 {lo, hi}, %1; setp.ge.s32 p, hi, 0; selp.f64 %0, %1, 0d0000000000000000, p; }" : "=d"(r) : "d"(d));
    return r;
}
__device__ __forceinline__ double neg_part_s(double d)
{
    double r;
    asm("{ .reg .pred p; .reg .b32 lo, hi; mov.b64 {lo, hi}, %1; setp.lt.s32 p, hi, 0; selp.f64 %0, %1, 0d0000000000000000, p; }" : "=d"(r) : "d"(d));
    return r;
}
__device__ __forceinline__ float pos_part_s(float d) { return fmaxf(d, 0.0f); }
__device__ __forceinline__ float neg_part_s(float d) { return fminf(d, 0.0f); }

// Cubic spline of the fused pair kernel: the same piecewise polynomial written with clamped terms,
//   W/alpha = (2-q)+^3 / 4 - (1-q)+^3,   W'/alpha = -3/4 (2-q)+^2 + 3 (1-q)+^2,
// identical to CubicSpline.py:10-70 on every branch (q <= 1, 1 < q <= 2, q > 2) up to rounding of O(1e-16) terms, without
// evaluating both branches and selecting.
// CLAMP_OUTER = false (pair.cu, PAIR_LEAN): the caller knows r^2 <= 4 h^2, so 2 - q can be negative by rounding only (a few
// 1e-16: 1e-47 in W, 1e-31 in W') and the outer term is used as it stands.
// ISIGN: the inner clamp by the sign bit (pos_part_s).  The *_a form takes the normalisation alpha = kernel_alpha(inv_h) from
// the caller (uniform smoothing length: a loop constant, pair.cu PAIR_UH).
template <typename Real, bool CLAMP_OUTER = true, bool ISIGN = false>
__device__ __forceinline__ void cubic_pair_a(Real q, Real alpha, Real inv_h, Real inv_r, Real &w, Real &g)
{
    const Real t2 = CLAMP_OUTER ? pos_part(Real(2) - q) : Real(2) - q, t1 = ISIGN ? pos_part_s(Real(1) - q) : pos_part(Real(1) - q);
    const Real s2 = t2 * t2, s1 = t1 * t1;
    const Real wv = fma(-s1, t1, Real(0.25) * s2 * t2);
    const Real gv = fma(Real(3), s1, Real(-0.75) * s2);
    w = alpha * wv;
    g = alpha * gv * inv_h * inv_r;
}
template <typename Real, bool CLAMP_OUTER = true, bool ISIGN = false>
__device__ __forceinline__ void cubic_pair(Real q, Real inv_h, Real inv_r, Real &w, Real &g)
{
    cubic_pair_a<Real, CLAMP_OUTER, ISIGN>(q, Real(10.0 / (7.0 * PI_D)) * inv_h * inv_h, inv_h, inv_r, w, g);
}

