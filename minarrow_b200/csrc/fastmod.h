// Float remainder (`%` on f32/f64 = C fmod; src/kernels/arithmetic/std.rs:150, simd.rs:404,471) without the libm loop.
//
// fmod is exact by definition (the result is representable), so any evaluation that is exact is bit-identical to libm's.
// CUDA's fmod is a shift-and-subtract loop of ~100 instructions and made `f64 % f64` compute-bound (3.6 TB/s, r01n).
// When the quotient is small enough to be an exactly representable integer — |a| / |b| < 2^(mantissa bits), which is every
// row of ordinary data — one division, one truncation and one fused multiply-add give the same bits:
//
//     q  = trunc(RN(|a| / |b|))            in { floor(t), floor(t) + 1 },  t = |a| / |b|   (RN is monotone, q < 2^p exact)
//     r  = fma(-q, |b|, |a|)               exact: |r| < |b| and r is a multiple of the quantum of |b|
//     r += |b|  if r < 0                   exact for the same reason; undoes the rounded-up quotient
//     result = copysign(r, a)              fmod keeps the sign of the dividend, also for a zero result
//
// Everything else (NaN, Inf dividend, zero divisor, huge quotients) takes the library routine.
//
// Shared by host (tests/cpp/test_fastmod.cpp checks it against libm's fmod/fmodf with g++) and device.
#pragma once
#include <cmath>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MNR_FM_HD __host__ __device__ __forceinline__
#else
#define MNR_FM_HD inline
#endif

namespace mnr {

// trunc(q) for 0 <= q < 2^p: add 2^p rounding toward zero, subtract it again (two full-rate adds instead of a
// conversion-pipe round on the device).
MNR_FM_HD double trunc_small(double q) {
#if defined(__CUDA_ARCH__)
    return __dadd_rz(q, 4503599627370496.0) - 4503599627370496.0;
#else
    return trunc(q);
#endif
}
MNR_FM_HD float trunc_small(float q) {
#if defined(__CUDA_ARCH__)
    return __fadd_rz(q, 8388608.0f) - 8388608.0f;
#else
    return truncf(q);
#endif
}
MNR_FM_HD double fmod_lib(double a, double b) { return fmod(a, b); }
MNR_FM_HD float fmod_lib(float a, float b) { return fmodf(a, b); }
MNR_FM_HD double fma_exact(double a, double b, double c) { return fma(a, b, c); }
MNR_FM_HD float fma_exact(float a, float b, float c) { return fmaf(a, b, c); }

template <typename F> struct FastModLimit;
template <> struct FastModLimit<double> { static constexpr double value = 4503599627370496.0; };   // 2^52
template <> struct FastModLimit<float> { static constexpr float value = 8388608.0f; };              // 2^23

template <typename F> MNR_FM_HD F fast_fmod(F a, F b) {
    const F ax = fabs(a), bx = fabs(b);
    if (ax < bx) return a;                             // also b = +-Inf with finite a, and a = +-0 with b != 0
    if (!(ax >= bx)) return fmod_lib(a, b);            // a NaN on either side
    const F t = ax / bx;                               // bx = 0 -> Inf or NaN; ax = Inf -> Inf or NaN
    if (!(t < FastModLimit<F>::value)) return fmod_lib(a, b);
    const F q = trunc_small(t);
    F r = fma_exact(-q, bx, ax);
    if (r < 0) r += bx;
    return copysign(r, a);
}

}  // namespace mnr
