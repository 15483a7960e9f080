// Float64 Power = exp(b * ln a) (src/kernels/arithmetic/std.rs:153-154) without the two libm calls.
//
// The reference evaluates `(b * a.ln()).exp()` through the platform libm; the result is libm-dependent and its own test
// accepts a relative tolerance of 1e-12 (arithmetic/mod.rs:328-340).  CUDA's log() + exp() cost ~120 FP64-pipe instructions
// per row, which made `f64 ** f64` the one FP64-bound kernel of the path (4.1 TB/s, 0.64 of the copy peak; a B200 SM issues
// ~70 FP64 operations per clock).  For the ordinary case — a normal positive finite base and |b ln a| < 700, so the result is a
// normal double — the same expression is evaluated here with ~50 FP64 operations:
//   ln a   fdlibm's e_log.c scheme (argument reduced to [sqrt(2)/2, sqrt(2)), s = f / (2 + f), degree-14 odd polynomial in s,
//          k ln2 added in two pieces), with the division replaced by MUFU.RCP64H + two Newton steps (relative error ~2^-52);
//   b * x  one rounded multiply, exactly as the reference;
//   exp    k = rint(x log2 e) by the 1.5 * 2^52 trick, r = x - k ln2 in two pieces, degree-12 Taylor polynomial (|r| <= 0.347:
//          truncation 2e-16), scaled by 2^k through the exponent field.
// Zero / negative / subnormal / Inf / NaN bases, NaN exponents and results outside the double range are folded in without a
// branch (selects on ln a, clamps on b ln a, a two-factor 2^k), giving what libm's exp(b * log(a)) gives.  Measured against exp(b * log(a)) in long double over 2e7 random
// pairs (tests/cpp/test_fastpow.cpp): relative difference <= 3e-14 for |b ln a| <= 100 and <= 2.3e-13 up to 700 (a ~1-ulp
// ln a is multiplied by b: the difference grows like |b ln a| * 2^-52, as it does between any two libms) — inside the
// 1e-12 bar with a factor of 4 to spare in the worst corner and 35 where the parity tests live (|b ln a| <= 74).
//
// Shared by device and host (the unit test builds this header with g++ and a software stand-in for MUFU.RCP64H).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MNR_FP_HD __host__ __device__ __forceinline__
#else
#define MNR_FP_HD inline
#endif

namespace mnr {

MNR_FP_HD double fp_from_bits(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
MNR_FP_HD uint64_t fp_bits(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }

// ~20-bit reciprocal seed: MUFU.RCP64H on the device, a float division on the host (same precision class).
MNR_FP_HD double fp_rcp_seed(double d) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    return r;
#else
    return (double)(1.0f / (float)d);
#endif
}

MNR_FP_HD double fast_pow_f64(double a, double b) {
    // Branch-free: every lane of a warp runs the same ~55 FP64 operations whatever its operands are (a column of mixed-sign
    // bases would otherwise send half of every warp through a second code path).  Non-ordinary bases are folded in with
    // selects on ln a; the rest is IEEE arithmetic on Inf / NaN.
    const uint64_t ia0 = fp_bits(a);
    const bool sub = (ia0 >> 52) == 0 && ia0 != 0;                 // positive subnormal: scale into the normal range
    const double as = sub ? a * 18014398509481984.0 : a;           // 2^54
    const uint64_t ia = fp_bits(as);
    const uint32_t hx = (uint32_t)(ia >> 32);
    // ---- ln a (fdlibm e_log.c) ----
    int k = (int)(hx >> 20) - 1023 - (sub ? 54 : 0);
    const uint32_t hm = hx & 0x000fffffu;
    const uint32_t i = (hm + 0x95f64u) & 0x100000u;
    const double m = fp_from_bits(((uint64_t)(hm | (i ^ 0x3ff00000u)) << 32) | (ia & 0xffffffffull));   // [sqrt(2)/2, sqrt(2))
    k += (int)(i >> 20);
    const double f = m - 1.0;
    const double d = 2.0 + f;
    double rc = fp_rcp_seed(d);
    double e = fma(-d, rc, 1.0);
    rc = fma(rc, e, rc);
    e = fma(-d, rc, 1.0);
    rc = fma(rc, e, rc);
    const double s = f * rc;
    const double dk = (double)k;
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01), 2.857142874366239149e-01), 6.666666666666735130e-01);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    double lg = dk * 6.93147180369123816490e-01 - ((hfsq - (s * (hfsq + R) + dk * 1.90821492927058770002e-10)) - f);
    // ln of the non-ordinary bases, exactly what libm returns: ln(+-0) = -Inf, ln(+Inf) = +Inf, ln(negative) = ln(NaN) = NaN
    const double inf = fp_from_bits(0x7ff0000000000000ull);
    if (a == 0.0) lg = -inf;
    if (ia0 == 0x7ff0000000000000ull) lg = inf;
    if (!(a >= 0.0)) lg = fp_from_bits(0x7ff8000000000000ull);
    double x = b * lg;                                              // one rounded multiply, as in the reference
    // ---- exp x ---- (x = NaN stays NaN through the polynomial; beyond the clamps the result is +Inf / 0 either way)
    x = x < -746.0 ? -746.0 : x;
    x = x > 710.0 ? 710.0 : x;
    const double kd = fma(x, 1.4426950408889634074, 6755399441055744.0);
    const int n = (int)(uint32_t)fp_bits(kd);
    const double kn = kd - 6755399441055744.0;
    double r = fma(-kn, 6.93147180369123816490e-01, x);
    r = fma(-kn, 1.90821492927058770002e-10, r);
    double p = 2.08767569878680989792e-09;                        // 1/12!
    p = fma(p, r, 2.50521083854417187751e-08);                    // 1/11!
    p = fma(p, r, 2.75573192239858906526e-07);                    // 1/10!
    p = fma(p, r, 2.75573192239858906526e-06);                    // 1/9!
    p = fma(p, r, 2.48015873015873015873e-05);                    // 1/8!
    p = fma(p, r, 1.98412698412698412698e-04);                    // 1/7!
    p = fma(p, r, 1.38888888888888888889e-03);                    // 1/6!
    p = fma(p, r, 8.33333333333333333333e-03);                    // 1/5!
    p = fma(p, r, 4.16666666666666666667e-02);                    // 1/4!
    p = fma(p, r, 1.66666666666666666667e-01);                    // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    // 2^n in two factors so that results near the ends of the range (and subnormal ones) scale without a spurious overflow
    const int n1 = n >> 1, n2 = n - n1;
    return p * fp_from_bits((uint64_t)(int64_t)(n1 + 1023) << 52) * fp_from_bits((uint64_t)(int64_t)(n2 + 1023) << 52);
}

}  // namespace mnr
