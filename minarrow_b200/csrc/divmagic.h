// Integer division by a column-wide scalar without a divide instruction.
//
// `array / scalar`, `array % scalar` and FloorDiv by a scalar (routing/broadcast.rs:87-112 broadcasts the scalar, then
// int_dense_body / int_masked_body divide element by element, src/kernels/arithmetic/std.rs:54-77) meet the same divisor
// in every row.  GPUs have no integer divider: a 64-bit signed `/` is a ~150-instruction routine, which turns a
// memory-bound pass into a compute-bound one (2.2 TB/s measured for i64).  The divisor is known on the host before the
// launch, so the host computes a multiplicative inverse once (Granlund & Montgomery, "Division by Invariant Integers
// using Multiplication", PLDI 1994, figures 4.1 and 5.1) and the kernel evaluates every quotient with one high multiply,
// two shifts and two adds — exact for every dividend, truncating toward zero like Rust's `/`.
//
// Shared by host (magic computation, unit test: tests/cpp/test_divmagic.cpp builds this header with g++) and device.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MNR_HD __host__ __device__ __forceinline__
#else
#define MNR_HD inline
#endif

namespace mnr {

// Unsigned N-bit: q = (t + ((n - t) >> s1)) >> s2 with t = mulhi(m, n).
// Signed   N-bit: q0 = n + mulhi(m, n);  q0 = (q0 >> s1) - (n >> (N-1));  q = (q0 ^ dsign) - dsign.
struct DivMagic {
    uint64_t m;      // N-bit magic multiplier (two's complement for the signed form), zero-extended
    uint32_t s1, s2;
};

MNR_HD uint32_t mulhi_u(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
MNR_HD int32_t mulhi_s(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * (int64_t)b) >> 32); }
MNR_HD uint64_t mulhi_u(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}
MNR_HD int64_t mulhi_s(int64_t a, int64_t b) {
#if defined(__CUDA_ARCH__)
    return __mul64hi(a, b);
#else
    return (int64_t)(((__int128)a * (__int128)b) >> 64);
#endif
}

// ---- per-row evaluation (W = uint32_t / int32_t / uint64_t / int64_t; narrower columns are widened to 32 bits) ----
template <typename W> MNR_HD W div_by_magic(W n, W d, const DivMagic& k);

template <> MNR_HD uint32_t div_by_magic<uint32_t>(uint32_t n, uint32_t, const DivMagic& k) {
    const uint32_t t = mulhi_u((uint32_t)k.m, n);
    return (t + ((n - t) >> k.s1)) >> k.s2;
}
template <> MNR_HD uint64_t div_by_magic<uint64_t>(uint64_t n, uint64_t, const DivMagic& k) {
    const uint64_t t = mulhi_u(k.m, n);
    return (t + ((n - t) >> k.s1)) >> k.s2;
}
template <> MNR_HD int32_t div_by_magic<int32_t>(int32_t n, int32_t d, const DivMagic& k) {
    const uint32_t q0 = (uint32_t)n + (uint32_t)mulhi_s((int32_t)(uint32_t)k.m, n);
    const uint32_t q1 = (uint32_t)((int32_t)q0 >> k.s1) - (uint32_t)(n >> 31);
    const uint32_t ds = (uint32_t)(d >> 31);
    return (int32_t)((q1 ^ ds) - ds);
}
template <> MNR_HD int64_t div_by_magic<int64_t>(int64_t n, int64_t d, const DivMagic& k) {
    const uint64_t q0 = (uint64_t)n + (uint64_t)mulhi_s((int64_t)k.m, n);
    const uint64_t q1 = (uint64_t)((int64_t)q0 >> k.s1) - (uint64_t)(n >> 63);
    const uint64_t ds = (uint64_t)(d >> 63);
    return (int64_t)((q1 ^ ds) - ds);
}

// ---- host side: the magic numbers (d != 0) ------------------------------------------------------------------------
inline int ceil_log2_u64(uint64_t x) {   // smallest l with 2^l >= x, x >= 1
    return x <= 1 ? 0 : 64 - __builtin_clzll(x - 1);
}
// Unsigned N-bit divisor (N = 32 or 64).  Figure 4.1: l = ceil(log2 d), m' = floor(2^N (2^l - d) / d) + 1.
inline DivMagic div_magic_unsigned(uint64_t d, int nbits) {
    const int l = ceil_log2_u64(d);
    const unsigned __int128 two_l = (unsigned __int128)1 << l;
    const unsigned __int128 m = (((two_l - d) << nbits) / d) + 1;
    DivMagic k;
    k.m = (uint64_t)m & (nbits == 64 ? ~0ull : 0xFFFFFFFFull);
    k.s1 = l < 1 ? l : 1;
    k.s2 = l > 1 ? l - 1 : 0;
    return k;
}
// Signed N-bit divisor.  Figure 5.1: l = max(ceil(log2 |d|), 1), m' = 1 + floor(2^(N+l-1) / |d|) - 2^N.
inline DivMagic div_magic_signed(int64_t d, int nbits) {
    const uint64_t ad = d < 0 ? 0ull - (uint64_t)d : (uint64_t)d;
    int l = ceil_log2_u64(ad);
    if (l < 1) l = 1;
    const unsigned __int128 m = (unsigned __int128)1 + (((unsigned __int128)1 << (nbits + l - 1)) / ad) - ((unsigned __int128)1 << nbits);
    DivMagic k;
    k.m = (uint64_t)m & (nbits == 64 ? ~0ull : 0xFFFFFFFFull);
    k.s1 = (uint32_t)(l - 1);
    k.s2 = 0;
    return k;
}

}  // namespace mnr
