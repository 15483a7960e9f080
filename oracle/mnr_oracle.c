/*
 * mnr_oracle.c — CPU restatement of Minarrow's columnar hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product (minarrow_b200/) never
 * links, imports or calls anything in oracle/.
 *
 * Every function restates one reference function and cites it (paths relative to the reference
 * checkout, pbower/minarrow v0.10.1).  The reference is Rust (nightly, portable_simd) and cannot be
 * compiled in this image (no cargo/rustc), so this is a "port" oracle.  It is pinned against the
 * reference's own in-tree known-answer tests (tests/golden/reference_kats.json, transcribed with
 * file:line) — see tests/test_oracle_golden.py.  Pieces of the path whose arithmetic lives in
 * third-party code that is absent from the reference tree are marked "parity unpinned" below:
 *   - core::simd (Rust std, nightly-2026-04-17): lane wrap, MIN/-1 guard, reduce_sum order.
 *   - libm ln/exp used by Power (tolerance-only in the reference's own tests).
 *   - min/max/avg/count reductions (live in the downstream `simd-kernels` crate, not in-tree):
 *     defined in DESIGN.md, cross-checked against pyarrow.compute in tests.
 *
 * Build: gcc -O2 -fwrapv -fno-strict-aliasing -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off: the reference never fuses a*b+c outside the dedicated FMA entry points
 * (src/kernels/arithmetic/simd.rs:401-409 vs :620).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_ERR_LENGTH_MISMATCH (-2)
#define ORC_ERR_DIVIDE_BY_ZERO (-10)

/* ArithmeticOperator discriminants, src/enums/operators.rs:19-48 (declaration order). */
enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_DIV = 3, OP_REM = 4, OP_POW = 5, OP_FLOORDIV = 6 };
/* LogicalOperator, src/enums/operators.rs:88-104. */
enum { LOP_AND = 0, LOP_OR = 1, LOP_XOR = 2 };

/* ---- Bitmask primitives: src/structs/bitmask.rs ------------------------------------------- */

/* Bitmask::get_unchecked, bitmask.rs:745-748: bit i = byte i>>3, bit i&7 (LSB first). */
static inline int bit_get(const uint8_t *bits, size_t i) { return (bits[i >> 3] >> (i & 7)) & 1; }
/* Bitmask::set_unchecked, bitmask.rs:248-258. */
static inline void bit_set(uint8_t *bits, size_t i, int v) {
    uint8_t b = (uint8_t)(1u << (i & 7));
    if (v) bits[i >> 3] |= b; else bits[i >> 3] &= (uint8_t)~b;
}
/* Bitmask::mask_trailing_bits, bitmask.rs:83-90 == kernels/bitmask/mod.rs:141-150. */
static inline void clear_trailing(uint8_t *bits, size_t len) {
    if (len == 0 || (len & 7) == 0) return;
    bits[(len + 7) / 8 - 1] &= (uint8_t)((1u << (len & 7)) - 1);
}
size_t orc_mask_bytes(size_t len) { return (len + 7) / 8; }

/* Bitmask::new_set_all, bitmask.rs:94-105. */
void orc_bits_new_set_all(uint8_t *out, size_t len, int set) {
    memset(out, set ? 0xFF : 0, (len + 7) / 8);
    clear_trailing(out, len);
}

/* Bitmask::count_ones, bitmask.rs:393-406 (byte-wise, last byte masked). */
uint64_t orc_bits_count_ones(const uint8_t *bits, size_t len) {
    size_t full = len / 8;
    uint64_t c = 0;
    for (size_t i = 0; i < full; ++i) c += (uint64_t)__builtin_popcount(bits[i]);
    size_t rem = len & 7;
    if (rem) c += (uint64_t)__builtin_popcount(bits[full] & ((1u << rem) - 1));
    return c;
}
/* Bitmask::count_zeros / null_count, bitmask.rs:409-417. */
uint64_t orc_bits_null_count(const uint8_t *bits, size_t len) { return len - orc_bits_count_ones(bits, len); }

/* Bitmask::union (bitwise OR), bitmask.rs:661-669; intersect (AND) :673-681; invert :685-692. */
void orc_bits_union(const uint8_t *a, const uint8_t *b, size_t len, uint8_t *out) {
    size_t nb = (len + 7) / 8;
    for (size_t i = 0; i < nb; ++i) out[i] = a[i] | b[i];
    clear_trailing(out, len);
}
void orc_bits_intersect(const uint8_t *a, const uint8_t *b, size_t len, uint8_t *out) {
    size_t nb = (len + 7) / 8;
    for (size_t i = 0; i < nb; ++i) out[i] = a[i] & b[i];
    clear_trailing(out, len);
}
void orc_bits_invert(const uint8_t *a, size_t len, uint8_t *out) {
    size_t nb = (len + 7) / 8;
    for (size_t i = 0; i < nb; ++i) out[i] = (uint8_t)~a[i];
    clear_trailing(out, len);
}

/* merge_bitmasks_to_new, src/kernels/bitmask/mod.rs:171-197: per-row AND; a missing side is all-valid.
 * Returns 1 if an output mask exists (either input present), 0 for (None, None). */
int orc_merge_bitmasks_to_new(const uint8_t *l, const uint8_t *r, size_t len, uint8_t *out) {
    if (!l && !r) return 0;
    orc_bits_new_set_all(out, len, 1);
    for (size_t i = 0; i < len; ++i) {
        int v = (l ? bit_get(l, i) : 1) && (r ? bit_get(r, i) : 1);
        bit_set(out, i, v);
    }
    clear_trailing(out, len);
    return 1;
}

/* Bitmask::slice_clone, bitmask.rs:604-626 (true bit-granular window copy). */
void orc_bits_slice_clone(const uint8_t *src, size_t offset, size_t len, uint8_t *out) {
    orc_bits_new_set_all(out, len, 0);
    for (size_t i = 0; i < len; ++i) if (bit_get(src, offset + i)) bit_set(out, i, 1);
    clear_trailing(out, len);
}

/* ---- Bitmask kernels: src/kernels/bitmask/{simd,std}.rs ------------------------------------ */

/* bitmask_binop_simd, bitmask/simd.rs:95-139 == bitmask_binop_std, std.rs:73-94.
 * The window starts at BYTE offset/8 of each operand (bitmask_window_bytes, mod.rs:124-128: sub-byte
 * offsets are floored) and is processed as ceil(len/64) u64 words; the result keeps ceil(len/8) bytes
 * with the slack bits of the last byte cleared (clear_trailing_bits).  Restated byte-wise, which is
 * the same function of the first ceil(len/8) window bytes. */
void orc_bitmask_binop(const uint8_t *lhs, size_t lhs_off, const uint8_t *rhs, size_t rhs_off,
                       size_t len, int op, uint8_t *out) {
    if (len == 0) return;
    const uint8_t *lp = lhs + lhs_off / 8, *rp = rhs + rhs_off / 8;
    size_t nb = (len + 7) / 8;
    for (size_t i = 0; i < nb; ++i) {
        uint8_t a = lp[i], b = rp[i];
        out[i] = op == LOP_AND ? (a & b) : op == LOP_OR ? (a | b) : (a ^ b);
    }
    clear_trailing(out, len);
}
/* bitmask_unop_simd (NOT), bitmask/simd.rs:169-203 == std.rs:97-114. */
void orc_bitmask_not(const uint8_t *src, size_t off, size_t len, uint8_t *out) {
    if (len == 0) return;
    const uint8_t *sp = src + off / 8;
    size_t nb = (len + 7) / 8;
    for (size_t i = 0; i < nb; ++i) out[i] = (uint8_t)~sp[i];
    clear_trailing(out, len);
}

/* Little-endian u64 word k of a byte buffer, reading only bytes < nbytes (others read as 0).  The
 * reference casts the byte pointer to *const u64 (bitmask.rs:266-268); bytes past ceil(len/8) are
 * either zero padding or masked off by every caller below. */
static inline uint64_t word_at(const uint8_t *bits, size_t nbytes, size_t k) {
    uint64_t w = 0;
    size_t base = k * 8;
    for (size_t j = 0; j < 8 && base + j < nbytes; ++j) w |= (uint64_t)bits[base + j] << (8 * j);
    return w;
}

/* popcount_mask_simd, bitmask/simd.rs:596-645 == popcount_mask, std.rs:279-296.
 * Window = words [offset/64, offset/64 + ceil(len/64)); last word masked to len%64 bits.
 * `mask_len` is the logical bit length of the whole mask (bounds for the byte reads). */
uint64_t orc_popcount_mask(const uint8_t *bits, size_t mask_len, size_t offset, size_t len) {
    if (len == 0) return 0;
    size_t nbytes = (mask_len + 7) / 8;
    size_t n_words = (len + 63) / 64, ws = offset / 64;
    uint64_t acc = 0;
    for (size_t k = 0; k < n_words; ++k) {
        uint64_t w = word_at(bits, nbytes, ws + k);
        if (k == n_words - 1 && (len % 64) != 0) w &= ((uint64_t)1 << (len % 64)) - 1;
        acc += (uint64_t)__builtin_popcountll(w);
    }
    return acc;
}

/* all_true_mask_simd, bitmask/simd.rs:648-692 == all_true_mask, std.rs:300-332 == Bitmask::all_true :311-326. */
int orc_all_true_mask(const uint8_t *bits, size_t len) {
    if (len == 0) return 1;
    return orc_bits_count_ones(bits, len) == len;
}
/* all_false_mask_simd, bitmask/simd.rs:695-736. */
int orc_all_false_mask(const uint8_t *bits, size_t len) {
    if (len == 0) return 1;
    return orc_bits_count_ones(bits, len) == 0;
}

/* eq_mask_simd, bitmask/simd.rs:402-450: out = !(a ^ b) over word-aligned windows, trailing bits
 * cleared.  Offsets must be multiples of 64 (the reference panics otherwise): returns -1 then. */
int orc_eq_mask(const uint8_t *a, size_t a_len, size_t ao, const uint8_t *b, size_t b_len, size_t bo,
                size_t len, uint8_t *out) {
    if (len == 0) return 0;
    if (ao % 64 || bo % 64) return -1;
    size_t an = (a_len + 7) / 8, bn = (b_len + 7) / 8, nb = (len + 7) / 8;
    size_t n_words = (len + 63) / 64;
    for (size_t k = 0; k < n_words; ++k) {
        uint64_t w = ~(word_at(a, an, ao / 64 + k) ^ word_at(b, bn, bo / 64 + k));
        for (size_t j = 0; j < 8 && k * 8 + j < nb; ++j) out[k * 8 + j] = (uint8_t)(w >> (8 * j));
    }
    clear_trailing(out, len);
    return 0;
}
/* ne_mask_simd, bitmask/simd.rs:468-472: !eq_mask (Bitmask Not = invert + mask_trailing_bits). */
int orc_ne_mask(const uint8_t *a, size_t a_len, size_t ao, const uint8_t *b, size_t b_len, size_t bo,
                size_t len, uint8_t *out) {
    int rc = orc_eq_mask(a, a_len, ao, b, b_len, bo, len, out);
    if (rc) return rc;
    size_t nb = (len + 7) / 8;
    for (size_t i = 0; i < nb; ++i) out[i] = (uint8_t)~out[i];
    clear_trailing(out, len);
    return 0;
}
/* all_eq_mask_simd, bitmask/simd.rs:511-581: equality of the logical bits of two word-aligned windows. */
int orc_all_eq_mask(const uint8_t *a, size_t a_len, size_t ao, const uint8_t *b, size_t b_len, size_t bo,
                    size_t len) {
    if (len == 0) return 1;
    size_t an = (a_len + 7) / 8, bn = (b_len + 7) / 8;
    size_t n_words = (len + 63) / 64, trailing = len & 63;
    for (size_t k = 0; k < n_words; ++k) {
        uint64_t wa = word_at(a, an, ao / 64 + k), wb = word_at(b, bn, bo / 64 + k);
        if (k == n_words - 1 && trailing) {
            uint64_t m = ((uint64_t)1 << trailing) - 1;
            wa &= m; wb &= m;
        }
        if (wa != wb) return 0;
    }
    return 1;
}
/* all_ne_mask_simd, bitmask/simd.rs:490-494: literally !all_eq (NOT "every bit differs"); restated as is. */
int orc_all_ne_mask(const uint8_t *a, size_t a_len, size_t ao, const uint8_t *b, size_t b_len, size_t bo,
                    size_t len) {
    return !orc_all_eq_mask(a, a_len, ao, b, b_len, bo, len);
}
/* in_mask_simd, bitmask/simd.rs:327-375: boolean set membership.  rhs scanned from word rhs_off/64. */
void orc_in_mask(const uint8_t *lhs, size_t lhs_off, const uint8_t *rhs, size_t rhs_len, size_t rhs_off,
                 size_t len, uint8_t *out) {
    if (len == 0) return;
    size_t rn = (rhs_len + 7) / 8, n_words = (len + 63) / 64, trailing = len & 63;
    uint64_t any_set = 0, any_unset = 0;
    for (size_t k = 0; k < n_words; ++k) {
        uint64_t w = word_at(rhs, rn, rhs_off / 64 + k);
        if (k == n_words - 1 && trailing) {
            uint64_t vm = ((uint64_t)1 << trailing) - 1;
            w &= vm; any_set |= w; any_unset |= (~w) & vm;
        } else { any_set |= w; any_unset |= ~w; }
        if (any_set && any_unset) break;
    }
    if (any_set && any_unset) orc_bits_new_set_all(out, len, 1);
    else if (any_set) orc_bits_slice_clone(lhs, lhs_off, len, out);
    else if (any_unset) orc_bitmask_not(lhs, lhs_off, len, out);
    else orc_bits_new_set_all(out, len, 0);
}
/* not_in_mask_simd, bitmask/simd.rs:393-398. */
void orc_not_in_mask(const uint8_t *lhs, size_t lhs_off, const uint8_t *rhs, size_t rhs_len, size_t rhs_off,
                     size_t len, uint8_t *out) {
    if (len == 0) return;
    orc_in_mask(lhs, lhs_off, rhs, rhs_len, rhs_off, len, out);
    size_t nb = (len + 7) / 8;
    for (size_t i = 0; i < nb; ++i) out[i] = (uint8_t)~out[i];
    clear_trailing(out, len);
}

/* ---- Integer element-wise: src/kernels/arithmetic/{std,simd}.rs ----------------------------- */

/* rhs.to_u32().unwrap_or(0), std.rs:67,116 / simd.rs:96: negative or > u32::MAX exponent => 0. */
#define EXP_U32_SIGNED(x) (((x) < 0 || (uint64_t)(x) > 0xFFFFFFFFull) ? 0u : (uint32_t)(x))
#define EXP_U32_UNSIGNED(x) (((uint64_t)(x) > 0xFFFFFFFFull) ? 0u : (uint32_t)(x))

/* Wrapping integer power: repeated wrapping_mul (simd.rs:94-100) == PrimInt::pow with release-mode
 * wrap (std.rs:67).  Square-and-multiply gives the same residue mod 2^bits. */
#define DEF_IPOW(NAME, UT)                                                      \
    static inline UT NAME(UT base, uint32_t e) {                                \
        UT acc = 1;                                                             \
        while (e) { if (e & 1) acc = (UT)(acc * base); base = (UT)(base * base); e >>= 1; } \
        return acc;                                                             \
    }
DEF_IPOW(ipow_u8, uint8_t) DEF_IPOW(ipow_u16, uint16_t) DEF_IPOW(ipow_u32, uint32_t) DEF_IPOW(ipow_u64, uint64_t)

/* One integer element.  `*ok` = 0 when the divisor is zero under Div/Rem/FloorDiv.
 * Add/Sub/Mul: two's-complement wrap (std.rs:50-52, simd.rs:72-74).
 * Div/Rem: truncate toward zero, remainder has the dividend's sign (Rust `/`, `%`).
 * ASSUMPTION (parity unpinned, core::simd): MIN / -1 = MIN and MIN % -1 = 0 (the SIMD body swaps the
 * divisor for 1; the reference's scalar tails would panic on overflow).
 * FloorDiv: std.rs:68-77. */
#define DEF_INT_ELEM(T, UT, SIGNED, TMIN, IPOW, EXPOF)                                         \
    static inline T elem_##T(int op, T l, T r, int *ok) {                                      \
        *ok = 1;                                                                               \
        switch (op) {                                                                          \
        case OP_ADD: return (T)((UT)l + (UT)r);                                                \
        case OP_SUB: return (T)((UT)l - (UT)r);                                                \
        case OP_MUL: return (T)((UT)l * (UT)r);                                                \
        case OP_DIV:                                                                           \
            if (r == 0) { *ok = 0; return 0; }                                                 \
            if (SIGNED && l == (T)(TMIN) && r == (T)-1) return l;                              \
            return (T)(l / r);                                                                 \
        case OP_REM:                                                                           \
            if (r == 0) { *ok = 0; return 0; }                                                 \
            if (SIGNED && l == (T)(TMIN) && r == (T)-1) return 0;                              \
            return (T)(l % r);                                                                 \
        case OP_POW: return (T)IPOW((UT)l, EXPOF(r));                                          \
        case OP_FLOORDIV: {                                                                    \
            if (r == 0) { *ok = 0; return 0; }                                                 \
            if (SIGNED && l == (T)(TMIN) && r == (T)-1) return l;                              \
            T d = (T)(l / r), m = (T)(l % r);                                                  \
            if (SIGNED && m != 0 && ((l ^ r) < 0)) return (T)((UT)d - 1);                      \
            return d;                                                                          \
        }                                                                                      \
        }                                                                                      \
        return 0;                                                                              \
    }

typedef int8_t i8; typedef uint8_t u8; typedef int16_t i16; typedef uint16_t u16;
typedef int32_t i32; typedef uint32_t u32; typedef int64_t i64; typedef uint64_t u64;
typedef float f32; typedef double f64;

DEF_INT_ELEM(i8, uint8_t, 1, INT8_MIN, ipow_u8, EXP_U32_SIGNED)
DEF_INT_ELEM(u8, uint8_t, 0, 0, ipow_u8, EXP_U32_UNSIGNED)
DEF_INT_ELEM(i16, uint16_t, 1, INT16_MIN, ipow_u16, EXP_U32_SIGNED)
DEF_INT_ELEM(u16, uint16_t, 0, 0, ipow_u16, EXP_U32_UNSIGNED)
DEF_INT_ELEM(i32, uint32_t, 1, INT32_MIN, ipow_u32, EXP_U32_SIGNED)
DEF_INT_ELEM(u32, uint32_t, 0, 0, ipow_u32, EXP_U32_UNSIGNED)
DEF_INT_ELEM(i64, uint64_t, 1, INT64_MIN, ipow_u64, EXP_U32_SIGNED)
DEF_INT_ELEM(u64, uint64_t, 0, 0, ipow_u64, EXP_U32_UNSIGNED)

/* apply_int_<T>, src/kernels/arithmetic/dispatch.rs:65-133 (instantiated :376-387).
 *   mask == NULL  -> int_dense_body (std.rs:41-80 / simd.rs:52-113): a zero divisor "panics" =>
 *                    ORC_ERR_DIVIDE_BY_ZERO, out contents unspecified; no output mask.
 *   mask != NULL  -> int_masked_body (std.rs:86-138 / simd.rs:118-370): invalid row => value 0,
 *                    validity 0; valid row with zero divisor under Div/Rem/FloorDiv => value 0,
 *                    validity 0; out_mask has `n` bits, slack bits zero (new_set_all, dispatch.rs:92).
 * lhs_n != rhs_n => LengthMismatch (confirm_equal_len, utils.rs:163-171). */
#define DEF_APPLY_INT(T)                                                                        \
    int orc_apply_int_##T(const T *lhs, size_t lhs_n, const T *rhs, size_t rhs_n, int op,        \
                          const uint8_t *mask, T *out, uint8_t *out_mask) {                     \
        if (lhs_n != rhs_n) return ORC_ERR_LENGTH_MISMATCH;                                     \
        size_t n = lhs_n;                                                                       \
        if (!mask) {                                                                            \
            for (size_t i = 0; i < n; ++i) {                                                    \
                int ok; out[i] = elem_##T(op, lhs[i], rhs[i], &ok);                             \
                if (!ok) return ORC_ERR_DIVIDE_BY_ZERO;                                         \
            }                                                                                   \
            return ORC_OK;                                                                      \
        }                                                                                       \
        orc_bits_new_set_all(out_mask, n, 1);                                                   \
        for (size_t i = 0; i < n; ++i) {                                                        \
            if (bit_get(mask, i)) {                                                             \
                int ok; out[i] = elem_##T(op, lhs[i], rhs[i], &ok);                             \
                bit_set(out_mask, i, ok);                                                       \
            } else { out[i] = 0; bit_set(out_mask, i, 0); }                                     \
        }                                                                                       \
        return ORC_OK;                                                                          \
    }
DEF_APPLY_INT(i8) DEF_APPLY_INT(u8) DEF_APPLY_INT(i16) DEF_APPLY_INT(u16)
DEF_APPLY_INT(i32) DEF_APPLY_INT(u32) DEF_APPLY_INT(i64) DEF_APPLY_INT(u64)

/* ---- Float element-wise ------------------------------------------------------------------- */

/* float_dense_body_std, std.rs:144-157 (same expressions as simd.rs:401-409,564-572).
 * Rem = Rust `%` on floats = C fmod.  Power = exp(b * ln(a)) through libm (tolerance-only parity:
 * arithmetic/mod.rs:328-340).  FloorDiv = floor(a / b). */
static inline f64 elem_f64(int op, f64 a, f64 b) {
    switch (op) {
    case OP_ADD: return a + b;
    case OP_SUB: return a - b;
    case OP_MUL: return a * b;
    case OP_DIV: return a / b;
    case OP_REM: return fmod(a, b);
    case OP_POW: return exp(b * log(a));
    case OP_FLOORDIV: return floor(a / b);
    }
    return 0.0;
}
static inline f32 elem_f32(int op, f32 a, f32 b) {
    switch (op) {
    case OP_ADD: return a + b;
    case OP_SUB: return a - b;
    case OP_MUL: return a * b;
    case OP_DIV: return a / b;
    case OP_REM: return fmodf(a, b);
    case OP_POW: return expf(b * logf(a));
    case OP_FLOORDIV: return floorf(a / b);
    }
    return 0.0f;
}

/* apply_float_<T>, dispatch.rs:138-206.  mask == NULL -> dense, no output mask.  mask != NULL ->
 * float_masked_body (std.rs:163-194 / simd.rs:376-505): invalid => +0.0 and validity 0, otherwise the
 * IEEE result stays valid (Inf/NaN included); out_mask == first n bits of the input mask. */
#define DEF_APPLY_FLOAT(T)                                                                      \
    int orc_apply_float_##T(const T *lhs, size_t lhs_n, const T *rhs, size_t rhs_n, int op,      \
                            const uint8_t *mask, T *out, uint8_t *out_mask) {                   \
        if (lhs_n != rhs_n) return ORC_ERR_LENGTH_MISMATCH;                                     \
        size_t n = lhs_n;                                                                       \
        if (!mask) { for (size_t i = 0; i < n; ++i) out[i] = elem_##T(op, lhs[i], rhs[i]); return ORC_OK; } \
        orc_bits_new_set_all(out_mask, n, 1);                                                   \
        for (size_t i = 0; i < n; ++i) {                                                        \
            if (bit_get(mask, i)) { out[i] = elem_##T(op, lhs[i], rhs[i]); bit_set(out_mask, i, 1); } \
            else { out[i] = (T)0; bit_set(out_mask, i, 0); }                                    \
        }                                                                                       \
        return ORC_OK;                                                                          \
    }
DEF_APPLY_FLOAT(f32) DEF_APPLY_FLOAT(f64)

/* apply_fma_<T>, dispatch.rs:211-290; bodies simd.rs:594-751 / std.rs:198-230: a.mul_add(b, c), one
 * rounding.  (The unaligned fallback at dispatch.rs:266,280 uses unfused a*b+c; Vec64 inputs are
 * always aligned, so the fused form is canonical.  `fused` = 0 selects the fallback expression.) */
#define DEF_APPLY_FMA(T, FMA)                                                                   \
    int orc_apply_fma_##T(const T *lhs, size_t lhs_n, const T *rhs, size_t rhs_n, const T *acc,  \
                          size_t acc_n, const uint8_t *mask, int fused, T *out, uint8_t *out_mask) { \
        if (lhs_n != rhs_n || lhs_n != acc_n) return ORC_ERR_LENGTH_MISMATCH;                   \
        size_t n = lhs_n;                                                                       \
        if (mask) orc_bits_new_set_all(out_mask, n, 1);                                         \
        for (size_t i = 0; i < n; ++i) {                                                        \
            if (!mask || bit_get(mask, i)) {                                                    \
                if (fused) out[i] = FMA(lhs[i], rhs[i], acc[i]);                                \
                else { volatile T p = lhs[i] * rhs[i]; out[i] = p + acc[i]; }                   \
            } else { out[i] = (T)0; bit_set(out_mask, i, 0); }                                  \
        }                                                                                       \
        return ORC_OK;                                                                          \
    }
DEF_APPLY_FMA(f32, fmaf) DEF_APPLY_FMA(f64, fma)

/* ---- Routing helpers: src/kernels/routing ---------------------------------------------------- */

/* broadcast_length_1_array, routing/broadcast.rs:25-47: vec64![a.data[0]; len]. */
#define DEF_FILL(T) void orc_broadcast_fill_##T(T v, size_t len, T *out) { for (size_t i = 0; i < len; ++i) out[i] = v; }
DEF_FILL(i32) DEF_FILL(u32) DEF_FILL(i64) DEF_FILL(u64) DEF_FILL(f32) DEF_FILL(f64)

/* int -> float promotion, routing/arithmetic.rs:244-269: `x as f64` / `x as f32` element-wise. */
void orc_cast_i32_f64(const i32 *in, size_t n, f64 *out) { for (size_t i = 0; i < n; ++i) out[i] = (f64)in[i]; }
void orc_cast_i32_f32(const i32 *in, size_t n, f32 *out) { for (size_t i = 0; i < n; ++i) out[i] = (f32)in[i]; }

/* ---- Sums: benches/benchmark_parallel_simd.rs, benches/hotloop_benchmark_simd.rs -------------- */

/* simd_sum_i64<LANES>, benchmark_parallel_simd.rs:44-60: LANES strided accumulators, reduce_sum,
 * scalar tail.  i64 addition wraps (release build), so any order gives the same bits. */
i64 orc_simd_sum_i64(const i64 *d, size_t n, int lanes) {
    u64 acc[64] = {0};
    size_t chunks = n / (size_t)lanes;
    for (size_t i = 0; i < chunks; ++i)
        for (int j = 0; j < lanes; ++j) acc[j] += (u64)d[i * (size_t)lanes + (size_t)j];
    u64 r = 0;
    for (int j = 0; j < lanes; ++j) r += acc[j];
    for (size_t i = chunks * (size_t)lanes; i < n; ++i) r += (u64)d[i];
    return (i64)r;
}
/* simd_sum_f64<LANES>, benchmark_parallel_simd.rs:63-79.  reduce_sum order (core::simd, parity
 * unpinned): lanes added sequentially in lane order starting from lane 0. */
f64 orc_simd_sum_f64(const f64 *d, size_t n, int lanes) {
    f64 acc[64] = {0};
    size_t chunks = n / (size_t)lanes;
    for (size_t i = 0; i < chunks; ++i)
        for (int j = 0; j < lanes; ++j) acc[j] += d[i * (size_t)lanes + (size_t)j];
    f64 r = acc[0];
    for (int j = 1; j < lanes; ++j) r += acc[j];
    for (size_t i = chunks * (size_t)lanes; i < n; ++i) r += d[i];
    return r;
}
/* simd_sum_i64 4x-unrolled, hotloop_benchmark_simd.rs:56-114. */
i64 orc_hotloop_sum_i64(const i64 *d, size_t n, int lanes) {
    u64 a1[64] = {0}, a2[64] = {0}, a3[64] = {0}, a4[64] = {0}, acc[64];
    size_t L = (size_t)lanes, chunks = n / L, unrolled = chunks / 4;
    for (size_t i = 0; i < unrolled; ++i) {
        size_t base = i * 4 * L;
        for (size_t j = 0; j < L; ++j) {
            a1[j] += (u64)d[base + j]; a2[j] += (u64)d[base + L + j];
            a3[j] += (u64)d[base + 2 * L + j]; a4[j] += (u64)d[base + 3 * L + j];
        }
    }
    for (size_t j = 0; j < L; ++j) acc[j] = a1[j] + a2[j] + a3[j] + a4[j];
    for (size_t i = unrolled * 4; i < chunks; ++i)
        for (size_t j = 0; j < L; ++j) acc[j] += (u64)d[i * L + j];
    u64 r = 0;
    for (size_t j = 0; j < L; ++j) r += acc[j];
    for (size_t i = chunks * L; i < n; ++i) r += (u64)d[i];
    return (i64)r;
}
/* simd_sum_f64 4x-unrolled, hotloop_benchmark_simd.rs:117-174: ((acc1+acc2)+acc3)+acc4 per lane,
 * leftover chunks, lanes added sequentially from 0.0, then the tail. */
f64 orc_hotloop_sum_f64(const f64 *d, size_t n, int lanes) {
    f64 a1[64] = {0}, a2[64] = {0}, a3[64] = {0}, a4[64] = {0}, acc[64];
    size_t L = (size_t)lanes, chunks = n / L, unrolled = chunks / 4;
    for (size_t i = 0; i < unrolled; ++i) {
        size_t base = i * 4 * L;
        for (size_t j = 0; j < L; ++j) {
            a1[j] += d[base + j]; a2[j] += d[base + L + j];
            a3[j] += d[base + 2 * L + j]; a4[j] += d[base + 3 * L + j];
        }
    }
    for (size_t j = 0; j < L; ++j) acc[j] = ((a1[j] + a2[j]) + a3[j]) + a4[j];
    for (size_t i = unrolled * 4; i < chunks; ++i)
        for (size_t j = 0; j < L; ++j) acc[j] += d[i * L + j];
    f64 r = 0.0;
    for (size_t j = 0; j < L; ++j) r += acc[j];
    for (size_t i = chunks * L; i < n; ++i) r += d[i];
    return r;
}

/* rayon_simd_sum_i64 / _f64, benchmark_parallel_simd.rs:81-97: par_chunks(1<<20).map(simd_sum).sum().
 * rayon's combine tree depends on work stealing; the oracle adds chunk partials in chunk order (one
 * valid schedule).  `threads` > 1 runs the per-chunk sums under OpenMP (dynamic schedule, like rayon's
 * work stealing) and still combines in chunk order, so the value is thread-count independent. */
#define PAR_CHUNK ((size_t)1 << 20)
i64 orc_rayon_simd_sum_i64(const i64 *d, size_t n, int lanes, int threads) {
    size_t nchunks = (n + PAR_CHUNK - 1) / PAR_CHUNK;
    u64 total = 0;
    (void)threads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads > 0 ? threads : 1) reduction(+ : total)
    for (size_t c = 0; c < nchunks; ++c) {
        size_t lo = c * PAR_CHUNK, hi = lo + PAR_CHUNK < n ? lo + PAR_CHUNK : n;
        total += (u64)orc_simd_sum_i64(d + lo, hi - lo, lanes);
    }
    return (i64)total;
}
f64 orc_rayon_simd_sum_f64(const f64 *d, size_t n, int lanes, int threads, f64 *partials /* nchunks */) {
    size_t nchunks = (n + PAR_CHUNK - 1) / PAR_CHUNK;
    (void)threads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads > 0 ? threads : 1)
    for (size_t c = 0; c < nchunks; ++c) {
        size_t lo = c * PAR_CHUNK, hi = lo + PAR_CHUNK < n ? lo + PAR_CHUNK : n;
        partials[c] = orc_simd_sum_f64(d + lo, hi - lo, lanes);
    }
    f64 total = 0.0;
    for (size_t c = 0; c < nchunks; ++c) total += partials[c];
    return total;
}

/* ---- Null-aware aggregates (DESIGN.md "A.6"; parity unpinned — defined by this project) -------- */
/* The reference has no masked reduction in-tree (they live in the downstream simd-kernels crate).
 * Definition: invalid rows are skipped regardless of the stored value; count = number of valid rows;
 * integer sums wrap in 64 bits (i32/u32 widen to i64/u64 first, like pyarrow.compute.sum); float sums
 * accumulate in f64 in index order (the oracle's order; the GPU documents its own and is compared
 * within 1e-12 relative); min/max skip invalid rows and, for floats, NaN; min prefers -0.0 over +0.0,
 * max prefers +0.0; count == 0 (or no non-NaN value) leaves min/max at their identities
 * (ints: TYPE_MAX / TYPE_MIN; floats: NaN). */
typedef struct { i64 sum; i64 min; i64 max; u64 count; } orc_agg_i64;
typedef struct { u64 sum; u64 min; u64 max; u64 count; } orc_agg_u64;
typedef struct { f64 sum; f64 min; f64 max; u64 count; } orc_agg_f64;

#define DEF_STATS_INT(T, ST, ACC, TMAXV, TMINV)                                                 \
    void orc_stats_##T(const T *d, size_t n, const uint8_t *validity, ST *out) {                \
        ACC sum = 0; T mn = TMAXV, mx = TMINV; u64 cnt = 0;                                     \
        for (size_t i = 0; i < n; ++i) {                                                        \
            if (validity && !bit_get(validity, i)) continue;                                    \
            sum += (ACC)d[i]; if (d[i] < mn) mn = d[i]; if (d[i] > mx) mx = d[i]; ++cnt;        \
        }                                                                                       \
        out->sum = sum; out->min = mn; out->max = mx; out->count = cnt;                         \
    }
/* -fwrapv makes the signed i64 accumulate wrap like Rust's wrapping_add. */
DEF_STATS_INT(i8, orc_agg_i64, i64, INT8_MAX, INT8_MIN)
DEF_STATS_INT(i16, orc_agg_i64, i64, INT16_MAX, INT16_MIN)
DEF_STATS_INT(u8, orc_agg_u64, u64, UINT8_MAX, 0)
DEF_STATS_INT(u16, orc_agg_u64, u64, UINT16_MAX, 0)
DEF_STATS_INT(i32, orc_agg_i64, i64, INT32_MAX, INT32_MIN)
DEF_STATS_INT(i64, orc_agg_i64, i64, INT64_MAX, INT64_MIN)
DEF_STATS_INT(u32, orc_agg_u64, u64, UINT32_MAX, 0)
DEF_STATS_INT(u64, orc_agg_u64, u64, UINT64_MAX, 0)

#define DEF_STATS_FLOAT(T)                                                                      \
    void orc_stats_##T(const T *d, size_t n, const uint8_t *validity, orc_agg_f64 *out) {     \
        f64 sum = 0.0, mn = NAN, mx = NAN; u64 cnt = 0;                                         \
        for (size_t i = 0; i < n; ++i) {                                                        \
            if (validity && !bit_get(validity, i)) continue;                                    \
            f64 v = (f64)d[i]; sum += v; ++cnt;                                                 \
            if (v != v) continue;                                                               \
            if (mn != mn || v < mn || (v == mn && signbit(v) && !signbit(mn))) mn = v;          \
            if (mx != mx || v > mx || (v == mx && !signbit(v) && signbit(mx))) mx = v;          \
        }                                                                                       \
        out->sum = sum; out->min = mn; out->max = mx; out->count = cnt;                         \
    }
DEF_STATS_FLOAT(f32) DEF_STATS_FLOAT(f64)

/* Masked i64 sum + count, the C2 workload, structured like the reference's parallel bench
 * (par_chunks(1<<20) -> per-chunk sum -> combine; benchmark_parallel_simd.rs:81-89) so that it can
 * serve as the timed CPU baseline.  Validity is consumed a u64 word at a time (chunks are 2^20 rows,
 * a multiple of 64).  All host threads via OpenMP when threads > 1. */
#if defined(__AVX2__)
#include <immintrin.h>
#define NL(b) ((b) ? -1ll : 0ll)
#define NIB(k) {{NL((k) & 1), NL((k) & 2), NL((k) & 4), NL((k) & 8)}}
static const union { long long q[4]; __m256i v; } NIBBLE_LANES_U[16] = {NIB(0), NIB(1), NIB(2), NIB(3), NIB(4), NIB(5), NIB(6), NIB(7),
                                                                        NIB(8), NIB(9), NIB(10), NIB(11), NIB(12), NIB(13), NIB(14), NIB(15)};
#define NIBBLE_LANES(k) (NIBBLE_LANES_U[(k)].v)
#endif
void orc_par_masked_sum_i64(const i64 *d, size_t n, const uint8_t *validity, int threads,
                            i64 *out_sum, u64 *out_count) {
    size_t nchunks = (n + PAR_CHUNK - 1) / PAR_CHUNK;
    u64 total = 0, count = 0;
    (void)threads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads > 0 ? threads : 1) reduction(+ : total, count)
    for (size_t c = 0; c < nchunks; ++c) {
        size_t lo = c * PAR_CHUNK, hi = lo + PAR_CHUNK < n ? lo + PAR_CHUNK : n;
        u64 s0 = 0, s1 = 0, s2 = 0, s3 = 0, cn = 0;
        if (!validity) {
            size_t i = lo;
            for (; i + 4 <= hi; i += 4) { s0 += (u64)d[i]; s1 += (u64)d[i + 1]; s2 += (u64)d[i + 2]; s3 += (u64)d[i + 3]; }
            for (; i < hi; ++i) s0 += (u64)d[i];
            cn = hi - lo;
        } else {
            size_t i = lo;
#if defined(__AVX2__)
            /* 4 x i64 lanes like the reference's Simd<i64, 4> (benchmark_parallel_simd.rs:44-60); a validity nibble
             * selects the lanes through a 16-entry table of lane masks. */
            __m256i a0 = _mm256_setzero_si256(), a1 = a0, a2 = a0, a3 = a0;
            for (; i + 64 <= hi; i += 64) {
                u64 w; memcpy(&w, validity + i / 8, 8);
                cn += (u64)__builtin_popcountll(w);
                const __m256i *p = (const __m256i *)(d + i);
                for (int q = 0; q < 16; q += 4) {
                    a0 = _mm256_add_epi64(a0, _mm256_and_si256(_mm256_loadu_si256(p + q), NIBBLE_LANES((w >> (4 * q)) & 15)));
                    a1 = _mm256_add_epi64(a1, _mm256_and_si256(_mm256_loadu_si256(p + q + 1), NIBBLE_LANES((w >> (4 * q + 4)) & 15)));
                    a2 = _mm256_add_epi64(a2, _mm256_and_si256(_mm256_loadu_si256(p + q + 2), NIBBLE_LANES((w >> (4 * q + 8)) & 15)));
                    a3 = _mm256_add_epi64(a3, _mm256_and_si256(_mm256_loadu_si256(p + q + 3), NIBBLE_LANES((w >> (4 * q + 12)) & 15)));
                }
            }
            {
                u64 t[4];
                _mm256_storeu_si256((__m256i *)t, _mm256_add_epi64(_mm256_add_epi64(a0, a1), _mm256_add_epi64(a2, a3)));
                s0 += t[0]; s1 += t[1]; s2 += t[2]; s3 += t[3];
            }
#endif
            for (; i + 64 <= hi; i += 64) {
                u64 w; memcpy(&w, validity + i / 8, 8);
                cn += (u64)__builtin_popcountll(w);
                for (int b = 0; b < 64; b += 4) {
                    s0 += (u64)d[i + b] & (0 - ((w >> b) & 1));
                    s1 += (u64)d[i + b + 1] & (0 - ((w >> (b + 1)) & 1));
                    s2 += (u64)d[i + b + 2] & (0 - ((w >> (b + 2)) & 1));
                    s3 += (u64)d[i + b + 3] & (0 - ((w >> (b + 3)) & 1));
                }
            }
            for (; i < hi; ++i) if (bit_get(validity, i)) { s0 += (u64)d[i]; ++cn; }
        }
        total += s0 + s1 + s2 + s3; count += cn;
    }
    *out_sum = (i64)total; *out_count = count;
}

/* Masked f64 element-wise add over all host cores (CPU baseline for C3; the reference leaf itself is
 * single-threaded, apply_float_f64; the chunked loop mirrors route_super_array_broadcast running one
 * leaf call per 2^20-row chunk). */
void orc_par_apply_float_f64(const f64 *lhs, const f64 *rhs, size_t n, int op, const uint8_t *mask,
                             int threads, f64 *out, uint8_t *out_mask) {
    size_t nchunks = (n + PAR_CHUNK - 1) / PAR_CHUNK;
    (void)threads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads > 0 ? threads : 1)
    for (size_t c = 0; c < nchunks; ++c) {
        size_t lo = c * PAR_CHUNK, len = lo + PAR_CHUNK < n ? PAR_CHUNK : n - lo;
        orc_apply_float_f64(lhs + lo, len, rhs + lo, len, op, mask ? mask + lo / 8 : NULL, out + lo,
                            mask ? out_mask + lo / 8 : NULL);
    }
}

/* simd_eq_mask_u8/_u16/_u32/_u64, src/kernels/bitmask/simd.rs:741-788: bit j = ((data[j] & field_mask) == target);
 * the SIMD body and the scalar tail compute the same predicate, so one loop restates both. */
#define DEF_EQ_MASK(T)                                                                          \
    void orc_simd_eq_mask_##T(const T *data, size_t n, T field_mask, T target, uint8_t *out) {  \
        memset(out, 0, (n + 7) / 8);                                                            \
        for (size_t j = 0; j < n; ++j)                                                          \
            if ((T)(data[j] & field_mask) == target) out[j / 8] |= (uint8_t)(1u << (j % 8));    \
    }
DEF_EQ_MASK(u8) DEF_EQ_MASK(u16) DEF_EQ_MASK(u32) DEF_EQ_MASK(u64)

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
