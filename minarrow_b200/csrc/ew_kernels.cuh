// Null-aware element-wise arithmetic in one HBM pass.
//
// Replaces int_dense_body / int_masked_body / float_dense_body / float_masked_body / fma bodies
// (src/kernels/arithmetic/simd.rs:52-751, std.rs:41-230) and, fused in, the caller-side validity merge
// (merge_bitmasks_to_new, src/kernels/bitmask/mod.rs:171-197; Bitmask::union, src/structs/bitmask.rs:661-669)
// and the scalar broadcast that the reference materialises (routing/broadcast.rs:25-47).
//
// Mapping: a lane owns one 16-byte vector (2 x 64-bit or 4 x 32-bit rows); a warp owns 32*U consecutive
// vectors, so each load/store instruction of a warp covers 512 contiguous bytes and the U output validity
// bytes-groups of a warp are contiguous.  Validity is read as the byte holding the lane's rows and written
// by one lane per byte after a shuffle gather — never a read-modify-write as in write_simd_mask_bits
// (src/utils.rs:255-283), never more than ceil(len/8) bytes.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "divmagic.h"
#include "fastmod.h"
#include "fastpow.h"
#include "internal.h"

namespace mnr {

// CLS_SDIV: integer Div / Rem / FloorDiv by a non-zero column-wide scalar, evaluated with the host-computed
// multiplicative inverse (divmagic.h) — chosen by the launcher, never by op_class().
enum { CLS_CHEAP = 0, CLS_DIV = 1, CLS_POW = 2, CLS_REM = 3, CLS_SDIV = 4 };

__host__ __device__ constexpr int op_class(bool is_float, int op) {
    return (op == MNR_ADD || op == MNR_SUB || op == MNR_MUL) ? CLS_CHEAP
           : (op == MNR_POW)                                  ? CLS_POW
           : (is_float && op == MNR_REM)                      ? CLS_REM
                                                              : CLS_DIV;
}

// ---- one element ---------------------------------------------------------------------------------------
// Truncating quotient of two 8/16-bit integers (r != 0) through the f32 pipe.  The compiler's `/` on the promoted ints
// is the generic 32-bit routine (~30 instructions: two conversions, a reciprocal, two correction steps); operands below
// 2^16 need none of the corrections:  uq = trunc((|l| + 0.5) * rcp(|r|)).  With t' = (|l| + 0.5) / |r|, both
// t' - floor(|l| / |r|) and floor(|l| / |r|) + 1 - t' are >= 0.5 / |r|, while the computed product is off by less than
// t' * 2^-21 <= 2^-5 / |r| (MUFU.RCP: 1 ulp, one rounding in the multiply) — the truncation can not cross an integer.
// int -> float goes through the 2^23 exponent trick (one integer op + FADD); unsigned quotients come back the same way
// (FADD.RZ + LOP3), signed ones through one truncating conversion.  MIN / -1 comes out as 2^(bits-1), which wraps to MIN
// in the caller's cast (the same wrapped quotient as the wide types).  Checked over the whole operand domain:
// tests/test_gpu_narrow_division.py (8-bit, and every multiple boundary of 16-bit), tests/sweep_div16.py (all 2^32 pairs).
template <bool SIGNED> __device__ __forceinline__ int narrow_quot(int l, int r) {
    float rc;
    if constexpr (SIGNED) {
        // Signed operands keep their signs through the float pipe: 1.5 * 2^23 + x is exact for |x| < 2^22, the half is
        // added away from zero (copysign), and the conversion truncates toward zero — no |x|, no sign restore.
        const float fa = __fsub_rn(__int_as_float(0x4B400000 + l), 12582912.0f);   // l, exact
        const float fb = __fsub_rn(__int_as_float(0x4B400000 + r), 12582912.0f);   // r, exact
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(fb));
        return __float2int_rz(__fmul_rn(__fadd_rn(fa, copysignf(0.5f, fa)), rc));    // r = 0: +-Inf saturates, nulled by the caller
    } else {
        const uint32_t al = (uint32_t)l, ar = (uint32_t)r;
        const float fa = __fsub_rn(__uint_as_float(0x4B000000u | al), 8388607.5f);   // l + 0.5, exact
        const float fb = __fsub_rn(__uint_as_float(0x4B000000u | ar), 8388608.0f);   // r, exact
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(fb));
        return (int)(__float_as_uint(__fadd_rz(__fmul_rn(fa, rc), 8388608.0f)) & 0x007FFFFFu);
    }
}

// Integers: Add/Sub/Mul wrap; Div/Rem truncate, FloorDiv per std.rs:68-77; MIN / -1 = MIN and
// MIN % -1 = 0 (core::simd's guard; DESIGN.md assumption); zero divisor => ok = false, value 0.
// Power: exponent = rhs.to_u32().unwrap_or(0) (std.rs:67), wrapping repeated multiply (simd.rs:94-100)
// evaluated by squaring — same residue mod 2^bits.
template <typename T, int CLS>
__device__ __forceinline__ T int_elem(int op, T l, T r, bool& ok, const DivMagic& dm) {
    using UT = typename std::make_unsigned<T>::type;
    ok = true;
    if constexpr (CLS == CLS_CHEAP) {
        const UT a = (UT)l, b = (UT)r;
        return (T)(UT)(op == MNR_ADD ? a + b : op == MNR_SUB ? a - b : a * b);
    } else if constexpr ((CLS == CLS_DIV || (CLS == CLS_SDIV && !std::is_signed<T>::value)) && sizeof(T) <= 2) {
        // 8/16-bit columns: branch-free (a zero divisor flows through as rcp = Inf -> quotient bits 0 and is nulled by `ok`),
        // in 32-bit registers, narrowed once at the end.  An UNSIGNED broadcast divisor (CLS_SDIV: non-zero by construction)
        // takes the same route — its reciprocal is loop-invariant, five full-rate instructions per quotient (r01y: u8 1.85 ->
        // 2.48 TB/s, u16 3.35 -> 3.60).  A signed one keeps the multiplicative inverse below: the |l| / sign-restore steps
        // cost more than the high multiply they replace (r01y: i16 3.63 -> 2.65 TB/s on this route, reverted).
        ok = CLS == CLS_SDIV || r != 0;
        const int q = narrow_quot<std::is_signed<T>::value>((int)l, (int)r);
        const int m = (int)l - q * (int)r;
        int res = op == MNR_DIV ? q : m;
        if (op == MNR_FLOORDIV) res = (std::is_signed<T>::value && m != 0 && (((int)l ^ (int)r) < 0)) ? q - 1 : q;
        return ok ? (T)res : (T)0;
    } else if constexpr (CLS == CLS_DIV && sizeof(T) == 4) {
        // 32-bit columns, branch-free: magnitudes through the unsigned divide (|MIN| = 2^31 is an ordinary u32, and 2^31 / 1
        // negated wraps back to MIN, so MIN / -1 = MIN and MIN % -1 = 0 need no test), a zero divisor is replaced by 1 and the
        // row nulled by `ok`.  The per-row `if (r == 0)` / `if (l == MIN && r == -1)` exits this replaces cost a divergence
        // region (BSSY / BSYNC + both sides executed) per row: ~50 instructions per row, 0.76-0.82 of the copy peak (r02z).
        ok = r != 0;
        const uint32_t ul = (uint32_t)l, ur = (uint32_t)r;
        uint32_t al = ul, ar = ur, sq = 0;
        if constexpr (std::is_signed<T>::value) {
            const uint32_t sl = (uint32_t)((int32_t)l >> 31), sr = (uint32_t)((int32_t)r >> 31);
            al = (ul ^ sl) - sl;
            ar = (ur ^ sr) - sr;
            sq = sl ^ sr;
        }
        const uint32_t uq = al / (ok ? ar : 1u);
        const uint32_t q = (uq ^ sq) - sq;
        const uint32_t m = ul - q * ur;
        uint32_t res = op == MNR_DIV ? q : m;
        if (op == MNR_FLOORDIV) res = (m != 0 && sq != 0) ? q - 1 : q;
        return ok ? (T)res : (T)0;
    } else if constexpr (CLS == CLS_DIV || CLS == CLS_SDIV) {
        T q;
        if constexpr (CLS == CLS_SDIV) {
            // r is the same non-zero scalar in every row: one high multiply instead of a divide.  8/16-bit columns are
            // widened to 32 bits; the wrapped quotient of MIN / -1 comes out of the same arithmetic.
            using W = typename std::conditional<sizeof(T) == 8, T,
                                                typename std::conditional<std::is_signed<T>::value, int32_t, uint32_t>::type>::type;
            q = (T)div_by_magic<W>((W)l, (W)r, dm);
        } else {
            if (r == 0) { ok = false; return 0; }
            if constexpr (std::is_signed<T>::value && sizeof(T) == 8) {
                // The 64-bit signed divide routine has no short path for small negative operands; the unsigned one does
                // (both high words zero -> 32-bit divide).  |l| / |r| with the signs restored is the same truncating
                // quotient, and 2^63 / 1 wraps back to MIN for MIN / -1.
                const uint64_t al = l < 0 ? 0ull - (uint64_t)l : (uint64_t)l, ar = r < 0 ? 0ull - (uint64_t)r : (uint64_t)r;
                const uint64_t uq = al / ar;
                q = (T)(((l ^ r) < 0) ? 0ull - uq : uq);
            } else {
                if constexpr (std::is_signed<T>::value) {
                    if (l == (T)((UT)1 << (sizeof(T) * 8 - 1)) && r == (T)-1) return op == MNR_REM ? (T)0 : l;
                }
                q = (T)(l / r);
            }
        }
        const T m = (T)((UT)l - (UT)((UT)q * (UT)r));
        if (op == MNR_DIV) return q;
        if (op == MNR_REM) return m;
        if constexpr (std::is_signed<T>::value) {
            if (m != 0 && ((l ^ r) < 0)) return (T)((UT)q - 1);
        }
        return q;
    } else {
        uint32_t e;
        if constexpr (std::is_signed<T>::value) e = (r < 0 || (uint64_t)r > 0xFFFFFFFFull) ? 0u : (uint32_t)r;
        else e = ((uint64_t)r > 0xFFFFFFFFull) ? 0u : (uint32_t)r;
        UT base = (UT)l, acc = 1;
        while (e) {
            if (e & 1u) acc = (UT)(acc * base);
            base = (UT)(base * base);
            e >>= 1;
        }
        return (T)acc;
    }
}

// Floats: single IEEE operations, no contraction (the TU is built with -fmad=false), `%` = fmod,
// Power = exp(b * ln a), FloorDiv = floor(a / b)  (std.rs:144-157).
template <typename T, int CLS>
__device__ __forceinline__ T float_elem(int op, T a, T b) {
    if constexpr (CLS == CLS_CHEAP) {
        return op == MNR_ADD ? a + b : op == MNR_SUB ? a - b : a * b;
    } else if constexpr (CLS == CLS_DIV) {
        const T q = a / b;
        return op == MNR_DIV ? q : floor(q);
    } else if constexpr (CLS == CLS_REM) {
        return fast_fmod<T>(a, b);   // exact like fmod; the libm loop only for NaN / Inf / zero divisors / huge quotients
    } else {
        if constexpr (std::is_same<T, double>::value) return fast_pow_f64(a, b);   // same expression, ~45 instead of ~120 FP64 operations
        else return exp(b * log(a));
    }
}

// 8/16-bit Add / Sub / Mul on a whole 32-bit word (4 or 2 elements): wrapping arithmetic has the same bits for signed
// and unsigned lanes.  Per element the generic path pays a byte extract, the operation, a select and a byte insert; a
// 1-byte column must move ~2e12 rows/s to stay on the HBM roofline, which leaves room for ~3 instructions per row.
template <int ESZ, int OP> __device__ __forceinline__ uint32_t packed_cheap_word(uint32_t a, uint32_t b) {
    if constexpr (ESZ == 1) {
        if constexpr (OP == MNR_ADD) return __vadd4(a, b);
        if constexpr (OP == MNR_SUB) return __vsub4(a, b);
        // the low byte of a product depends on the low bytes of the factors only: shift, multiply, keep byte 0
        const uint32_t p0 = a * b, p1 = (a >> 8) * (b >> 8), p2 = (a >> 16) * (b >> 16), p3 = (a >> 24) * (b >> 24);
        return __byte_perm(__byte_perm(p0, p1, 0x0040), __byte_perm(p2, p3, 0x0040), 0x5410);
    } else {
        if constexpr (OP == MNR_ADD) return __vadd2(a, b);
        if constexpr (OP == MNR_SUB) return __vsub2(a, b);
        return __byte_perm(a * b, (a >> 16) * (b >> 16), 0x5410);
    }
}
// One vector (NW words) with the operator fixed at compile time; `bits` = validity of the vector's elements.
// (a / b are register arrays: the scalar side is chosen per VALUE with has_a / has_b — choosing between POINTERS would force
// the arrays into local memory, ncu r01w: 64 LDL per tile and 60 % DRAM utilisation.)
template <int ESZ, int OP, bool MASKED, int NW>
__device__ __forceinline__ void packed_cheap_vec(const uint32_t (&a)[NW], bool has_a, const uint32_t (&b)[NW], bool has_b, uint32_t sword,
                                                 uint32_t bits, uint32_t (&o)[NW]) {
    constexpr int EPW = 4 / ESZ;
    uint32_t even = 0, odd = 0;   // 1-byte elements: nibble planes of the validity word, one PRMT per word picks nibble j
    if constexpr (MASKED && ESZ == 1) { even = bits & 0x0F0F0F0Fu; odd = (bits >> 4) & 0x0F0F0F0Fu; }
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        uint32_t r = packed_cheap_word<ESZ, OP>(has_a ? a[j] : sword, has_b ? b[j] : sword);
        if constexpr (MASKED) {
            if constexpr (ESZ == 1) r &= ((__byte_perm((j & 1) ? odd : even, 0u, 0x4440u | (uint32_t)(j >> 1)) * 0x00204081u) & 0x01010101u) * 0xFFu;
            else r &= expand_valid_word<ESZ>(bits >> (j * EPW));
        }
        o[j] = r;
    }
}

// ---- 8/16-bit integer Div / Rem / FloorDiv of two columns, SIMD-in-register --------------------------------------------
// The per-element path costs ~35 instructions per row (ncu r01zz: ALU pipe 72 %, issue slots 70 %): a byte extract and a
// sign extension per operand, the quotient, `q * r`, three selects for the operator, the zero-divisor test, the validity
// select, the validity bit and a byte insert.  Only the quotient has to be per row.  Here a 32-bit word (4 or 2 lanes) is
// the unit for everything else:
//   * signed lanes become magnitudes with one packed subtract ((x ^ s) - s, s = 0xFF.. in negative lanes); |MIN| stays
//     0x80.. and is read as the unsigned 2^(bits-1), so MIN / -1 comes out as 2^(bits-1) = MIN after the sign is restored
//     (the wrapped quotient every other width produces);
//   * one PRMT per operand both extracts lane k and plants it in the mantissa of 2^23 (the exponent constant is PRMT's second
//     source), one FADD removes the bias: the exact float |l| + 0.5 or |r|.  MUFU.RCP, FMUL and FADD.RZ(2^23) then leave
//     floor(|l| / |r|) in the low mantissa bits — bit for bit the arithmetic of narrow_quot<false>, whose exactness is
//     enumerated over all 2^32 (dividend, divisor) pairs of u16 (tests/sweep_div16.py).  A zero divisor gives Inf -> bits 0;
//   * the quotient lanes are packed back with 0.75 PRMT per lane; remainder = |l| - q * |r| with the packed multiply of
//     packed_cheap_word; signs are restored per word (quotient: sign(l) ^ sign(r); remainder: sign(l); FloorDiv: q - 1 where
//     the remainder is non-zero and the signs differ, std.rs:72-75);
//   * zero divisors and input validity are lane masks; the output word is ANDed once and the output validity bits are
//     squeezed out of the mask with one multiply.
// ~10 (unsigned Div) to ~15 (signed Rem / FloorDiv) instructions per row.
template <int ESZ> __device__ __forceinline__ uint32_t lane_neg_mask(uint32_t x) {
    // PRMT with selector bit 3 set replicates the sign bit of the selected byte: one instruction per word
    uint32_t d;
    if constexpr (ESZ == 1) asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(d) : "r"(x));
    else asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(d) : "r"(x));
    return d;
}
template <int ESZ> __device__ __forceinline__ uint32_t lane_nonzero_mask(uint32_t x) {   // 0xFF.. in lanes that are != 0
    if constexpr (ESZ == 1) return (((((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) >> 7) & 0x01010101u) * 0xFFu;
    else return (((((x & 0x7FFF7FFFu) + 0x7FFF7FFFu) | x) >> 15) & 0x00010001u) * 0xFFFFu;
}
template <int ESZ> __device__ __forceinline__ uint32_t lane_sub(uint32_t a, uint32_t b) {
    if constexpr (ESZ == 1) return __vsub4(a, b);
    else return __vsub2(a, b);
}
// t + inc per lane for inc in {0, 1}, without carries between lanes (4 instructions; the emulated packed subtract costs more)
template <int ESZ> __device__ __forceinline__ uint32_t lane_add01(uint32_t t, uint32_t inc) {
    constexpr uint32_t L = ESZ == 1 ? 0x7F7F7F7Fu : 0x7FFF7FFFu;
    return ((t & L) + inc) ^ (t & ~L);
}
// lane k of `w` as the float bits of 2^23 + lane (PRMT with the exponent constant as second source)
template <int ESZ, int K> __device__ __forceinline__ float lane_biased(uint32_t w) {
    constexpr uint32_t sel = ESZ == 1 ? (0x7540u | (uint32_t)K) : (K == 0 ? 0x7510u : 0x7532u);
    return __uint_as_float(__byte_perm(w, 0x4B000000u, sel));
}
template <int ESZ> __device__ __forceinline__ uint32_t lane_valid_to_bits(uint32_t m) {   // lane mask -> EPW validity bits
    if constexpr (ESZ == 1) return ((m & 0x01010101u) * 0x01020408u) >> 24 & 0xFu;
    else { const uint32_t x = m & 0x00010001u; return (x | (x >> 15)) & 3u; }
}

// A broadcast scalar divisor (the same value in every lane): its magnitude, sign mask and reciprocal are computed once
// per thread, outside the loops, and a row costs one PRMT, FADD, FMUL and FADD.RZ.
struct PackedDivisor {
    uint32_t ra;    // |r| in every lane
    uint32_t sr;    // 0xFF.. in every lane iff r < 0
    float rc;       // MUFU.RCP(|r|): the same instruction on the same operand as the two-column path -> the same bits
};
template <int ESZ, bool SIGNED> __device__ __forceinline__ PackedDivisor packed_divisor(uint32_t sword) {
    constexpr uint32_t ONE = ESZ == 1 ? 0x01010101u : 0x00010001u;
    PackedDivisor d;
    d.sr = SIGNED ? lane_neg_mask<ESZ>(sword) : 0u;
    d.ra = SIGNED ? (sword ^ d.sr) + (d.sr & ONE) : sword;
    const float fb = __fsub_rn(lane_biased<ESZ, 0>(d.ra), 8388608.0f);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(d.rc) : "f"(fb));
    return d;
}

// One 32-bit word of lanes.  `vm` (in/out): lanes that are valid on input -> lanes valid on output (zero divisors removed).
// SCALAR_R: the divisor is the broadcast scalar `ds` (rw is ignored).
template <int ESZ, bool SIGNED, int OP, bool SCALAR_R = false>
__device__ __forceinline__ uint32_t packed_div_word(uint32_t lw, uint32_t rw, uint32_t& vm, const PackedDivisor& ds = PackedDivisor{}) {
    constexpr int EPW = 4 / ESZ;
    uint32_t sl = 0, sq = 0, la = lw, ra = SCALAR_R ? ds.ra : rw;
    if constexpr (SIGNED) {
        sl = lane_neg_mask<ESZ>(lw);
        const uint32_t sr = SCALAR_R ? ds.sr : lane_neg_mask<ESZ>(rw);
        sq = sl ^ sr;
        // |x| = (x ^ s) + 1 in negative lanes: ~x <= 2^(bits-1) - 1 there, so the increment never carries into the next lane
        constexpr uint32_t ONE = ESZ == 1 ? 0x01010101u : 0x00010001u;
        la = (lw ^ sl) + (sl & ONE);
        if constexpr (!SCALAR_R) ra = (rw ^ sr) + (sr & ONE);
    }
    uint32_t t[EPW];
#pragma unroll
    for (int k = 0; k < EPW; ++k) {
        float fa, fb = 0.0f;
        if (k == 0) { fa = lane_biased<ESZ, 0>(la); if constexpr (!SCALAR_R) fb = lane_biased<ESZ, 0>(ra); }
        else if (k == 1) { fa = lane_biased<ESZ, 1>(la); if constexpr (!SCALAR_R) fb = lane_biased<ESZ, 1>(ra); }
        else if (k == 2) { fa = lane_biased<ESZ, (EPW > 2 ? 2 : 0)>(la); if constexpr (!SCALAR_R) fb = lane_biased<ESZ, (EPW > 2 ? 2 : 0)>(ra); }
        else { fa = lane_biased<ESZ, (EPW > 2 ? 3 : 0)>(la); if constexpr (!SCALAR_R) fb = lane_biased<ESZ, (EPW > 2 ? 3 : 0)>(ra); }
        fa = __fsub_rn(fa, 8388607.5f);   // |l| + 0.5, exact
        float rc;
        if constexpr (SCALAR_R) rc = ds.rc;
        else {
            fb = __fsub_rn(fb, 8388608.0f);   // |r|, exact
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(fb));
        }
        t[k] = __float_as_uint(__fadd_rz(__fmul_rn(fa, rc), 8388608.0f));
    }
    uint32_t uq;
    if constexpr (ESZ == 1) uq = __byte_perm(__byte_perm(t[0], t[1], 0x0040), __byte_perm(t[2], t[3], 0x0040), 0x5410);
    else uq = __byte_perm(t[0], t[1], 0x5410);
    if constexpr (!SCALAR_R) vm &= lane_nonzero_mask<ESZ>(rw);   // a scalar divisor on this path is non-zero by construction (api.cu)
    if constexpr (OP == MNR_DIV && !SIGNED) return uq;
    if constexpr (OP == MNR_FLOORDIV && !SIGNED) return uq;
    constexpr uint32_t ONE = ESZ == 1 ? 0x01010101u : 0x00010001u;
    if constexpr (OP == MNR_DIV) return lane_add01<ESZ>(uq ^ sq, sq & ONE);                 // -q = ~q + 1 in the lanes whose signs differ
    const uint32_t um = la - packed_cheap_word<ESZ, MNR_MUL>(uq, ra);   // |l| - q |r| >= 0 in every lane: a plain subtract never borrows across lanes
    if constexpr (OP == MNR_REM) return SIGNED ? lane_add01<ESZ>(um ^ sl, sl & ONE) : um;   // the remainder takes the dividend's sign
    // FloorDiv, signed (std.rs:72-75): where the signs differ the result is -q - (remainder != 0) = ~q + (remainder == 0)
    return lane_add01<ESZ>(uq ^ sq, sq & ~lane_nonzero_mask<ESZ>(um) & ONE);
}

// One vector (NW words); returns the output validity bits of its lanes (MASKED) / whether a zero divisor was met (dense).
template <int ESZ, bool SIGNED, int OP, bool MASKED, int NW, bool SCALAR_NZ = false>
__device__ __forceinline__ uint32_t packed_div_vec(const uint32_t (&a)[NW], bool has_a, const uint32_t (&b)[NW], bool has_b, uint32_t sword,
                                                   uint32_t bits, uint32_t (&o)[NW], const PackedDivisor& ds = PackedDivisor{}) {
    constexpr int EPW = 4 / ESZ;
    uint32_t even = 0, odd = 0, ob = 0;
    if constexpr (MASKED && ESZ == 1) { even = bits & 0x0F0F0F0Fu; odd = (bits >> 4) & 0x0F0F0F0Fu; }
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        uint32_t vm = 0xFFFFFFFFu;
        if constexpr (MASKED) {
            if constexpr (ESZ == 1) vm = ((__byte_perm((j & 1) ? odd : even, 0u, 0x4440u | (uint32_t)(j >> 1)) * 0x00204081u) & 0x01010101u) * 0xFFu;
            else vm = expand_valid_word<ESZ>(bits >> (j * EPW));
        }
        uint32_t r;
        if (SCALAR_NZ) r = packed_div_word<ESZ, SIGNED, OP, true>(a[j], 0u, vm, ds);
        else r = packed_div_word<ESZ, SIGNED, OP>(has_a ? a[j] : sword, has_b ? b[j] : sword, vm);
        if constexpr (MASKED) { o[j] = r & vm; ob |= lane_valid_to_bits<ESZ>(vm) << (j * EPW); }
        else { o[j] = r; ob |= ~vm; }   // dense: any zero divisor raises the flag (the reference panics)
    }
    return ob;
}

// ---- 8-bit integer Power through a shared-memory table ---------------------------------------------------------------------
// base^e mod 256 by repeated squaring is ~35 instructions per row for general exponents (and data-dependent).  In Z/256 the
// whole function is tiny: an ODD base has multiplicative order dividing 64, so b^e = b^(e mod 64); an EVEN base gives 0 as
// soon as e >= 8 (2^8 | b^8).  Table: 128 half-bases x (64 odd-exponent entries + 16 even-exponent entries) = 10 KB of
// shared memory, built by the block once (resident grid), then ONE byte load per row.  Signed columns use the same bits
// (wrapping multiply); a negative exponent is `rhs.to_u32().unwrap_or(0)` = 0 (std.rs:67) -> 1.
constexpr int kPowLutRow = 80;   // per half-base: [0, 64) odd base, exponent mod 64; [64, 80) even base, exponent min(e, 8) (+ padding)
template <int BLOCK> __device__ __forceinline__ void build_pow8_lut(uint8_t* lut) {
    // one thread per (half-base, parity): the powers of one base by running product, 64 (odd) or 16 (even) multiplies
    for (int i = threadIdx.x; i < 256; i += BLOCK) {
        const uint32_t x = (uint32_t)i >> 1, odd = (uint32_t)i & 1u, base = 2 * x + odd;
        uint8_t* row = lut + x * kPowLutRow + (odd ? 0 : 64);
        uint32_t acc = 1;
        const int cnt = odd ? 64 : 16;
        for (int t = 0; t < cnt; ++t) {
            row[t] = (uint8_t)((!odd && t >= 8) ? 0u : acc);
            acc = (acc * base) & 0xFFu;
        }
    }
    __syncthreads();
}
template <bool SIGNED, bool MASKED, int NW>
__device__ __forceinline__ void packed_pow8_vec(const uint8_t* __restrict__ lut, const uint32_t (&a)[NW], bool has_a, const uint32_t (&b)[NW], bool has_b,
                                                uint32_t sword, uint32_t bits, uint32_t (&o)[NW]) {
    uint32_t even = 0, odd = 0;
    if constexpr (MASKED) { even = bits & 0x0F0F0F0Fu; odd = (bits >> 4) & 0x0F0F0F0Fu; }
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        const uint32_t lw = has_a ? a[j] : sword;
        uint32_t ew = has_b ? b[j] : sword;
        if constexpr (SIGNED) ew &= ~lane_neg_mask<1>(ew);                       // negative exponent -> 0
        const uint32_t oddm = (lw & 0x01010101u) * 0xFFu;                         // lanes with an odd base
        const uint32_t tw = ((ew & 0x3F3F3F3Fu) & oddm) | ((0x40404040u + __vminu4(ew, 0x08080808u)) & ~oddm);
        const uint32_t xw = (lw >> 1) & 0x7F7F7F7Fu;                              // half-base
        uint32_t r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = lut[((xw >> (8 * k)) & 0xFFu) * kPowLutRow + ((tw >> (8 * k)) & 0xFFu)];
        uint32_t w = __byte_perm(__byte_perm(r[0], r[1], 0x0040), __byte_perm(r[2], r[3], 0x0040), 0x5410);
        if constexpr (MASKED) w &= ((__byte_perm((j & 1) ? odd : even, 0u, 0x4440u | (uint32_t)(j >> 1)) * 0x00204081u) & 0x01010101u) * 0xFFu;
        o[j] = w;
    }
}

// ---- 16-bit integer Power: square-and-multiply with a thread-uniform trip count -------------------------------------------
// The per-element loop of int_elem ran ~80 instructions per row on 16-bit columns (r02z: 2.8 TB/s, 0.435 of the copy peak):
// every row has its own `while (e)`, so a warp replays the loop body for each distinct bit length, each product is narrowed
// back to 16 bits, and the operands are extracted and sign-extended one by one.  Here the exponent bits are tested in place
// in the packed operand word (bit t of the low lane, bit t + 16 of the high lane), the products run in 32-bit registers without
// narrowing (the low 16 bits of a product depend on the low 16 bits of the factors only; the high lane is `word >> 16`), and
// the loop runs max-bit-length-of-the-vector times for all 16 rows of a thread: 2 IMAD + 1 LOP3 per row and bit.
// A negative exponent is `rhs.to_u32().unwrap_or(0)` = 0 (std.rs:67) -> 1.
template <bool SIGNED, bool MASKED, int NW>
__device__ __forceinline__ void packed_pow16_vec(const uint32_t (&a)[NW], bool has_a, const uint32_t (&b)[NW], bool has_b, uint32_t sword,
                                                 uint32_t bits, uint32_t (&o)[NW]) {
    constexpr int CW = NW < 4 ? NW : 4;   // four words (8 rows) per loop: 5 live registers per word, the tile's other operands stay in registers
#pragma unroll
    for (int c = 0; c < NW; c += CW) {
        uint32_t ew[CW], blo[CW], bhi[CW], alo[CW], ahi[CW];
        uint32_t any = 0;
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            uint32_t e = has_b ? b[c + j] : sword;
            if constexpr (SIGNED) e &= ~lane_neg_mask<2>(e);
            ew[j] = e;
            any |= e;
            const uint32_t lw = has_a ? a[c + j] : sword;
            blo[j] = lw;
            bhi[j] = lw >> 16;
            alo[j] = 1u;
            ahi[j] = 1u;
        }
        any = (any | (any >> 16)) & 0xFFFFu;
        uint32_t mlo = 1u, mhi = 0x10000u;
        while (any) {
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                if (ew[j] & mlo) alo[j] *= blo[j];
                blo[j] *= blo[j];
                if (ew[j] & mhi) ahi[j] *= bhi[j];
                bhi[j] *= bhi[j];
            }
            mlo <<= 1;
            mhi <<= 1;
            any >>= 1;
        }
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            uint32_t w = __byte_perm(alo[j], ahi[j], 0x5410);
            if constexpr (MASKED) w &= expand_valid_word<2>(bits >> ((c + j) * 2));
            o[c + j] = w;
        }
    }
}

// 32 / 64-bit integer Power of one vector: the same thread-uniform loop (no per-row divergence), exponents shifted in place.
template <typename T> __device__ __forceinline__ uint32_t pow_exponent(T r) {   // rhs.to_u32().unwrap_or(0), std.rs:67
    if constexpr (std::is_signed<T>::value) return (r < 0 || (uint64_t)r > 0xFFFFFFFFull) ? 0u : (uint32_t)r;
    else return ((uint64_t)r > 0xFFFFFFFFull) ? 0u : (uint32_t)r;
}
template <typename UT, int VEC>
__device__ __forceinline__ void int_pow_vec(UT (&base)[VEC], uint32_t (&e)[VEC], UT (&acc)[VEC]) {
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < VEC; ++k) { any |= e[k]; acc[k] = 1; }
    while (any) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            if (e[k] & 1u) acc[k] = (UT)(acc[k] * base[k]);
            base[k] = (UT)(base[k] * base[k]);
            e[k] >>= 1;
        }
        any >>= 1;
    }
}

template <typename T, int CLS>
__device__ __forceinline__ T elem(int op, T l, T r, bool& ok, const DivMagic& dm) {
    if constexpr (Traits<T>::is_float) { ok = true; return float_elem<T, CLS>(op, l, r); }
    else return int_elem<T, CLS>(op, l, r, ok, dm);
}

// ---- validity helpers ------------------------------------------------------------------------------------
// Guarded variant of load_valid_bits: only bytes that hold rows < n are touched.
template <int NBITS>
__device__ __forceinline__ uint32_t load_valid_bits_guard(const uint8_t* __restrict__ mask, uint64_t row0, uint64_t n) {
    if (row0 >= n) return 0;
    uint32_t bits;
    if constexpr (NBITS <= 8) {
        bits = (ldg_u8(mask + (row0 >> 3)) >> (uint32_t)(row0 & 7)) & ((1u << NBITS) - 1u);
    } else {
        bits = 0;
        const uint8_t* p = mask + (row0 >> 3);
#pragma unroll
        for (int j = 0; j < NBITS / 8; ++j)
            if (row0 + 8ull * j < n) bits |= ldg_u8(p + j) << (8 * j);
    }
    const uint64_t left = n - row0;
    if (left < NBITS) bits &= (1u << (uint32_t)left) - 1u;
    return bits;
}

// Gather the per-lane validity bits of a warp into bytes and store them.  VEC rows per lane; lanes hold
// consecutive row groups starting at a multiple of 32*VEC rows.  `row0` is the lane's first row; lanes whose
// rows are all >= n pass bits = 0 and store nothing.
template <int VEC>
__device__ __forceinline__ void store_valid_bits(uint8_t* __restrict__ out_mask, uint64_t row0, uint64_t n, uint32_t bits) {
    if constexpr (VEC < 8) {
        constexpr int LPB = 8 / VEC;   // lanes per output byte
        uint32_t x = bits;
#pragma unroll
        for (int s = 1; s < LPB; s <<= 1) x |= __shfl_down_sync(0xffffffffu, x, s) << (VEC * s);
        if (((threadIdx.x & 31) % LPB) == 0 && row0 < n) out_mask[row0 >> 3] = (uint8_t)x;
    } else {
        // 16 / 32 rows per lane (2- and 1-byte columns): one 16- / 32-bit store when the lane's rows are all inside the column
        // and the address allows it (row0 is a multiple of VEC, so only the base pointer matters) — four byte stores per
        // lane were a visible share of the LSU work of the 1-byte kernels
        uint8_t* p = out_mask + (row0 >> 3);
        if (row0 + VEC <= n && (reinterpret_cast<uintptr_t>(p) & (VEC / 8 - 1)) == 0) {
            if constexpr (VEC == 32) *reinterpret_cast<uint32_t*>(p) = bits;
            else if constexpr (VEC == 16) *reinterpret_cast<uint16_t*>(p) = (uint16_t)bits;
            else {
#pragma unroll
                for (int j = 0; j < VEC / 8; ++j) p[j] = (uint8_t)(bits >> (8 * j));
            }
            return;
        }
#pragma unroll
        for (int j = 0; j < VEC / 8; ++j)
            if (row0 + 8ull * j < n) out_mask[(row0 >> 3) + j] = (uint8_t)(bits >> (8 * j));
    }
}

template <typename T, typename VecT> struct VecU {
    static constexpr int VEC = sizeof(VecT) / sizeof(T);
    union { VecT v; T e[VEC]; };
};

struct EwDev {
    const void* lhs;
    const void* rhs;
    uint64_t scalar_bits;
    const uint8_t* lmask;
    const uint8_t* rmask;
    int mask_or;
    void* out;
    uint8_t* out_mask;
    uint64_t n;
    unsigned int* div0_flag;
    int op;
    int sdiv;            // 1: the scalar is a non-zero integer divisor and `magic` is its inverse (CLS_SDIV launches only)
    DivMagic magic;
};

template <typename T> __device__ __forceinline__ T scalar_from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));   // little-endian: the low bytes hold the element
    return v;
}

template <int VEC, bool GUARD>
__device__ __forceinline__ uint32_t merged_bits(const EwDev& a, uint64_t row0) {
    uint32_t m;
    if constexpr (VEC == 32 && !GUARD) {
        // 1-byte elements, two masks: one loop-invariant test of both base pointers instead of an alignment test per load
        if (a.lmask && a.rmask && ((reinterpret_cast<uintptr_t>(a.lmask) | reinterpret_cast<uintptr_t>(a.rmask)) & 3u) == 0) {
            const uint32_t x = ldg_u32(a.lmask + (row0 >> 3)), y = ldg_u32(a.rmask + (row0 >> 3));
            return a.mask_or ? (x | y) : (x & y);
        }
    }
    if (a.lmask && a.rmask) {
        const uint32_t x = GUARD ? load_valid_bits_guard<VEC>(a.lmask, row0, a.n) : load_valid_bits<VEC>(a.lmask, row0);
        const uint32_t y = GUARD ? load_valid_bits_guard<VEC>(a.rmask, row0, a.n) : load_valid_bits<VEC>(a.rmask, row0);
        m = a.mask_or ? (x | y) : (x & y);
    } else {
        const uint8_t* p = a.lmask ? a.lmask : a.rmask;
        m = GUARD ? load_valid_bits_guard<VEC>(p, row0, a.n) : load_valid_bits<VEC>(p, row0);
    }
    return m;
}

// Element-wise binary kernel.  TL/TR: stored operand types (== T except for the cast-on-load promotion).
template <typename T, typename TL, typename TR, typename VecT, int CLS, bool MASKED, int BLOCK, int U>
__device__ __forceinline__ void ew_binary_body(const EwDev& a) {
    constexpr int VEC = sizeof(VecT) / sizeof(T);
    // Operand vectors carry VEC elements of their own (possibly narrower) type.
    struct alignas(sizeof(TL) * VEC) LV { TL e[VEC]; };
    struct alignas(sizeof(TR) * VEC) RV { TR e[VEC]; };
    using LVec = typename std::conditional<std::is_same<TL, T>::value, VecT, LV>::type;
    using RVec = typename std::conditional<std::is_same<TR, T>::value, VecT, RV>::type;

    const LVec* __restrict__ lp = static_cast<const LVec*>(a.lhs);
    const RVec* __restrict__ rp = static_cast<const RVec*>(a.rhs);
    VecT* __restrict__ op_ = static_cast<VecT*>(a.out);
    const int op = a.op;
    const uint64_t n = a.n;
    const uint64_t nvec_full = n / VEC;
    constexpr uint64_t WTILE = 32ull * U;
    const uint64_t ntiles = nvec_full / WTILE;
    const uint64_t warps = (uint64_t)gridDim.x * (BLOCK / 32);
    const uint64_t gwarp = (uint64_t)blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const T sval = scalar_from_bits<T>(a.scalar_bits);
    // SIMD-in-register path: 8-bit integers, cheap operators, same stored type on both sides, vector loads.  (Measured
    // for 16-bit columns too: the per-element path already runs at the HBM rate there — 7.07 TB/s for a two-mask add —
    // and the packed one is slower, 5.5 TB/s, because expanding 2 validity bits costs as much as the two selects it saves.)
    constexpr bool PACKED = CLS == CLS_CHEAP && !Traits<T>::is_float && sizeof(T) == 1 && std::is_same<TL, T>::value &&
                            std::is_same<TR, T>::value && sizeof(VecT) >= 16;
    // ... and the division family of 8/16-bit columns (packed_div_vec above).
    constexpr bool PACKED_DIV = CLS == CLS_DIV && !Traits<T>::is_float && sizeof(T) <= 2 && std::is_same<TL, T>::value &&
                                std::is_same<TR, T>::value && sizeof(VecT) >= 16;
    // ... and Power of 1-byte columns (packed_pow8_vec above: one shared-memory table lookup per row).
    constexpr bool PACKED_POW = CLS == CLS_POW && !Traits<T>::is_float && sizeof(T) == 1 && std::is_same<TL, T>::value &&
                                std::is_same<TR, T>::value && sizeof(VecT) >= 16;
    // ... Power of 2-byte columns (packed_pow16_vec) and of 4 / 8-byte columns (int_pow_vec): thread-uniform square-and-multiply.
    constexpr bool PACKED_POW16 = CLS == CLS_POW && !Traits<T>::is_float && sizeof(T) == 2 && std::is_same<TL, T>::value &&
                                  std::is_same<TR, T>::value && sizeof(VecT) >= 16;
    constexpr bool VEC_POW = CLS == CLS_POW && !Traits<T>::is_float && sizeof(T) >= 4 && std::is_same<TL, T>::value &&
                             std::is_same<TR, T>::value;
    __shared__ uint8_t pow_lut[PACKED_POW ? 128 * kPowLutRow : 1];
    if constexpr (PACKED_POW) build_pow8_lut<BLOCK>(pow_lut);
    const uint32_t sword = sizeof(T) == 1 ? (uint32_t)(uint8_t)a.scalar_bits * 0x01010101u : (uint32_t)(uint16_t)a.scalar_bits * 0x00010001u;
    const DivMagic dm = a.magic;
    bool div0 = false;
    // packed division by a non-zero broadcast scalar on the right: magnitude, sign and reciprocal of the divisor, once
    bool scalar_nz = false;
    PackedDivisor pds{};
    if constexpr (PACKED_DIV) {
        scalar_nz = rp == nullptr && lp != nullptr && (sizeof(T) == 1 ? (uint8_t)a.scalar_bits != 0 : (uint16_t)a.scalar_bits != 0);
        if (scalar_nz) pds = packed_divisor<(sizeof(T) <= 2 ? (int)sizeof(T) : 1), Traits<T>::is_signed>(sword);
    }

    for (uint64_t t = gwarp; t < ntiles; t += warps) {
        const uint64_t v0 = t * WTILE + lane;
        union LU { LVec v; TL e[VEC]; } L[U];
        union RU { RVec v; TR e[VEC]; } R[U];
        uint32_t mb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (lp) L[u].v = ldg_stream(lp + v0 + 32ull * u);
#pragma unroll
        for (int u = 0; u < U; ++u) if (rp) R[u].v = ldg_stream(rp + v0 + 32ull * u);
        if constexpr (MASKED) {
#pragma unroll
            for (int u = 0; u < U; ++u) mb[u] = merged_bits<VEC, false>(a, (v0 + 32ull * u) * VEC);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            VecU<T, VecT> O;
            uint32_t ob = 0;
            if constexpr (PACKED) {
                // 8/16-bit add / sub / mul: 32 bits at a time; these operators never null a row, so the output validity
                // is the merged input validity and the values of invalid rows are cleared with the expanded mask.
                constexpr int NW = sizeof(VecT) / 4;
                union { VecT v; uint32_t w[NW]; } PL, PR, PO;
                memcpy(&PL.v, &L[u].v, sizeof(VecT));
                memcpy(&PR.v, &R[u].v, sizeof(VecT));
                const bool ha = lp != nullptr, hb = rp != nullptr;
                const uint32_t vb = MASKED ? mb[u] : 0u;
                if (op == MNR_ADD) packed_cheap_vec<sizeof(T), MNR_ADD, MASKED, NW>(PL.w, ha, PR.w, hb, sword, vb, PO.w);
                else if (op == MNR_SUB) packed_cheap_vec<sizeof(T), MNR_SUB, MASKED, NW>(PL.w, ha, PR.w, hb, sword, vb, PO.w);
                else packed_cheap_vec<sizeof(T), MNR_MUL, MASKED, NW>(PL.w, ha, PR.w, hb, sword, vb, PO.w);
                O.v = PO.v;
                if constexpr (MASKED) ob = mb[u];
            } else if constexpr (PACKED_DIV) {
                constexpr int NW = sizeof(VecT) / 4;
                constexpr int ESZ = sizeof(T) <= 2 ? (int)sizeof(T) : 1;
                constexpr bool SG = Traits<T>::is_signed;
                union { VecT v; uint32_t w[NW]; } PL, PR, PO;
                memcpy(&PL.v, &L[u].v, sizeof(VecT));
                memcpy(&PR.v, &R[u].v, sizeof(VecT));
                const bool ha = lp != nullptr, hb = rp != nullptr;
                const uint32_t vb = MASKED ? mb[u] : 0u;
                uint32_t r;
                if (scalar_nz) {   // column (op) non-zero scalar: the divisor's reciprocal lives in registers (kernel-uniform branch)
                    if (op == MNR_DIV) r = packed_div_vec<ESZ, SG, MNR_DIV, MASKED, NW, true>(PL.w, true, PR.w, false, sword, vb, PO.w, pds);
                    else if (op == MNR_REM) r = packed_div_vec<ESZ, SG, MNR_REM, MASKED, NW, true>(PL.w, true, PR.w, false, sword, vb, PO.w, pds);
                    else r = packed_div_vec<ESZ, SG, MNR_FLOORDIV, MASKED, NW, true>(PL.w, true, PR.w, false, sword, vb, PO.w, pds);
                } else if (op == MNR_DIV) r = packed_div_vec<ESZ, SG, MNR_DIV, MASKED, NW>(PL.w, ha, PR.w, hb, sword, vb, PO.w);
                else if (op == MNR_REM) r = packed_div_vec<ESZ, SG, MNR_REM, MASKED, NW>(PL.w, ha, PR.w, hb, sword, vb, PO.w);
                else r = packed_div_vec<ESZ, SG, MNR_FLOORDIV, MASKED, NW>(PL.w, ha, PR.w, hb, sword, vb, PO.w);
                O.v = PO.v;
                if constexpr (MASKED) ob = r;
                else div0 |= r != 0;
            } else if constexpr (PACKED_POW) {
                constexpr int NW = sizeof(VecT) / 4;
                union { VecT v; uint32_t w[NW]; } PL, PR, PO;
                memcpy(&PL.v, &L[u].v, sizeof(VecT));
                memcpy(&PR.v, &R[u].v, sizeof(VecT));
                packed_pow8_vec<Traits<T>::is_signed, MASKED, NW>(pow_lut, PL.w, lp != nullptr, PR.w, rp != nullptr, sword, MASKED ? mb[u] : 0u, PO.w);
                O.v = PO.v;
                if constexpr (MASKED) ob = mb[u];      // Power never nulls a row: output validity = merged input validity
            } else if constexpr (PACKED_POW16) {
                constexpr int NW = sizeof(VecT) / 4;
                union { VecT v; uint32_t w[NW]; } PL, PR, PO;
                memcpy(&PL.v, &L[u].v, sizeof(VecT));
                memcpy(&PR.v, &R[u].v, sizeof(VecT));
                packed_pow16_vec<Traits<T>::is_signed, MASKED, NW>(PL.w, lp != nullptr, PR.w, rp != nullptr, sword, MASKED ? mb[u] : 0u, PO.w);
                O.v = PO.v;
                if constexpr (MASKED) ob = mb[u];
            } else if constexpr (VEC_POW) {
                using UT = typename std::make_unsigned<T>::type;
                UT base[VEC], acc[VEC];
                uint32_t e[VEC];
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    base[k] = (UT)(lp ? (T)L[u].e[k] : sval);
                    e[k] = pow_exponent<T>(rp ? (T)R[u].e[k] : sval);
                }
                int_pow_vec<UT, VEC>(base, e, acc);
#pragma unroll
                for (int k = 0; k < VEC; ++k) O.e[k] = (!MASKED || ((mb[u] >> k) & 1u)) ? (T)acc[k] : (T)0;
                if constexpr (MASKED) ob = mb[u];
            } else
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const T l = lp ? (T)L[u].e[k] : sval;
                const T r = rp ? (T)R[u].e[k] : sval;
                bool ok;
                const T val = elem<T, CLS>(op, l, r, ok, dm);
                if constexpr (MASKED) {
                    const bool valid = ((mb[u] >> k) & 1u) && ok;
                    O.e[k] = valid ? val : (T)0;
                    ob |= (uint32_t)valid << k;
                } else {
                    O.e[k] = val;
                    div0 |= !ok;
                }
            }
            stg_stream(op_ + v0 + 32ull * u, O.v);
            if constexpr (MASKED) store_valid_bits<VEC>(a.out_mask, (v0 + 32ull * u) * VEC, n, ob);
        }
    }

    // Guarded remainder: vectors past the last full warp tile, including a final partial vector.
    const uint64_t nvec_ceil = (n + VEC - 1) / VEC;
    for (uint64_t vb = ntiles * WTILE + gwarp * 32ull; vb < nvec_ceil; vb += warps * 32ull) {
        const uint64_t v = vb + lane;
        const uint64_t row0 = v * VEC;
        const int nrows = row0 >= n ? 0 : (n - row0 >= (uint64_t)VEC ? VEC : (int)(n - row0));
        VecU<T, VecT> O;
        uint32_t ob = 0;
        if (nrows > 0) {
            union LU { LVec v; TL e[VEC]; } L;
            union RU { RVec v; TR e[VEC]; } R;
            if (nrows == VEC) {
                if (lp) L.v = ldg_stream(lp + v);
                if (rp) R.v = ldg_stream(rp + v);
            } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    if (k < nrows) {
                        if (lp) L.e[k] = static_cast<const TL*>(a.lhs)[row0 + k];
                        if (rp) R.e[k] = static_cast<const TR*>(a.rhs)[row0 + k];
                    } else {
                        if (lp) L.e[k] = (TL)1;
                        if (rp) R.e[k] = (TR)1;
                    }
                }
            }
            uint32_t m = 0;
            if constexpr (MASKED) m = merged_bits<VEC, true>(a, row0);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const T l = lp ? (T)L.e[k] : sval;
                const T r = rp ? (T)R.e[k] : sval;
                bool ok;
                const T val = elem<T, CLS>(op, l, r, ok, dm);
                if constexpr (MASKED) {
                    const bool valid = ((m >> k) & 1u) && ok && k < nrows;
                    O.e[k] = valid ? val : (T)0;
                    ob |= (uint32_t)valid << k;
                } else {
                    O.e[k] = val;
                    if (k < nrows) div0 |= !ok;
                }
            }
            if (nrows == VEC) stg_stream(op_ + v, O.v);
            else {
#pragma unroll
                for (int k = 0; k < VEC; ++k) if (k < nrows) static_cast<T*>(a.out)[row0 + k] = O.e[k];
            }
        }
        if constexpr (MASKED) store_valid_bits<VEC>(a.out_mask, row0, n, ob);
    }
    if constexpr (!MASKED && !Traits<T>::is_float && CLS == CLS_DIV) {
        if (div0) *a.div0_flag = 1u;
    }
}

template <typename T, typename TL, typename TR, typename VecT, int CLS, bool MASKED, int BLOCK, int U, int MINB = 1>
__global__ void __launch_bounds__(BLOCK, MINB) ew_binary_kernel(const __grid_constant__ EwDev a) {
    ew_binary_body<T, TL, TR, VecT, CLS, MASKED, BLOCK, U>(a);
}

// One launch, many chunks (SuperArray / SuperTable fan-out): blockIdx.y selects the descriptor, which is staged in
// shared memory so the body reads it exactly like kernel parameters.  Blocks beyond a short segment's tiles fall
// through the grid-stride loops without touching memory.
template <typename T, typename TL, typename TR, typename VecT, int CLS, bool MASKED, int BLOCK, int U, int MINB = 1>
__global__ void __launch_bounds__(BLOCK, MINB) ew_binary_batch_kernel(const EwDev* __restrict__ segs) {
    __shared__ EwDev a;
    static_assert(sizeof(EwDev) % 8 == 0, "EwDev is copied as 64-bit words");
    if (threadIdx.x < sizeof(EwDev) / 8)
        reinterpret_cast<uint64_t*>(&a)[threadIdx.x] = reinterpret_cast<const uint64_t*>(segs + blockIdx.y)[threadIdx.x];
    __syncthreads();
    ew_binary_body<T, TL, TR, VecT, CLS, MASKED, BLOCK, U>(a);
}

// FMA: out = fma(a, b, c) with one rounding (apply_fma_*, dispatch.rs:211-290; simd.rs:620,714).
template <typename T, typename VecT, bool MASKED, int BLOCK, int U>
__global__ void __launch_bounds__(BLOCK)
ew_fma_kernel(const T* __restrict__ pa, const T* __restrict__ pb, const T* __restrict__ pc,
              const uint8_t* __restrict__ mask, T* __restrict__ out, uint8_t* __restrict__ out_mask, uint64_t n) {
    constexpr int VEC = sizeof(VecT) / sizeof(T);
    const VecT* __restrict__ va = reinterpret_cast<const VecT*>(pa);
    const VecT* __restrict__ vb_ = reinterpret_cast<const VecT*>(pb);
    const VecT* __restrict__ vc = reinterpret_cast<const VecT*>(pc);
    VecT* __restrict__ vo = reinterpret_cast<VecT*>(out);
    const uint64_t nvec_ceil = (n + VEC - 1) / VEC;
    const uint64_t warps = (uint64_t)gridDim.x * (BLOCK / 32);
    const uint64_t gwarp = (uint64_t)blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    constexpr uint64_t WTILE = 32ull * U;
    const uint64_t ntiles = (n / VEC) / WTILE;
    for (uint64_t t = gwarp; t < ntiles; t += warps) {
        const uint64_t v0 = t * WTILE + lane;
        VecU<T, VecT> A[U], B[U], C[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            A[u].v = ldg_stream(va + v0 + 32ull * u);
            B[u].v = ldg_stream(vb_ + v0 + 32ull * u);
            C[u].v = ldg_stream(vc + v0 + 32ull * u);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t row0 = (v0 + 32ull * u) * VEC;
            uint32_t m = 0;
            if constexpr (MASKED) m = load_valid_bits<VEC>(mask, row0);
            VecU<T, VecT> O;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const T val = fma(A[u].e[k], B[u].e[k], C[u].e[k]);
                O.e[k] = (!MASKED || ((m >> k) & 1u)) ? val : (T)0;
            }
            stg_stream(vo + v0 + 32ull * u, O.v);
            if constexpr (MASKED) store_valid_bits<VEC>(out_mask, row0, n, m);
        }
    }
    for (uint64_t vb0 = ntiles * WTILE + gwarp * 32ull; vb0 < nvec_ceil; vb0 += warps * 32ull) {
        const uint64_t v = vb0 + lane, row0 = v * VEC;
        uint32_t m = 0;
        if constexpr (MASKED) m = load_valid_bits_guard<VEC>(mask, row0, n);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            if (row0 + k < n) {
                const T val = fma(pa[row0 + k], pb[row0 + k], pc[row0 + k]);
                out[row0 + k] = (!MASKED || ((m >> k) & 1u)) ? val : (T)0;
            }
        }
        if constexpr (MASKED) store_valid_bits<VEC>(out_mask, row0, n, m);
    }
}

}  // namespace mnr
