// Shared device/host helpers for the minarrow_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/minarrow_b200.h"

namespace mnr {

// ---- streaming global loads / stores ---------------------------------------------------------------
// Every column is touched exactly once per kernel, so loads bypass L1 allocation (read-only path) and
// stores are evict-first: the 126 MB L2 is not polluted with data that will not be re-read.

struct alignas(16) V16 { uint64_t x, y; };                 // one 128-bit vector
struct alignas(32) V32 { uint64_t x, y, z, w; };           // one 256-bit vector (LDG.256, sm_100+)

// MNR_LD_POLICY / MNR_ST_POLICY select the cache hints (tools/sweep.cu builds one binary per combination;
// the library default is the winner of that sweep, profiles/).
#ifndef MNR_LD_POLICY
#define MNR_LD_POLICY 0
#endif
#ifndef MNR_ST_POLICY
#define MNR_ST_POLICY 0
#endif
#if MNR_LD_POLICY == 0
#define MNR_LD "ld.global.nc.L1::no_allocate"
#elif MNR_LD_POLICY == 1
#define MNR_LD "ld.global.nc.L1::no_allocate.L2::256B"
#elif MNR_LD_POLICY == 2
#define MNR_LD "ld.global.nc"
#else
#define MNR_LD "ld.global.nc.L1::evict_first.L2::256B"
#endif
#if MNR_ST_POLICY == 0
#define MNR_ST "st.global.cs"
#elif MNR_ST_POLICY == 1
#define MNR_ST "st.global"
#else
#define MNR_ST "st.global.L1::no_allocate"
#endif
__device__ __forceinline__ V16 ldg_stream(const V16* p) {
    V16 r;
    asm volatile(MNR_LD ".v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ V32 ldg_stream(const V32* p) {
    V32 r;
    asm volatile(MNR_LD ".v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.x), "=l"(r.y), "=l"(r.z), "=l"(r.w) : "l"(p));
    return r;
}
// Scalar fallback (pointer not 16-byte aligned: an ArrayV window at an odd offset) — the analogue of the
// reference's alignment check falling back to its scalar body (dispatch.rs:86,108-111).
template <typename S> __device__ __forceinline__ S ldg_stream(const S* p) { return *p; }
template <typename S> __device__ __forceinline__ void stg_stream(S* p, const S& v) { *p = v; }
__device__ __forceinline__ void stg_stream(V16* p, const V16& v) {
    asm volatile(MNR_ST ".v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream(V32* p, const V32& v) {
    asm volatile(MNR_ST ".v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v.x), "l"(v.y), "l"(v.z), "l"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint32_t ldg_u8(const uint8_t* p) {
    uint32_t r;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_u16(const uint8_t* p) {
    uint32_t r;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_u32(const uint8_t* p) {
    uint32_t r;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint64_t ldg_u64(const uint64_t* p) {
    uint64_t r;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}

// Validity bits of rows [row0, row0 + NBITS) as the low NBITS bits of the result.  NBITS ∈ {2,4,8,16,32}
// and row0 is a multiple of NBITS, so the bits never straddle the loaded unit.  Byte-granular loads keep
// every access inside ceil(len/8) bytes and need no alignment of the mask pointer (Arrow validity buffers
// handed over by a foreign producer are only guaranteed byte addressability).
template <int NBITS>
__device__ __forceinline__ uint32_t load_valid_bits(const uint8_t* __restrict__ mask, uint64_t row0) {
    if constexpr (NBITS <= 8) {
        uint32_t b = ldg_u8(mask + (row0 >> 3));
        return (b >> (uint32_t)(row0 & 7)) & ((1u << NBITS) - 1u);
    } else if constexpr (NBITS == 16) {
        const uint8_t* p = mask + (row0 >> 3);
        return ldg_u8(p) | (ldg_u8(p + 1) << 8);
    } else {
        const uint8_t* p = mask + (row0 >> 3);
        // 32 rows per lane = 1-byte elements in a 256-bit vector: one 32-bit load when the address allows it (warp-uniform:
        // row0 is a multiple of 32).  Measured (r01v, i8 one-mask add): 6.9 TB/s with it, 4.9 without; the same shortcut for
        // the 16-row case made 2-byte columns SLOWER (7.07 -> 5.0 TB/s), so that case keeps its two byte loads.
        if ((reinterpret_cast<uintptr_t>(p) & 3u) == 0) return ldg_u32(p);
        return ldg_u8(p) | (ldg_u8(p + 1) << 8) | (ldg_u8(p + 2) << 16) | (ldg_u8(p + 3) << 24);
    }
}

// The low 4 (1-byte elements) or 2 (2-byte elements) validity bits of `bits` -> a 32-bit word with 0xFF / 0xFFFF in
// the lanes of valid elements (SIMD-in-register paths for 8/16-bit columns).  Higher bits of `bits` are ignored.
template <int ESZ> __device__ __forceinline__ uint32_t expand_valid_word(uint32_t bits) {
    if constexpr (ESZ == 1) return (((bits & 15u) * 0x00204081u) & 0x01010101u) * 0xFFu;
    else return (((bits & 3u) * 0x00008001u) & 0x00010001u) * 0xFFFFu;
}

__device__ __forceinline__ bool row_valid(const uint8_t* __restrict__ mask, uint64_t row) {
    return (ldg_u8(mask + (row >> 3)) >> (uint32_t)(row & 7)) & 1u;
}

// ---- dtype traits ------------------------------------------------------------------------------------
template <typename T> struct Traits;
#define MNR_TRAITS(T, ACC, ISF, ISS)                                        \
    template <> struct Traits<T> {                                          \
        using Acc = ACC;                                                    \
        static constexpr bool is_float = ISF;                               \
        static constexpr bool is_signed = ISS;                              \
    };
MNR_TRAITS(int8_t, int64_t, false, true)
MNR_TRAITS(uint8_t, uint64_t, false, false)
MNR_TRAITS(int16_t, int64_t, false, true)
MNR_TRAITS(uint16_t, uint64_t, false, false)
MNR_TRAITS(int32_t, int64_t, false, true)
MNR_TRAITS(uint32_t, uint64_t, false, false)
MNR_TRAITS(int64_t, int64_t, false, true)
MNR_TRAITS(uint64_t, uint64_t, false, false)
MNR_TRAITS(float, double, true, true)
MNR_TRAITS(double, double, true, true)
#undef MNR_TRAITS

// Fixed launch geometry.  These are constants (not occupancy queries) so that the order of a float
// sum — and therefore its bits — depends only on (len, dtype), never on the device or driver.
constexpr int kSMs = 148;            // B200: 2 dies x 74 SMs

static inline size_t dtype_size(mnr_dtype d) {
    switch (d) {
        case MNR_I8: case MNR_U8: return 1;
        case MNR_I16: case MNR_U16: return 2;
        case MNR_I32: case MNR_U32: case MNR_F32: return 4;
        default: return 8;
    }
}

}  // namespace mnr
