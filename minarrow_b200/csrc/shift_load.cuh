// Realigning loads: a bit (or byte) stream that starts anywhere, read as ALIGNED 128-bit vectors and shifted into place
// in registers.  Shared by the bitmask kernels (bits.cu: windows at arbitrary bit offsets) and the device consolidate
// (concat.cu: chunks appended at arbitrary element offsets).
#pragma once
#include "common.cuh"

namespace mnr {

constexpr int kShiftU = 4;   // vectors in flight per lane per operand

// A window that starts at bit `pos` of buffer `p` is described by the 16-byte block holding its first byte (`base`) and
// the bit distance S in [0, 128) from the start of that block to the window's first bit.  Output vector v (window bits
// [128 v, 128 v + 128)) = bits [S, S + 128) of the aligned vectors v and v + 1.
struct ShiftSrc {
    const V16* base;
    uint32_t S;
};

__device__ __forceinline__ V16 ldg_v16_cached(const V16* p) {   // L1-allocating: the neighbouring lane re-reads this vector
    V16 r;
    asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}

// Bits [S, S + 128) of the 256-bit pair (cur, nxt).
__device__ __forceinline__ V16 funnel128(const V16& cur, const V16& nxt, uint32_t S) {
    union { V16 v[2]; uint32_t w[8]; } u;
    u.v[0] = cur;
    u.v[1] = nxt;
    const uint32_t r = S & 31u;
    union { V16 v; uint32_t w[4]; } o;
    switch (S >> 5) {   // warp-uniform
        case 0:
#pragma unroll
            for (int k = 0; k < 4; ++k) o.w[k] = __funnelshift_r(u.w[k], u.w[k + 1], r);
            break;
        case 1:
#pragma unroll
            for (int k = 0; k < 4; ++k) o.w[k] = __funnelshift_r(u.w[k + 1], u.w[k + 2], r);
            break;
        case 2:
#pragma unroll
            for (int k = 0; k < 4; ++k) o.w[k] = __funnelshift_r(u.w[k + 2], u.w[k + 3], r);
            break;
        default:
#pragma unroll
            for (int k = 0; k < 4; ++k) o.w[k] = __funnelshift_r(u.w[k + 3], u.w[k + 4], r);
            break;
    }
    return o.v;
}

__device__ __forceinline__ V16 shfl_down1(const V16& v) {
    V16 r;
    r.x = __shfl_down_sync(0xffffffffu, (unsigned long long)v.x, 1);
    r.y = __shfl_down_sync(0xffffffffu, (unsigned long long)v.y, 1);
    return r;
}
__device__ __forceinline__ V16 shfl_lane0(const V16& v) {
    V16 r;
    r.x = __shfl_sync(0xffffffffu, (unsigned long long)v.x, 0);
    r.y = __shfl_sync(0xffffffffu, (unsigned long long)v.y, 0);
    return r;
}

// One warp tile = 32*kShiftU consecutive output vectors starting at vt.  Every aligned source vector is loaded ONCE
// (streaming, coalesced); a lane takes the vector that follows its own from the next lane with a shuffle — lane 31 from
// lane 0's vector of the next group, the last group from one extra vector that lane 0 loads.
__device__ __forceinline__ void load_shifted_tile(const ShiftSrc& s, uint64_t vt, int lane, V16 (&out)[kShiftU]) {
    V16 x[kShiftU + 1];
#pragma unroll
    for (int u = 0; u < kShiftU; ++u) x[u] = ldg_stream(s.base + vt + lane + 32ull * u);
    if (s.S == 0) {
#pragma unroll
        for (int u = 0; u < kShiftU; ++u) out[u] = x[u];
        return;
    }
    x[kShiftU] = V16{0, 0};
    if (lane == 0) x[kShiftU] = ldg_stream(s.base + vt + 32ull * kShiftU);
#pragma unroll
    for (int u = 0; u < kShiftU; ++u) {
        V16 nxt = shfl_down1(x[u]);
        const V16 wrap = shfl_lane0(x[u + 1]);
        if (lane == 31) nxt = wrap;
        out[u] = funnel128(x[u], nxt, s.S);
    }
}

__device__ __forceinline__ V16 load_shifted(const ShiftSrc& s, uint64_t v) {
    union { V16 v[2]; uint32_t w[8]; } u;
    u.v[0] = ldg_v16_cached(s.base + v);
    if (s.S) u.v[1] = ldg_v16_cached(s.base + v + 1);
    else u.v[1] = V16{0, 0};
    const uint32_t r = s.S & 31u;
    union { V16 v; uint32_t w[4]; } o;
    switch (s.S >> 5) {   // warp-uniform
        case 0:
#pragma unroll
            for (int k = 0; k < 4; ++k) o.w[k] = __funnelshift_r(u.w[k], u.w[k + 1], r);
            break;
        case 1:
#pragma unroll
            for (int k = 0; k < 4; ++k) o.w[k] = __funnelshift_r(u.w[k + 1], u.w[k + 2], r);
            break;
        case 2:
#pragma unroll
            for (int k = 0; k < 4; ++k) o.w[k] = __funnelshift_r(u.w[k + 2], u.w[k + 3], r);
            break;
        default:
#pragma unroll
            for (int k = 0; k < 4; ++k) o.w[k] = __funnelshift_r(u.w[k + 3], u.w[k + 4], r);
            break;
    }
    return o.v;
}

}  // namespace mnr
