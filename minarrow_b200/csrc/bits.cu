// Bitmask logic, validity merge and popcount over (mask, offset, len) windows.
//
// Replaces bitmask_binop_simd / bitmask_unop_simd / eq_mask_simd / popcount_mask_simd / all_true / all_false
// (src/kernels/bitmask/simd.rs:95-207,402-450,596-737) and Bitmask::{union,intersect,invert,count_ones,
// slice_clone} (src/structs/bitmask.rs:393-406,604-626,661-692).
//
// Two paths in one kernel, like the reference's aligned-SIMD body + scalar tail:
//   * vector body: byte-aligned windows whose start addresses are 16-byte aligned move as 128-bit vectors,
//     4 in flight per lane, streaming loads and evict-first stores;
//   * byte path: everything else (arbitrary bit offsets, the last < 16 bytes, the partial last byte) is
//     produced one output byte per thread by funnel-shifting two adjacent input bytes.
// Output is exactly ceil(len/8) bytes with the slack bits of the last byte zero (clear_trailing_bits,
// src/kernels/bitmask/mod.rs:141-150); reads never leave ceil(total_bits/8) bytes of either input.
#include "common.cuh"
#include "internal.h"

namespace mnr {

enum { B_AND = 0, B_OR = 1, B_XOR = 2, B_XNOR = 3, B_NOT = 4, B_COPY = 5 };

__device__ __forceinline__ uint32_t fetch_byte(const uint8_t* __restrict__ p, uint64_t bitpos, uint64_t nbytes) {
    const uint64_t j = bitpos >> 3;
    const uint32_t s = (uint32_t)(bitpos & 7);
    const uint32_t lo = j < nbytes ? (uint32_t)p[j] : 0u;
    if (s == 0) return lo;
    const uint32_t hi = (j + 1 < nbytes) ? (uint32_t)p[j + 1] : 0u;
    return ((lo >> s) | (hi << (8 - s))) & 0xffu;
}

template <typename W> __device__ __forceinline__ W bit_op(int op, W a, W b) {
    switch (op) {
        case B_AND: return a & b;
        case B_OR: return a | b;
        case B_XOR: return a ^ b;
        case B_XNOR: return ~(a ^ b);
        case B_NOT: return ~a;
        default: return a;
    }
}

constexpr int kBBlock = 256, kBU = 4;

__global__ void __launch_bounds__(kBBlock)
bits_op_kernel(int op, const uint8_t* __restrict__ a, uint64_t a_pos, uint64_t a_nbytes, const uint8_t* __restrict__ b,
               uint64_t b_pos, uint64_t b_nbytes, uint64_t len, uint8_t* __restrict__ out, uint64_t nvec) {
    const bool two = op <= B_XNOR;
    // ---- vector body: nvec 16-byte vectors (0 when the window is not vector-eligible) ----
    if (nvec) {
        const V16* __restrict__ va = reinterpret_cast<const V16*>(a + (a_pos >> 3));
        const V16* __restrict__ vb = two ? reinterpret_cast<const V16*>(b + (b_pos >> 3)) : nullptr;
        V16* __restrict__ vo = reinterpret_cast<V16*>(out);
        const uint64_t warps = (uint64_t)gridDim.x * (kBBlock / 32);
        const uint64_t gwarp = (uint64_t)blockIdx.x * (kBBlock / 32) + (threadIdx.x >> 5);
        const int lane = threadIdx.x & 31;
        constexpr uint64_t WTILE = 32ull * kBU;
        const uint64_t ntiles = nvec / WTILE;
        for (uint64_t t = gwarp; t < ntiles; t += warps) {
            const uint64_t v0 = t * WTILE + lane;
            V16 x[kBU], y[kBU];
#pragma unroll
            for (int u = 0; u < kBU; ++u) x[u] = ldg_stream(va + v0 + 32ull * u);
            if (two) {
#pragma unroll
                for (int u = 0; u < kBU; ++u) y[u] = ldg_stream(vb + v0 + 32ull * u);
            }
#pragma unroll
            for (int u = 0; u < kBU; ++u) {
                V16 r;
                r.x = bit_op(op, x[u].x, two ? y[u].x : (uint64_t)0);
                r.y = bit_op(op, x[u].y, two ? y[u].y : (uint64_t)0);
                stg_stream(vo + v0 + 32ull * u, r);
            }
        }
        for (uint64_t v = ntiles * WTILE + (uint64_t)blockIdx.x * kBBlock + threadIdx.x; v < nvec;
             v += (uint64_t)gridDim.x * kBBlock) {
            const V16 x = ldg_stream(va + v);
            V16 y{0, 0};
            if (two) y = ldg_stream(vb + v);
            V16 r;
            r.x = bit_op(op, x.x, y.x);
            r.y = bit_op(op, x.y, y.y);
            stg_stream(vo + v, r);
        }
    }
    // ---- byte path: output bytes [16*nvec, ceil(len/8)) ----
    const uint64_t nbytes = (len + 7) >> 3;
    for (uint64_t i = nvec * 16 + (uint64_t)blockIdx.x * kBBlock + threadIdx.x; i < nbytes;
         i += (uint64_t)gridDim.x * kBBlock) {
        const uint32_t x = fetch_byte(a, a_pos + 8 * i, a_nbytes);
        const uint32_t y = two ? fetch_byte(b, b_pos + 8 * i, b_nbytes) : 0u;
        uint32_t r = bit_op(op, x, y) & 0xffu;
        if (i == nbytes - 1 && (len & 7)) r &= (1u << (uint32_t)(len & 7)) - 1u;
        out[i] = (uint8_t)r;
    }
}

static bool vec_eligible(const uint8_t* p, uint64_t pos) {
    return (pos & 7) == 0 && ((reinterpret_cast<uintptr_t>(p) + (pos >> 3)) & 15u) == 0;
}

cudaError_t launch_bits_op(int op, const uint8_t* a, uint64_t a_pos, uint64_t a_total, const uint8_t* b, uint64_t b_pos,
                           uint64_t b_total, uint64_t len, uint8_t* out, cudaStream_t s) {
    if (len == 0) return cudaSuccess;
    const bool two = op <= B_XNOR;
    const bool vec = vec_eligible(a, a_pos) && (!two || vec_eligible(b, b_pos)) &&
                     (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    const uint64_t nvec = vec ? (len >> 3) / 16 : 0;   // only whole bytes fully inside the window
    const uint64_t nbytes = (len + 7) >> 3;
    uint64_t blocks;
    if (nvec) {
        const uint64_t tiles = (nvec + 32ull * kBU - 1) / (32ull * kBU);
        blocks = (tiles + (kBBlock / 32) - 1) / (kBBlock / 32);
    } else {
        blocks = (nbytes + kBBlock - 1) / kBBlock;
    }
    if (blocks < 1) blocks = 1;
    if (blocks > (uint64_t)kSMs * 64) blocks = (uint64_t)kSMs * 64;
    bits_op_kernel<<<(unsigned)blocks, kBBlock, 0, s>>>(op, a, a_pos, (a_total + 7) >> 3, b, b_pos, (b_total + 7) >> 3,
                                                         len, out, nvec);
    return cudaGetLastError();
}

// Popcount of a window, optionally of (a xor b) — the latter answers all_eq (simd.rs:511-581) in one pass.
__global__ void __launch_bounds__(kBBlock)
bits_popcount_kernel(const uint8_t* __restrict__ a, uint64_t a_pos, uint64_t a_nbytes, const uint8_t* __restrict__ b,
                     uint64_t b_pos, uint64_t b_nbytes, uint64_t len, uint64_t nvec, unsigned long long* __restrict__ result) {
    const bool two = b != nullptr;
    unsigned long long acc = 0;
    if (nvec) {
        const V16* __restrict__ va = reinterpret_cast<const V16*>(a + (a_pos >> 3));
        const V16* __restrict__ vb = two ? reinterpret_cast<const V16*>(b + (b_pos >> 3)) : nullptr;
        const uint64_t warps = (uint64_t)gridDim.x * (kBBlock / 32);
        const uint64_t gwarp = (uint64_t)blockIdx.x * (kBBlock / 32) + (threadIdx.x >> 5);
        const int lane = threadIdx.x & 31;
        constexpr uint64_t WTILE = 32ull * kBU;
        const uint64_t ntiles = nvec / WTILE;
        for (uint64_t t = gwarp; t < ntiles; t += warps) {
            const uint64_t v0 = t * WTILE + lane;
            V16 x[kBU], y[kBU];
#pragma unroll
            for (int u = 0; u < kBU; ++u) x[u] = ldg_stream(va + v0 + 32ull * u);
            if (two) {
#pragma unroll
                for (int u = 0; u < kBU; ++u) y[u] = ldg_stream(vb + v0 + 32ull * u);
            }
#pragma unroll
            for (int u = 0; u < kBU; ++u) {
                const uint64_t p = two ? (x[u].x ^ y[u].x) : x[u].x, q = two ? (x[u].y ^ y[u].y) : x[u].y;
                acc += (unsigned)__popcll(p) + (unsigned)__popcll(q);
            }
        }
        for (uint64_t v = ntiles * WTILE + (uint64_t)blockIdx.x * kBBlock + threadIdx.x; v < nvec;
             v += (uint64_t)gridDim.x * kBBlock) {
            V16 x = ldg_stream(va + v);
            if (two) { const V16 y = ldg_stream(vb + v); x.x ^= y.x; x.y ^= y.y; }
            acc += (unsigned)__popcll(x.x) + (unsigned)__popcll(x.y);
        }
    }
    const uint64_t nbytes = (len + 7) >> 3;
    for (uint64_t i = nvec * 16 + (uint64_t)blockIdx.x * kBBlock + threadIdx.x; i < nbytes;
         i += (uint64_t)gridDim.x * kBBlock) {
        uint32_t x = fetch_byte(a, a_pos + 8 * i, a_nbytes);
        if (two) x ^= fetch_byte(b, b_pos + 8 * i, b_nbytes);
        if (i == nbytes - 1 && (len & 7)) x &= (1u << (uint32_t)(len & 7)) - 1u;
        acc += (unsigned)__popc(x);
    }
    // block reduce (integer: order irrelevant) -> one atomic per block
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    __shared__ unsigned long long sm[kBBlock / 32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < kBBlock / 32; ++w) t += sm[w];
        if (t) atomicAdd(result, t);
    }
}

cudaError_t launch_bits_popcount(const uint8_t* a, uint64_t a_pos, uint64_t a_total, const uint8_t* b, uint64_t b_pos,
                                 uint64_t b_total, uint64_t len, unsigned long long* result, cudaStream_t s) {
    if (len == 0) return cudaSuccess;
    const bool vec = vec_eligible(a, a_pos) && (!b || vec_eligible(b, b_pos));
    const uint64_t nvec = vec ? (len >> 3) / 16 : 0;
    const uint64_t nbytes = (len + 7) >> 3;
    uint64_t blocks;
    if (nvec) {
        const uint64_t tiles = (nvec + 32ull * kBU - 1) / (32ull * kBU);
        blocks = (tiles + (kBBlock / 32) - 1) / (kBBlock / 32);
    } else {
        blocks = (nbytes + kBBlock - 1) / kBBlock;
    }
    if (blocks < 1) blocks = 1;
    if (blocks > (uint64_t)kSMs * 8) blocks = (uint64_t)kSMs * 8;
    bits_popcount_kernel<<<(unsigned)blocks, kBBlock, 0, s>>>(a, a_pos, (a_total + 7) >> 3, b, b_pos, (b_total + 7) >> 3,
                                                               len, nvec, result);
    return cudaGetLastError();
}

}  // namespace mnr
