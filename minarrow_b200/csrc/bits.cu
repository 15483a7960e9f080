// Bitmask logic, validity merge and popcount over (mask, offset, len) windows.
//
// Replaces bitmask_binop_simd / bitmask_unop_simd / eq_mask_simd / popcount_mask_simd / all_true / all_false
// (src/kernels/bitmask/simd.rs:95-207,402-450,596-737) and Bitmask::{union,intersect,invert,count_ones,
// slice_clone} (src/structs/bitmask.rs:393-406,604-626,661-692).
//
// Three paths, like the reference's aligned-SIMD body + scalar tail plus the bit-shifted window it does not have:
//   * vector body: byte-aligned windows whose start addresses are 16/32-byte aligned move as 128/256-bit vectors,
//     4 in flight per lane, streaming loads and evict-first stores;
//   * shifted body (bits_shift_kernel): any other window start — an odd byte, an arbitrary bit offset — is read as
//     ALIGNED 128-bit vectors from the 16-byte block that holds its first byte and realigned in registers with a
//     128-bit funnel shift (two adjacent vectors -> one output vector), so an unaligned window streams at the same
//     rate as an aligned one instead of one byte per thread;
//   * byte path: the few output bytes whose aligned source vectors would leave the buffer, the last < 16 bytes and
//     the partial last byte: one output byte per thread from two adjacent input bytes.
// Output is exactly ceil(len/8) bytes with the slack bits of the last byte zero (clear_trailing_bits,
// src/kernels/bitmask/mod.rs:141-150); reads never leave ceil(total_bits/8) bytes of either input.
#include <algorithm>

#include "common.cuh"
#include "internal.h"
#include "shift_load.cuh"

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

constexpr int kBBlock = 256, kBU = kShiftU;

template <typename VecT>
__global__ void __launch_bounds__(kBBlock)
bits_op_kernel(int op, const uint8_t* __restrict__ a, uint64_t a_pos, uint64_t a_nbytes, const uint8_t* __restrict__ b,
               uint64_t b_pos, uint64_t b_nbytes, uint64_t len, uint8_t* __restrict__ out, uint64_t nvec) {
    const bool two = op <= B_XNOR;
    // ---- vector body: nvec 16-byte vectors (0 when the window is not vector-eligible) ----
    if (nvec) {
        const VecT* __restrict__ va = reinterpret_cast<const VecT*>(a + (a_pos >> 3));
        const VecT* __restrict__ vb = two ? reinterpret_cast<const VecT*>(b + (b_pos >> 3)) : nullptr;
        VecT* __restrict__ vo = reinterpret_cast<VecT*>(out);
        const uint64_t warps = (uint64_t)gridDim.x * (kBBlock / 32);
        const uint64_t gwarp = (uint64_t)blockIdx.x * (kBBlock / 32) + (threadIdx.x >> 5);
        const int lane = threadIdx.x & 31;
        constexpr uint64_t WTILE = 32ull * kBU;
        constexpr int NW = sizeof(VecT) / 8;
        union VU { VecT v; uint64_t w[NW]; };
        const uint64_t ntiles = nvec / WTILE;
        for (uint64_t t = gwarp; t < ntiles; t += warps) {
            const uint64_t v0 = t * WTILE + lane;
            VU x[kBU], y[kBU];
#pragma unroll
            for (int u = 0; u < kBU; ++u) x[u].v = ldg_stream(va + v0 + 32ull * u);
            if (two) {
#pragma unroll
                for (int u = 0; u < kBU; ++u) y[u].v = ldg_stream(vb + v0 + 32ull * u);
            }
#pragma unroll
            for (int u = 0; u < kBU; ++u) {
                VU r;
#pragma unroll
                for (int k = 0; k < NW; ++k) r.w[k] = bit_op(op, x[u].w[k], two ? y[u].w[k] : (uint64_t)0);
                stg_stream(vo + v0 + 32ull * u, r.v);
            }
        }
        for (uint64_t v = ntiles * WTILE + (uint64_t)blockIdx.x * kBBlock + threadIdx.x; v < nvec;
             v += (uint64_t)gridDim.x * kBBlock) {
            VU x, y, r;
            x.v = ldg_stream(va + v);
#pragma unroll
            for (int k = 0; k < NW; ++k) y.w[k] = 0;
            if (two) y.v = ldg_stream(vb + v);
#pragma unroll
            for (int k = 0; k < NW; ++k) r.w[k] = bit_op(op, x.w[k], y.w[k]);
            stg_stream(vo + v, r.v);
        }
    }
    // ---- byte path: output bytes [sizeof(VecT)*nvec, ceil(len/8)) ----
    const uint64_t nbytes = (len + 7) >> 3;
    for (uint64_t i = nvec * sizeof(VecT) + (uint64_t)blockIdx.x * kBBlock + threadIdx.x; i < nbytes;
         i += (uint64_t)gridDim.x * kBBlock) {
        const uint32_t x = fetch_byte(a, a_pos + 8 * i, a_nbytes);
        const uint32_t y = two ? fetch_byte(b, b_pos + 8 * i, b_nbytes) : 0u;
        uint32_t r = bit_op(op, x, y) & 0xffu;
        if (i == nbytes - 1 && (len & 7)) r &= (1u << (uint32_t)(len & 7)) - 1u;
        out[i] = (uint8_t)r;
    }
}

// Block sum -> one partial per block -> the last block to arrive (atomic ticket) adds the partials and stores the result
// (and, for the synchronous API, a second copy straight into mapped pinned host memory).  Shared by both popcount kernels.
__device__ __forceinline__ void popcount_finish(unsigned long long acc, unsigned long long* __restrict__ partials,
                                                unsigned int* __restrict__ ticket, unsigned long long* __restrict__ result,
                                                unsigned long long* __restrict__ result_host) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    __shared__ unsigned long long sm[kBBlock / 32];
    __shared__ bool is_last;
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < kBBlock / 32; ++w) t += sm[w];
        partials[blockIdx.x] = t;
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    unsigned long long t = 0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += kBBlock) {
        unsigned long long v;
        asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(partials + i));
        t += v;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long r = 0;
#pragma unroll
        for (int w = 0; w < kBBlock / 32; ++w) r += sm[w];
        *result = r;
        if (result_host) *result_host = r;   // one 8-byte store into mapped pinned host memory; the host polls the slot (api.cu popcount_sync)
        *ticket = 0;
    }
}

// Output vectors [v_lo, v_hi) through the shifted body, every other output byte through the byte path.
// POPC: count the bits of (a op b) instead of storing them (op = B_COPY for one operand, B_XOR for all_eq).
template <bool POPC>
__global__ void __launch_bounds__(kBBlock, 4)   // <= 64 registers: 4 blocks/SM (ncu r01s: 72 registers, 3 blocks, 34 % warps active, 70 % DRAM)
bits_shift_kernel(int op, const uint8_t* __restrict__ a, uint64_t a_pos, uint64_t a_nbytes, ShiftSrc sa,
                  const uint8_t* __restrict__ b, uint64_t b_pos, uint64_t b_nbytes, ShiftSrc sb, uint64_t len,
                  uint8_t* __restrict__ out, uint64_t v_lo, uint64_t v_hi, unsigned long long* __restrict__ partials,
                  unsigned int* __restrict__ ticket, unsigned long long* __restrict__ result,
                  unsigned long long* __restrict__ result_host) {
    const bool two = op <= B_XNOR;
    unsigned long long acc = 0;
    V16* __restrict__ vo = reinterpret_cast<V16*>(out);
    const uint64_t nthreads = (uint64_t)gridDim.x * kBBlock;
    const uint64_t gtid = (uint64_t)blockIdx.x * kBBlock + threadIdx.x;
    // full warp tiles: single-load + shuffle form
    constexpr uint64_t WTILE = 32ull * kBU;
    const uint64_t ntiles = (v_hi - v_lo) / WTILE;
    const uint64_t warps = nthreads / 32, gwarp = gtid / 32;
    const int lane = threadIdx.x & 31;
    for (uint64_t t = gwarp; t < ntiles; t += warps) {
        const uint64_t vt = v_lo + t * WTILE;
        V16 x[kBU], y[kBU];
        load_shifted_tile(sa, vt, lane, x);
        if (two) load_shifted_tile(sb, vt, lane, y);
#pragma unroll
        for (int u = 0; u < kBU; ++u) {
            V16 r;
            r.x = bit_op(op, x[u].x, two ? y[u].x : (uint64_t)0);
            r.y = bit_op(op, x[u].y, two ? y[u].y : (uint64_t)0);
            if constexpr (POPC) acc += (unsigned)__popcll(r.x) + (unsigned)__popcll(r.y);
            else stg_stream(vo + vt + lane + 32ull * u, r);
        }
    }
    // vectors past the last full warp tile (< 32*kBU of them): two loads per vector
    const uint64_t span = nthreads * kBU;
    for (uint64_t v0 = v_lo + ntiles * WTILE + (gtid / 32) * (32ull * kBU) + (gtid & 31); v0 < v_hi; v0 += span) {
        V16 x[kBU], y[kBU];
#pragma unroll
        for (int u = 0; u < kBU; ++u)
            if (v0 + 32ull * u < v_hi) x[u] = load_shifted(sa, v0 + 32ull * u);
        if (two) {
#pragma unroll
            for (int u = 0; u < kBU; ++u)
                if (v0 + 32ull * u < v_hi) y[u] = load_shifted(sb, v0 + 32ull * u);
        }
#pragma unroll
        for (int u = 0; u < kBU; ++u) {
            if (v0 + 32ull * u < v_hi) {
                V16 r;
                r.x = bit_op(op, x[u].x, two ? y[u].x : (uint64_t)0);
                r.y = bit_op(op, x[u].y, two ? y[u].y : (uint64_t)0);
                if constexpr (POPC) acc += (unsigned)__popcll(r.x) + (unsigned)__popcll(r.y);
                else stg_stream(vo + v0 + 32ull * u, r);
            }
        }
    }
    // byte path: output bytes [0, 16 v_lo) and [16 v_hi, ceil(len/8))
    const uint64_t nbytes = (len + 7) >> 3;
    const uint64_t head = v_lo * 16, tail0 = v_hi * 16;
    const uint64_t nedge = head + (nbytes - tail0);
    for (uint64_t e = gtid; e < nedge; e += nthreads) {
        const uint64_t i = e < head ? e : tail0 + (e - head);
        const uint32_t x = fetch_byte(a, a_pos + 8 * i, a_nbytes);
        const uint32_t y = two ? fetch_byte(b, b_pos + 8 * i, b_nbytes) : 0u;
        uint32_t r = bit_op(op, x, y) & 0xffu;
        if (i == nbytes - 1 && (len & 7)) r &= (1u << (uint32_t)(len & 7)) - 1u;
        if constexpr (POPC) acc += (unsigned)__popc(r);
        else out[i] = (uint8_t)r;
    }
    if constexpr (POPC) popcount_finish(acc, partials, ticket, result, result_host);
}

// Host side of ShiftSrc + the range of output vectors whose two aligned source vectors lie inside the buffer.
struct ShiftPlan {
    ShiftSrc src;
    uint64_t v_lo, v_hi;
};
static ShiftPlan plan_shift(const uint8_t* p, uint64_t pos, uint64_t total_bytes) {
    const uintptr_t first = reinterpret_cast<uintptr_t>(p) + (pos >> 3);
    const uintptr_t base = first & ~(uintptr_t)15;
    ShiftPlan pl;
    pl.src.base = reinterpret_cast<const V16*>(base);
    pl.src.S = (uint32_t)((first - base) * 8 + (pos & 7));
    pl.v_lo = base < reinterpret_cast<uintptr_t>(p) ? 1 : 0;
    const uintptr_t end = reinterpret_cast<uintptr_t>(p) + total_bytes;
    const uint64_t whole = end > base ? (uint64_t)(end - base) / 16 : 0;      // aligned vectors fully inside the buffer
    const uint64_t need = pl.src.S ? 1 : 0;                                  // output vector v also reads vector v + 1
    pl.v_hi = whole > need ? whole - need : 0;
    return pl;
}

static bool vec_eligible(const uint8_t* p, uint64_t pos, unsigned align = 16) {
    return (pos & 7) == 0 && ((reinterpret_cast<uintptr_t>(p) + (pos >> 3)) & (align - 1)) == 0;
}

cudaError_t launch_bits_op(int op, const uint8_t* a, uint64_t a_pos, uint64_t a_total, const uint8_t* b, uint64_t b_pos,
                           uint64_t b_total, uint64_t len, uint8_t* out, cudaStream_t s) {
    if (len == 0) return cudaSuccess;
    const bool two = op <= B_XNOR;
    auto elig = [&](unsigned al) {
        return vec_eligible(a, a_pos, al) && (!two || vec_eligible(b, b_pos, al)) && (reinterpret_cast<uintptr_t>(out) & (al - 1)) == 0;
    };
    const unsigned vbytes = elig(32) ? 32 : elig(16) ? 16 : 0;   // 256-bit vectors when every window start allows it
    const uint64_t nvec = vbytes ? (len >> 3) / vbytes : 0;      // only whole bytes fully inside the window
    const uint64_t nbytes = (len + 7) >> 3;
    if (!vbytes && (reinterpret_cast<uintptr_t>(out) & 15u) == 0 && nbytes >= 64) {
        // unaligned / bit-offset window: aligned loads + funnel shift
        const ShiftPlan pa = plan_shift(a, a_pos, (a_total + 7) >> 3);
        ShiftPlan pb = pa;
        if (two) pb = plan_shift(b, b_pos, (b_total + 7) >> 3);
        const uint64_t v_lo = std::max(pa.v_lo, pb.v_lo);
        uint64_t v_hi = std::min(std::min(pa.v_hi, pb.v_hi), (len >> 3) / 16);
        if (v_hi < v_lo) v_hi = v_lo;
        const uint64_t work = std::max<uint64_t>(((v_hi - v_lo) + kBU - 1) / kBU, 64);   // one warp tile per warp
        uint64_t blocks = (work + kBBlock - 1) / kBBlock;
        if (blocks > 0x7fffffffull) blocks = 0x7fffffffull;
        bits_shift_kernel<false><<<(unsigned)blocks, kBBlock, 0, s>>>(op, a, a_pos, (a_total + 7) >> 3, pa.src, b, b_pos,
                                                                      (b_total + 7) >> 3, pb.src, len, out, v_lo, v_hi, nullptr,
                                                                      nullptr, nullptr, nullptr);
        return cudaGetLastError();
    }
    uint64_t blocks;
    if (nvec) {
        const uint64_t tiles = (nvec + 32ull * kBU - 1) / (32ull * kBU);
        blocks = (tiles + (kBBlock / 32) - 1) / (kBBlock / 32);
    } else {
        blocks = (nbytes + kBBlock - 1) / kBBlock;
    }
    if (blocks < 1) blocks = 1;
    if (blocks > 0x7fffffffull) blocks = 0x7fffffffull;
    if (vbytes == 32)
        bits_op_kernel<V32><<<(unsigned)blocks, kBBlock, 0, s>>>(op, a, a_pos, (a_total + 7) >> 3, b, b_pos, (b_total + 7) >> 3,
                                                                  len, out, nvec);
    else
        bits_op_kernel<V16><<<(unsigned)blocks, kBBlock, 0, s>>>(op, a, a_pos, (a_total + 7) >> 3, b, b_pos, (b_total + 7) >> 3,
                                                                  len, out, nvec);
    return cudaGetLastError();
}

// Popcount of a window, optionally of (a xor b) — the latter answers all_eq (simd.rs:511-581) in one pass.
__global__ void __launch_bounds__(kBBlock)
bits_popcount_kernel(const uint8_t* __restrict__ a, uint64_t a_pos, uint64_t a_nbytes, const uint8_t* __restrict__ b,
                     uint64_t b_pos, uint64_t b_nbytes, uint64_t len, uint64_t nvec, unsigned long long* __restrict__ partials,
                     unsigned int* __restrict__ ticket, unsigned long long* __restrict__ result,
                     unsigned long long* __restrict__ result_host) {
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
    popcount_finish(acc, partials, ticket, result, result_host);
}

int popcount_max_grid() { return kSMs * 8; }

cudaError_t launch_bits_popcount(const uint8_t* a, uint64_t a_pos, uint64_t a_total, const uint8_t* b, uint64_t b_pos,
                                 uint64_t b_total, uint64_t len, unsigned long long* partials, unsigned int* ticket,
                                 unsigned long long* result, unsigned long long* result_host, cudaStream_t s) {
    if (len == 0) return cudaSuccess;
    const bool vec = vec_eligible(a, a_pos) && (!b || vec_eligible(b, b_pos));
    const uint64_t nvec = vec ? (len >> 3) / 16 : 0;
    const uint64_t nbytes = (len + 7) >> 3;
    if (!vec && nbytes >= 64) {
        const ShiftPlan pa = plan_shift(a, a_pos, (a_total + 7) >> 3);
        ShiftPlan pb = pa;
        if (b) pb = plan_shift(b, b_pos, (b_total + 7) >> 3);
        const uint64_t v_lo = std::max(pa.v_lo, pb.v_lo);
        uint64_t v_hi = std::min(std::min(pa.v_hi, pb.v_hi), (len >> 3) / 16);
        if (v_hi < v_lo) v_hi = v_lo;
        const uint64_t work = std::max<uint64_t>(((v_hi - v_lo) + kBU - 1) / kBU, 64);
        uint64_t blocks = (work + kBBlock - 1) / kBBlock;
        if (blocks > (uint64_t)kSMs * 8) blocks = (uint64_t)kSMs * 8;
        bits_shift_kernel<true><<<(unsigned)blocks, kBBlock, 0, s>>>(b ? B_XOR : B_COPY, a, a_pos, (a_total + 7) >> 3, pa.src, b, b_pos,
                                                                     (b_total + 7) >> 3, pb.src, len, nullptr, v_lo, v_hi, partials,
                                                                     ticket, result, result_host);
        return cudaGetLastError();
    }
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
                                                               len, nvec, partials, ticket, result, result_host);
    return cudaGetLastError();
}

}  // namespace mnr
