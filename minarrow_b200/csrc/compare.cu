// Typed compare -> bitmask: out bit i = ((data[i] & field_mask) == target).
// Replaces simd_eq_mask_u8 / _u16 / _u32 / _u64 (src/kernels/bitmask/simd.rs:741-788): one streaming read of the
// column, 1 bit per row written; the reference ORs sub-byte results into zeroed bytes, here a lane group assembles
// each output byte with a shuffle gather and one lane stores it (no read-modify-write, no pre-zeroing).
#include "ew_kernels.cuh"

namespace mnr {

constexpr int kCBlock = 256, kCU = 4;

// 8/16-bit elements, 32 bits at a time: (w & mask) ^ target is zero exactly in the equal lanes; the classic
// zero-lane test ((x & 0x7f..) + 0x7f.. | x has the lane's top bit set iff the lane is non-zero) turns that into one bit
// per lane, and a multiply gathers the 4 byte flags into a nibble.  ~2 instructions per row instead of ~5.
template <int ESZ> __device__ __forceinline__ uint32_t eq_bits_word(uint32_t w, uint32_t fm, uint32_t tg) {
    const uint32_t x = (w & fm) ^ tg;
    if constexpr (ESZ == 1) {
        const uint32_t nz = (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
        const uint32_t eq = (nz ^ 0x80808080u) >> 7;          // bit 0 of byte k = lane k equal
        return ((eq * 0x00204081u) >> 21) & 15u;              // bits 0, 8, 16, 24 -> bits 21..24 of the product
    } else {
        const uint32_t nz = (((x & 0x7fff7fffu) + 0x7fff7fffu) | x) & 0x80008000u;
        const uint32_t eq = (nz ^ 0x80008000u) >> 15;         // bits 0 and 16
        return (eq | (eq >> 15)) & 3u;
    }
}

template <typename T, typename VecT>
__global__ void __launch_bounds__(kCBlock)
eq_mask_kernel(const T* __restrict__ data, uint64_t n, T field_mask, T target, uint8_t* __restrict__ out) {
    constexpr int VEC = sizeof(VecT) / sizeof(T);
    const VecT* __restrict__ vp = reinterpret_cast<const VecT*>(data);
    const uint64_t nvec_ceil = (n + VEC - 1) / VEC;
    constexpr uint64_t WTILE = 32ull * kCU;
    const uint64_t ntiles = (n / VEC) / WTILE;
    const uint64_t warps = (uint64_t)gridDim.x * (kCBlock / 32);
    const uint64_t gwarp = (uint64_t)blockIdx.x * (kCBlock / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    for (uint64_t t = gwarp; t < ntiles; t += warps) {
        const uint64_t v0 = t * WTILE + lane;
        VecU<T, VecT> x[kCU];
#pragma unroll
        for (int u = 0; u < kCU; ++u) x[u].v = ldg_stream(vp + v0 + 32ull * u);
#pragma unroll
        for (int u = 0; u < kCU; ++u) {
            uint32_t bits = 0;
            if constexpr (sizeof(T) <= 2 && sizeof(VecT) >= 16) {
                constexpr int NW = sizeof(VecT) / 4, EPW = 4 / sizeof(T);
                constexpr uint32_t REP = sizeof(T) == 1 ? 0x01010101u : 0x00010001u;
                union { VecT v; uint32_t w[NW]; } p;
                p.v = x[u].v;
#pragma unroll
                for (int j = 0; j < NW; ++j)
                    bits |= eq_bits_word<sizeof(T)>(p.w[j], (uint32_t)field_mask * REP, (uint32_t)target * REP) << (j * EPW);
            } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k) bits |= (uint32_t)((T)(x[u].e[k] & field_mask) == target) << k;
            }
            store_valid_bits<VEC>(out, (v0 + 32ull * u) * VEC, n, bits);
        }
    }
    for (uint64_t vb = ntiles * WTILE + gwarp * 32ull; vb < nvec_ceil; vb += warps * 32ull) {
        const uint64_t row0 = (vb + lane) * VEC;
        uint32_t bits = 0;
#pragma unroll
        for (int k = 0; k < VEC; ++k)
            if (row0 + k < n) bits |= (uint32_t)((T)(data[row0 + k] & field_mask) == target) << k;
        store_valid_bits<VEC>(out, row0, n, bits);
    }
}

template <typename T>
static cudaError_t eq_t(const void* data, uint64_t n, uint64_t fm, uint64_t tg, uint8_t* out, cudaStream_t s) {
    const uintptr_t p = reinterpret_cast<uintptr_t>(data);
    const T* d = static_cast<const T*>(data);
    auto grid = [&](int vec) {
        const uint64_t nvec = (n + vec - 1) / vec, tiles = (nvec + 32ull * kCU - 1) / (32ull * kCU);
        uint64_t b = (tiles + kCBlock / 32 - 1) / (kCBlock / 32);
        return (unsigned)(b < 1 ? 1 : b > 0x7fffffffull ? 0x7fffffffull : b);
    };
    if ((p & 31u) == 0) eq_mask_kernel<T, V32><<<grid(32 / sizeof(T)), kCBlock, 0, s>>>(d, n, (T)fm, (T)tg, out);
    else if ((p & 15u) == 0) eq_mask_kernel<T, V16><<<grid(16 / sizeof(T)), kCBlock, 0, s>>>(d, n, (T)fm, (T)tg, out);
    else eq_mask_kernel<T, T><<<grid(1), kCBlock, 0, s>>>(d, n, (T)fm, (T)tg, out);
    return cudaGetLastError();
}

cudaError_t launch_eq_mask(int elem_bytes, const void* data, uint64_t n, uint64_t field_mask, uint64_t target, uint8_t* out,
                           cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    switch (elem_bytes) {
        case 1: return eq_t<uint8_t>(data, n, field_mask, target, out, s);
        case 2: return eq_t<uint16_t>(data, n, field_mask, target, out, s);
        case 4: return eq_t<uint32_t>(data, n, field_mask, target, out, s);
        case 8: return eq_t<uint64_t>(data, n, field_mask, target, out, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace mnr
