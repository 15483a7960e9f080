// Device-side consolidate / concat of chunks.
//
// SuperArray::consolidate and rechunk (src/structs/chunked/super_array.rs:674-787, src/traits/consolidate.rs:61-69)
// are built from Array::concat = MaskedArray::append_array (src/macros.rs:311-341): values are appended, validity is
// appended bit by bit at an arbitrary bit offset (Bitmask::extend_from_bitmask, src/structs/bitmask.rs:523-553), a chunk
// without a mask counts as all-valid, and the result has a mask iff any chunk had one.  Here all chunks move in two
// launches: a batched value copy (blockIdx.y = chunk) and a destination-byte-centric validity gather, so no output
// byte is written twice and chunk boundaries inside a byte need no atomics.
#include "common.cuh"
#include "internal.h"

namespace mnr {

constexpr int kKBlock = 256;

template <int ES>
__global__ void __launch_bounds__(kKBlock) concat_values_kernel(const ConcatSeg* __restrict__ segs, char* __restrict__ out) {
    const ConcatSeg s = segs[blockIdx.y];
    const uint64_t bytes = s.rows * ES;
    const char* __restrict__ src = static_cast<const char*>(s.data);
    char* __restrict__ dst = out + s.row0 * ES;
    const uint64_t tid = (uint64_t)blockIdx.x * kKBlock + threadIdx.x, nthr = (uint64_t)gridDim.x * kKBlock;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0) {
        const uint64_t nvec = bytes / 16;
        const V16* __restrict__ vs = reinterpret_cast<const V16*>(src);
        V16* __restrict__ vd = reinterpret_cast<V16*>(dst);
        uint64_t v = tid;
        for (; v + 3 * nthr < nvec; v += 4 * nthr) {   // 4 independent 128-bit loads in flight
            const V16 a = ldg_stream(vs + v), b = ldg_stream(vs + v + nthr), c = ldg_stream(vs + v + 2 * nthr), d = ldg_stream(vs + v + 3 * nthr);
            stg_stream(vd + v, a); stg_stream(vd + v + nthr, b); stg_stream(vd + v + 2 * nthr, c); stg_stream(vd + v + 3 * nthr, d);
        }
        for (; v < nvec; v += nthr) stg_stream(vd + v, ldg_stream(vs + v));
        for (uint64_t b = nvec * 16 + tid; b < bytes; b += nthr) dst[b] = src[b];
    } else {
        // element-granular copy (destination row offsets are only element-aligned in general)
        using E = typename std::conditional<ES == 8, uint64_t, typename std::conditional<ES == 4, uint32_t,
                  typename std::conditional<ES == 2, uint16_t, uint8_t>::type>::type>::type;
        const E* __restrict__ es = reinterpret_cast<const E*>(src);
        E* __restrict__ ed = reinterpret_cast<E*>(dst);
        for (uint64_t i = tid; i < s.rows; i += nthr) ed[i] = es[i];
    }
}

__device__ __forceinline__ uint32_t seg_byte(const ConcatSeg& s, uint64_t bit) {   // 8 validity bits of seg from `bit`
    if (!s.mask) return 0xffu;
    const uint64_t nbytes = (s.rows + 7) >> 3, j = bit >> 3;
    const uint32_t sh = (uint32_t)(bit & 7);
    const uint32_t lo = j < nbytes ? (uint32_t)s.mask[j] : 0u;
    if (sh == 0) return lo;
    const uint32_t hi = (j + 1 < nbytes) ? (uint32_t)s.mask[j + 1] : 0u;
    return ((lo >> sh) | (hi << (8 - sh))) & 0xffu;
}

__global__ void __launch_bounds__(kKBlock)
concat_bits_kernel(const ConcatSeg* __restrict__ segs, uint32_t nseg, uint64_t total_rows, uint8_t* __restrict__ out) {
    const uint64_t nbytes = (total_rows + 7) >> 3;
    for (uint64_t i = (uint64_t)blockIdx.x * kKBlock + threadIdx.x; i < nbytes; i += (uint64_t)gridDim.x * kKBlock) {
        const uint64_t r0 = i * 8;
        // last segment whose first row is <= r0 (segments are in row order; empty ones share a row0 and are skipped below)
        uint32_t lo = 0, hi = nseg;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (segs[mid].row0 <= r0) lo = mid; else hi = mid;
        }
        ConcatSeg s = segs[lo];
        uint32_t byte;
        if (r0 + 8 <= s.row0 + s.rows) {
            byte = seg_byte(s, r0 - s.row0);
        } else {
            byte = 0;
            uint32_t k = lo;
            for (int b = 0; b < 8; ++b) {
                const uint64_t r = r0 + b;
                if (r >= total_rows) break;
                while (r >= s.row0 + s.rows) s = segs[++k];
                const uint64_t o = r - s.row0;
                const uint32_t bit = s.mask ? ((uint32_t)s.mask[o >> 3] >> (uint32_t)(o & 7)) & 1u : 1u;
                byte |= bit << b;
            }
        }
        if (i == nbytes - 1 && (total_rows & 7)) byte &= (1u << (uint32_t)(total_rows & 7)) - 1u;
        out[i] = (uint8_t)byte;
    }
}

cudaError_t launch_concat(int elem_bytes, const ConcatSeg* segs, uint32_t nseg, uint64_t max_rows, uint64_t total_rows, void* out,
                          uint8_t* out_mask, cudaStream_t s) {
    if (nseg == 0 || total_rows == 0) return cudaSuccess;
    for (uint32_t off = 0; off < nseg; off += 65535) {
        const uint32_t cnt = nseg - off < 65535 ? nseg - off : 65535;
        uint64_t bx = (max_rows * elem_bytes / 16 + (uint64_t)kKBlock * 4 - 1) / ((uint64_t)kKBlock * 4);
        if (bx < 1) bx = 1;
        if (bx > (uint64_t)kSMs * 16) bx = (uint64_t)kSMs * 16;
        const dim3 grid((unsigned)bx, cnt, 1);
        char* o = static_cast<char*>(out);
        switch (elem_bytes) {
            case 1: concat_values_kernel<1><<<grid, kKBlock, 0, s>>>(segs + off, o); break;
            case 2: concat_values_kernel<2><<<grid, kKBlock, 0, s>>>(segs + off, o); break;
            case 4: concat_values_kernel<4><<<grid, kKBlock, 0, s>>>(segs + off, o); break;
            default: concat_values_kernel<8><<<grid, kKBlock, 0, s>>>(segs + off, o); break;
        }
    }
    if (out_mask) {
        const uint64_t nbytes = (total_rows + 7) >> 3;
        uint64_t blocks = (nbytes + kKBlock - 1) / kKBlock;
        if (blocks > (uint64_t)kSMs * 32) blocks = (uint64_t)kSMs * 32;
        concat_bits_kernel<<<(unsigned)blocks, kKBlock, 0, s>>>(segs, nseg, total_rows, out_mask);
    }
    return cudaGetLastError();
}

}  // namespace mnr
