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
#include "shift_load.cuh"

namespace mnr {

constexpr int kKBlock = 256;
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

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
        // Source and destination disagree modulo 16 bytes (a chunk appended after an odd number of rows).  Stores stay
        // ALIGNED 128-bit vectors; the source is read as aligned vectors too and realigned in registers by a byte-granular
        // funnel shift (shift_load.cuh) — the copy runs at the aligned rate for every element size instead of one
        // element per lane (0.9 TB/s for 1-byte elements).  The bytes before the first aligned destination vector, and
        // the vectors whose aligned source neighbours would leave the chunk, are copied byte by byte.
        const uintptr_t d0 = reinterpret_cast<uintptr_t>(dst);
        const uint64_t head = umin64((16 - (d0 & 15u)) & 15u, bytes);
        const uint64_t nvd = (bytes - head) / 16;                          // aligned destination vectors
        const uintptr_t sfirst = reinterpret_cast<uintptr_t>(src) + head;   // source of destination vector 0
        const uintptr_t sbase = sfirst & ~(uintptr_t)15;
        ShiftSrc sh;
        sh.base = reinterpret_cast<const V16*>(sbase);
        sh.S = (uint32_t)(sfirst - sbase) * 8u;
        const uint64_t k_lo = umin64(sbase < reinterpret_cast<uintptr_t>(src) ? 1 : 0, nvd);
        const uint64_t whole = (uint64_t)(reinterpret_cast<uintptr_t>(src) + bytes - sbase) / 16;   // aligned source vectors inside the chunk
        const uint64_t need = sh.S ? 1 : 0;
        uint64_t k_hi = umin64(whole > need ? whole - need : 0, nvd);
        if (k_hi < k_lo) k_hi = k_lo;
        V16* __restrict__ vd = reinterpret_cast<V16*>(dst + head);
        constexpr uint64_t WTILE = 32ull * kShiftU;
        const uint64_t ntiles = (k_hi - k_lo) / WTILE;
        const uint64_t warps = nthr / 32, gwarp = tid / 32;
        const int lane = threadIdx.x & 31;
        for (uint64_t t = gwarp; t < ntiles; t += warps) {
            const uint64_t kt = k_lo + t * WTILE;
            V16 x[kShiftU];
            load_shifted_tile(sh, kt, lane, x);
#pragma unroll
            for (int u = 0; u < kShiftU; ++u) stg_stream(vd + kt + lane + 32ull * u, x[u]);
        }
        for (uint64_t k = k_lo + ntiles * WTILE + tid; k < k_hi; k += nthr) stg_stream(vd + k, load_shifted(sh, k));
        // edge bytes: [0, head + 16 k_lo) and [head + 16 k_hi, bytes)
        const uint64_t e0 = head + 16 * k_lo, e1 = head + 16 * k_hi;
        for (uint64_t b = tid; b < e0 + (bytes - e1); b += nthr) {
            const uint64_t i = b < e0 ? b : e1 + (b - e0);
            dst[i] = src[i];
        }
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

// One output byte (rows [8 i, 8 i + 8)) the careful way: any number of segment boundaries inside the byte.
__device__ __forceinline__ uint32_t concat_byte(const ConcatSeg* __restrict__ segs, uint32_t nseg, uint64_t total_rows, uint64_t i) {
    const uint64_t r0 = i * 8;
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
    if (i == ((total_rows + 7) >> 3) - 1 && (total_rows & 7)) byte &= (1u << (uint32_t)(total_rows & 7)) - 1u;
    return byte;
}

// Destination-vector-centric gather: one thread = one 16-byte output vector (128 rows).  A vector that lies inside one
// segment — all but a handful — is five aligned 32-bit loads of that segment's mask and a funnel shift by the bit
// distance between source and destination (Bitmask::extend_from_bitmask walks this bit by bit, bitmask.rs:523-553); a
// vector that straddles segments, or whose aligned source words would leave the segment's mask, goes byte by byte.
__global__ void __launch_bounds__(kKBlock)
concat_bits_vec_kernel(const ConcatSeg* __restrict__ segs, uint32_t nseg, uint64_t total_rows, uint8_t* __restrict__ out) {
    const uint64_t nbytes = (total_rows + 7) >> 3;
    const uint64_t nvec = (nbytes + 15) / 16;
    for (uint64_t v = (uint64_t)blockIdx.x * kKBlock + threadIdx.x; v < nvec; v += (uint64_t)gridDim.x * kKBlock) {
        const uint64_t r0 = v * 128;
        uint32_t lo = 0, hi = nseg;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (segs[mid].row0 <= r0) lo = mid; else hi = mid;
        }
        const ConcatSeg s = segs[lo];
        bool fast = r0 + 128 <= s.row0 + s.rows && r0 + 128 <= total_rows;
        union { V16 v; uint32_t w[4]; } o;
        if (fast && s.mask) {
            const uint64_t off = r0 - s.row0;
            const uintptr_t first = reinterpret_cast<uintptr_t>(s.mask) + (off >> 3);
            const uintptr_t base = first & ~(uintptr_t)3;
            const uintptr_t end = reinterpret_cast<uintptr_t>(s.mask) + ((s.rows + 7) >> 3);
            if (base >= reinterpret_cast<uintptr_t>(s.mask) && base + 20 <= end) {
                const uint32_t S = (uint32_t)((first - base) * 8 + (off & 7));   // 0..31
                const uint8_t* bp = reinterpret_cast<const uint8_t*>(base);
                uint32_t w[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) w[k] = ldg_u32(bp + 4 * k);
#pragma unroll
                for (int k = 0; k < 4; ++k) o.w[k] = __funnelshift_r(w[k], w[k + 1], S);
            } else {
                fast = false;
            }
        } else if (fast) {
            o.w[0] = o.w[1] = o.w[2] = o.w[3] = 0xffffffffu;   // chunk without a mask = all valid
        }
        if (fast) {
            *reinterpret_cast<V16*>(out + v * 16) = o.v;
        } else {
            const uint64_t b1 = (v + 1) * 16 < nbytes ? (v + 1) * 16 : nbytes;
            for (uint64_t i = v * 16; i < b1; ++i) out[i] = (uint8_t)concat_byte(segs, nseg, total_rows, i);
        }
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
        if ((reinterpret_cast<uintptr_t>(out_mask) & 15u) == 0) {
            uint64_t blocks = ((nbytes + 15) / 16 + kKBlock - 1) / kKBlock;
            if (blocks > (uint64_t)kSMs * 32) blocks = (uint64_t)kSMs * 32;
            concat_bits_vec_kernel<<<(unsigned)blocks, kKBlock, 0, s>>>(segs, nseg, total_rows, out_mask);
        } else {
            uint64_t blocks = (nbytes + kKBlock - 1) / kKBlock;
            if (blocks > (uint64_t)kSMs * 32) blocks = (uint64_t)kSMs * 32;
            concat_bits_kernel<<<(unsigned)blocks, kKBlock, 0, s>>>(segs, nseg, total_rows, out_mask);
        }
    }
    return cudaGetLastError();
}

}  // namespace mnr
