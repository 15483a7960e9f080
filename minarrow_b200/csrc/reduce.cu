// Launchers for the null-aware aggregate kernels (reduce_kernels.cuh).
#include "internal.h"
#include "reduce_kernels.cuh"

namespace mnr {

// Library configuration, chosen from the on-device sweep in profiles/ (tools/sweep_kernels.cu):
// 256-thread blocks, >= 4 resident blocks per SM (<= 64 registers), 4 x 128-bit loads in flight per lane.
// Grid = min(tiles, 148 SMs x 4 blocks): one resident wave, every warp owns the same number of tiles.
constexpr int kRBlock = 256, kRMinB = 4, kRU = 4;

int reduce_max_grid() { return kSMs * kRMinB; }

template <typename T, typename VecT, bool MASKED, bool MINMAX>
static cudaError_t launch_one(const void* data, const uint8_t* mask, uint64_t n, AggRaw* partials,
                              unsigned int* ticket, AggRaw* out, cudaStream_t s) {
    constexpr int VEC = sizeof(VecT) / sizeof(T);
    const uint64_t nvec = n / VEC;
    const uint64_t tile = (uint64_t)kRBlock * kRU;
    // Narrow types carry 8-16 slot accumulators per lane: give them 128 registers (2 blocks/SM) instead of 64.
    constexpr int MINB = sizeof(T) >= 4 ? kRMinB : 2;
    const uint64_t cap = (uint64_t)kSMs * MINB;
    uint64_t blocks = (nvec + tile - 1) / tile;
    if (blocks < 1) blocks = 1;
    if (blocks > cap) blocks = cap;
    reduce_stats_kernel<T, VecT, MASKED, MINMAX, kRBlock, MINB, kRU>
        <<<(unsigned)blocks, kRBlock, 0, s>>>(static_cast<const T*>(data), mask, n, partials, ticket, out);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_t(const void* data, const uint8_t* mask, uint64_t n, bool minmax, AggRaw* partials,
                            unsigned int* ticket, AggRaw* out, cudaStream_t s) {
    const bool vec_ok = (reinterpret_cast<uintptr_t>(data) & 15u) == 0;
#define MNR_GO(V)                                                                                         \
    do {                                                                                                  \
        if (mask) return minmax ? launch_one<T, V, true, true>(data, mask, n, partials, ticket, out, s)   \
                                : launch_one<T, V, true, false>(data, mask, n, partials, ticket, out, s); \
        return minmax ? launch_one<T, V, false, true>(data, mask, n, partials, ticket, out, s)            \
                      : launch_one<T, V, false, false>(data, mask, n, partials, ticket, out, s);          \
    } while (0)
    if (vec_ok) MNR_GO(V16);
    MNR_GO(T);
#undef MNR_GO
}

cudaError_t launch_reduce_stats(mnr_dtype dt, const void* data, const uint8_t* mask, uint64_t n, bool minmax,
                                AggRaw* partials, unsigned int* ticket, AggRaw* out, cudaStream_t s) {
    switch (dt) {
        case MNR_I8: return launch_t<int8_t>(data, mask, n, minmax, partials, ticket, out, s);
        case MNR_U8: return launch_t<uint8_t>(data, mask, n, minmax, partials, ticket, out, s);
        case MNR_I16: return launch_t<int16_t>(data, mask, n, minmax, partials, ticket, out, s);
        case MNR_U16: return launch_t<uint16_t>(data, mask, n, minmax, partials, ticket, out, s);
        case MNR_I32: return launch_t<int32_t>(data, mask, n, minmax, partials, ticket, out, s);
        case MNR_U32: return launch_t<uint32_t>(data, mask, n, minmax, partials, ticket, out, s);
        case MNR_I64: return launch_t<int64_t>(data, mask, n, minmax, partials, ticket, out, s);
        case MNR_U64: return launch_t<uint64_t>(data, mask, n, minmax, partials, ticket, out, s);
        case MNR_F32: return launch_t<float>(data, mask, n, minmax, partials, ticket, out, s);
        case MNR_F64: return launch_t<double>(data, mask, n, minmax, partials, ticket, out, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace mnr
