// Launchers for the null-aware aggregate kernels (reduce_kernels.cuh).
#include "internal.h"
#include "reduce_kernels.cuh"

namespace mnr {

// Launch geometry from the on-device sweep (tools/sweep.cu, profiles/r01c_sweep.md), 1e9-row masked i64:
//   sum + count      : 128-bit loads, 4 in flight per lane, 256 threads, 4 blocks/SM (64 regs)  -> 7.35 TB/s
//   + min / max      : 256-bit loads, 2 in flight per lane, same block shape                     -> 6.67 TB/s (128-bit: 5.6)
// Grid = min(tiles, 148 SMs x resident blocks): a function of (len, dtype, alignment tier) only — never an
// occupancy query — so a float sum is bit-reproducible run to run and device to device.
constexpr int kRBlock = 256;
template <typename T> struct RMinB { static constexpr int value = 4; };
template <typename VecT, bool MINMAX> struct RU { static constexpr int value = (sizeof(VecT) == 32) ? 2 : 4; };

int reduce_tier(const void* data, bool minmax);

template <typename T, typename VecT, bool MINMAX> static uint32_t nblk_of(uint64_t n) {
    constexpr int VEC = sizeof(VecT) / sizeof(T);
    const uint64_t nvec = n / VEC;
    const uint64_t tile = (uint64_t)kRBlock * RU<VecT, MINMAX>::value;
    const uint64_t cap = (uint64_t)kSMs * RMinB<T>::value;
    uint64_t blocks = (nvec + tile - 1) / tile;
    if (blocks < 1) blocks = 1;
    if (blocks > cap) blocks = cap;
    return (uint32_t)blocks;
}

// Reductions are launched with the programmatic-stream-serialization attribute (see pdl_* in reduce_kernels.cuh): the
// kernel after a reduction may be scheduled while the reduction drains.  The kernel itself waits for its predecessor
// (griddepcontrol.wait) before it reads anything (default) or before it touches the scratch it shares with it (flags bit
// kReduceLateWait — only when the caller vouches that the column is at rest).
template <typename T, typename VecT, bool MASKED, bool MINMAX>
static cudaError_t launch_one(const void* data, const uint8_t* mask, uint64_t n, AggRaw* partials, unsigned int* ticket,
                              AggRaw* out, AggRaw* out_host, const XchgDev& x, int flags, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nblk_of<T, VecT, MINMAX>(n), 1, 1);
    cfg.blockDim = dim3(kRBlock, 1, 1);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (flags & kReduceNoPdl) ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, reduce_stats_kernel<T, VecT, MASKED, MINMAX, kRBlock, RMinB<T>::value, RU<VecT, MINMAX>::value>,
                              static_cast<const T*>(data), mask, n, partials, ticket, out, out_host, x, flags & ~kReduceNoPdl);
}

template <typename T, typename VecT, bool MASKED, bool MINMAX>
static cudaError_t launch_batch_one(const ReduceSeg* segs, uint32_t nseg, uint32_t max_blk, AggRaw* partials,
                                    unsigned int* tickets, AggRaw* outs, const FoldArgs& f, const XchgDev& x, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(max_blk, nseg, 1);
    cfg.blockDim = dim3(kRBlock, 1, 1);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // the launches of one batched call overlap (api.cu reduce_batch_launch)
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, reduce_stats_batch_kernel<T, VecT, MASKED, MINMAX, kRBlock, RMinB<T>::value, RU<VecT, MINMAX>::value>,
                              segs, partials, tickets, outs, max_blk, f, x);
}

// (tier, masked, minmax) -> instantiation.  The 256-bit tier exists only with min/max.
#define MNR_REDUCE_DISPATCH(CALL)                                        \
    do {                                                                 \
        if (minmax) {                                                    \
            if (tier == 2) { if (masked) CALL(V32, true, true); else CALL(V32, false, true); } \
            if (tier == 1) { if (masked) CALL(V16, true, true); else CALL(V16, false, true); } \
            if (masked) CALL(T, true, true); else CALL(T, false, true);  \
        } else {                                                         \
            if (tier >= 1) { if (masked) CALL(V16, true, false); else CALL(V16, false, false); } \
            if (masked) CALL(T, true, false); else CALL(T, false, false); \
        }                                                                \
    } while (0)

template <typename T>
static cudaError_t launch_t(const void* data, const uint8_t* mask, uint64_t n, bool minmax, AggRaw* partials,
                            unsigned int* ticket, AggRaw* out, AggRaw* out_host, const XchgDev& x, int flags, cudaStream_t s) {
    const int tier = reduce_tier(data, minmax);
    const bool masked = mask != nullptr;
#define CALL1(V, M, X) return launch_one<T, V, M, X>(data, mask, n, partials, ticket, out, out_host, x, flags, s)
    MNR_REDUCE_DISPATCH(CALL1);
#undef CALL1
}

template <typename T>
static uint32_t nblk_t(uint64_t n, int tier, bool minmax) {
    if (minmax) {
        if (tier == 2) return nblk_of<T, V32, true>(n);
        if (tier == 1) return nblk_of<T, V16, true>(n);
        return nblk_of<T, T, true>(n);
    }
    return tier >= 1 ? nblk_of<T, V16, false>(n) : nblk_of<T, T, false>(n);
}

template <typename T>
static cudaError_t batch_t(int tier, bool masked, bool minmax, const ReduceSeg* segs, uint32_t nseg, uint32_t max_blk,
                           AggRaw* partials, unsigned int* tickets, AggRaw* outs, const FoldArgs& f, const XchgDev& x, cudaStream_t s) {
#define CALLB(V, M, X) return launch_batch_one<T, V, M, X>(segs, nseg, max_blk, partials, tickets, outs, f, x, s)
    MNR_REDUCE_DISPATCH(CALLB);
#undef CALLB
}

// Force-load every instantiation of this element type (CUDA loads kernels lazily, and loading one needs the device to
// be idle: a first launch issued while a peer's kernel is already spinning on this rank's flag would deadlock against
// it on a shared device).  cudaFuncGetAttributes loads the function.
template <typename T, typename VecT, bool MASKED, bool MINMAX>
static cudaError_t preload_one() {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, reduce_stats_kernel<T, VecT, MASKED, MINMAX, kRBlock, RMinB<T>::value, RU<VecT, MINMAX>::value>);
    if (e != cudaSuccess) return e;
    return cudaFuncGetAttributes(&a, reduce_stats_batch_kernel<T, VecT, MASKED, MINMAX, kRBlock, RMinB<T>::value, RU<VecT, MINMAX>::value>);
}
template <typename T>
static cudaError_t preload_t() {
    cudaError_t e = cudaSuccess;
#define PL(V, M, X) if (e == cudaSuccess) e = preload_one<T, V, M, X>()
    PL(V32, true, true); PL(V32, false, true);
    PL(V16, true, true); PL(V16, false, true); PL(V16, true, false); PL(V16, false, false);
    PL(T, true, true); PL(T, false, true); PL(T, true, false); PL(T, false, false);
#undef PL
    return e;
}

// This file is compiled once per element type (-DMNR_RED_DTYPE=<mnr_dtype code>) plus a front (-DMNR_RED_DTYPE=100)
// that switches on the runtime dtype; see the Makefile.
#define MNR_RED_DECL(NAME)                                                                                                   \
    cudaError_t reduce_single_##NAME(const void*, const uint8_t*, uint64_t, bool, AggRaw*, unsigned int*, AggRaw*, AggRaw*, \
                                     const XchgDev&, int, cudaStream_t);                                                     \
    uint32_t reduce_nblk_##NAME(uint64_t, int, bool);                                                                        \
    cudaError_t reduce_batch_##NAME(int, bool, bool, const ReduceSeg*, uint32_t, uint32_t, AggRaw*, unsigned int*, AggRaw*,  \
                                    const FoldArgs&, const XchgDev&, cudaStream_t);                                          \
    cudaError_t reduce_preload_##NAME();
#define MNR_RED_DEF(NAME, T)                                                                                                  \
    cudaError_t reduce_single_##NAME(const void* data, const uint8_t* mask, uint64_t n, bool minmax, AggRaw* partials,       \
                                     unsigned int* ticket, AggRaw* out, AggRaw* out_host, const XchgDev& x, int flags,       \
                                     cudaStream_t s) {                                                                       \
        return launch_t<T>(data, mask, n, minmax, partials, ticket, out, out_host, x, flags, s);                             \
    }                                                                                                                         \
    uint32_t reduce_nblk_##NAME(uint64_t n, int tier, bool minmax) { return nblk_t<T>(n, tier, minmax); }                    \
    cudaError_t reduce_batch_##NAME(int tier, bool masked, bool minmax, const ReduceSeg* segs, uint32_t nseg,                \
                                    uint32_t max_blk, AggRaw* partials, unsigned int* tickets, AggRaw* outs,                 \
                                    const FoldArgs& f, const XchgDev& x, cudaStream_t s) {                                   \
        return batch_t<T>(tier, masked, minmax, segs, nseg, max_blk, partials, tickets, outs, f, x, s);                      \
    }                                                                                                                         \
    cudaError_t reduce_preload_##NAME() { return preload_t<T>(); }

#if MNR_RED_DTYPE == 0
MNR_RED_DEF(i32, int32_t)
#elif MNR_RED_DTYPE == 1
MNR_RED_DEF(u32, uint32_t)
#elif MNR_RED_DTYPE == 2
MNR_RED_DEF(i64, int64_t)
#elif MNR_RED_DTYPE == 3
MNR_RED_DEF(u64, uint64_t)
#elif MNR_RED_DTYPE == 4
MNR_RED_DEF(f32, float)
#elif MNR_RED_DTYPE == 5
MNR_RED_DEF(f64, double)
#elif MNR_RED_DTYPE == 6
MNR_RED_DEF(i8, int8_t)
#elif MNR_RED_DTYPE == 7
MNR_RED_DEF(u8, uint8_t)
#elif MNR_RED_DTYPE == 8
MNR_RED_DEF(i16, int16_t)
#elif MNR_RED_DTYPE == 9
MNR_RED_DEF(u16, uint16_t)
#elif MNR_RED_DTYPE == 100
MNR_RED_DECL(i8) MNR_RED_DECL(u8) MNR_RED_DECL(i16) MNR_RED_DECL(u16) MNR_RED_DECL(i32)
MNR_RED_DECL(u32) MNR_RED_DECL(i64) MNR_RED_DECL(u64) MNR_RED_DECL(f32) MNR_RED_DECL(f64)

#define MNR_DTYPE_SWITCH(dt, FN, ...)                      \
    switch (dt) {                                          \
        case MNR_I8: return FN##_i8(__VA_ARGS__);          \
        case MNR_U8: return FN##_u8(__VA_ARGS__);          \
        case MNR_I16: return FN##_i16(__VA_ARGS__);        \
        case MNR_U16: return FN##_u16(__VA_ARGS__);        \
        case MNR_I32: return FN##_i32(__VA_ARGS__);        \
        case MNR_U32: return FN##_u32(__VA_ARGS__);        \
        case MNR_I64: return FN##_i64(__VA_ARGS__);        \
        case MNR_U64: return FN##_u64(__VA_ARGS__);        \
        case MNR_F32: return FN##_f32(__VA_ARGS__);        \
        case MNR_F64: return FN##_f64(__VA_ARGS__);        \
    }

int reduce_max_grid() { return kSMs * 4; }

// tier: 0 = element loads, 1 = 128-bit, 2 = 256-bit (only used with min/max)
int reduce_tier(const void* data, bool minmax) {
    const uintptr_t p = reinterpret_cast<uintptr_t>(data);
    if (minmax && (p & 31u) == 0) return 2;
    return (p & 15u) == 0 ? 1 : 0;
}

cudaError_t launch_reduce_stats(mnr_dtype dt, const void* data, const uint8_t* mask, uint64_t n, bool minmax,
                                AggRaw* partials, unsigned int* ticket, AggRaw* out, AggRaw* out_host, cudaStream_t s, uint32_t host_seq) {
    const XchgDev none{};
    const int flags = (int)(host_seq << kReduceHostSeqShift);
    MNR_DTYPE_SWITCH(dt, reduce_single, data, mask, n, minmax, partials, ticket, out, out_host, none, flags, s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_reduce_stats_xchg(mnr_dtype dt, const void* data, const uint8_t* mask, uint64_t n, bool minmax,
                                     AggRaw* partials, unsigned int* ticket, AggRaw* out, AggRaw* out_host,
                                     const XchgDev& x, bool late_wait, bool pdl, cudaStream_t s) {
    const int flags = pdl ? (late_wait ? kReduceLateWait : 0) : kReduceNoPdl;
    MNR_DTYPE_SWITCH(dt, reduce_single, data, mask, n, minmax, partials, ticket, out, out_host, x, flags, s);
    return cudaErrorInvalidValue;
}

uint32_t reduce_nblk(mnr_dtype dt, uint64_t n, int tier, bool minmax) {
    MNR_DTYPE_SWITCH(dt, reduce_nblk, n, tier, minmax);
    return 1;
}

cudaError_t launch_reduce_stats_batch(mnr_dtype dt, int tier, bool masked, bool minmax, const ReduceSeg* segs,
                                      uint32_t nseg, uint32_t max_blk, AggRaw* partials, unsigned int* tickets,
                                      AggRaw* outs, const FoldArgs& f, const XchgDev& x, cudaStream_t s) {
    MNR_DTYPE_SWITCH(dt, reduce_batch, tier, masked, minmax, segs, nseg, max_blk, partials, tickets, outs, f, x, s);
    return cudaErrorInvalidValue;
}

cudaError_t reduce_preload_all() {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, fold_exchange_kernel<kRBlock>);
#define PLD(NAME) if (e == cudaSuccess) e = reduce_preload_##NAME()
    PLD(i8); PLD(u8); PLD(i16); PLD(u16); PLD(i32); PLD(u32); PLD(i64); PLD(u64); PLD(f32); PLD(f64);
#undef PLD
    return e;
}

cudaError_t launch_fold_exchange(const AggRaw* outs, const FoldArgs& f, const XchgDev& x, cudaStream_t s) {
    fold_exchange_kernel<kRBlock><<<1, kRBlock, 0, s>>>(outs, f, x);
    return cudaGetLastError();
}
#else
#error "compile reduce.cu with -DMNR_RED_DTYPE=<0..9 | 100>"
#endif

}  // namespace mnr
