// Launch dispatch for the element-wise kernels (ew_kernels.cuh).
#include "ew_kernels.cuh"

namespace mnr {

constexpr int kEBlock = 256, kEU = 4;
int g_ew_grid_cap = 0;   // 0: one warp tile per warp (grid covers the column); >0: persistent grid of that many blocks

static unsigned ew_grid(uint64_t n, int vec) {
    const uint64_t nvec = (n + vec - 1) / vec;
    const uint64_t tiles = (nvec + 32ull * kEU - 1) / (32ull * kEU);
    uint64_t blocks = (tiles + (kEBlock / 32) - 1) / (kEBlock / 32);
    if (blocks < 1) blocks = 1;
    if (g_ew_grid_cap > 0 && blocks > (uint64_t)g_ew_grid_cap) blocks = (uint64_t)g_ew_grid_cap;
    if (blocks > 0x7fffffffull) blocks = 0x7fffffffull;
    return (unsigned)blocks;
}

static EwDev to_dev(const EwArgs& a) {
    EwDev d;
    d.lhs = a.lhs; d.rhs = a.rhs; d.scalar_bits = a.scalar_bits; d.lmask = a.lmask; d.rmask = a.rmask;
    d.mask_or = a.mask_or; d.out = a.out; d.out_mask = a.out_mask; d.n = a.n; d.div0_flag = a.div0_flag; d.op = a.op;
    return d;
}

template <typename T, typename TL, typename TR, typename VecT, int CLS>
static cudaError_t go(const EwArgs& a, cudaStream_t s) {
    constexpr int VEC = sizeof(VecT) / sizeof(T);
    const bool masked = a.lmask || a.rmask;
    const unsigned grid = ew_grid(a.n, VEC);
    if (masked) ew_binary_kernel<T, TL, TR, VecT, CLS, true, kEBlock, kEU><<<grid, kEBlock, 0, s>>>(to_dev(a));
    else ew_binary_kernel<T, TL, TR, VecT, CLS, false, kEBlock, kEU><<<grid, kEBlock, 0, s>>>(to_dev(a));
    return cudaGetLastError();
}

template <typename T, typename TL, typename TR, int CLS>
static cudaError_t go_align(const EwArgs& a, cudaStream_t s) {
    constexpr int VEC = 16 / sizeof(T);
    auto ok = [](const void* p, size_t align) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % align) == 0; };
    const bool vec_ok = ok(a.lhs, sizeof(TL) * VEC) && ok(a.rhs, sizeof(TR) * VEC) && ok(a.out, 16);
    if (vec_ok) return go<T, TL, TR, V16, CLS>(a, s);
    return go<T, TL, TR, T, CLS>(a, s);
}

template <typename T>
static cudaError_t go_t(const EwArgs& a, cudaStream_t s) {
    switch (op_class(Traits<T>::is_float, a.op)) {
        case CLS_CHEAP: return go_align<T, T, T, CLS_CHEAP>(a, s);
        case CLS_DIV: return go_align<T, T, T, CLS_DIV>(a, s);
        case CLS_POW: return go_align<T, T, T, CLS_POW>(a, s);
        case CLS_REM:
            if constexpr (Traits<T>::is_float) return go_align<T, T, T, CLS_REM>(a, s);
            break;
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_ew_binary(const EwArgs& a, cudaStream_t s) {
    switch (a.dtype) {
        case MNR_I8: return go_t<int8_t>(a, s);
        case MNR_U8: return go_t<uint8_t>(a, s);
        case MNR_I16: return go_t<int16_t>(a, s);
        case MNR_U16: return go_t<uint16_t>(a, s);
        case MNR_I32: return go_t<int32_t>(a, s);
        case MNR_U32: return go_t<uint32_t>(a, s);
        case MNR_I64: return go_t<int64_t>(a, s);
        case MNR_U64: return go_t<uint64_t>(a, s);
        case MNR_F32: return go_t<float>(a, s);
        case MNR_F64: return go_t<double>(a, s);
    }
    return cudaErrorInvalidValue;
}

template <typename T, typename TL, typename TR>
static cudaError_t go_promote(const EwArgs& a, cudaStream_t s) {
    switch (op_class(true, a.op)) {
        case CLS_CHEAP: return go_align<T, TL, TR, CLS_CHEAP>(a, s);
        case CLS_DIV: return go_align<T, TL, TR, CLS_DIV>(a, s);
        case CLS_POW: return go_align<T, TL, TR, CLS_POW>(a, s);
        case CLS_REM: return go_align<T, TL, TR, CLS_REM>(a, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_ew_promote(const EwArgs& a, mnr_dtype lt, mnr_dtype rt, cudaStream_t s) {
    if (a.dtype == MNR_F64) {
        if (lt == MNR_I32 && rt == MNR_F64) return go_promote<double, int32_t, double>(a, s);
        if (lt == MNR_F64 && rt == MNR_I32) return go_promote<double, double, int32_t>(a, s);
    } else if (a.dtype == MNR_F32) {
        if (lt == MNR_I32 && rt == MNR_F32) return go_promote<float, int32_t, float>(a, s);
        if (lt == MNR_F32 && rt == MNR_I32) return go_promote<float, float, int32_t>(a, s);
    }
    return cudaErrorInvalidValue;
}

template <typename T>
static cudaError_t fma_t(const void* a, const void* b, const void* c, const uint8_t* mask, void* out, uint8_t* out_mask,
                         uint64_t n, cudaStream_t s) {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    const bool vec_ok = al(a) && al(b) && al(c) && al(out);
    const int vec = vec_ok ? 16 / (int)sizeof(T) : 1;
    const unsigned grid = ew_grid(n, vec);
    const T *pa = static_cast<const T*>(a), *pb = static_cast<const T*>(b), *pc = static_cast<const T*>(c);
    T* po = static_cast<T*>(out);
    if (vec_ok) {
        if (mask) ew_fma_kernel<T, V16, true, kEBlock, kEU><<<grid, kEBlock, 0, s>>>(pa, pb, pc, mask, po, out_mask, n);
        else ew_fma_kernel<T, V16, false, kEBlock, kEU><<<grid, kEBlock, 0, s>>>(pa, pb, pc, mask, po, out_mask, n);
    } else {
        if (mask) ew_fma_kernel<T, T, true, kEBlock, kEU><<<grid, kEBlock, 0, s>>>(pa, pb, pc, mask, po, out_mask, n);
        else ew_fma_kernel<T, T, false, kEBlock, kEU><<<grid, kEBlock, 0, s>>>(pa, pb, pc, mask, po, out_mask, n);
    }
    return cudaGetLastError();
}

cudaError_t launch_ew_fma(mnr_dtype dt, const void* a, const void* b, const void* c, const uint8_t* mask, void* out,
                          uint8_t* out_mask, uint64_t n, cudaStream_t s) {
    if (dt == MNR_F32) return fma_t<float>(a, b, c, mask, out, out_mask, n, s);
    if (dt == MNR_F64) return fma_t<double>(a, b, c, mask, out, out_mask, n, s);
    return cudaErrorInvalidValue;
}

}  // namespace mnr
