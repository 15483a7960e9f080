// Launch dispatch for the element-wise kernels (ew_kernels.cuh).
//
// This file is compiled once per element type (-DMNR_EW_DTYPE=<mnr_dtype code>, see the Makefile) so the template
// instantiations build in parallel; -DMNR_EW_DTYPE=100 builds the type-independent front (dtype switch, promotion,
// FMA).  Launch geometry comes from the on-device sweep (tools/sweep.cu, profiles/r01c_sweep.md):
//   * cheap ops (add/sub/mul, incl. scalar broadcast): 256-bit vectors, 4 in flight per operand, 128-thread blocks,
//     one warp tile per warp (grid covers the column)                       -> f64 masked add 7.08-7.12 TB/s
//   * div / floordiv / rem / pow (ALU-heavy): 128-bit vectors, 2 in flight, 256 threads, <= 64 registers
//     (4 blocks/SM), persistent grid of one resident wave                  -> f64 masked div 6.5 TB/s
#include "ew_kernels.cuh"

namespace mnr {

// MINB = 4 caps the kernel at 128 registers (4 blocks/SM).  1-byte columns carry the packed SIMD-in-register path (~150
// registers uncapped) and gain from the cap; 4/8-byte kernels fit 128 registers on their own, the cap only matters for
// their batched variants (C5 scalar broadcast 6.30 -> 6.77 TB/s, r01j -> r01s); 2-byte columns LOSE with it (r01v:
// two-mask add 7.07 TB/s uncapped, 5.0 capped), so they stay uncapped.
template <int ESZ> struct CfgCheap { static constexpr int BLOCK = 128, U = 4, MINB = ESZ == 2 ? 1 : 4; static constexpr bool RESIDENT = false; using Wide = V32; };
struct CfgHeavy { static constexpr int BLOCK = 256, U = 2, MINB = 4; static constexpr bool RESIDENT = true; using Wide = V16; };
template <int CLS, int ESZ> struct CfgOf { using type = CfgHeavy; };
template <int ESZ> struct CfgOf<CLS_CHEAP, ESZ> { using type = CfgCheap<ESZ>; };
template <int ESZ> struct CfgOf<CLS_SDIV, ESZ> { using type = CfgCheap<ESZ>; };   // a multiply-high per row: memory-bound like add
// 64-bit column / scalar is ~115 instructions per row (a 64 x 64 -> 128-bit signed high multiply out of 32-bit IMADs): with
// CfgCheap's 128 registers only 16 warps per SM are resident and the issue slots sit at 51 % (ncu r01zz: long-scoreboard
// + fixed-latency stalls, DRAM 63 %).  Half the loads in flight and <= 85 registers (6 blocks/SM) is worth +9 % on the
// masked kernels (i64 5.08 -> 5.52 TB/s, FloorDiv 4.49 -> 4.93; tools/sdiv64_exp.py, profiles/r01zz_sdiv64_exp.txt); the
// dense kernels lose 5 % with it and keep CfgCheap.  Also measured and dropped: 256 thr x 2 x 128-bit at <= 64 registers
// (resident grid 5.05, covering grid 4.25) and 256 thr x 2 x 256-bit at <= 85 registers (5.30).
struct CfgSdiv64 { static constexpr int BLOCK = 128, U = 2, MINB = 6; static constexpr bool RESIDENT = false; using Wide = V32; };
// Float Div / FloorDiv: the division sequence (~10 FMA-pipe instructions + a range check per row) sits between the cheap
// and the heavy classes, and the best geometry depends on how many streams a row touches (tools/fdiv_exp.py,
// profiles/r01zz_fdiv_exp.txt, 1 GiB per operand, GB/s; CfgHeavy = 256 thr x 2 x 128-bit, <= 64 regs, resident grid):
//                          CfgHeavy   CfgFdiv2   CfgFdiv3
//   f64 scalar, masked       5 469      5 573      6 149      f32: 5 264 / 5 879 / 6 242
//   f64 scalar, dense        6 109      6 090      6 070      f32: 6 049 / 6 537 / 6 209
//   f64 two masks            6 555      6 845      5 939      f32: 6 453 / 6 365 / 6 769
//   f64 dense                6 367      7 038      6 258      f32: 6 371 / 6 965 / 6 150
// (also measured: CfgCheap's 128 thr x 4 x 256-bit — never the best; CfgHeavy with a covering grid — always the worst.)
struct CfgFdiv2 { static constexpr int BLOCK = 128, U = 2, MINB = 6; static constexpr bool RESIDENT = false; using Wide = V32; };
struct CfgFdiv3 { static constexpr int BLOCK = 256, U = 2, MINB = 3; static constexpr bool RESIDENT = true; using Wide = V32; };
// f64 Power (fastpow.h: ~55 FP64 operations per row, long dependent chains, 16-byte values): the 85-register geometries spill
// (72 bytes of stack at U = 2 x 256-bit).  Candidates without spills: one vector in flight, or 128 registers.
struct CfgPow4 { static constexpr int BLOCK = 128, U = 1, MINB = 6; static constexpr bool RESIDENT = false; using Wide = V32; };
struct CfgPow5 { static constexpr int BLOCK = 128, U = 2, MINB = 4; static constexpr bool RESIDENT = false; using Wide = V32; };


static EwDev to_dev(const EwArgs& a) {
    EwDev d;
    d.lhs = a.lhs; d.rhs = a.rhs; d.scalar_bits = a.scalar_bits; d.lmask = a.lmask; d.rmask = a.rmask;
    d.mask_or = a.mask_or; d.out = a.out; d.out_mask = a.out_mask; d.n = a.n; d.div0_flag = a.div0_flag; d.op = a.op;
    d.sdiv = a.sdiv; d.magic.m = a.magic_m; d.magic.s1 = a.magic_s1; d.magic.s2 = a.magic_s2;
    return d;
}

template <int BLOCK, int U, int MINB, bool RESIDENT>
static unsigned ew_grid(uint64_t n, int vec, int grid_cap = 0) {
    const uint64_t nvec = (n + vec - 1) / vec;
    const uint64_t tiles = (nvec + 32ull * U - 1) / (32ull * U);
    uint64_t blocks = (tiles + (BLOCK / 32) - 1) / (BLOCK / 32);
    if (blocks < 1) blocks = 1;
    if (RESIDENT && blocks > (uint64_t)kSMs * MINB) blocks = (uint64_t)kSMs * MINB;
    if (grid_cap > 0 && blocks > (uint64_t)grid_cap) blocks = (uint64_t)grid_cap;
    if (blocks > 0x7fffffffull) blocks = 0x7fffffffull;
    return (unsigned)blocks;
}

template <typename T, typename TL, typename TR, typename VecT, int CLS, typename Cfg = typename CfgOf<CLS, (int)sizeof(T)>::type>
static cudaError_t go(const EwArgs& a, cudaStream_t s) {
    constexpr int VEC = sizeof(VecT) / sizeof(T);
    const bool masked = a.lmask || a.rmask;
    const unsigned grid = ew_grid<Cfg::BLOCK, Cfg::U, Cfg::MINB, Cfg::RESIDENT>(a.n, VEC, a.k.grid_cap);
    if (masked) ew_binary_kernel<T, TL, TR, VecT, CLS, true, Cfg::BLOCK, Cfg::U, Cfg::MINB><<<grid, Cfg::BLOCK, 0, s>>>(to_dev(a));
    else ew_binary_kernel<T, TL, TR, VecT, CLS, false, Cfg::BLOCK, Cfg::U, Cfg::MINB><<<grid, Cfg::BLOCK, 0, s>>>(to_dev(a));
    return cudaGetLastError();
}

template <typename T, typename VecT, int CLS>
static cudaError_t go_batch(bool masked, const EwDev* segs, uint32_t nseg, uint64_t max_n, int grid_cap, cudaStream_t s) {
    using Cfg = typename CfgOf<CLS, (int)sizeof(T)>::type;
    constexpr int VEC = sizeof(VecT) / sizeof(T);
    const dim3 grid(ew_grid<Cfg::BLOCK, Cfg::U, Cfg::MINB, Cfg::RESIDENT>(max_n, VEC, grid_cap), nseg, 1);
    if (masked) ew_binary_batch_kernel<T, T, T, VecT, CLS, true, Cfg::BLOCK, Cfg::U, Cfg::MINB><<<grid, Cfg::BLOCK, 0, s>>>(segs);
    else ew_binary_batch_kernel<T, T, T, VecT, CLS, false, Cfg::BLOCK, Cfg::U, Cfg::MINB><<<grid, Cfg::BLOCK, 0, s>>>(segs);
    return cudaGetLastError();
}

// tier 2 = the class's wide vector (256-bit for cheap ops), tier 1 = 128-bit.  Unaligned items are launched one by one.
template <typename T, int CLS>
static cudaError_t go_batch_tier(int tier, bool masked, const EwDev* segs, uint32_t nseg, uint64_t max_n, int grid_cap, cudaStream_t s) {
    using Wide = typename CfgOf<CLS, (int)sizeof(T)>::type::Wide;
    if (tier == 2) return go_batch<T, Wide, CLS>(masked, segs, nseg, max_n, grid_cap, s);
    return go_batch<T, V16, CLS>(masked, segs, nseg, max_n, grid_cap, s);
}

template <typename T>
static cudaError_t go_batch_t(int op, int tier, bool masked, bool sdiv, const EwDev* segs, uint32_t nseg, uint64_t max_n, int grid_cap, cudaStream_t s) {
    if constexpr (!Traits<T>::is_float) {
        if (sdiv) return go_batch_tier<T, CLS_SDIV>(tier, masked, segs, nseg, max_n, grid_cap, s);
    }
    switch (op_class(Traits<T>::is_float, op)) {
        case CLS_CHEAP: return go_batch_tier<T, CLS_CHEAP>(tier, masked, segs, nseg, max_n, grid_cap, s);
        case CLS_DIV: return go_batch_tier<T, CLS_DIV>(tier, masked, segs, nseg, max_n, grid_cap, s);
        case CLS_POW: return go_batch_tier<T, CLS_POW>(tier, masked, segs, nseg, max_n, grid_cap, s);
        case CLS_REM:
            if constexpr (Traits<T>::is_float) return go_batch_tier<T, CLS_REM>(tier, masked, segs, nseg, max_n, grid_cap, s);
            break;
    }
    return cudaErrorInvalidValue;
}

// Alignment tiers, like the reference's "64-byte aligned -> SIMD body, else scalar body" (dispatch.rs:86,108-111):
// widest vector every operand pointer allows, else 128-bit, else element-wise loads.
template <typename T, typename TL, typename TR, int CLS, typename Cfg = typename CfgOf<CLS, (int)sizeof(T)>::type>
static cudaError_t go_align(const EwArgs& a, cudaStream_t s) {
    using Wide = typename Cfg::Wide;
    auto ok = [](const void* p, size_t align) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % align) == 0; };
    constexpr int VW = sizeof(Wide) / sizeof(T);
    if (sizeof(Wide) > 16 && a.k.max_tier >= 2 && ok(a.lhs, sizeof(TL) * VW) && ok(a.rhs, sizeof(TR) * VW) && ok(a.out, sizeof(Wide)))
        return go<T, TL, TR, Wide, CLS, Cfg>(a, s);
    constexpr int V = 16 / sizeof(T);
    if (ok(a.lhs, sizeof(TL) * V) && ok(a.rhs, sizeof(TR) * V) && ok(a.out, 16)) return go<T, TL, TR, V16, CLS, Cfg>(a, s);
    return go<T, TL, TR, T, CLS, Cfg>(a, s);
}

// Integer Div / Rem / FloorDiv of two columns, float Rem, Power.  CfgHeavy was tuned on f64 division in r01c; with the
// same two alternatives as above the wider types and Power gain 4-18 % (tools/heavy_exp.py, profiles/r01zz_heavy_exp.txt,
// 1 GiB per operand, GB/s, CfgHeavy / CfgFdiv2 / CfgFdiv3):
//   u64 div two masks 6 187 / 6 988 / 6 634     i64 div two masks 5 307 / 5 789 / 5 854     i64 div dense 6 202 / 6 740 / 5 756
//   u32 div two masks 5 544 / 5 847 / 5 556     i32 div two masks 5 039 / 5 050 / 5 214     i32 div dense 5 784 / 6 243 / 5 963
//   i16 div two masks 4 735 / 4 352 / 4 170     i8 floordiv two masks 2 041 / 2 145 / 1 961
//   f64 rem two masks 6 141 / 6 701 / 6 387     f32 rem two masks 5 792 / 5 596 / 5 451
//   pow two masks: i64 5 845 / 6 877 / 6 453, u32 5 538 / 5 803 / 5 555, i16 2 468 / 2 810 / 2 512, f32 5 883 / 6 764 / 6 036,
//                  f64 3 675 / 4 106 / 3 930
// ew_heavy_cfg: 0 = the choice below, 1 = CfgHeavy, 2 / 3 = force CfgFdiv2 / CfgFdiv3.
template <typename T, int CLS>
static cudaError_t go_heavy(const EwArgs& a, cudaStream_t s) {
    int cfg = a.k.heavy_cfg;
    if (cfg == 0) {
        const bool masked = a.lmask || a.rmask;
        if (CLS == CLS_POW && sizeof(T) > 1) cfg = 2;
        else if (CLS == CLS_REM) cfg = sizeof(T) == 8 ? 2 : 1;                       // float remainder
        else if (sizeof(T) == 8) cfg = 2;
        else if (sizeof(T) == 4) cfg = (masked && Traits<T>::is_signed) ? 3 : 2;
        else if (CLS == CLS_POW && sizeof(T) == 1) cfg = 3;   // table in shared memory, built once per block: resident grid
        else cfg = 2;   // 8/16-bit columns: the packed division path (r02d: 128 thr x 2 x 256-bit wins every shape)
    }
    if (cfg == 2) return go_align<T, T, T, CLS, CfgFdiv2>(a, s);
    if (cfg == 3) return go_align<T, T, T, CLS, CfgFdiv3>(a, s);
    if constexpr (CLS == CLS_POW && std::is_same<T, double>::value) {
        if (cfg == 4) return go_align<T, T, T, CLS, CfgPow4>(a, s);
        if (cfg == 5) return go_align<T, T, T, CLS, CfgPow5>(a, s);
    }
    return go_align<T, T, T, CLS>(a, s);
}

template <typename T>
static cudaError_t go_t(const EwArgs& a, cudaStream_t s) {
    if constexpr (!Traits<T>::is_float) {
        if (a.sdiv) {
            if constexpr (sizeof(T) == 8) {
                // 0 = masked -> CfgSdiv64, dense -> CfgCheap; 1 / 2 force one of them (tools/sdiv64_exp.py)
                const bool masked = a.lmask || a.rmask;
                if (a.k.sdiv64_cfg == 2 || (a.k.sdiv64_cfg == 0 && masked)) return go_align<T, T, T, CLS_SDIV, CfgSdiv64>(a, s);
            }
            return go_align<T, T, T, CLS_SDIV>(a, s);
        }
    }
    switch (op_class(Traits<T>::is_float, a.op)) {
        case CLS_CHEAP:
            if constexpr (sizeof(T) == 1) {
                // Masked add / sub / mul on 1-byte columns.  At CfgCheap's 4 x 256-bit per thread the kernel holds 128
                // registers (16 warps per SM) and each warp alternates between a long load phase and 150 instructions of
                // lane work: ncu (profiles/r02k_u8_masked_add_ncu.md) shows 20 % issue utilisation, 23 % warps active,
                // DRAM 63 %.  Two 256-bit loads per thread at <= 85 registers keep 24 warps resident and lift array (+)
                // array with two masks from 5.2 to 6.9-7.0 TB/s; array (+) scalar is a little faster as it was.
                // ew_cheap8_cfg: 0 = this choice, 1 = CfgCheap, 2 / 3 = CfgFdiv2 / CfgFdiv3.
                int cfg = a.k.cheap8_cfg;
                if (cfg == 0) cfg = (a.lhs && a.rhs) ? 2 : 1;
                if (cfg == 2 && (a.lmask || a.rmask)) return go_align<T, T, T, CLS_CHEAP, CfgFdiv2>(a, s);
                if (cfg == 3 && (a.lmask || a.rmask)) return go_align<T, T, T, CLS_CHEAP, CfgFdiv3>(a, s);
            }
            return go_align<T, T, T, CLS_CHEAP>(a, s);
        case CLS_DIV:
            if constexpr (Traits<T>::is_float) {
                // ew_fdiv_cfg: 0 = the table above, 1 = CfgHeavy, 2 / 3 = force CfgFdiv2 / CfgFdiv3
                int cfg = a.k.fdiv_cfg;
                if (cfg == 0) {
                    const bool masked = a.lmask || a.rmask, scalar = !a.lhs || !a.rhs;
                    cfg = scalar ? (masked ? 3 : 2) : (masked ? (sizeof(T) == 8 ? 2 : 3) : 2);
                }
                if (cfg == 2) return go_align<T, T, T, CLS_DIV, CfgFdiv2>(a, s);
                if (cfg == 3) return go_align<T, T, T, CLS_DIV, CfgFdiv3>(a, s);
            }
            if constexpr (Traits<T>::is_float) return go_align<T, T, T, CLS_DIV>(a, s);
            else return go_heavy<T, CLS_DIV>(a, s);
        case CLS_POW: return go_heavy<T, CLS_POW>(a, s);
        case CLS_REM:
            if constexpr (Traits<T>::is_float) return go_heavy<T, CLS_REM>(a, s);
            break;
    }
    return cudaErrorInvalidValue;
}

#define MNR_EW_ENTRY(NAME, T)                                                                                         \
    cudaError_t NAME(const EwArgs& a, cudaStream_t s) { return go_t<T>(a, s); }                                       \
    cudaError_t NAME##_batch(int op, int tier, bool masked, bool sdiv, const EwDev* segs, uint32_t nseg,              \
                             uint64_t max_n, int grid_cap, cudaStream_t s) {                                          \
        return go_batch_t<T>(op, tier, masked, sdiv, segs, nseg, max_n, grid_cap, s);                                 \
    }

#if MNR_EW_DTYPE == 6
MNR_EW_ENTRY(launch_ew_i8, int8_t)
#elif MNR_EW_DTYPE == 7
MNR_EW_ENTRY(launch_ew_u8, uint8_t)
#elif MNR_EW_DTYPE == 8
MNR_EW_ENTRY(launch_ew_i16, int16_t)
#elif MNR_EW_DTYPE == 9
MNR_EW_ENTRY(launch_ew_u16, uint16_t)
#elif MNR_EW_DTYPE == 0
MNR_EW_ENTRY(launch_ew_i32, int32_t)
#elif MNR_EW_DTYPE == 1
MNR_EW_ENTRY(launch_ew_u32, uint32_t)
#elif MNR_EW_DTYPE == 2
MNR_EW_ENTRY(launch_ew_i64, int64_t)
#elif MNR_EW_DTYPE == 3
MNR_EW_ENTRY(launch_ew_u64, uint64_t)
#elif MNR_EW_DTYPE == 4
MNR_EW_ENTRY(launch_ew_f32, float)
#elif MNR_EW_DTYPE == 5
MNR_EW_ENTRY(launch_ew_f64, double)
#elif MNR_EW_DTYPE == 100

#define MNR_EW_DECL(NAME)                                 \
    cudaError_t NAME(const EwArgs&, cudaStream_t);         \
    cudaError_t NAME##_batch(int, int, bool, bool, const EwDev*, uint32_t, uint64_t, int, cudaStream_t);
MNR_EW_DECL(launch_ew_i8) MNR_EW_DECL(launch_ew_u8) MNR_EW_DECL(launch_ew_i16) MNR_EW_DECL(launch_ew_u16)
MNR_EW_DECL(launch_ew_i32) MNR_EW_DECL(launch_ew_u32) MNR_EW_DECL(launch_ew_i64) MNR_EW_DECL(launch_ew_u64)
MNR_EW_DECL(launch_ew_f32) MNR_EW_DECL(launch_ew_f64)

// 2: every pointer allows the op class's wide vector; 1: 128-bit; 0: element loads (not batched).
int ew_batch_tier(mnr_dtype dt, int op, bool sdiv, const void* lhs, const void* rhs, const void* out, int max_tier) {
    const bool is_float = dt == MNR_F32 || dt == MNR_F64;
    const unsigned wide = ((sdiv || op_class(is_float, op) == CLS_CHEAP) && max_tier >= 2) ? 32u : 16u;
    auto ok = [](const void* p, unsigned al) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & (al - 1)) == 0; };
    if (wide > 16 && ok(lhs, wide) && ok(rhs, wide) && ok(out, wide)) return 2;
    if (ok(lhs, 16) && ok(rhs, 16) && ok(out, 16)) return wide == 16 ? 2 : 1;
    return 0;
}

cudaError_t launch_ew_batch(mnr_dtype dt, int op, int tier, bool masked, bool sdiv, const EwDev* segs, uint32_t nseg,
                            uint64_t max_n, int grid_cap, cudaStream_t s) {
    switch (dt) {
        case MNR_I8: return launch_ew_i8_batch(op, tier, masked, sdiv, segs, nseg, max_n, grid_cap, s);
        case MNR_U8: return launch_ew_u8_batch(op, tier, masked, sdiv, segs, nseg, max_n, grid_cap, s);
        case MNR_I16: return launch_ew_i16_batch(op, tier, masked, sdiv, segs, nseg, max_n, grid_cap, s);
        case MNR_U16: return launch_ew_u16_batch(op, tier, masked, sdiv, segs, nseg, max_n, grid_cap, s);
        case MNR_I32: return launch_ew_i32_batch(op, tier, masked, sdiv, segs, nseg, max_n, grid_cap, s);
        case MNR_U32: return launch_ew_u32_batch(op, tier, masked, sdiv, segs, nseg, max_n, grid_cap, s);
        case MNR_I64: return launch_ew_i64_batch(op, tier, masked, sdiv, segs, nseg, max_n, grid_cap, s);
        case MNR_U64: return launch_ew_u64_batch(op, tier, masked, sdiv, segs, nseg, max_n, grid_cap, s);
        case MNR_F32: return launch_ew_f32_batch(op, tier, masked, sdiv, segs, nseg, max_n, grid_cap, s);
        case MNR_F64: return launch_ew_f64_batch(op, tier, masked, sdiv, segs, nseg, max_n, grid_cap, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_ew_binary(const EwArgs& a, cudaStream_t s) {
    switch (a.dtype) {
        case MNR_I8: return launch_ew_i8(a, s);
        case MNR_U8: return launch_ew_u8(a, s);
        case MNR_I16: return launch_ew_i16(a, s);
        case MNR_U16: return launch_ew_u16(a, s);
        case MNR_I32: return launch_ew_i32(a, s);
        case MNR_U32: return launch_ew_u32(a, s);
        case MNR_I64: return launch_ew_i64(a, s);
        case MNR_U64: return launch_ew_u64(a, s);
        case MNR_F32: return launch_ew_f32(a, s);
        case MNR_F64: return launch_ew_f64(a, s);
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

constexpr int kFBlock = 128, kFU = 2;

template <typename T, typename VecT>
static cudaError_t fma_v(const T* pa, const T* pb, const T* pc, const uint8_t* mask, T* po, uint8_t* out_mask, uint64_t n,
                         cudaStream_t s) {
    constexpr int VEC = sizeof(VecT) / sizeof(T);
    const unsigned grid = ew_grid<kFBlock, kFU, 1, false>(n, VEC);
    if (mask) ew_fma_kernel<T, VecT, true, kFBlock, kFU><<<grid, kFBlock, 0, s>>>(pa, pb, pc, mask, po, out_mask, n);
    else ew_fma_kernel<T, VecT, false, kFBlock, kFU><<<grid, kFBlock, 0, s>>>(pa, pb, pc, mask, po, out_mask, n);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t fma_t(const void* a, const void* b, const void* c, const uint8_t* mask, void* out, uint8_t* out_mask,
                         uint64_t n, cudaStream_t s) {
    auto al = [&](unsigned m) {
        return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                 reinterpret_cast<uintptr_t>(out)) & m) == 0;
    };
    const T *pa = static_cast<const T*>(a), *pb = static_cast<const T*>(b), *pc = static_cast<const T*>(c);
    T* po = static_cast<T*>(out);
    if (al(31u)) return fma_v<T, V32>(pa, pb, pc, mask, po, out_mask, n, s);
    if (al(15u)) return fma_v<T, V16>(pa, pb, pc, mask, po, out_mask, n, s);
    return fma_v<T, T>(pa, pb, pc, mask, po, out_mask, n, s);
}

cudaError_t launch_ew_fma(mnr_dtype dt, const void* a, const void* b, const void* c, const uint8_t* mask, void* out,
                          uint8_t* out_mask, uint64_t n, cudaStream_t s) {
    if (dt == MNR_F32) return fma_t<float>(a, b, c, mask, out, out_mask, n, s);
    if (dt == MNR_F64) return fma_t<double>(a, b, c, mask, out, out_mask, n, s);
    return cudaErrorInvalidValue;
}
#else
#error "compile elementwise.cu with -DMNR_EW_DTYPE=<0..9 | 100>"
#endif

}  // namespace mnr
