// Null-aware column aggregates (sum / min / max / count) in one HBM pass.
//
// Replaces the sums the reference defines in its benches (benches/benchmark_parallel_simd.rs:44-97,
// benches/hotloop_benchmark_simd.rs:56-174) and adds the null-skipping + count/min/max the north-star asks
// for (definition: DESIGN.md "A.6"; the reference keeps those kernels in the downstream simd-kernels crate).
//
// Shape of the computation (this is also the documented float summation order):
//   * the column is cut into 16-byte (or 32-byte) vectors; warp w of the grid owns "warp tiles" of
//     32*U consecutive vectors, tiles w, w+W, w+2W, ... (W = warps in the grid);
//   * a lane adds the vectors it loads in index order into VEC per-slot accumulators (slot k = element k
//     of the vector — the analogue of the reference's SIMD lanes), then adds the slots in slot order;
//   * lanes combine by xor-butterfly (offsets 16,8,4,2,1), warps combine in warp order inside the block,
//     every block stores one partial, and the last block to finish (atomic ticket) folds the partials:
//     thread t adds partials t, t+BLOCK, ... in index order, then the same block reduction.
//   Grid and block sizes are compile-time constants or functions of len only, so a float sum is
//   bit-reproducible run to run and device to device.
// Validity costs 1 bit per row: each lane reads the byte that holds its vector's bits (no alignment or
// padding requirement on the mask), zeroes invalid rows with a select — never a multiply, stored values
// at null slots may be NaN/Inf — and popcounts its bits for `count`.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace mnr {

struct alignas(16) AggRaw {   // device image of mnr_agg
    uint64_t sum, mn, mx, count;
};

template <typename T> struct MinMax {
    // Identity of min / max when nothing qualifies (DESIGN.md "A.6").
    __device__ static T min_identity();
    __device__ static T max_identity();
};
#define MNR_MM_INT(T, MX, MN)                                        \
    template <> struct MinMax<T> {                                   \
        __device__ static T min_identity() { return MX; }            \
        __device__ static T max_identity() { return MN; }            \
    };
MNR_MM_INT(int8_t, INT8_MAX, INT8_MIN) MNR_MM_INT(uint8_t, UINT8_MAX, 0)
MNR_MM_INT(int16_t, INT16_MAX, INT16_MIN) MNR_MM_INT(uint16_t, UINT16_MAX, 0)
MNR_MM_INT(int32_t, INT32_MAX, INT32_MIN) MNR_MM_INT(uint32_t, UINT32_MAX, 0)
MNR_MM_INT(int64_t, INT64_MAX, INT64_MIN) MNR_MM_INT(uint64_t, UINT64_MAX, 0)
#undef MNR_MM_INT
template <> struct MinMax<float> {
    __device__ static float min_identity() { return __int_as_float(0x7fc00000); }
    __device__ static float max_identity() { return __int_as_float(0x7fc00000); }
};
template <> struct MinMax<double> {
    __device__ static double min_identity() { return __longlong_as_double(0x7ff8000000000000ll); }
    __device__ static double max_identity() { return __longlong_as_double(0x7ff8000000000000ll); }
};

// min/max combine.  Integers: plain.  Floats: NaN never wins (NaN doubles as "empty") and -0.0 < +0.0, so the result
// does not depend on the order in which equal zeros are met.  That is exactly the hardware's IEEE-754 minNum/maxNum
// (FMNMX / DMNMX: the non-NaN operand wins, -0.0 orders below +0.0), one instruction instead of four compares.
template <typename T> __device__ __forceinline__ T comb_min(T a, T b) {
    if constexpr (std::is_same<T, float>::value) return fminf(a, b);
    else if constexpr (std::is_same<T, double>::value) return fmin(a, b);
    else return b < a ? b : a;
}
template <typename T> __device__ __forceinline__ T comb_max(T a, T b) {
    if constexpr (std::is_same<T, float>::value) return fmaxf(a, b);
    else if constexpr (std::is_same<T, double>::value) return fmax(a, b);
    else return b > a ? b : a;
}

template <typename A> __device__ __forceinline__ uint64_t acc_bits(A v) {
    if constexpr (sizeof(A) == 8 && !std::is_floating_point<A>::value) return (uint64_t)v;
    else return (uint64_t)__double_as_longlong((double)v);
}
template <typename A> __device__ __forceinline__ A acc_from_bits(uint64_t b) {
    if constexpr (std::is_floating_point<A>::value) return (A)__longlong_as_double((long long)b);
    else return (A)b;
}
// min/max travel widened to the accumulator type (i64 / u64 / f64), as mnr_agg stores them.
template <typename T> __device__ __forceinline__ uint64_t mm_bits(T v) {
    using A = typename Traits<T>::Acc;
    return acc_bits<A>((A)v);
}
template <typename T> __device__ __forceinline__ T mm_from_bits(uint64_t b) {
    using A = typename Traits<T>::Acc;
    return (T)acc_from_bits<A>(b);
}

template <typename T, bool MINMAX> struct Partial {
    using A = typename Traits<T>::Acc;
    A sum;
    T mn, mx;
    uint64_t cnt;
    __device__ __forceinline__ void init() {
        sum = (A)0; cnt = 0;
        mn = MinMax<T>::min_identity(); mx = MinMax<T>::max_identity();
    }
    __device__ __forceinline__ void merge(const Partial& o) {
        if constexpr (Traits<T>::is_float) sum = sum + o.sum;
        else sum = (A)((uint64_t)sum + (uint64_t)o.sum);
        cnt += o.cnt;
        if constexpr (MINMAX) { mn = comb_min(mn, o.mn); mx = comb_max(mx, o.mx); }
    }
    __device__ __forceinline__ Partial shfl_xor(int off) const {
        Partial r;
        r.sum = acc_from_bits<A>(__shfl_xor_sync(0xffffffffu, (unsigned long long)acc_bits<A>(sum), off));
        r.cnt = __shfl_xor_sync(0xffffffffu, (unsigned long long)cnt, off);
        if constexpr (MINMAX) {
            r.mn = mm_from_bits<T>(__shfl_xor_sync(0xffffffffu, (unsigned long long)mm_bits<T>(mn), off));
            r.mx = mm_from_bits<T>(__shfl_xor_sync(0xffffffffu, (unsigned long long)mm_bits<T>(mx), off));
        } else { r.mn = mn; r.mx = mx; }
        return r;
    }
    __device__ __forceinline__ AggRaw raw() const { return AggRaw{acc_bits<A>(sum), mm_bits<T>(mn), mm_bits<T>(mx), cnt}; }
    __device__ __forceinline__ void from_raw(const AggRaw& r) {
        sum = acc_from_bits<A>(r.sum); cnt = r.count; mn = mm_from_bits<T>(r.mn); mx = mm_from_bits<T>(r.mx);
    }
};

// Block-wide combine; result valid in thread 0.  Deterministic: butterfly in lanes, warp order in smem.
template <typename P, int BLOCK> __device__ __forceinline__ P block_combine(P p, AggRaw* smem) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) p.merge(p.shfl_xor(off));
    constexpr int NW = BLOCK / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) smem[warp] = p.raw();
    __syncthreads();
    if (warp == 0) {
        P q; q.init();
        if (lane == 0) {
#pragma unroll 1
            for (int w = 0; w < NW; ++w) { P t; t.from_raw(smem[w]); if (w == 0) q = t; else q.merge(t); }
        }
        p = q;
    }
    __syncthreads();
    return p;
}

template <typename T, typename VecT, bool MASKED, bool MINMAX>
__device__ __forceinline__ void accum_vec(const VecT& v, uint32_t bits, typename Traits<T>::Acc* slot,
                                          Partial<T, MINMAX>& p) {
    constexpr int VEC = sizeof(VecT) / sizeof(T);
    using A = typename Traits<T>::Acc;
    union { VecT v; T e[VEC]; } u;
    u.v = v;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const bool ok = !MASKED || ((bits >> k) & 1u);
        const T x = u.e[k];
        if constexpr (Traits<T>::is_float) slot[k] = slot[k] + (ok ? (A)x : (A)0);
        else slot[k] = (A)((uint64_t)slot[k] + (ok ? (uint64_t)(A)x : 0ull));
        if constexpr (MINMAX) {
            if constexpr (Traits<T>::is_float) {   // invalid row -> NaN, which minNum/maxNum skip: no branch, no predicate
                const T xm = ok ? x : MinMax<T>::min_identity();
                p.mn = comb_min(p.mn, xm); p.mx = comb_max(p.mx, xm);
            } else {
                p.mn = comb_min(p.mn, ok ? x : MinMax<T>::min_identity());
                p.mx = comb_max(p.mx, ok ? x : MinMax<T>::max_identity());
            }
        }
    }
    if constexpr (MASKED) p.cnt += (uint64_t)__popc(bits);
}

// ---- 8- and 16-bit integers: packed (SIMD-in-register) accumulation -------------------------------------------
// A 1-byte column has to be consumed at ~7e12 rows/s to stay on the HBM roofline; the per-element path above costs
// ~10 integer instructions per row (bit extract, select, 64-bit add, two selects + two compares), several times what
// the SMs can issue at that rate.  Integer sums are order-free, so narrow columns are processed 32 bits at a time:
//   * the word's 4 (or 2) validity bits are expanded to a byte (halfword) mask with one multiply-and-mask;
//   * sum: IDP.4A / IDP.2A dot product of the masked word with 0x01..01 into a 32-bit accumulator, folded into
//     the 64-bit sum once per vector (a vector adds at most 32 * 255 or 16 * 65535, far from 2^31);
//   * min / max: invalid lanes are replaced by the identity with one LOP3; 16-bit columns fold with VIMNMX.{S,U}16x2.
//     8-bit columns are compared as the HIGH byte of a 16-bit lane — whatever sits in the low byte can only break ties
//     between equal high bytes — so bytes 1 and 3 are in place in the word as it is and bytes 0 and 2 after `<< 8`:
//     VIMNMX3.{S,U}16x2(acc, w, w << 8) takes in four rows with one shift, no widening PRMTs (four per word before:
//     r02z had the masked 8-bit sum+min+max at 0.80 of the copy peak, ALU-pipe bound).
// About 2.7 instructions per row for an 8-bit column with min/max, 1.3 without.  Same results as the per-element
// path: wrapping 64-bit sums and integer min/max do not depend on the order of combination.
template <typename T> struct NarrowState {
    static constexpr bool kSigned = Traits<T>::is_signed;
    static constexpr int EPW = 4 / sizeof(T);   // elements per 32-bit word
    uint64_t sum;
    uint32_t mn2, mx2;                          // running min / max, two 16-bit lanes
    __device__ __forceinline__ void init() {
        sum = 0;
        // 8-bit columns keep the running value in the high byte of each lane (see above)
        const uint32_t idmin = sizeof(T) == 1 ? (uint32_t)(uint8_t)MinMax<T>::min_identity() << 8 : (uint16_t)(int16_t)MinMax<T>::min_identity();
        const uint32_t idmax = sizeof(T) == 1 ? (uint32_t)(uint8_t)MinMax<T>::max_identity() << 8 : (uint16_t)(int16_t)MinMax<T>::max_identity();
        mn2 = idmin | (idmin << 16);
        mx2 = idmax | (idmax << 16);
    }
    static __device__ __forceinline__ uint32_t min2(uint32_t a, uint32_t b) { return kSigned ? __vmins2(a, b) : __vminu2(a, b); }
    static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) { return kSigned ? __vmaxs2(a, b) : __vmaxu2(a, b); }
    static __device__ __forceinline__ uint32_t min3(uint32_t a, uint32_t b, uint32_t c) {
        return kSigned ? __vimin3_s16x2(a, b, c) : __vimin3_u16x2(a, b, c);
    }
    static __device__ __forceinline__ uint32_t max3(uint32_t a, uint32_t b, uint32_t c) {
        return kSigned ? __vimax3_s16x2(a, b, c) : __vimax3_u16x2(a, b, c);
    }

    // One 32-bit word = EPW elements; `bits` = their validity in the low EPW bits (ignored when !MASKED).
    template <bool MASKED, bool MINMAX> __device__ __forceinline__ void add_word(uint32_t w, uint32_t bits, uint32_t& acc32) {
        if constexpr (sizeof(T) == 1 && MASKED) {
            // `bits` is a clean nibble here (accum_vec_packed extracts it with one PRMT).  The sum needs no select at all:
            // the dot product against the 0/1 byte pattern skips invalid lanes; the 0xFF mask is only built for min/max.
            const uint32_t s01 = (bits * 0x00204081u) & 0x01010101u;
            if constexpr (kSigned) acc32 = (uint32_t)__dp4a((int)w, (int)s01, (int)acc32);
            else acc32 = __dp4a(w, s01, acc32);
            if constexpr (MINMAX) {
                const uint32_t m = s01 * 0xFFu;
                const uint32_t wmn = kSigned ? ((w & m) | (0x7f7f7f7fu & ~m)) : (w | ~m);
                const uint32_t wmx = kSigned ? ((w & m) | (0x80808080u & ~m)) : (w & m);
                mn2 = min3(mn2, wmn, wmn << 8);
                mx2 = max3(mx2, wmx, wmx << 8);
            }
            return;
        }
        uint32_t m = 0xFFFFFFFFu;
        if constexpr (MASKED) m = expand_valid_word<sizeof(T)>(bits);
        const uint32_t wz = w & m;   // invalid lanes -> 0
        if constexpr (sizeof(T) == 1) {
            if constexpr (kSigned) acc32 = (uint32_t)__dp4a((int)wz, 0x01010101, (int)acc32);
            else acc32 = __dp4a(wz, 0x01010101u, acc32);
        } else {
            if constexpr (kSigned) acc32 = (uint32_t)__dp2a_lo((int)wz, 0x00000101, (int)acc32);
            else acc32 = __dp2a_lo(wz, 0x00000101u, acc32);
        }
        if constexpr (MINMAX) {
            constexpr uint32_t IDMIN = sizeof(T) == 1 ? (kSigned ? 0x7f7f7f7fu : 0xffffffffu) : (kSigned ? 0x7fff7fffu : 0xffffffffu);
            constexpr uint32_t IDMAX = sizeof(T) == 1 ? (kSigned ? 0x80808080u : 0u) : (kSigned ? 0x80008000u : 0u);
            const uint32_t wmn = MASKED ? (wz | (IDMIN & ~m)) : w;
            const uint32_t wmx = MASKED ? (wz | (IDMAX & ~m)) : w;
            if constexpr (sizeof(T) == 1) {
                mn2 = min3(mn2, wmn, wmn << 8);
                mx2 = max3(mx2, wmx, wmx << 8);
            } else {
                mn2 = min2(mn2, wmn);
                mx2 = max2(mx2, wmx);
            }
        }
    }
    // Fold a 32-bit per-vector accumulator into the wrapping 64-bit sum.
    __device__ __forceinline__ void fold(uint32_t acc32) {
        if constexpr (kSigned) sum += (uint64_t)(int64_t)(int32_t)acc32;
        else sum += (uint64_t)acc32;
    }
    __device__ __forceinline__ T min_value() const {
        constexpr int SH = sizeof(T) == 1 ? 8 : 0;
        if constexpr (kSigned) { const int16_t a = (int16_t)(mn2 & 0xffffu), b = (int16_t)(mn2 >> 16); return (T)((a < b ? a : b) >> SH); }
        else { const uint16_t a = (uint16_t)(mn2 & 0xffffu), b = (uint16_t)(mn2 >> 16); return (T)((a < b ? a : b) >> SH); }
    }
    __device__ __forceinline__ T max_value() const {
        constexpr int SH = sizeof(T) == 1 ? 8 : 0;
        if constexpr (kSigned) { const int16_t a = (int16_t)(mx2 & 0xffffu), b = (int16_t)(mx2 >> 16); return (T)((a > b ? a : b) >> SH); }
        else { const uint16_t a = (uint16_t)(mx2 & 0xffffu), b = (uint16_t)(mx2 >> 16); return (T)((a > b ? a : b) >> SH); }
    }
};

template <typename T, typename VecT> struct UsesPacked {
    static constexpr bool value = !Traits<T>::is_float && sizeof(T) <= 2 && sizeof(VecT) >= 16;
};

template <typename T, typename VecT, bool MASKED, bool MINMAX>
__device__ __forceinline__ void accum_vec_packed(const VecT& v, uint32_t bits, NarrowState<T>& ns, uint64_t& cnt) {
    constexpr int NW = sizeof(VecT) / 4;
    constexpr int EPW = NarrowState<T>::EPW;
    union { VecT v; uint32_t w[NW]; } u;
    u.v = v;
    uint32_t acc32 = 0;
    if constexpr (sizeof(T) == 1 && MASKED) {
        // nibble j of `bits` = byte j/2 of the even- or odd-nibble plane: two ops per vector + one PRMT per word
        // (a shift + mask per word would be two ALU ops; this path is ALU-bound)
        const uint32_t even = bits & 0x0F0F0F0Fu, odd = (bits >> 4) & 0x0F0F0F0Fu;
#pragma unroll
        for (int j = 0; j < NW; ++j)
            ns.template add_word<MASKED, MINMAX>(u.w[j], __byte_perm((j & 1) ? odd : even, 0u, 0x4440u | (uint32_t)(j >> 1)), acc32);
    } else {
#pragma unroll
        for (int j = 0; j < NW; ++j) ns.template add_word<MASKED, MINMAX>(u.w[j], MASKED ? (bits >> (j * EPW)) : 0u, acc32);
    }
    ns.fold(acc32);
    if constexpr (MASKED) cnt += (uint64_t)__popc(bits);
}

// One launch = whole column -> one mnr_agg (two-level finish inside the launch via an atomic ticket).
__device__ __forceinline__ AggRaw load_partial(const AggRaw* p) {   // L2-coherent read of another block's partial
    AggRaw r;
    asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(r.sum), "=l"(r.mn) : "l"(p));
    asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(r.mx), "=l"(r.count) : "l"(&p->mx));
    return r;
}

// ---- fused cross-GPU finish over NVLink peer memory ------------------------------------------------------------
// Every rank owns a mailbox in its own HBM, mapped into every peer (CUDA IPC between processes, plain peer access inside
// one process).  Slot (parity p, source rank s) lives at ((p * kMaxPeers) + s) * slot_bytes: bytes 0..7 a flag holding the
// epoch, bytes 32.. up to `max_aggs` AggRaw images (one per column of the exchange).  The finishing block of rank r's
// reduction stores its aggregates + flag into slot (epoch & 1, r) of EVERY rank's mailbox (P2P stores through NVSwitch),
// waits until the `world` slots of its own mailbox carry this epoch, and folds them in rank order — the same fold on every
// rank, so all ranks hold identical bits.  One kernel: the reduction, the all-gather of partials and the combine; no
// NCCL call, no host round trip.  Parity double-buffering is race-free: a rank can only complete epoch e+1 after every
// peer has finished epoch e (it needs their e+1 flags, which each peer posts only after its own epoch-e kernel is done).
constexpr int kMaxPeers = 16;
constexpr int kXchgHeader = 32;            // flag + padding in front of the aggregates of a slot
struct XchgDev {
    int world;                 // 0 = no exchange (plain single-GPU finish)
    int rank;
    unsigned long long epoch;  // 1, 2, 3, ... identical sequence on every rank
    unsigned int* err;         // device word, set to 1 if a peer's flag never arrives (bounded spin)
    unsigned long long* done;  // device word: the last epoch whose exchange has finished on this rank (orders overlapped launches)
    unsigned int slot_bytes;   // kXchgHeader + 32 * max_aggs, rounded up to 64
    char* mailbox[kMaxPeers];  // mailbox[r]: rank r's mailbox as mapped in this process
};
// A result whose exchange timed out carries this count (no column has 2^64 - 1 rows), so an asynchronous caller that
// never looks at mnr_xchg_status still cannot mistake it for an aggregate.
constexpr unsigned long long kAggPoison = ~0ull;

__device__ __forceinline__ void st_sys_v2(void* p, uint64_t a, uint64_t b) {
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void st_release_sys(void* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const void* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ AggRaw ld_sys_agg(const void* p) {
    AggRaw r;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];" : "=l"(r.sum), "=l"(r.mn) : "l"(p) : "memory");
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];" : "=l"(r.mx), "=l"(r.count) : "l"(static_cast<const char*>(p) + 16) : "memory");
    return r;
}
__device__ __forceinline__ void st_sys_agg(void* p, const AggRaw& a) {
    st_sys_v2(p, a.sum, a.mn);
    st_sys_v2(static_cast<char*>(p) + 16, a.mx, a.count);
}

// Programmatic dependent launch (griddepcontrol): reductions are launched with the programmatic-stream-serialization
// attribute, so the NEXT kernel on the stream may be scheduled as soon as every block of this one has started, and its
// blocks take over SM slots as ours exit — the launch ramp, the ticket finish and the cross-GPU flag wait of reduction k
// overlap the streaming phase of reduction k+1.  `pdl_wait` blocks until the previous kernel on the stream has completed
// and its writes are visible; both instructions are no-ops for a kernel launched without the attribute.
// In the overlapped ("late") form of the exchange reduction NO block takes that wait: streaming blocks never depend on the
// previous launch, partials / tickets alternate between two buffers by epoch parity, and the launches are ordered through
// the exchange's `done` word instead — a block checks `done >= epoch - 2` before it writes its partial (the buffer's previous
// user has folded; practically never spins) and the finishing block alone waits for `done >= epoch - 1` before it posts to
// the mailboxes and writes the result, then publishes `done = epoch`.  So a rank may run up to two reductions ahead of the
// slowest peer, and the per-step max-over-ranks jitter of a lock-step collective is absorbed instead of added.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Post `n` aggregates (`mine`: shared or global memory written before the call by this block) to every rank, raise the
// epoch flag, and wait until every rank's flag for this epoch sits in the own mailbox.  Called by all threads of the
// finishing block.  Returns false (block-uniform) on a timeout; afterwards xchg_slot(x, src, g) is rank src's aggregate g.
template <int BLOCK>
__device__ __forceinline__ bool xchg_exchange(const XchgDev& x, const AggRaw* mine, const unsigned n) {
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    const size_t par = (size_t)(x.epoch & 1ull) * kMaxPeers;
    __syncthreads();
    for (unsigned t = threadIdx.x; t < (unsigned)x.world * n; t += BLOCK) {
        const unsigned peer = t / n, g = t - peer * n;
        st_sys_agg(x.mailbox[peer] + (par + (size_t)x.rank) * x.slot_bytes + kXchgHeader + (size_t)g * 32, mine[g]);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < (unsigned)x.world) {
        st_release_sys(x.mailbox[threadIdx.x] + (par + (size_t)x.rank) * x.slot_bytes, x.epoch);
        const char* src = x.mailbox[x.rank] + (par + threadIdx.x) * x.slot_bytes;
        const long long t0 = clock64();
        while (ld_acquire_sys(src) != x.epoch) {
            if (clock64() - t0 > (20ll << 30)) { timed_out = 1; atomicExch(x.err, 1u); break; }   // ~10 s: a peer never launched
        }
    }
    __syncthreads();
    return timed_out == 0;
}
__device__ __forceinline__ AggRaw xchg_slot(const XchgDev& x, unsigned src, unsigned g) {
    const size_t par = (size_t)(x.epoch & 1ull) * kMaxPeers;
    return ld_sys_agg(x.mailbox[x.rank] + (par + src) * x.slot_bytes + kXchgHeader + (size_t)g * 32);
}

// Run-time typed combine of two aggregate images (the batched fold handles columns of different dtypes in one block).
// kind: 0 signed, 1 unsigned, 2 float.  Same arithmetic as Partial<T>::merge / mnr_agg_combine: integer sums wrap, float
// sums add in double, float min/max are IEEE minNum/maxNum (NaN = "empty" never wins, -0.0 < +0.0).
__device__ __forceinline__ void agg_merge_rt(int kind, AggRaw& a, const AggRaw& b) {
    a.count += b.count;
    if (kind == 2) {
        a.sum = (uint64_t)__double_as_longlong(__longlong_as_double((long long)a.sum) + __longlong_as_double((long long)b.sum));
        a.mn = (uint64_t)__double_as_longlong(fmin(__longlong_as_double((long long)a.mn), __longlong_as_double((long long)b.mn)));
        a.mx = (uint64_t)__double_as_longlong(fmax(__longlong_as_double((long long)a.mx), __longlong_as_double((long long)b.mx)));
    } else if (kind == 1) {
        a.sum += b.sum;
        a.mn = b.mn < a.mn ? b.mn : a.mn;
        a.mx = b.mx > a.mx ? b.mx : a.mx;
    } else {
        a.sum += b.sum;
        a.mn = (int64_t)b.mn < (int64_t)a.mn ? b.mn : a.mn;
        a.mx = (int64_t)b.mx > (int64_t)a.mx ? b.mx : a.mx;
    }
}

// Body shared by the single-column kernel and the batched (one launch, many chunks) kernel.  `bid` / `nblk` are the
// block's index and the number of blocks working on THIS column; nblk is a function of (len, dtype) only, so a column
// reduced inside a batch gives the same bits as the same column reduced alone.
// Returns true in the block that produced the column's aggregate (the finishing block), false in all others.
// `late_wait`: the programmatic-dependency wait is taken after the streaming phase (the caller guarantees the column is
// not being written by the previous kernel on the stream) instead of by the kernel's first instruction.
template <typename T, typename VecT, bool MASKED, bool MINMAX, int BLOCK, int U>
__device__ __forceinline__ bool reduce_stats_body(const T* __restrict__ data, const uint8_t* __restrict__ mask, uint64_t n,
                                                  AggRaw* __restrict__ partials, unsigned int* __restrict__ ticket,
                                                  AggRaw* __restrict__ out, AggRaw* __restrict__ out_host,
                                                  const unsigned int bid, const unsigned int nblk, const XchgDev& x,
                                                  const bool late_wait, const uint32_t host_seq = 0) {
    constexpr int VEC = sizeof(VecT) / sizeof(T);
    using A = typename Traits<T>::Acc;
    using P = Partial<T, MINMAX>;
    __shared__ AggRaw smem[BLOCK / 32];
    __shared__ bool is_last;

    constexpr bool PACKED = UsesPacked<T, VecT>::value;   // 8/16-bit integers: SIMD-in-register accumulation
    constexpr int NSLOT = PACKED ? 1 : VEC;
    P p; p.init();
    A slot[NSLOT];
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) slot[k] = (A)0;
    NarrowState<typename std::conditional<PACKED, T, int8_t>::type> ns;
    if constexpr (PACKED) ns.init();

    const VecT* __restrict__ vp = reinterpret_cast<const VecT*>(data);
    const uint64_t nvec = n / VEC;
    constexpr uint64_t WTILE = 32ull * U;                       // vectors per warp tile
    const uint64_t ntiles = nvec / WTILE;
    const uint64_t warps = (uint64_t)nblk * (BLOCK / 32);
    const uint64_t gwarp = (uint64_t)bid * (BLOCK / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;

    for (uint64_t t = gwarp; t < ntiles; t += warps) {
        const uint64_t v0 = t * WTILE + lane;
        VecT v[U];
        uint32_t bits[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldg_stream(vp + v0 + 32ull * u);
        if constexpr (MASKED) {
#pragma unroll
            for (int u = 0; u < U; ++u) bits[u] = load_valid_bits<VEC>(mask, (v0 + 32ull * u) * VEC);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if constexpr (PACKED) accum_vec_packed<T, VecT, MASKED, MINMAX>(v[u], MASKED ? bits[u] : 0u, ns, p.cnt);
            else accum_vec<T, VecT, MASKED, MINMAX>(v[u], MASKED ? bits[u] : 0u, slot, p);
        }
    }
    // Vectors past the last full warp tile (< 32*U of them), spread over the grid's first threads.
    {
        const uint64_t gtid = (uint64_t)bid * BLOCK + threadIdx.x;
        for (uint64_t v = ntiles * WTILE + gtid; v < nvec; v += (uint64_t)nblk * BLOCK) {
            uint32_t b = 0;
            if constexpr (MASKED) b = load_valid_bits<VEC>(mask, v * VEC);
            if constexpr (PACKED) accum_vec_packed<T, VecT, MASKED, MINMAX>(ldg_stream(vp + v), b, ns, p.cnt);
            else accum_vec<T, VecT, MASKED, MINMAX>(ldg_stream(vp + v), b, slot, p);
        }
        // Rows past the last full vector (< VEC of them): one thread, row by row.
        if (gtid == 0) {
            for (uint64_t r = nvec * VEC; r < n; ++r) {
                const bool ok = !MASKED || row_valid(mask, r);
                if (ok) {
                    const T x = data[r];
                    if constexpr (Traits<T>::is_float) slot[0] = slot[0] + (A)x;
                    else slot[0] = (A)((uint64_t)slot[0] + (uint64_t)(A)x);
                    if constexpr (MINMAX) { p.mn = comb_min(p.mn, x); p.mx = comb_max(p.mx, x); }
                    if constexpr (MASKED) p.cnt += 1;
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) {
        if constexpr (Traits<T>::is_float) p.sum = p.sum + slot[k];
        else p.sum = (A)((uint64_t)p.sum + (uint64_t)slot[k]);
    }
    if constexpr (PACKED) {
        p.sum = (A)((uint64_t)p.sum + ns.sum);
        if constexpr (MINMAX) { p.mn = comb_min(p.mn, ns.min_value()); p.mx = comb_max(p.mx, ns.max_value()); }
    }

    p = block_combine<P, BLOCK>(p, smem);
    // Everything below touches memory shared with earlier launches on the stream (partials, ticket, out, mailboxes); in the
    // late form the order is kept through x.done (see pdl_* above), otherwise the kernel already waited for its predecessor.
    P q;
    if (nblk == 1) {
        q = p;   // small column: the only block is the finishing block — no partials, no ticket, no fences
    } else {
        if (threadIdx.x == 0) {
            if (late_wait) while (ld_acquire_gpu(x.done) + 2 < x.epoch) {}   // this parity's buffer: its previous user has folded
            partials[bid] = p.raw();
            __threadfence();
            const unsigned int done = atomicAdd(ticket, 1u);
            is_last = (done == nblk - 1);
        }
        __syncthreads();
        if (!is_last) return false;
        __threadfence();
        q.init();
        bool first = true;
        for (unsigned int i = threadIdx.x; i < nblk; i += BLOCK) {
            P t; t.from_raw(load_partial(partials + i));
            if (first) { q = t; first = false; } else q.merge(t);
        }
        q = block_combine<P, BLOCK>(q, smem);
    }
    bool poisoned = false;
    if (x.world > 0) {   // fused cross-GPU finish (block-uniform branch; only the finishing block gets here)
        __shared__ AggRaw mine;
        if (threadIdx.x == 0) {
            if constexpr (!MASKED) q.cnt = n;
            mine = q.raw();
            if (late_wait) while (ld_acquire_gpu(x.done) + 1 < x.epoch) {}   // the previous reduction has left the mailboxes and `out`
        }
        const bool ok = xchg_exchange<BLOCK>(x, &mine, 1u);   // starts with a __syncthreads
        if (threadIdx.x == 0) {
            if (ok) {
                P acc; acc.from_raw(xchg_slot(x, 0, 0));
                for (int r = 1; r < x.world; ++r) { P t; t.from_raw(xchg_slot(x, r, 0)); acc.merge(t); }   // rank order
                q = acc;
            } else poisoned = true;   // a peer never answered: the error word is set, the result is marked unusable
        }
    } else if (threadIdx.x == 0) {
        if constexpr (!MASKED) q.cnt = n;
    }
    if (threadIdx.x == 0) {
        AggRaw r = q.raw();
        if (poisoned) r.count = kAggPoison;
        *out = r;
        if (out_host) {
            // second copy straight into mapped pinned host memory (synchronous APIs: no D2H memcpy, no system fence on the
            // latency path).  host_seq == 0: plain image, visible to the host once the kernel has completed.  host_seq != 0:
            // every 32-bit word of the image travels in its own 8-byte store next to the call's sequence number, so the host
            // can poll the slot and take the result the moment all eight flags match — it does not wait for the stream to
            // drain (api.cu reduce_sync; 8-byte stores are single PCIe writes, the framing NCCL's LL protocol relies on).
            if (host_seq) {
                const uint32_t* w = reinterpret_cast<const uint32_t*>(&r);
                uint64_t* o = reinterpret_cast<uint64_t*>(out_host);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(o + i), "l"(((uint64_t)host_seq << 32) | w[i]) : "memory");
            } else *out_host = r;
        }
        if (nblk > 1) *ticket = 0;   // re-arm for the next launch on this stream
        if (x.world > 0) { __threadfence(); st_release_gpu(x.done, x.epoch); }   // last: the successor may post / write now
    }
    return true;
}

constexpr int kReduceLateWait = 1;   // flags bit 0 of reduce_stats_kernel
constexpr int kReduceNoPdl = 2;      // host-side only: launch without the programmatic-launch attribute
constexpr int kReduceHostSeqShift = 8;   // flags bits 8..31: sequence number of a polled host result (0 = plain image, see out_host)

template <typename T, typename VecT, bool MASKED, bool MINMAX, int BLOCK, int MINB, int U>
__global__ void __launch_bounds__(BLOCK, MINB)
reduce_stats_kernel(const T* __restrict__ data, const uint8_t* __restrict__ mask, uint64_t n,
                    AggRaw* __restrict__ partials, unsigned int* __restrict__ ticket, AggRaw* __restrict__ out,
                    AggRaw* __restrict__ out_host, const XchgDev x, const int flags) {
    pdl_launch_dependents();
    const bool late = (flags & kReduceLateWait) != 0;
    if (!late) pdl_wait();
    reduce_stats_body<T, VecT, MASKED, MINMAX, BLOCK, U>(data, mask, n, partials, ticket, out, out_host, blockIdx.x, gridDim.x, x, late,
                                                         (uint32_t)flags >> kReduceHostSeqShift);
}

// One launch, many columns/chunks (SuperArray / SuperTable fan-out, broadcast/super_table.rs:38-73 walks them one
// by one): blockIdx.y selects the segment, blockIdx.x the block inside it; blocks beyond a segment's own count exit.
struct ReduceSeg {
    const void* data;
    const uint8_t* mask;
    uint64_t n;
    uint32_t nblk;      // blocks working on this segment ( = the single-launch grid for this length)
    uint32_t out_index; // slot in `outs`
};

// Optional second stage of a batched call (a sharded SuperArray / SuperTable reduction): when the LAST chunk aggregate of
// the whole call (all launches of it, counted by `gticket`) has been written, that block folds the chunk aggregates per
// column in chunk order, exchanges the per-column partials with the other ranks (xchg_exchange) and folds those in rank
// order.  n_groups <= BLOCK.  With x.world == 0 it is a purely local per-column fold.
struct FoldArgs {
    unsigned int* gticket;     // NULL = no second stage
    uint32_t total_segs;       // chunk aggregates of the whole call
    uint32_t n_groups;         // columns
    const uint32_t* grp_off;   // [n_groups + 1] offsets into grp_idx
    const uint32_t* grp_idx;   // slots in `outs`, grouped by column, chunk order inside a column
    const AggRaw* identity;    // [n_groups] aggregate of an empty column of that dtype (a rank may hold no chunk of it)
    const uint8_t* kind;       // [n_groups] 0 signed / 1 unsigned / 2 float
    AggRaw* local;             // [n_groups] scratch: this rank's per-column partials
    AggRaw* result;            // [n_groups] device result
    AggRaw* result_host;       // [n_groups] mapped pinned host copy, or NULL
};

template <int BLOCK>
__device__ __forceinline__ void fold_and_exchange(const FoldArgs& f, const AggRaw* __restrict__ outs, const XchgDev& x) {
    const unsigned g = threadIdx.x;
    AggRaw acc{};
    int kind = 0;
    if (g < f.n_groups) {
        kind = f.kind[g];
        acc = f.identity[g];
        for (uint32_t i = f.grp_off[g]; i < f.grp_off[g + 1]; ++i) agg_merge_rt(kind, acc, load_partial(outs + f.grp_idx[i]));
        f.local[g] = acc;
    }
    bool ok = true;
    if (x.world > 0) {
        ok = xchg_exchange<BLOCK>(x, f.local, f.n_groups);   // starts with a __syncthreads: f.local is complete
        if (g < f.n_groups && ok) {
            acc = xchg_slot(x, 0, g);
            for (int r = 1; r < x.world; ++r) agg_merge_rt(kind, acc, xchg_slot(x, (unsigned)r, g));   // rank order
        }
    }
    if (g < f.n_groups) {
        if (!ok) acc.count = kAggPoison;
        f.result[g] = acc;
        if (f.result_host) f.result_host[g] = acc;
    }
    if (threadIdx.x == 0) {
        *f.gticket = 0;   // re-arm
        if (x.world > 0) { __threadfence(); st_release_gpu(x.done, x.epoch); }
    }
}

template <typename T, typename VecT, bool MASKED, bool MINMAX, int BLOCK, int MINB, int U>
__global__ void __launch_bounds__(BLOCK, MINB)
reduce_stats_batch_kernel(const ReduceSeg* __restrict__ segs, AggRaw* __restrict__ partials, unsigned int* __restrict__ tickets,
                          AggRaw* __restrict__ outs, uint32_t max_blk, const FoldArgs f, const XchgDev x) {
    pdl_launch_dependents();   // the next launch of the same batched call may start filling the SMs (no data dependency)
    const ReduceSeg s = segs[blockIdx.y];
    if (blockIdx.x >= s.nblk) return;
    const bool fin = reduce_stats_body<T, VecT, MASKED, MINMAX, BLOCK, U>(static_cast<const T*>(s.data), s.mask, s.n,
                                                                        partials + (size_t)blockIdx.y * max_blk, tickets + blockIdx.y,
                                                                        outs + s.out_index, nullptr, blockIdx.x, s.nblk, XchgDev{}, false);
    if (!fin || f.gticket == nullptr) return;
    __shared__ bool all_done;
    if (threadIdx.x == 0) {
        __threadfence();
        all_done = atomicAdd(f.gticket, 1u) == f.total_segs - 1;
    }
    __syncthreads();
    if (!all_done) return;
    __threadfence();
    fold_and_exchange<BLOCK>(f, outs, x);
}

// A rank that owns no chunk of the sharded container still takes part in the exchange with identity aggregates.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1) fold_exchange_kernel(const AggRaw* __restrict__ outs, const FoldArgs f, const XchgDev x) {
    fold_and_exchange<BLOCK>(f, outs, x);
}

}  // namespace mnr
