// minarrow_b200.hpp — C++17 host side above the C ABI (include/minarrow_b200.h), mirroring the reference's interface
// for the hot path: same names, argument meaning and error behaviour, so tests read like the reference's own.
//
//   apply_int_i32/u32/i64/u64, apply_float_f32/f64   src/kernels/arithmetic/dispatch.rs:74-79,147-152,376-402
//   apply_datetime_i32/u32/i64/u64                   src/kernels/arithmetic/dispatch.rs:300-372,420-427
//   apply_fma_f32/f64                                dispatch.rs:221-226,404-418
//   and_masks / or_masks / xor_masks / not_mask / popcount_mask / all_true_mask / all_false_mask / merge_bitmasks_to_new
//                                                    src/kernels/bitmask/dispatch.rs:96-295, bitmask/mod.rs:171-197
//   Bitmask, IntegerArray<T>, FloatArray<T>, Vec64   src/structs/bitmask.rs:66-71, variants/integer.rs:105-111,
//                                                    variants/float.rs:109-116, vec64 crate (64-byte aligned Vec)
//   DeviceBuffer<T>, DeviceBitmask                   the device-resident types the north-star adds beside Vec64 / Bitmask
//
// Rust `Result<_, KernelError>` -> a thrown `KernelError` (kind = variant name); the reference's dense-integer
// divide-by-zero *panic* (std.rs:54-55) is kind "DivideByZero".  Header-only; link with -lminarrow_b200.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <initializer_list>
#include <new>
#include <optional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "minarrow_b200.h"

namespace minarrow_b200 {

enum class ArithmeticOperator : int { Add = 0, Subtract, Multiply, Divide, Remainder, Power, FloorDiv };   // operators.rs:19-48
enum class LogicalOperator : int { And = 0, Or, Xor };                                                        // operators.rs:88-104

struct KernelError : std::runtime_error {
    int code;
    std::string kind;
    KernelError(int c, std::string k, const std::string& msg) : std::runtime_error(k + ": " + msg), code(c), kind(std::move(k)) {}
};

inline const char* kind_of(int rc) {
    switch (rc) {
        case MNR_ERR_TYPE_MISMATCH: return "TypeMismatch";
        case MNR_ERR_LENGTH_MISMATCH: return "LengthMismatch";
        case MNR_ERR_BROADCASTING: return "BroadcastingError";
        case MNR_ERR_OPERATOR_MISMATCH: return "OperatorMismatch";
        case MNR_ERR_UNSUPPORTED_TYPE: return "UnsupportedType";
        case MNR_ERR_COLUMN_NOT_FOUND: return "ColumnNotFound";
        case MNR_ERR_INVALID_ARGUMENTS: return "InvalidArguments";
        case MNR_ERR_PLAN: return "Plan";
        case MNR_ERR_OUT_OF_BOUNDS: return "OutOfBounds";
        case MNR_ERR_DIVIDE_BY_ZERO: return "DivideByZero";
        case MNR_ERR_NO_DEVICE: return "NoDevice";
        case MNR_ERR_OUT_OF_MEMORY: return "OutOfMemory";
        default: return "Cuda";
    }
}
inline void check(int rc) {
    if (rc != MNR_OK) throw KernelError(rc, kind_of(rc), mnr_last_error());
}

// ---- Vec64: 64-byte aligned vector (vec64 crate) ----------------------------------------------------------------
template <class T> struct Aligned64 {
    using value_type = T;
    Aligned64() = default;
    template <class U> Aligned64(const Aligned64<U>&) {}
    T* allocate(size_t n) {
        void* p = nullptr;
        const size_t bytes = n * sizeof(T);
        if (posix_memalign(&p, 64, bytes != 0 ? bytes : 64) != 0) throw std::bad_alloc();
        return static_cast<T*>(p);
    }
    void deallocate(T* p, size_t) { free(p); }
    template <class U> bool operator==(const Aligned64<U>&) const { return true; }
    template <class U> bool operator!=(const Aligned64<U>&) const { return false; }
};
template <class T> using Vec64 = std::vector<T, Aligned64<T>>;

// ---- Bitmask (Arrow layout: LSB first, 1 = valid / true, slack bits zero) ------------------------------------------
struct Bitmask {
    Vec64<uint8_t> bits;
    size_t len = 0;
    static Bitmask new_set_all(size_t n, bool v) {                       // bitmask.rs:94-105
        Bitmask m;
        m.len = n;
        m.bits.assign((n + 7) / 8, v ? 0xFF : 0);
        if (v && (n & 7)) m.bits.back() &= uint8_t((1u << (n & 7)) - 1u);
        return m;
    }
    static Bitmask from_bools(std::initializer_list<bool> b) {          // bitmask.rs:349-365
        Bitmask m = new_set_all(b.size(), false);
        size_t i = 0;
        for (bool x : b) { if (x) m.bits[i >> 3] |= uint8_t(1u << (i & 7)); ++i; }
        return m;
    }
    bool get(size_t i) const { return (bits[i >> 3] >> (i & 7)) & 1; }
    size_t count_ones() const { size_t c = 0; for (uint8_t x : bits) c += (size_t)__builtin_popcount(x); return c; }
    size_t null_count() const { return len - count_ones(); }
    std::vector<bool> to_bools() const { std::vector<bool> v(len); for (size_t i = 0; i < len; ++i) v[i] = get(i); return v; }
};
using BitmaskVT = std::tuple<const Bitmask&, size_t, size_t>;           // (&Bitmask, offset, len), aliases.rs

template <class T> struct IntegerArray {
    Vec64<T> data;
    std::optional<Bitmask> null_mask;
    size_t len() const { return data.size(); }
    bool is_empty() const { return data.empty(); }
};
template <class T> struct FloatArray {
    Vec64<T> data;
    std::optional<Bitmask> null_mask;
    size_t len() const { return data.size(); }
    bool is_empty() const { return data.empty(); }
};

// DatetimeArray<T> {data, null_mask, time_unit} (structs/variants/datetime/mod.rs:90-140): integer offsets from the epoch.
template <class T> struct DatetimeArray {
    Vec64<T> data;
    std::optional<Bitmask> null_mask;
    std::optional<std::string> time_unit;
    size_t len() const { return data.size(); }
    bool is_empty() const { return data.empty(); }
    static DatetimeArray from_slice(std::initializer_list<T> v, std::optional<std::string> unit = std::nullopt) {
        DatetimeArray a;
        a.data.assign(v.begin(), v.end());
        a.time_unit = std::move(unit);
        return a;
    }
};
template <class T> using DatetimeAVT = std::tuple<const DatetimeArray<T>&, size_t, size_t>;   // (&DatetimeArray, offset, len), aliases.rs

template <class T> struct DType;
template <> struct DType<int32_t> { static constexpr mnr_dtype code = MNR_I32; };
template <> struct DType<uint32_t> { static constexpr mnr_dtype code = MNR_U32; };
template <> struct DType<int64_t> { static constexpr mnr_dtype code = MNR_I64; };
template <> struct DType<uint64_t> { static constexpr mnr_dtype code = MNR_U64; };
template <> struct DType<float> { static constexpr mnr_dtype code = MNR_F32; };
template <> struct DType<double> { static constexpr mnr_dtype code = MNR_F64; };
template <> struct DType<int8_t> { static constexpr mnr_dtype code = MNR_I8; };
template <> struct DType<uint8_t> { static constexpr mnr_dtype code = MNR_U8; };
template <> struct DType<int16_t> { static constexpr mnr_dtype code = MNR_I16; };
template <> struct DType<uint16_t> { static constexpr mnr_dtype code = MNR_U16; };

// ---- context --------------------------------------------------------------------------------------------------------
class Context {
public:
    explicit Context(int device = 0) { check(mnr_ctx_create(device, &h_)); }
    ~Context() { mnr_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    mnr_ctx* get() const { return h_; }
    uint64_t launch_count() const { return mnr_ctx_launch_count(h_); }
    void synchronize() { check(mnr_ctx_synchronize(h_)); }
    static Context& thread_default() { thread_local Context c(0); return c; }   // one context per host thread
private:
    mnr_ctx* h_ = nullptr;
};

// ---- leaf kernels: host slices in, fresh typed array out ---------------------------------------------------------------
namespace detail {
template <class T, class Arr>
Arr apply(Context& ctx, const T* lhs, size_t ln, const T* rhs, size_t rn, ArithmeticOperator op, const Bitmask* mask) {
    Arr out;
    out.data.resize(ln == rn ? ln : 0);
    if (mask) out.null_mask = Bitmask::new_set_all(ln == rn ? ln : 0, false);     // Some(mask) iff a mask was passed (dispatch.rs:90-104)
    check(mnr_apply_host(ctx.get(), DType<T>::code, static_cast<mnr_op>(op), lhs, ln, rhs, rn, mask ? mask->bits.data() : nullptr,
                         out.data.data(), mask ? out.null_mask->bits.data() : nullptr));
    return out;
}
template <class T>
FloatArray<T> fma(Context& ctx, const T* a, size_t an, const T* b, size_t bn, const T* c, size_t cn, const Bitmask* mask) {
    FloatArray<T> out;
    out.data.resize((an == bn && an == cn) ? an : 0);
    if (mask) out.null_mask = Bitmask::new_set_all(out.data.size(), false);
    check(mnr_apply_fma_host(ctx.get(), DType<T>::code, a, an, b, bn, c, cn, mask ? mask->bits.data() : nullptr, out.data.data(),
                             mask ? out.null_mask->bits.data() : nullptr));
    return out;
}
}  // namespace detail

#define MNR_CPP_APPLY(NAME, T, ARR)                                                                                     \
    template <class L, class R>                                                                                         \
    ARR<T> NAME(const L& lhs, const R& rhs, ArithmeticOperator op, const Bitmask* mask = nullptr,                       \
                Context& ctx = Context::thread_default()) {                                                             \
        return detail::apply<T, ARR<T>>(ctx, lhs.data(), lhs.size(), rhs.data(), rhs.size(), op, mask);                 \
    }
MNR_CPP_APPLY(apply_int_i32, int32_t, IntegerArray)
MNR_CPP_APPLY(apply_int_u32, uint32_t, IntegerArray)
MNR_CPP_APPLY(apply_int_i64, int64_t, IntegerArray)
MNR_CPP_APPLY(apply_int_u64, uint64_t, IntegerArray)
MNR_CPP_APPLY(apply_int_i16, int16_t, IntegerArray)
MNR_CPP_APPLY(apply_int_u16, uint16_t, IntegerArray)
MNR_CPP_APPLY(apply_int_i8, int8_t, IntegerArray)
MNR_CPP_APPLY(apply_int_u8, uint8_t, IntegerArray)
MNR_CPP_APPLY(apply_float_f32, float, FloatArray)
MNR_CPP_APPLY(apply_float_f64, double, FloatArray)
#undef MNR_CPP_APPLY

template <class A, class B, class C_>
FloatArray<float> apply_fma_f32(const A& l, const B& r, const C_& acc, const Bitmask* mask = nullptr, Context& ctx = Context::thread_default()) {
    return detail::fma<float>(ctx, l.data(), l.size(), r.data(), r.size(), acc.data(), acc.size(), mask);
}
template <class A, class B, class C_>
FloatArray<double> apply_fma_f64(const A& l, const B& r, const C_& acc, const Bitmask* mask = nullptr, Context& ctx = Context::thread_default()) {
    return detail::fma<double>(ctx, l.data(), l.size(), r.data(), r.size(), acc.data(), acc.size(), mask);
}

// ---- device-resident types ------------------------------------------------------------------------------------------------
class DeviceBitmask {
public:
    DeviceBitmask(Context& ctx, const Bitmask& m) : ctx_(&ctx) { check(mnr_bits_upload(ctx.get(), m.bits.data(), m.len, &h_)); }
    DeviceBitmask(Context& ctx, mnr_bits* h) : ctx_(&ctx), h_(h) {}
    DeviceBitmask(DeviceBitmask&& o) noexcept : ctx_(o.ctx_), h_(o.h_) { o.h_ = nullptr; }
    DeviceBitmask(const DeviceBitmask&) = delete;
    ~DeviceBitmask() { if (h_) mnr_bits_free(h_); }
    size_t len() const { return mnr_bits_len(h_); }
    mnr_bits* get() const { return h_; }
    Bitmask download() const {
        Bitmask m = Bitmask::new_set_all(len(), false);
        check(mnr_bits_download(ctx_->get(), h_, m.bits.data()));
        return m;
    }
    uint64_t count_ones() const { uint64_t n = 0; check(mnr_bits_popcount(ctx_->get(), h_, 0, len(), &n)); return n; }
private:
    Context* ctx_;
    mnr_bits* h_ = nullptr;
};

template <class T> class DeviceBuffer {
public:
    template <class V> DeviceBuffer(Context& ctx, const V& host) : ctx_(&ctx) {
        check(mnr_buf_upload(ctx.get(), DType<T>::code, host.data(), host.size(), &h_));
    }
    DeviceBuffer(Context& ctx, mnr_buf* h) : ctx_(&ctx), h_(h) {}
    DeviceBuffer(DeviceBuffer&& o) noexcept : ctx_(o.ctx_), h_(o.h_) { o.h_ = nullptr; }
    DeviceBuffer(const DeviceBuffer&) = delete;
    ~DeviceBuffer() { if (h_) mnr_buf_free(h_); }
    size_t len() const { return mnr_buf_len(h_); }
    mnr_buf* get() const { return h_; }
    Vec64<T> download() const {
        Vec64<T> v(len());
        check(mnr_buf_download(ctx_->get(), h_, v.data()));
        return v;
    }
    // Fused null-aware arithmetic with the validity merge inside the kernel (AND = merge_bitmasks_to_new, OR = Bitmask::union).
    std::pair<DeviceBuffer<T>, std::optional<DeviceBitmask>> binary(ArithmeticOperator op, const DeviceBuffer<T>& rhs,
                                                                    const DeviceBitmask* lm = nullptr, const DeviceBitmask* rm = nullptr,
                                                                    mnr_mask_mode mode = MNR_MASK_AND) const {
        mnr_buf* ob = nullptr; mnr_bits* om = nullptr;
        check(mnr_ew_binary(ctx_->get(), static_cast<mnr_op>(op), h_, rhs.h_, lm ? lm->get() : nullptr, rm ? rm->get() : nullptr, mode, &ob, &om));
        std::optional<DeviceBitmask> m;
        if (om) m.emplace(*ctx_, om);
        return {DeviceBuffer<T>(*ctx_, ob), std::move(m)};
    }
    mnr_agg stats(const DeviceBitmask* validity = nullptr) const {
        mnr_agg a{};
        check(mnr_reduce_stats(ctx_->get(), h_, validity ? validity->get() : nullptr, &a));
        return a;
    }
private:
    Context* ctx_;
    mnr_buf* h_ = nullptr;
};

// ---- datetime delegation: apply_datetime_{i32,u32,i64,u64} (dispatch.rs:300-372, 420-427) -----------------------------------
// The integer kernels over the two data windows; output validity = merge_bitmasks_to_new(lhs.null_mask, rhs.null_mask, llen),
// i.e. bits [0, llen) of each ARRAY's mask (the reference does not offset them by the window start, :321-322), fused into the
// launch as two mask operands.  Some(mask) iff either side has one.
namespace detail {
template <class T>
DatetimeArray<T> apply_datetime(Context& ctx, DatetimeAVT<T> lhs, DatetimeAVT<T> rhs, ArithmeticOperator op) {
    const auto& [la, lo, ll] = lhs;
    const auto& [ra, ro, rl] = rhs;
    if (ll != rl) throw KernelError(MNR_ERR_LENGTH_MISMATCH, "LengthMismatch", "apply_datetime: length mismatch");
    if (lo + ll > la.data.size() || ro + rl > ra.data.size()) throw KernelError(MNR_ERR_OUT_OF_BOUNDS, "OutOfBounds", "apply_datetime: window leaves the array");
    DatetimeArray<T> out;
    out.time_unit = la.time_unit;
    const bool any_mask = la.null_mask.has_value() || ra.null_mask.has_value();
    if (ll == 0) { if (any_mask) out.null_mask = Bitmask::new_set_all(0, false); return out; }
    auto up_mask = [&](const std::optional<Bitmask>& m) -> std::optional<DeviceBitmask> {
        if (!m) return std::nullopt;
        if (m->len < ll) throw KernelError(MNR_ERR_INVALID_ARGUMENTS, "InvalidArguments", "Bitmask too short in merge");
        Bitmask w = Bitmask::new_set_all(ll, false);
        std::copy(m->bits.begin(), m->bits.begin() + (ll + 7) / 8, w.bits.begin());
        if (ll & 7) w.bits.back() &= uint8_t((1u << (ll & 7)) - 1u);
        return DeviceBitmask(ctx, w);
    };
    std::optional<DeviceBitmask> lm = up_mask(la.null_mask), rm = up_mask(ra.null_mask);
    mnr_buf *L = nullptr, *R = nullptr;
    check(mnr_buf_upload(ctx.get(), DType<T>::code, la.data.data() + lo, ll, &L));
    DeviceBuffer<T> dl(ctx, L);
    check(mnr_buf_upload(ctx.get(), DType<T>::code, ra.data.data() + ro, rl, &R));
    DeviceBuffer<T> dr(ctx, R);
    auto [ob, om] = dl.binary(op, dr, lm ? &*lm : nullptr, rm ? &*rm : nullptr, MNR_MASK_AND);
    out.data = ob.download();
    if (om) out.null_mask = om->download();
    return out;
}
}  // namespace detail
inline DatetimeArray<int32_t> apply_datetime_i32(DatetimeAVT<int32_t> l, DatetimeAVT<int32_t> r, ArithmeticOperator op, Context& ctx = Context::thread_default()) { return detail::apply_datetime<int32_t>(ctx, l, r, op); }
inline DatetimeArray<uint32_t> apply_datetime_u32(DatetimeAVT<uint32_t> l, DatetimeAVT<uint32_t> r, ArithmeticOperator op, Context& ctx = Context::thread_default()) { return detail::apply_datetime<uint32_t>(ctx, l, r, op); }
inline DatetimeArray<int64_t> apply_datetime_i64(DatetimeAVT<int64_t> l, DatetimeAVT<int64_t> r, ArithmeticOperator op, Context& ctx = Context::thread_default()) { return detail::apply_datetime<int64_t>(ctx, l, r, op); }
inline DatetimeArray<uint64_t> apply_datetime_u64(DatetimeAVT<uint64_t> l, DatetimeAVT<uint64_t> r, ArithmeticOperator op, Context& ctx = Context::thread_default()) { return detail::apply_datetime<uint64_t>(ctx, l, r, op); }

// ---- bitmask kernels over host masks (BitmaskVT windows) ------------------------------------------------------------------
namespace detail {
inline Bitmask binop(Context& ctx, LogicalOperator op, BitmaskVT l, BitmaskVT r) {
    const auto& [lm, lo, ll] = l;
    const auto& [rm, ro, rl] = r;
    if (ll != rl) throw KernelError(MNR_ERR_LENGTH_MISMATCH, "LengthMismatch", "bitmask_binop: window lengths differ");
    // raw host pointers cross the C ABI below: the windows must lie inside their masks' own storage (the reference would
    // panic on the slice, bitmask/mod.rs:124-139)
    auto inside = [](const Bitmask& m, size_t off, size_t len) { return len == 0 || (off / 8 + (len + 7) / 8 <= m.bits.size() && off + len <= m.bits.size() * 8); };
    if (!inside(lm, lo, ll) || !inside(rm, ro, rl)) throw KernelError(MNR_ERR_OUT_OF_BOUNDS, "OutOfBounds", "bitmask_binop: window leaves the mask");
    Bitmask out = Bitmask::new_set_all(ll, false);
    check(mnr_bitmask_binop_host(ctx.get(), static_cast<mnr_logical_op>(op), lm.bits.data(), lo, rm.bits.data(), ro, ll, out.bits.data()));
    return out;
}
}  // namespace detail
inline Bitmask and_masks(BitmaskVT l, BitmaskVT r, Context& ctx = Context::thread_default()) { return detail::binop(ctx, LogicalOperator::And, l, r); }
inline Bitmask or_masks(BitmaskVT l, BitmaskVT r, Context& ctx = Context::thread_default()) { return detail::binop(ctx, LogicalOperator::Or, l, r); }
inline Bitmask xor_masks(BitmaskVT l, BitmaskVT r, Context& ctx = Context::thread_default()) { return detail::binop(ctx, LogicalOperator::Xor, l, r); }
inline Bitmask not_mask(BitmaskVT s, Context& ctx = Context::thread_default()) {
    const auto& [m, off, len] = s;
    DeviceBitmask d(ctx, m);
    mnr_bits* o = nullptr;
    check(mnr_bits_not(ctx.get(), d.get(), off, len, &o));
    return DeviceBitmask(ctx, o).download();
}
inline size_t popcount_mask(BitmaskVT s, Context& ctx = Context::thread_default()) {
    const auto& [m, off, len] = s;
    DeviceBitmask d(ctx, m);
    uint64_t n = 0;
    check(mnr_bits_popcount(ctx.get(), d.get(), off, len, &n));
    return (size_t)n;
}
inline bool all_true_mask(const Bitmask& m, Context& ctx = Context::thread_default()) {
    DeviceBitmask d(ctx, m); int r = 0; check(mnr_bits_all_true(ctx.get(), d.get(), &r)); return r != 0;
}
inline bool all_false_mask(const Bitmask& m, Context& ctx = Context::thread_default()) {
    DeviceBitmask d(ctx, m); int r = 0; check(mnr_bits_all_false(ctx.get(), d.get(), &r)); return r != 0;
}
// merge_bitmasks_to_new (bitmask/mod.rs:171-197): AND of the two validity masks; a missing side means "no nulls".
inline std::optional<Bitmask> merge_bitmasks_to_new(const Bitmask* l, const Bitmask* r, size_t len, Context& ctx = Context::thread_default()) {
    if (!l && !r) return std::nullopt;
    std::optional<DeviceBitmask> dl, dr;
    if (l) dl.emplace(ctx, *l);
    if (r) dr.emplace(ctx, *r);
    mnr_bits* o = nullptr;
    check(mnr_bits_merge(ctx.get(), dl ? dl->get() : nullptr, dr ? dr->get() : nullptr, len, MNR_MASK_AND, &o));
    return DeviceBitmask(ctx, o).download();
}

// ---- reductions over host columns -----------------------------------------------------------------------------------------
template <class V> mnr_agg stats(const V& data, const Bitmask* validity = nullptr, bool with_minmax = true, Context& ctx = Context::thread_default()) {
    using T = typename V::value_type;
    mnr_agg a{};
    check(mnr_stats_host(ctx.get(), DType<T>::code, data.data(), data.size(), validity ? validity->bits.data() : nullptr, with_minmax ? 1 : 0, &a));
    return a;
}

}  // namespace minarrow_b200
