// minarrow_b200_containers.hpp — the callers of the leaf kernels, C++17 over the C ABI: the router and the container
// fan-out routes of the reference with the same names, argument order and error behaviour.
//
//   Array / NumericArray / ArrayV            src/enums/array.rs:118-126, enums/collections/numeric_array.rs:81-101,
//                                            structs/views/array_view.rs  ((Array, offset, len))
//   resolve_binary_arithmetic                src/kernels/routing/arithmetic.rs:214-406
//   maybe_broadcast_scalar_array             src/kernels/routing/broadcast.rs:87-112 (length-1 operand)
//   SuperArray + route_super_array_broadcast src/structs/chunked/super_array.rs:96-103, kernels/broadcast/super_array.rs:180-249
//   Table + broadcast_table_with_operator    src/structs/table.rs:103-115, kernels/broadcast/table.rs:31-62
//   SuperTable + broadcast_super_table_with_operator   structs/chunked/super_table.rs:78-83, broadcast/super_table.rs:38-73
//   table/array/scalar fan-outs              broadcast/table.rs:230-261 and neighbours
//
// What differs from the reference is performance only:
//   * a length-1 operand travels in the kernel arguments instead of being materialised `len` times
//     (broadcast_length_1_array, routing/broadcast.rs:25-47);
//   * (i32, f64) / (i32, f32) pairs are cast on load inside the kernel instead of through two copied Vec64s
//     (routing/arithmetic.rs:244-269);
//   * the SuperArray route's per-chunk validity merge (Bitmask::union, super_array.rs:214-230) is fused into the
//     arithmetic kernel (MNR_MASK_OR), and the chunk loop ("TODO: Parallelise", super_array.rs:193) is ONE batched
//     launch per (dtype, alignment, masked) class (mnr_ew_binary_batch);
//   * super_array_stats: per-chunk partial aggregates in one batched launch, folded in chunk order (mnr_agg_combine).
// `MinarrowError::ShapeError` is a thrown KernelError of kind "ShapeError".
#pragma once
#include <algorithm>
#include <cstring>
#include <memory>
#include <variant>

#include "minarrow_b200.hpp"

namespace minarrow_b200 {

template <class T> using ArcInt = std::shared_ptr<const IntegerArray<T>>;
template <class T> using ArcFloat = std::shared_ptr<const FloatArray<T>>;
// NumericArray (numeric_array.rs:81-101): Arc-wrapped typed arrays behind one tag; cloning is a reference-count bump.
using NumericArray = std::variant<ArcInt<int32_t>, ArcInt<int64_t>, ArcInt<uint32_t>, ArcInt<uint64_t>, ArcFloat<float>, ArcFloat<double>,
                                  ArcInt<int8_t>, ArcInt<int16_t>, ArcInt<uint8_t>, ArcInt<uint16_t>>;

inline constexpr int MNR_ERR_SHAPE = -100;   // MinarrowError::ShapeError (src/enums/error.rs): a host-side error, never returned by the C ABI

class Array {
public:
    NumericArray v;
    Array() : v(ArcInt<int32_t>(std::make_shared<IntegerArray<int32_t>>())) {}
    template <class T> explicit Array(IntegerArray<T> a) : v(ArcInt<T>(std::make_shared<const IntegerArray<T>>(std::move(a)))) {}
    template <class T> explicit Array(FloatArray<T> a) : v(ArcFloat<T>(std::make_shared<const FloatArray<T>>(std::move(a)))) {}
    static Array from_int32(IntegerArray<int32_t> a) { return Array(std::move(a)); }      // array.rs Array::from_int32 ...
    static Array from_int64(IntegerArray<int64_t> a) { return Array(std::move(a)); }
    static Array from_uint32(IntegerArray<uint32_t> a) { return Array(std::move(a)); }
    static Array from_uint64(IntegerArray<uint64_t> a) { return Array(std::move(a)); }
    static Array from_float32(FloatArray<float> a) { return Array(std::move(a)); }
    static Array from_float64(FloatArray<double> a) { return Array(std::move(a)); }
    template <class T> static Array from_slice(std::initializer_list<T> vals, const Bitmask* mask = nullptr) {
        if constexpr (std::is_floating_point<T>::value) {
            FloatArray<T> a; a.data.assign(vals.begin(), vals.end()); if (mask) a.null_mask = *mask; return Array(std::move(a));
        } else {
            IntegerArray<T> a; a.data.assign(vals.begin(), vals.end()); if (mask) a.null_mask = *mask; return Array(std::move(a));
        }
    }
    size_t len() const { return std::visit([](const auto& p) { return p->data.size(); }, v); }
    mnr_dtype dtype() const {
        return std::visit([](const auto& p) { return DType<typename std::decay_t<decltype(p->data)>::value_type>::code; }, v);
    }
    const void* data_ptr() const { return std::visit([](const auto& p) { return static_cast<const void*>(p->data.data()); }, v); }
    const Bitmask* null_mask() const {
        return std::visit([](const auto& p) -> const Bitmask* { return p->null_mask ? &*p->null_mask : nullptr; }, v);
    }
    // `if let Array::NumericArray(NumericArray::Int32(arr)) = result` -> result.values<int32_t>() (nullptr for another variant)
    template <class T> const Vec64<T>* values() const {
        if constexpr (std::is_floating_point<T>::value) {
            if (auto p = std::get_if<ArcFloat<T>>(&v)) return &(*p)->data;
        } else {
            if (auto p = std::get_if<ArcInt<T>>(&v)) return &(*p)->data;
        }
        return nullptr;
    }
};

// ArrayV = (Array, offset, len) window (array_view.rs).
struct ArrayV {
    Array array;
    size_t offset = 0, len = 0;
    ArrayV(Array a) : array(std::move(a)), offset(0), len(array.len()) {}                 // impl From<Array> for ArrayV
    ArrayV(Array a, size_t off, size_t n) : array(std::move(a)), offset(off), len(n) {
        if (off + n > array.len()) throw KernelError(MNR_ERR_OUT_OF_BOUNDS, "OutOfBounds", "ArrayV window exceeds the array");
    }
    static ArrayV make(Array a, size_t off, size_t n) { return ArrayV(std::move(a), off, n); }   // ArrayV::new
};

namespace detail {
inline size_t esize(mnr_dtype d) {
    switch (d) { case MNR_I8: case MNR_U8: return 1; case MNR_I16: case MNR_U16: return 2; case MNR_I32: case MNR_U32: case MNR_F32: return 4; default: return 8; }
}
inline bool routed_dtype(mnr_dtype d) {   // the six arms of arithmetic_dispatch (routing/arithmetic.rs:277-337)
    return d == MNR_I32 || d == MNR_I64 || d == MNR_U32 || d == MNR_U64 || d == MNR_F32 || d == MNR_F64;
}
// owning handles of device buffers / bitmasks (freed on scope exit, also when a later call throws)
struct BufH {
    mnr_buf* h = nullptr;
    BufH() = default;
    BufH(const BufH&) = delete;
    BufH(BufH&& o) noexcept : h(o.h) { o.h = nullptr; }
    BufH& operator=(BufH&& o) noexcept { if (this != &o) { if (h) mnr_buf_free(h); h = o.h; o.h = nullptr; } return *this; }
    ~BufH() { if (h) mnr_buf_free(h); }
};
struct BitsH {
    mnr_bits* h = nullptr;
    BitsH() = default;
    BitsH(const BitsH&) = delete;
    BitsH(BitsH&& o) noexcept : h(o.h) { o.h = nullptr; }
    BitsH& operator=(BitsH&& o) noexcept { if (this != &o) { if (h) mnr_bits_free(h); h = o.h; o.h = nullptr; } return *this; }
    ~BitsH() { if (h) mnr_bits_free(h); }
};

inline BufH upload_window(Context& ctx, const ArrayV& a) {
    BufH b;
    const mnr_dtype dt = a.array.dtype();
    check(mnr_buf_upload(ctx.get(), dt, static_cast<const char*>(a.array.data_ptr()) + a.offset * esize(dt), a.len, &b.h));
    return b;
}
inline BitsH upload_mask(Context& ctx, const Bitmask* m, size_t need) {
    BitsH b;
    if (!m) return b;
    if (m->len < need) throw KernelError(MNR_ERR_INVALID_ARGUMENTS, "InvalidArguments", "mask has " + std::to_string(m->len) + " bits, need " + std::to_string(need));
    check(mnr_bits_upload(ctx.get(), m->bits.data(), need, &b.h));
    return b;
}
template <class T> Array download_as(Context& ctx, mnr_buf* ob, mnr_bits* om) {
    BufH bo; bo.h = ob; BitsH mo; mo.h = om;
    const size_t n = mnr_buf_len(ob);
    std::optional<Bitmask> mask;
    if (om) { mask = Bitmask::new_set_all(n, false); check(mnr_bits_download(ctx.get(), om, mask->bits.data())); }
    if constexpr (std::is_floating_point<T>::value) {
        FloatArray<T> a; a.data.resize(n); check(mnr_buf_download(ctx.get(), ob, a.data.data())); a.null_mask = std::move(mask); return Array(std::move(a));
    } else {
        IntegerArray<T> a; a.data.resize(n); check(mnr_buf_download(ctx.get(), ob, a.data.data())); a.null_mask = std::move(mask); return Array(std::move(a));
    }
}
inline Array download(Context& ctx, mnr_dtype dt, mnr_buf* ob, mnr_bits* om) {
    switch (dt) {
        case MNR_I32: return download_as<int32_t>(ctx, ob, om);
        case MNR_I64: return download_as<int64_t>(ctx, ob, om);
        case MNR_U32: return download_as<uint32_t>(ctx, ob, om);
        case MNR_U64: return download_as<uint64_t>(ctx, ob, om);
        case MNR_F32: return download_as<float>(ctx, ob, om);
        case MNR_F64: return download_as<double>(ctx, ob, om);
        case MNR_I8: return download_as<int8_t>(ctx, ob, om);
        case MNR_I16: return download_as<int16_t>(ctx, ob, om);
        case MNR_U8: return download_as<uint8_t>(ctx, ob, om);
        default: return download_as<uint16_t>(ctx, ob, om);
    }
}
// The first element of a length-1 operand as one host element of `out_dt`.  Like broadcast_length_1_array it reads
// data[0] of the underlying array, not data[offset] (routing/broadcast.rs:29-46).
inline uint64_t scalar_bits(const Array& a, mnr_dtype out_dt) {
    uint64_t bits = 0;
    std::visit([&](const auto& p) {
        using S = typename std::decay_t<decltype(p->data)>::value_type;
        const S s = p->data.at(0);
        if (out_dt == MNR_F64) { const double d = (double)s; std::memcpy(&bits, &d, 8); }
        else if (out_dt == MNR_F32) { const float f = (float)s; std::memcpy(&bits, &f, 4); }
        else std::memcpy(&bits, &s, sizeof(S));
    }, a.v);
    return bits;
}
// i32 window -> float column on the host (`x as f64` / `x as f32`): only for the scalar-broadcast + promotion corner.
template <class F> Array cast_i32_window(const ArrayV& a) {
    const auto* src = a.array.values<int32_t>();
    FloatArray<F> out;
    out.data.resize(a.len);
    for (size_t i = 0; i < a.len; ++i) out.data[i] = (F)(*src)[a.offset + i];
    return Array(std::move(out));
}

// One chunk through the router with up to two validity masks merged inside the kernel.
inline Array route_chunk(Context& ctx, ArithmeticOperator op, const ArrayV& lhs, const ArrayV& rhs, const Bitmask* lmask, const Bitmask* rmask,
                         mnr_mask_mode mode) {
    const size_t l = lhs.len, r = rhs.len;
    if (l != r && l != 1 && r != 1)
        throw KernelError(MNR_ERR_LENGTH_MISMATCH, "LengthMismatch", "cannot broadcast arrays of length " + std::to_string(l) + " and " + std::to_string(r));
    const mnr_dtype lt = lhs.array.dtype(), rt = rhs.array.dtype();
    mnr_dtype out_dt = lt;
    bool promote = false;
    if (lt != rt) {
        const bool li = lt == MNR_I32, ri = rt == MNR_I32;
        if ((li && rt == MNR_F64) || (ri && lt == MNR_F64)) out_dt = MNR_F64;
        else if ((li && rt == MNR_F32) || (ri && lt == MNR_F32)) out_dt = MNR_F32;
        else throw KernelError(MNR_ERR_UNSUPPORTED_TYPE, "UnsupportedType", "Unsupported array type combination for arithmetic operations");
        promote = true;
    } else if (!routed_dtype(lt)) {
        throw KernelError(MNR_ERR_UNSUPPORTED_TYPE, "UnsupportedType", "Unsupported array type combination for arithmetic operations");
    }
    const size_t n = std::max(l, r);
    BitsH lm = upload_mask(ctx, lmask, n), rm = upload_mask(ctx, rmask, n);
    mnr_buf* ob = nullptr;
    mnr_bits* om = nullptr;
    if (l != r) {   // maybe_broadcast_scalar_array: the length-1 side becomes a kernel argument
        const bool scalar_is_lhs = l == 1;
        const ArrayV& arr = scalar_is_lhs ? rhs : lhs;
        const uint64_t sbits = scalar_bits((scalar_is_lhs ? lhs : rhs).array, out_dt);
        ArrayV arr_c = arr;
        if (promote && arr.array.dtype() == MNR_I32) arr_c = ArrayV(out_dt == MNR_F64 ? cast_i32_window<double>(arr) : cast_i32_window<float>(arr));
        BufH a = upload_window(ctx, arr_c);
        BitsH merged;
        const mnr_bits* m = lm.h ? lm.h : rm.h;
        if (lm.h && rm.h) { check(mnr_bits_merge(ctx.get(), lm.h, rm.h, n, mode, &merged.h)); m = merged.h; }
        check(mnr_ew_scalar(ctx.get(), static_cast<mnr_op>(op), a.h, &sbits, scalar_is_lhs ? 1 : 0, m, &ob, &om));
    } else {
        BufH a = upload_window(ctx, lhs), b = upload_window(ctx, rhs);
        if (promote) check(mnr_ew_binary_promote(ctx.get(), static_cast<mnr_op>(op), a.h, b.h, lm.h, rm.h, mode, &ob, &om));
        else check(mnr_ew_binary(ctx.get(), static_cast<mnr_op>(op), a.h, b.h, lm.h, rm.h, mode, &ob, &om));
    }
    return download(ctx, out_dt, ob, om);
}
}  // namespace detail

// resolve_binary_arithmetic (routing/arithmetic.rs:214-222): length-1 broadcast, dtype match / i32 -> float promotion,
// window slicing, then the leaf kernel.  `null_mask` is the single pre-merged mask of the leaf API indexed from bit 0;
// the operands' own masks are NOT consulted (that is the caller's job, exactly like the reference).
inline Array resolve_binary_arithmetic(ArithmeticOperator op, const ArrayV& lhs, const ArrayV& rhs, const Bitmask* null_mask = nullptr,
                                       Context& ctx = Context::thread_default()) {
    return detail::route_chunk(ctx, op, lhs, rhs, null_mask, nullptr, MNR_MASK_AND);
}
// broadcast_array_add & friends (kernels/broadcast/array.rs): thin names over the router.
inline Array broadcast_array_add(const ArrayV& l, const ArrayV& r, const Bitmask* m = nullptr, Context& ctx = Context::thread_default()) {
    return resolve_binary_arithmetic(ArithmeticOperator::Add, l, r, m, ctx);
}

// ---- SuperArray ----------------------------------------------------------------------------------------------------------
struct SuperArray {
    std::vector<Array> chunks_;
    static SuperArray from_chunks(std::vector<Array> c) { SuperArray s; s.chunks_ = std::move(c); return s; }
    const std::vector<Array>& chunks() const { return chunks_; }
    void push(Array a) { chunks_.push_back(std::move(a)); }
    size_t n_chunks() const { return chunks_.size(); }
    size_t len() const { size_t n = 0; for (const auto& c : chunks_) n += c.len(); return n; }
    std::vector<size_t> shape_1d() const { std::vector<size_t> s; for (const auto& c : chunks_) s.push_back(c.len()); return s; }
};

namespace detail {
inline std::string shape_str(const std::vector<size_t>& s) {
    std::string o = "[";
    for (size_t i = 0; i < s.size(); ++i) o += (i ? ", " : "") + std::to_string(s[i]);
    return o + "]";
}
}  // namespace detail

// route_super_array_broadcast (broadcast/super_array.rs:180-249): chunk i of lhs against chunk i of rhs; the chunk's
// validity is the override if given, else the union of the two chunks' masks, else the one that exists.
inline SuperArray route_super_array_broadcast(ArithmeticOperator op, const SuperArray& lhs, const SuperArray& rhs,
                                              const Bitmask* null_mask_override = nullptr, Context& ctx = Context::thread_default()) {
    if (rhs.n_chunks() < lhs.n_chunks())
        throw KernelError(MNR_ERR_SHAPE, "ShapeError", "Super Array broadcasting error - chunk count: LHS " + std::to_string(lhs.n_chunks()) + " RHS " + std::to_string(rhs.n_chunks()));
    const size_t nc = lhs.n_chunks();
    for (size_t i = 0; i < nc; ++i) {
        const size_t ll = lhs.chunks_[i].len(), rl = rhs.chunks_[i].len();
        if (ll != rl)
            throw KernelError(MNR_ERR_SHAPE, "ShapeError", "Super Array broadcasting error - Chunk: LHS " + std::to_string(ll) + " RHS " + std::to_string(rl) +
                                                              ", Shape: LHS " + detail::shape_str(lhs.shape_1d()) + " RHS " + detail::shape_str(rhs.shape_1d()));
    }
    SuperArray out;
    // Same-dtype chunk pairs on the six routed types share batched launches; the rest goes chunk by chunk.
    bool batchable = nc > 1;
    for (size_t i = 0; i < nc && batchable; ++i)
        batchable = lhs.chunks_[i].dtype() == rhs.chunks_[i].dtype() && detail::routed_dtype(lhs.chunks_[i].dtype()) && lhs.chunks_[i].len() > 0;
    if (!batchable) {
        for (size_t i = 0; i < nc; ++i) {
            const Bitmask* lm = null_mask_override ? null_mask_override : lhs.chunks_[i].null_mask();
            const Bitmask* rm = null_mask_override ? nullptr : rhs.chunks_[i].null_mask();
            try {
                out.push(detail::route_chunk(ctx, op, ArrayV(lhs.chunks_[i]), ArrayV(rhs.chunks_[i]), lm, rm, MNR_MASK_OR));
            } catch (const KernelError& e) {
                throw KernelError(e.code, e.kind, std::string("Super Array broadcasting error - Error: ") + e.what());
            }
        }
        return out;
    }
    std::vector<detail::BufH> L(nc), R(nc);
    std::vector<detail::BitsH> LM(nc), RM(nc);
    std::vector<const mnr_buf*> lp(nc), rp(nc);
    std::vector<const mnr_bits*> lmp(nc), rmp(nc);
    for (size_t i = 0; i < nc; ++i) {
        L[i] = detail::upload_window(ctx, ArrayV(lhs.chunks_[i]));
        R[i] = detail::upload_window(ctx, ArrayV(rhs.chunks_[i]));
        const size_t n = lhs.chunks_[i].len();
        LM[i] = detail::upload_mask(ctx, null_mask_override ? null_mask_override : lhs.chunks_[i].null_mask(), n);
        RM[i] = detail::upload_mask(ctx, null_mask_override ? nullptr : rhs.chunks_[i].null_mask(), n);
        lp[i] = L[i].h; rp[i] = R[i].h; lmp[i] = LM[i].h; rmp[i] = RM[i].h;
    }
    std::vector<mnr_buf*> ob(nc, nullptr);
    std::vector<mnr_bits*> om(nc, nullptr);
    try {
        check(mnr_ew_binary_batch(ctx.get(), static_cast<mnr_op>(op), nc, lp.data(), rp.data(), lmp.data(), rmp.data(), MNR_MASK_OR, ob.data(), om.data()));
    } catch (const KernelError& e) {
        throw KernelError(e.code, e.kind, std::string("Super Array broadcasting error - Error: ") + e.what());
    }
    for (size_t i = 0; i < nc; ++i) out.push(detail::download(ctx, lhs.chunks_[i].dtype(), ob[i], om[i]));
    return out;
}
inline SuperArray broadcast_super_array_add(const SuperArray& l, const SuperArray& r, const Bitmask* m = nullptr, Context& ctx = Context::thread_default()) {
    return route_super_array_broadcast(ArithmeticOperator::Add, l, r, m, ctx);
}

// Null-aware sum / count / min / max over all chunks: per-chunk partials in one batched launch per class, folded in
// chunk order (benches/benchmark_parallel_simd.rs:81-97 — par_chunks -> chunk sums -> combine).
inline mnr_agg super_array_stats(const SuperArray& a, bool with_minmax = true, Context& ctx = Context::thread_default()) {
    const size_t nc = a.n_chunks();
    if (nc == 0) throw KernelError(MNR_ERR_INVALID_ARGUMENTS, "InvalidArguments", "super_array_stats: no chunks");
    const mnr_dtype dt = a.chunks_[0].dtype();
    std::vector<detail::BufH> B(nc);
    std::vector<detail::BitsH> M(nc);
    std::vector<const mnr_buf*> bp(nc);
    std::vector<const mnr_bits*> mp(nc);
    for (size_t i = 0; i < nc; ++i) {
        if (a.chunks_[i].dtype() != dt) throw KernelError(MNR_ERR_TYPE_MISMATCH, "TypeMismatch", "super_array_stats: chunks differ in dtype");
        B[i] = detail::upload_window(ctx, ArrayV(a.chunks_[i]));
        M[i] = detail::upload_mask(ctx, a.chunks_[i].null_mask(), a.chunks_[i].len());
        bp[i] = B[i].h; mp[i] = M[i].h;
    }
    std::vector<mnr_agg> parts(nc);
    check(mnr_reduce_stats_batch(ctx.get(), nc, bp.data(), mp.data(), with_minmax ? 1 : 0, parts.data()));
    mnr_agg out{};
    check(mnr_agg_combine(dt, parts.data(), nc, &out));
    return out;
}

// ---- Table / SuperTable ------------------------------------------------------------------------------------------------------
struct Table {
    std::string name;
    std::vector<Array> cols;
    Table() = default;
    Table(std::string n, std::vector<Array> c) : name(std::move(n)), cols(std::move(c)) {}
    size_t n_cols() const { return cols.size(); }
    size_t n_rows() const { return cols.empty() ? 0 : cols[0].len(); }
    const Array* col_ix(size_t i) const { return i < cols.size() ? &cols[i] : nullptr; }
};

// broadcast_table_with_operator (table.rs:31-62): column i against column i through the router with NO mask.
inline Table broadcast_table_with_operator(ArithmeticOperator op, const Table& l, const Table& r, Context& ctx = Context::thread_default()) {
    if (l.n_cols() != r.n_cols())
        throw KernelError(MNR_ERR_SHAPE, "ShapeError", "Table column count mismatch: " + std::to_string(l.n_cols()) + " vs " + std::to_string(r.n_cols()));
    Table out;
    out.name = l.name;   // schema and name come from the left table
    for (size_t i = 0; i < l.n_cols(); ++i) out.cols.push_back(resolve_binary_arithmetic(op, ArrayV(l.cols[i]), ArrayV(r.cols[i]), nullptr, ctx));
    return out;
}
inline Table broadcast_table_add(const Table& l, const Table& r, Context& ctx = Context::thread_default()) {
    return broadcast_table_with_operator(ArithmeticOperator::Add, l, r, ctx);
}
// broadcast_table_add(lhs, rhs, null_mask) (table.rs:69-128): column i + column i through broadcast_array_add =
// resolve_binary_arithmetic(Add, l, r, null_mask) — the SAME optional mask for every column; the reference's BroadcastingError texts.
inline Table broadcast_table_add(const Table& l, const Table& r, const Bitmask* null_mask, Context& ctx = Context::thread_default()) {
    if (l.n_cols() != r.n_cols())
        throw KernelError(MNR_ERR_SHAPE, "BroadcastingError", "Table column count mismatch: LHS " + std::to_string(l.n_cols()) + " cols, RHS " + std::to_string(r.n_cols()) + " cols");
    if (l.n_rows() != r.n_rows())
        throw KernelError(MNR_ERR_SHAPE, "BroadcastingError", "Table row count mismatch: LHS " + std::to_string(l.n_rows()) + " rows, RHS " + std::to_string(r.n_rows()) + " rows");
    Table out;
    out.name = l.name;
    for (size_t i = 0; i < l.n_cols(); ++i)
        out.cols.push_back(resolve_binary_arithmetic(ArithmeticOperator::Add, ArrayV(l.cols[i]), ArrayV(r.cols[i]), null_mask, ctx));
    return out;
}
// table op array / array op table / table op scalar / scalar op table: operand order is significant.
inline Table broadcast_table_to_array(ArithmeticOperator op, const Table& t, const Array& a, Context& ctx = Context::thread_default()) {
    Table out; out.name = t.name;
    for (const auto& c : t.cols) out.cols.push_back(resolve_binary_arithmetic(op, ArrayV(c), ArrayV(a), nullptr, ctx));
    return out;
}
inline Table broadcast_array_to_table(ArithmeticOperator op, const Array& a, const Table& t, Context& ctx = Context::thread_default()) {
    Table out; out.name = t.name;
    for (const auto& c : t.cols) out.cols.push_back(resolve_binary_arithmetic(op, ArrayV(a), ArrayV(c), nullptr, ctx));
    return out;
}
// Scalar (src/enums/scalar.rs) -> a length-1 Array of its own type; a column of another type is UnsupportedType in the
// router, as in the reference.
template <class S> Table broadcast_table_to_scalar(ArithmeticOperator op, const Table& t, S scalar, Context& ctx = Context::thread_default()) {
    return broadcast_table_to_array(op, t, Array::from_slice<S>({scalar}), ctx);
}
template <class S> Table broadcast_scalar_to_table(ArithmeticOperator op, S scalar, const Table& t, Context& ctx = Context::thread_default()) {
    return broadcast_array_to_table(op, Array::from_slice<S>({scalar}), t, ctx);
}

struct SuperTable {
    std::vector<std::shared_ptr<const Table>> batches;
    std::string name;
    static SuperTable from_batches(std::vector<Table> b, std::string n = "") {
        SuperTable s; s.name = std::move(n);
        for (auto& t : b) s.batches.push_back(std::make_shared<const Table>(std::move(t)));
        return s;
    }
    size_t n_batches() const { return batches.size(); }
    size_t n_rows() const { size_t n = 0; for (const auto& b : batches) n += b->n_rows(); return n; }
    size_t n_cols() const { return batches.empty() ? 0 : batches[0]->n_cols(); }
};

// broadcast_super_table_with_operator (super_table.rs:38-73): batch by batch through the Table route.
inline SuperTable broadcast_super_table_with_operator(ArithmeticOperator op, const SuperTable& l, const SuperTable& r, Context& ctx = Context::thread_default()) {
    if (l.n_batches() != r.n_batches())
        throw KernelError(MNR_ERR_SHAPE, "ShapeError", "SuperTable chunk count mismatch: " + std::to_string(l.n_batches()) + " vs " + std::to_string(r.n_batches()));
    std::vector<Table> out;
    for (size_t i = 0; i < l.n_batches(); ++i) out.push_back(broadcast_table_with_operator(op, *l.batches[i], *r.batches[i], ctx));
    return SuperTable::from_batches(std::move(out));
}
// broadcast_super_table_add (table.rs:135-176): chunk i + chunk i through broadcast_table_add with the same optional mask; the
// result takes the first chunk's name, or "SuperTable".
inline SuperTable broadcast_super_table_add(const SuperTable& l, const SuperTable& r, const Bitmask* null_mask = nullptr, Context& ctx = Context::thread_default()) {
    if (l.n_batches() != r.n_batches())
        throw KernelError(MNR_ERR_SHAPE, "BroadcastingError", "SuperTable chunk count mismatch: LHS " + std::to_string(l.n_batches()) + " chunks, RHS " + std::to_string(r.n_batches()) + " chunks");
    std::vector<Table> out;
    for (size_t i = 0; i < l.n_batches(); ++i) {
        try {
            out.push_back(broadcast_table_add(*l.batches[i], *r.batches[i], null_mask, ctx));
        } catch (const KernelError& e) {
            throw KernelError(MNR_ERR_SHAPE, "BroadcastingError", "Chunk " + std::to_string(i) + " addition failed: " + e.what());
        }
    }
    return SuperTable::from_batches(std::move(out), (!l.batches.empty() && !l.batches[0]->name.empty()) ? l.batches[0]->name : std::string("SuperTable"));
}
template <class S> SuperTable broadcast_super_table_to_scalar(ArithmeticOperator op, const SuperTable& t, S scalar, Context& ctx = Context::thread_default()) {
    std::vector<Table> out;
    for (const auto& b : t.batches) out.push_back(broadcast_table_to_scalar(op, *b, scalar, ctx));
    return SuperTable::from_batches(std::move(out), t.name);
}

// broadcast_array_to_supertable (array.rs:236-252) / broadcast_supertable_to_array (super_table.rs): the array against every
// column of every batch, operand order kept.
inline SuperTable broadcast_array_to_supertable(ArithmeticOperator op, const Array& a, const SuperTable& st, Context& ctx = Context::thread_default()) {
    std::vector<Table> out;
    for (const auto& b : st.batches) out.push_back(broadcast_array_to_table(op, a, *b, ctx));
    return SuperTable::from_batches(std::move(out), st.name);
}
inline SuperTable broadcast_supertable_to_array(ArithmeticOperator op, const SuperTable& st, const Array& a, Context& ctx = Context::thread_default()) {
    std::vector<Table> out;
    for (const auto& b : st.batches) out.push_back(broadcast_table_to_array(op, *b, a, ctx));
    return SuperTable::from_batches(std::move(out), st.name);
}


// ---- device-resident containers: operands stay in HBM between calls --------------------------------------------------------
// The same containers holding mnr_buf / mnr_bits handles.  A route gathers every (chunk, column) leaf call and issues them
// through mnr_ew_binary_batch (one launch per (dtype, alignment, masked) class); results are device-resident again, so
// `table * table + table` never visits the host.  Views are free: DeviceArray::view is the ArrayV window (a pointer offset
// for the values, mnr_bits_slice for the validity), so SuperArrayV = DeviceSuperArray of views, TableV = DeviceTable::view.
struct DeviceArray {
    std::shared_ptr<mnr_buf> buf;
    std::shared_ptr<mnr_bits> mask;   // null = no validity
    static std::shared_ptr<mnr_buf> own(mnr_buf* b) { return std::shared_ptr<mnr_buf>(b, [](mnr_buf* p) { mnr_buf_free(p); }); }
    static std::shared_ptr<mnr_bits> own(mnr_bits* b) { return b ? std::shared_ptr<mnr_bits>(b, [](mnr_bits* p) { mnr_bits_free(p); }) : nullptr; }
    static DeviceArray from_host(Context& ctx, const Array& a) {
        DeviceArray d;
        detail::BufH b = detail::upload_window(ctx, ArrayV(a));
        d.buf = own(b.h); b.h = nullptr;
        detail::BitsH m = detail::upload_mask(ctx, a.null_mask(), a.len());
        d.mask = own(m.h); m.h = nullptr;
        return d;
    }
    Array to_host(Context& ctx) const {
        // detail::download takes ownership of the handles it is given: hand it fresh non-owning views
        mnr_buf* b = nullptr;
        check(mnr_buf_slice(buf.get(), 0, len(), &b));
        mnr_bits* m = nullptr;
        if (mask) check(mnr_bits_wrap(ctx.get(), mnr_bits_device_ptr(mask.get()), len(), &m));
        return detail::download(ctx, dtype(), b, m);
    }
    size_t len() const { return mnr_buf_len(buf.get()); }
    mnr_dtype dtype() const { return static_cast<mnr_dtype>(mnr_buf_dtype(buf.get())); }
    // ArrayV::new(array, offset, len): zero-copy values window (the parent stays alive through the deleter), validity
    // re-based to bit 0 at its exact bit offset (Bitmask::slice_clone on the device).
    DeviceArray view(Context& ctx, size_t offset, size_t n) const {
        if (offset + n > len()) throw KernelError(MNR_ERR_OUT_OF_BOUNDS, "OutOfBounds", "DeviceArray::view exceeds the array");
        DeviceArray v;
        mnr_buf* b = nullptr;
        check(mnr_buf_slice(buf.get(), offset, n, &b));
        std::shared_ptr<mnr_buf> parent = buf;
        v.buf = std::shared_ptr<mnr_buf>(b, [parent](mnr_buf* p) { mnr_buf_free(p); });
        if (mask) { mnr_bits* m = nullptr; check(mnr_bits_slice(ctx.get(), mask.get(), offset, n, &m)); v.mask = own(m); }
        return v;
    }
};
struct DeviceSuperArray {
    std::vector<DeviceArray> chunks;
    static DeviceSuperArray from_host(Context& ctx, const SuperArray& s) { DeviceSuperArray d; for (const auto& c : s.chunks()) d.chunks.push_back(DeviceArray::from_host(ctx, c)); return d; }
    SuperArray to_host(Context& ctx) const { SuperArray s; for (const auto& c : chunks) s.push(c.to_host(ctx)); return s; }
    size_t len() const { size_t n = 0; for (const auto& c : chunks) n += c.len(); return n; }
    size_t n_chunks() const { return chunks.size(); }
};
struct DeviceTable {
    std::string name;
    std::vector<DeviceArray> cols;
    static DeviceTable from_host(Context& ctx, const Table& t) { DeviceTable d; d.name = t.name; for (const auto& c : t.cols) d.cols.push_back(DeviceArray::from_host(ctx, c)); return d; }
    Table to_host(Context& ctx) const { Table t; t.name = name; for (const auto& c : cols) t.cols.push_back(c.to_host(ctx)); return t; }
    size_t n_cols() const { return cols.size(); }
    size_t n_rows() const { return cols.empty() ? 0 : cols[0].len(); }
    DeviceTable view(Context& ctx, size_t offset, size_t n) const { DeviceTable v; v.name = name; for (const auto& c : cols) v.cols.push_back(c.view(ctx, offset, n)); return v; }   // TableV
};
struct DeviceSuperTable {
    std::vector<DeviceTable> batches;
    std::string name;
    static DeviceSuperTable from_host(Context& ctx, const SuperTable& s) { DeviceSuperTable d; d.name = s.name; for (const auto& b : s.batches) d.batches.push_back(DeviceTable::from_host(ctx, *b)); return d; }
    SuperTable to_host(Context& ctx) const { std::vector<Table> b; for (const auto& t : batches) b.push_back(t.to_host(ctx)); return SuperTable::from_batches(std::move(b), name); }
    size_t n_batches() const { return batches.size(); }
    size_t n_cols() const { return batches.empty() ? 0 : batches[0].n_cols(); }
};

namespace detail {
struct Leaf { const DeviceArray* l; const DeviceArray* r; const mnr_bits* lm; const mnr_bits* rm; };
// All leaves of one container operation in ONE mnr_ew_binary_batch call.  Operands of a leaf must agree in dtype and length
// (the length-1 / promotion corners of the router stay on the host layer above: resolve_binary_arithmetic).
inline std::vector<DeviceArray> route_leaves(Context& ctx, ArithmeticOperator op, const std::vector<Leaf>& leaves, mnr_mask_mode mode) {
    const size_t n = leaves.size();
    std::vector<const mnr_buf*> l(n), r(n);
    std::vector<const mnr_bits*> lm(n), rm(n);
    for (size_t i = 0; i < n; ++i) {
        if (leaves[i].l->len() != leaves[i].r->len())
            throw KernelError(MNR_ERR_LENGTH_MISMATCH, "LengthMismatch", "cannot broadcast arrays of length " + std::to_string(leaves[i].l->len()) + " and " + std::to_string(leaves[i].r->len()));
        if (leaves[i].l->dtype() != leaves[i].r->dtype() || !routed_dtype(leaves[i].l->dtype()))
            throw KernelError(MNR_ERR_UNSUPPORTED_TYPE, "UnsupportedType", "Unsupported array type combination for arithmetic operations");
        l[i] = leaves[i].l->buf.get(); r[i] = leaves[i].r->buf.get(); lm[i] = leaves[i].lm; rm[i] = leaves[i].rm;
    }
    std::vector<mnr_buf*> ob(n, nullptr);
    std::vector<mnr_bits*> om(n, nullptr);
    if (n) check(mnr_ew_binary_batch(ctx.get(), static_cast<mnr_op>(op), n, l.data(), r.data(), lm.data(), rm.data(), mode, ob.data(), om.data()));
    std::vector<DeviceArray> out(n);
    for (size_t i = 0; i < n; ++i) { out[i].buf = DeviceArray::own(ob[i]); out[i].mask = DeviceArray::own(om[i]); }
    return out;
}
}  // namespace detail

// route_super_array_broadcast on the device: chunk i against chunk i, validity = union of the chunks' masks or the one present.
inline DeviceSuperArray route_super_array_broadcast(ArithmeticOperator op, const DeviceSuperArray& lhs, const DeviceSuperArray& rhs, Context& ctx = Context::thread_default()) {
    if (rhs.n_chunks() < lhs.n_chunks()) throw KernelError(MNR_ERR_SHAPE, "ShapeError", "Super Array broadcasting error - chunk count");
    std::vector<detail::Leaf> leaves;
    for (size_t i = 0; i < lhs.n_chunks(); ++i) {
        if (lhs.chunks[i].len() != rhs.chunks[i].len())
            throw KernelError(MNR_ERR_SHAPE, "ShapeError", "Super Array broadcasting error - Chunk: LHS " + std::to_string(lhs.chunks[i].len()) + " RHS " + std::to_string(rhs.chunks[i].len()));
        leaves.push_back({&lhs.chunks[i], &rhs.chunks[i], lhs.chunks[i].mask.get(), rhs.chunks[i].mask.get()});
    }
    DeviceSuperArray out;
    out.chunks = detail::route_leaves(ctx, op, leaves, MNR_MASK_OR);
    return out;
}

// union_array_superarray_masks (src/utils.rs:367-413) on the device: the chunk masks concatenated at bit granularity
// (mnr_concat's validity gather; a chunk without a mask counts as all valid once ANY chunk has one) OR-ed with the array's.
inline std::shared_ptr<mnr_bits> union_array_superarray_masks(const DeviceArray& array, const DeviceSuperArray& sa, Context& ctx = Context::thread_default()) {
    std::shared_ptr<mnr_bits> sa_mask;
    bool any = false;
    for (const auto& c : sa.chunks) any = any || c.mask;
    if (any) {
        std::vector<const mnr_buf*> b;
        std::vector<const mnr_bits*> m;
        for (const auto& c : sa.chunks) { b.push_back(c.buf.get()); m.push_back(c.mask.get()); }
        mnr_buf* ob = nullptr; mnr_bits* om = nullptr;
        check(mnr_concat(ctx.get(), b.size(), b.data(), m.data(), &ob, &om));
        mnr_buf_free(ob);
        sa_mask = DeviceArray::own(om);
    }
    if (array.mask && sa_mask) {
        if (mnr_bits_len(array.mask.get()) != mnr_bits_len(sa_mask.get()))
            throw KernelError(MNR_ERR_SHAPE, "ShapeError", "Mask lengths must match for union");
        mnr_bits* u = nullptr;
        check(mnr_bits_merge(ctx.get(), array.mask.get(), sa_mask.get(), array.len(), MNR_MASK_OR, &u));
        return DeviceArray::own(u);
    }
    return array.mask ? array.mask : sa_mask;
}

// create_aligned_chunks_from_array (src/utils.rs:417-481) on the device: `array` re-chunked to the SuperArray's chunk
// lengths; value chunks are zero-copy windows, every chunk carries its window of the FULL union mask.
inline DeviceSuperArray create_aligned_chunks_from_array(const DeviceArray& array, const DeviceSuperArray& sa, Context& ctx = Context::thread_default()) {
    if (array.len() != sa.len())
        throw KernelError(MNR_ERR_SHAPE, "ShapeError", "Array and SuperArray must have same total length for broadcasting: " + std::to_string(array.len()) + " vs " + std::to_string(sa.len()));
    DeviceArray carrier;
    carrier.buf = array.buf;
    carrier.mask = union_array_superarray_masks(array, sa, ctx);
    DeviceSuperArray out;
    size_t start = 0;
    for (const auto& c : sa.chunks) { out.chunks.push_back(carrier.view(ctx, start, c.len())); start += c.len(); }
    return out;
}

// Value::Array (op) Value::SuperArray and the mirrored arm (broadcast/mod.rs:1351-1361): re-chunk, then the SuperArray route.
inline DeviceSuperArray broadcast_array_to_superarray(ArithmeticOperator op, const DeviceArray& array, const DeviceSuperArray& sa, bool array_is_lhs = true,
                                                      Context& ctx = Context::thread_default()) {
    DeviceSuperArray aligned = create_aligned_chunks_from_array(array, sa, ctx);
    return array_is_lhs ? route_super_array_broadcast(op, aligned, sa, ctx) : route_super_array_broadcast(op, sa, aligned, ctx);
}
inline SuperArray create_aligned_chunks_from_array(const Array& array, const SuperArray& sa, Context& ctx = Context::thread_default()) {
    return create_aligned_chunks_from_array(DeviceArray::from_host(ctx, array), DeviceSuperArray::from_host(ctx, sa), ctx).to_host(ctx);
}
inline SuperArray broadcast_array_to_superarray(ArithmeticOperator op, const Array& array, const SuperArray& sa, Context& ctx = Context::thread_default()) {
    return broadcast_array_to_superarray(op, DeviceArray::from_host(ctx, array), DeviceSuperArray::from_host(ctx, sa), true, ctx).to_host(ctx);
}
inline SuperArray broadcast_superarray_to_array(ArithmeticOperator op, const SuperArray& sa, const Array& array, Context& ctx = Context::thread_default()) {
    return broadcast_array_to_superarray(op, DeviceArray::from_host(ctx, array), DeviceSuperArray::from_host(ctx, sa), false, ctx).to_host(ctx);
}

// Table / SuperTable routes on the device: NO mask (table.rs:54), every batch x column of the operation in one batched call.
inline DeviceTable broadcast_table_with_operator(ArithmeticOperator op, const DeviceTable& l, const DeviceTable& r, Context& ctx = Context::thread_default()) {
    if (l.n_cols() != r.n_cols())
        throw KernelError(MNR_ERR_SHAPE, "ShapeError", "Table column count mismatch: " + std::to_string(l.n_cols()) + " vs " + std::to_string(r.n_cols()));
    std::vector<detail::Leaf> leaves;
    for (size_t i = 0; i < l.n_cols(); ++i) leaves.push_back({&l.cols[i], &r.cols[i], nullptr, nullptr});
    DeviceTable out;
    out.name = l.name;
    out.cols = detail::route_leaves(ctx, op, leaves, MNR_MASK_AND);
    return out;
}
inline DeviceSuperTable broadcast_super_table_with_operator(ArithmeticOperator op, const DeviceSuperTable& l, const DeviceSuperTable& r, Context& ctx = Context::thread_default()) {
    if (l.n_batches() != r.n_batches())
        throw KernelError(MNR_ERR_SHAPE, "ShapeError", "SuperTable chunk count mismatch: " + std::to_string(l.n_batches()) + " vs " + std::to_string(r.n_batches()));
    std::vector<detail::Leaf> leaves;
    for (size_t b = 0; b < l.n_batches(); ++b) {
        if (l.batches[b].n_cols() != r.batches[b].n_cols())
            throw KernelError(MNR_ERR_SHAPE, "ShapeError", "Table column count mismatch: " + std::to_string(l.batches[b].n_cols()) + " vs " + std::to_string(r.batches[b].n_cols()));
        for (size_t i = 0; i < l.batches[b].n_cols(); ++i) leaves.push_back({&l.batches[b].cols[i], &r.batches[b].cols[i], nullptr, nullptr});
    }
    std::vector<DeviceArray> res = detail::route_leaves(ctx, op, leaves, MNR_MASK_AND);
    DeviceSuperTable out;
    out.name = l.name;
    size_t k = 0;
    for (const auto& b : l.batches) {
        DeviceTable t;
        t.name = b.name;
        for (size_t i = 0; i < b.n_cols(); ++i) t.cols.push_back(std::move(res[k++]));
        out.batches.push_back(std::move(t));
    }
    return out;
}

// Per-column {sum, min, max, count} over all batches of a device-resident SuperTable (BASELINE configs[4]): every chunk of
// every column in one batched call, folded per column in chunk order on the device.
inline std::vector<mnr_agg> super_table_stats(const DeviceSuperTable& st, bool with_minmax = true, Context& ctx = Context::thread_default()) {
    const size_t nc = st.n_cols();
    std::vector<const mnr_buf*> b;
    std::vector<const mnr_bits*> m;
    std::vector<uint32_t> col;
    std::vector<mnr_dtype> dts(nc);
    for (size_t c = 0; c < nc; ++c)
        for (const auto& t : st.batches) { b.push_back(t.cols[c].buf.get()); m.push_back(t.cols[c].mask.get()); col.push_back((uint32_t)c); dts[c] = t.cols[c].dtype(); }
    std::vector<mnr_agg> out(nc);
    check(mnr_reduce_stats_batch_exchange_sync(ctx.get(), nullptr, b.size(), b.data(), m.data(), with_minmax ? 1 : 0, nc, col.data(), dts.data(), out.data()));
    return out;
}

}  // namespace minarrow_b200
