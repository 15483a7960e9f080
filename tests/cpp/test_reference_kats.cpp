// The reference's own leaf-kernel tests, transcribed into C++ against include/minarrow_b200.hpp (same function names,
// same argument order, same expectations) and run on the GPU through the C ABI.
//
//   src/kernels/arithmetic/mod.rs:117-230  int_kernel_suite!  (i32, u32, i64, u64 + 8/16-bit)
//   src/kernels/arithmetic/mod.rs:293-367  float_kernel_suite! (f32 eps 1e-6, f64 eps 1e-12)
//   src/kernels/arithmetic/mod.rs:372-399  fma_f32 / fma_f64
//   src/kernels/arithmetic/mod.rs:401-409  merge_masks_correctness
//   src/kernels/arithmetic/mod.rs:507-537  simd int power (2^10 = 1024, 16 and 128 elements)
//   src/kernels/bitmask/simd.rs:817-945    and/or/xor/not, popcount [T,F,T,F,T,F,F,T] = 4, all_true / all_false
//   benches/hotloop_benchmark_simd.rs      sum(0..1000) = 499500
//
// Usage: test_reference_kats            run everything on cuda:0 (exit code = number of failed checks)
//        test_reference_kats --link     only prove that the binary links and the library loads (no GPU needed)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <numeric>

#include "minarrow_b200.hpp"

using namespace minarrow_b200;
using Op = ArithmeticOperator;

static int g_failed = 0, g_checks = 0;
#define CHECK(cond)                                                                     \
    do {                                                                                \
        ++g_checks;                                                                     \
        if (!(cond)) { ++g_failed; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)

template <class T> static bool same(const Vec64<T>& got, std::initializer_list<T> exp) {
    return got.size() == exp.size() && std::equal(got.begin(), got.end(), exp.begin());
}
static bool mask_is(const std::optional<Bitmask>& m, std::initializer_list<bool> exp) {
    if (!m || m->len != exp.size()) return false;
    size_t i = 0;
    for (bool b : exp) if (m->get(i++) != b) return false;
    return true;
}
static bool panics(const std::function<void()>& f) {       // std::panic::catch_unwind(..).is_err()
    try { f(); } catch (const KernelError& e) { return e.kind == "DivideByZero"; }
    return false;
}

template <class T, class F> static void int_kernel_suite(F apply) {
    // $fn_dense (mod.rs:120-177)
    Vec64<T> lhs{1, 4, 9, 16}, rhs{1, 2, 3, 4};
    auto out = apply(lhs, rhs, Op::Add, nullptr);
    CHECK(same<T>(out.data, {2, 6, 12, 20}) && !out.null_mask);
    CHECK(same<T>(apply(lhs, rhs, Op::Subtract, nullptr).data, {0, 2, 6, 12}));
    CHECK(same<T>(apply(lhs, rhs, Op::Multiply, nullptr).data, {1, 8, 27, 64}));
    CHECK(same<T>(apply(lhs, rhs, Op::Divide, nullptr).data, {1, 2, 3, 4}));
    CHECK(same<T>(apply(lhs, rhs, Op::Remainder, nullptr).data, {0, 0, 0, 0}));
    CHECK(same<T>(apply(lhs, rhs, Op::Power, nullptr).data, {1, 16, 729, (T)65536}));   // wrapping repeated multiply
    Vec64<T> zeros{0, 0, 0, 0};
    CHECK(panics([&] { apply(lhs, zeros, Op::Divide, nullptr); }));      // "Dense integer kernel division by zero must panic"
    CHECK(panics([&] { apply(lhs, zeros, Op::Remainder, nullptr); }));   // "Dense integer kernel remainder by zero must panic"
    // $fn_masked (mod.rs:179-219)
    Vec64<T> l2{10, 20, 30, 40}, r2{2, 0, 3, 5};
    Bitmask mask = Bitmask::from_bools({true, false, true, false});
    out = apply(l2, r2, Op::Divide, &mask);
    CHECK(same<T>(out.data, {5, 0, 10, 0}) && mask_is(out.null_mask, {true, false, true, false}));
    out = apply(l2, r2, Op::Remainder, &mask);
    CHECK(same<T>(out.data, {0, 0, 0, 0}) && mask_is(out.null_mask, {true, false, true, false}));
    Bitmask all = Bitmask::from_bools({true, true, true, true});
    Vec64<T> l3{100, 100, 100, 100}, r3{1, 0, 2, 0};
    out = apply(l3, r3, Op::Divide, &all);
    CHECK(same<T>(out.data, {100, 0, 50, 0}) && mask_is(out.null_mask, {true, false, true, false}));
    // $fn_empty (mod.rs:221-227)
    Vec64<T> e;
    CHECK(apply(e, e, Op::Add, nullptr).is_empty());
    // length mismatch is an Err, not a panic (confirm_equal_len, utils.rs:163-171)
    bool lm = false;
    try { apply(lhs, e, Op::Add, nullptr); } catch (const KernelError& k) { lm = k.kind == "LengthMismatch"; }
    CHECK(lm);
}

template <class T, class F> static void float_kernel_suite(F apply, T eps) {
    Vec64<T> lhs{1.0, 4.0, 9.0, 16.0}, rhs{0.5, 2.0, 3.0, 4.0};
    CHECK(same<T>(apply(lhs, rhs, Op::Add, nullptr).data, {1.5, 6.0, 12.0, 20.0}));
    CHECK(same<T>(apply(lhs, rhs, Op::Subtract, nullptr).data, {0.5, 2.0, 6.0, 12.0}));
    CHECK(same<T>(apply(lhs, rhs, Op::Multiply, nullptr).data, {0.5, 8.0, 27.0, 64.0}));
    CHECK(same<T>(apply(lhs, rhs, Op::Divide, nullptr).data, {2.0, 2.0, 3.0, 4.0}));
    auto rem = apply(lhs, rhs, Op::Remainder, nullptr);
    for (size_t i = 0; i < 4; ++i) CHECK(std::fabs(rem.data[i] - std::fmod(lhs[i], rhs[i])) < eps);
    auto pw = apply(lhs, rhs, Op::Power, nullptr);
    for (size_t i = 0; i < 4; ++i) CHECK(std::fabs(pw.data[i] - std::exp(rhs[i] * std::log(lhs[i]))) < eps * 64);   // values up to 65536
    Vec64<T> zeros{0.0, 0.0, 0.0, 0.0};
    auto dz = apply(lhs, zeros, Op::Divide, nullptr);        // "Float division by zero should yield Inf"
    for (T x : dz.data) CHECK(std::isinf(x));
    auto rz = apply(lhs, zeros, Op::Remainder, nullptr);     // "Float remainder by zero should yield NaN"
    for (T x : rz.data) CHECK(std::isnan(x));
    Bitmask mask = Bitmask::from_bools({true, false, true, false});
    auto m = apply(lhs, rhs, Op::Multiply, &mask);
    CHECK(same<T>(m.data, {0.5, 0.0, 27.0, 0.0}) && m.null_mask && m.null_mask->len == 4);
    Vec64<T> e;
    CHECK(apply(e, e, Op::Add, nullptr).is_empty());
}

int main(int argc, char** argv) {
    if (argc > 1 && !std::strcmp(argv[1], "--link")) {
        std::printf("abi %d, devices %d\n", mnr_abi_version(), mnr_device_count());
        return mnr_abi_version() == MNR_ABI_VERSION ? 0 : 1;
    }
    try {
        Context& ctx = Context::thread_default();
        int_kernel_suite<int32_t>([](auto&&... a) { return apply_int_i32(a...); });
        int_kernel_suite<uint32_t>([](auto&&... a) { return apply_int_u32(a...); });
        int_kernel_suite<int64_t>([](auto&&... a) { return apply_int_i64(a...); });
        int_kernel_suite<uint64_t>([](auto&&... a) { return apply_int_u64(a...); });
        int_kernel_suite<int16_t>([](auto&&... a) { return apply_int_i16(a...); });
        int_kernel_suite<uint16_t>([](auto&&... a) { return apply_int_u16(a...); });
        float_kernel_suite<float>([](auto&&... a) { return apply_float_f32(a...); }, 1e-6f);
        float_kernel_suite<double>([](auto&&... a) { return apply_float_f64(a...); }, 1e-12);

        {   // fma_f32 / fma_f64 (mod.rs:372-399)
            Vec64<float> l{1.0f, 2.0f, 3.0f}, r{4.0f, 5.0f, 6.0f}, acc{0.5f, 0.5f, 0.5f};
            CHECK(same<float>(apply_fma_f32(l, r, acc).data, {4.5f, 10.5f, 18.5f}));
            Bitmask mask = Bitmask::from_bools({true, false, true});
            auto o = apply_fma_f32(l, r, acc, &mask);
            CHECK(same<float>(o.data, {4.5f, 0.0f, 18.5f}) && mask_is(o.null_mask, {true, false, true}));
            Vec64<float> e;
            CHECK(apply_fma_f32(e, e, e).is_empty());
            Vec64<double> ld{1.0, 2.0, 3.0}, rd{4.0, 5.0, 6.0}, ad{0.5, 0.5, 0.5};
            CHECK(same<double>(apply_fma_f64(ld, rd, ad).data, {4.5, 10.5, 18.5}));
            auto od = apply_fma_f64(ld, rd, ad, &mask);
            CHECK(same<double>(od.data, {4.5, 0.0, 18.5}) && mask_is(od.null_mask, {true, false, true}));
        }
        {   // merge_masks_correctness (mod.rs:401-409)
            Bitmask a = Bitmask::from_bools({true, false, true, true}), b = Bitmask::from_bools({true, true, false, true});
            auto m = merge_bitmasks_to_new(&a, &b, 4);
            CHECK(mask_is(m, {true, false, false, true}));
        }
        {   // simd int power: 2^10 = 1024 for 16- and 128-element inputs (mod.rs:507-537)
            for (size_t n : {size_t(16), size_t(128)}) {
                Vec64<int32_t> base(n, 2), ex(n, 10);
                auto o = apply_int_i32(base, ex, Op::Power);
                bool all = true;
                for (auto v : o.data) all &= v == 1024;
                CHECK(all && o.data.size() == n);
            }
        }
        {   // bitmask suite (bitmask/simd.rs:817-945): and/or/xor/not on 8-bit patterns, popcount, all_true/all_false
            Bitmask a = Bitmask::from_bools({true, false, true, false, true, true, false, false});
            Bitmask b = Bitmask::from_bools({true, true, false, false, true, false, true, false});
            auto bits = [](const Bitmask& m) { return m.bits[0]; };
            CHECK(bits(and_masks({a, 0, 8}, {b, 0, 8})) == (bits(a) & bits(b)));
            CHECK(bits(or_masks({a, 0, 8}, {b, 0, 8})) == (bits(a) | bits(b)));
            CHECK(bits(xor_masks({a, 0, 8}, {b, 0, 8})) == (bits(a) ^ bits(b)));
            CHECK(bits(not_mask({a, 0, 8})) == (uint8_t)~bits(a));
            Bitmask p = Bitmask::from_bools({true, false, true, false, true, false, false, true});
            CHECK(popcount_mask({p, 0, 8}) == 4);
            CHECK(all_true_mask(Bitmask::new_set_all(64 * 8, true)) && !all_false_mask(Bitmask::new_set_all(64 * 8, true)));
            CHECK(all_false_mask(Bitmask::new_set_all(64 * 8, false)) && !all_true_mask(Bitmask::new_set_all(64 * 8, false)));
            Bitmask t = Bitmask::new_set_all(10, true);          // clear_trailing_bits: 10 bits -> last byte 0x03 (bitmask/mod.rs:240-287)
            CHECK(not_mask({Bitmask::new_set_all(10, false), 0, 10}).bits[1] == 0x03 && t.bits[1] == 0x03);
        }
        {   // bench self-checks: sum(0..1000) = 499500 (hotloop_benchmark_simd.rs); device-resident + null-aware aggregates
            Vec64<int64_t> v(1000);
            std::iota(v.begin(), v.end(), 0);
            mnr_agg a = stats(v, nullptr, true);
            CHECK(a.sum.i64 == 499500 && a.count == 1000 && a.min.i64 == 0 && a.max.i64 == 999);
            DeviceBuffer<int64_t> d(ctx, v);
            Bitmask even = Bitmask::new_set_all(1000, false);
            for (size_t i = 0; i < 1000; i += 2) even.bits[i >> 3] |= uint8_t(1u << (i & 7));
            DeviceBitmask dm(ctx, even);
            mnr_agg e = d.stats(&dm);
            CHECK(e.sum.i64 == 249500 && e.count == 500 && e.max.i64 == 998);
            auto [sq, sqm] = d.binary(Op::Multiply, d, &dm, nullptr);
            auto host = sq.download();
            CHECK(host[10] == 100 && host[11] == 0 && sqm && sqm->count_ones() == 500);
        }
        std::printf("%d checks, %d failed, %llu kernel launches\n", g_checks, g_failed, (unsigned long long)ctx.launch_count());
    } catch (const std::exception& e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 99;
    }
    return g_failed;
}
