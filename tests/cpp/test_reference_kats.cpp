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
#include <algorithm>
#include <climits>
#include <utility>
#include <vector>

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

// ---- paths beyond the reference's own vectors, sized to reach the vector / shifted / packed bodies -------------------
// (run under compute-sanitizer by tools/gpu_sanitize.sh: memcheck, racecheck, initcheck, synccheck).  Expectations are
// computed here on the host, bit by bit / row by row.
static uint64_t lcg_state = 12345;
static uint32_t lcg() { lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(lcg_state >> 33); }
static bool host_bit(const std::vector<uint8_t>& m, size_t i) { return (m[i >> 3] >> (i & 7)) & 1; }
#define OK(call) CHECK((call) == MNR_OK)

static void device_path_suite(Context& ctx) {
    mnr_ctx* c = ctx.get();
    const size_t NB = 100003;   // bits
    std::vector<uint8_t> ha((NB + 7) / 8), hb((NB + 7) / 8);
    for (auto& x : ha) x = (uint8_t)lcg();
    for (auto& x : hb) x = (uint8_t)lcg();
    ha.back() &= (uint8_t)((1u << (NB & 7)) - 1); hb.back() &= (uint8_t)((1u << (NB & 7)) - 1);
    mnr_bits *A = nullptr, *B = nullptr;
    OK(mnr_bits_upload(c, ha.data(), NB, &A));
    OK(mnr_bits_upload(c, hb.data(), NB, &B));
    {   // and_masks over windows at byte offsets (BitmaskVT offsets are floored to bytes, bitmask/mod.rs:124-128): shifted body
        for (auto [lo, ro] : {std::pair<size_t, size_t>{8, 16}, {24, 0}, {136, 40}, {5, 77}}) {
            const size_t len = NB - 200;
            mnr_bits* R = nullptr;
            OK(mnr_bits_binop(c, MNR_AND, A, lo, B, ro, len, &R));
            std::vector<uint8_t> got((len + 7) / 8);
            OK(mnr_bits_download(c, R, got.data()));
            bool same_bits = true;
            const size_t lb = lo / 8 * 8, rb = ro / 8 * 8;
            for (size_t i = 0; i < len; ++i) same_bits &= host_bit(got, i) == (host_bit(ha, lb + i) && host_bit(hb, rb + i));
            for (size_t i = len; i < got.size() * 8; ++i) same_bits &= !host_bit(got, i);   // slack bits zero
            CHECK(same_bits);
            mnr_bits_free(R);
        }
    }
    {   // Bitmask::slice_clone at arbitrary bit offsets + popcount of offset windows
        for (size_t off : {size_t(1), size_t(13), size_t(64), size_t(127), size_t(1001)}) {
            const size_t len = NB - 2000;
            mnr_bits* S = nullptr;
            OK(mnr_bits_slice(c, A, off, len, &S));
            std::vector<uint8_t> got((len + 7) / 8);
            OK(mnr_bits_download(c, S, got.data()));
            bool same_bits = true;
            size_t ones = 0;
            for (size_t i = 0; i < len; ++i) { same_bits &= host_bit(got, i) == host_bit(ha, off + i); ones += host_bit(ha, off + i); }
            CHECK(same_bits);
            uint64_t pc = 0;
            OK(mnr_bits_popcount(c, S, 0, len, &pc));
            CHECK(pc == ones);
            const size_t w = off / 64 * 64;   // popcount_mask floors the offset to a 64-bit word (simd.rs:596-645)
            size_t ones_w = 0;
            for (size_t i = 0; i < len; ++i) ones_w += host_bit(ha, w + i);
            OK(mnr_bits_popcount(c, A, off, len, &pc));
            CHECK(pc == ones_w);
            mnr_bits_free(S);
        }
    }
    {   // consolidate: ragged chunks (odd element and bit offsets), 1- and 4-byte elements, one chunk without validity
        const size_t lens[] = {4099, 0, 333, 70001, 7, 12345};
        std::vector<int8_t> h8; std::vector<int32_t> h32; std::vector<uint8_t> hv;
        std::vector<mnr_buf*> b8, b32; std::vector<mnr_bits*> vs;
        size_t k = 0;
        for (size_t n : lens) {
            std::vector<int8_t> d8(n); std::vector<int32_t> d32(n); std::vector<uint8_t> v((n + 7) / 8, 0);
            for (size_t i = 0; i < n; ++i) { d8[i] = (int8_t)lcg(); d32[i] = (int32_t)lcg(); const bool ok = k == 2 || (lcg() % 5) != 0; if (ok) v[i >> 3] |= uint8_t(1u << (i & 7)); hv.push_back(ok); }
            h8.insert(h8.end(), d8.begin(), d8.end()); h32.insert(h32.end(), d32.begin(), d32.end());
            mnr_buf *x8 = nullptr, *x32 = nullptr; mnr_bits* m = nullptr;
            OK(mnr_buf_upload(c, MNR_I8, d8.data(), n, &x8)); OK(mnr_buf_upload(c, MNR_I32, d32.data(), n, &x32));
            if (k != 2) OK(mnr_bits_upload(c, v.data(), n, &m));
            b8.push_back(x8); b32.push_back(x32); vs.push_back(m);
            ++k;
        }
        for (int pass = 0; pass < 2; ++pass) {
            mnr_buf* out = nullptr; mnr_bits* om = nullptr;
            OK(mnr_concat(c, b8.size(), pass ? b32.data() : b8.data(), vs.data(), &out, &om));
            const size_t total = h8.size();
            CHECK(out && mnr_buf_len(out) == total && om && mnr_bits_len(om) == total);
            if (pass) { std::vector<int32_t> g(total); OK(mnr_buf_download(c, out, g.data())); CHECK(g == h32); }
            else { std::vector<int8_t> g(total); OK(mnr_buf_download(c, out, g.data())); CHECK(g == h8); }
            std::vector<uint8_t> gm((total + 7) / 8);
            OK(mnr_bits_download(c, om, gm.data()));
            bool same_bits = true;
            for (size_t i = 0; i < total; ++i) same_bits &= host_bit(gm, i) == (hv[i] != 0);
            CHECK(same_bits);
            mnr_buf_free(out); mnr_bits_free(om);
        }
        for (auto* x : b8) mnr_buf_free(x);
        for (auto* x : b32) mnr_buf_free(x);
        for (auto* x : vs) mnr_bits_free(x);
    }
    {   // column / scalar through the multiplicative inverse, 8-bit packed add, 8-bit masked sum/min/max
        const size_t n = 70001;
        std::vector<int64_t> h(n); std::vector<int8_t> p(n), q(n); std::vector<uint8_t> v((n + 7) / 8, 0);
        for (size_t i = 0; i < n; ++i) {
            h[i] = (int64_t)(((uint64_t)lcg() << 32) | lcg()); p[i] = (int8_t)lcg(); q[i] = (int8_t)lcg();
            if (lcg() % 4) v[i >> 3] |= uint8_t(1u << (i & 7));
        }
        h[0] = INT64_MIN; h[1] = INT64_MAX; h[2] = -1; h[3] = 0;
        mnr_buf *H = nullptr, *P = nullptr, *Q = nullptr; mnr_bits* V = nullptr;
        OK(mnr_buf_upload(c, MNR_I64, h.data(), n, &H)); OK(mnr_buf_upload(c, MNR_I8, p.data(), n, &P));
        OK(mnr_buf_upload(c, MNR_I8, q.data(), n, &Q)); OK(mnr_bits_upload(c, v.data(), n, &V));
        for (int64_t d : {int64_t(7), int64_t(-86400), int64_t(1) << 40, int64_t(-1)}) {
            mnr_buf* o = nullptr; mnr_bits* om = nullptr;
            OK(mnr_ew_scalar(c, MNR_FLOORDIV, H, &d, 0, V, &o, &om));
            std::vector<int64_t> g(n);
            OK(mnr_buf_download(c, o, g.data()));
            bool same_vals = true;
            for (size_t i = 0; i < n; ++i) {
                int64_t e = 0;
                if (host_bit(v, i)) {
                    if (h[i] == INT64_MIN && d == -1) e = INT64_MIN;
                    else { const int64_t qq = h[i] / d, m = h[i] % d; e = (m != 0 && ((h[i] ^ d) < 0)) ? qq - 1 : qq; }
                }
                same_vals &= g[i] == e;
            }
            CHECK(same_vals);
            mnr_buf_free(o); mnr_bits_free(om);
        }
        {
            mnr_buf* o = nullptr; mnr_bits* om = nullptr;
            OK(mnr_ew_binary(c, MNR_ADD, P, Q, V, nullptr, MNR_MASK_AND, &o, &om));
            std::vector<int8_t> g(n);
            OK(mnr_buf_download(c, o, g.data()));
            bool same_vals = true;
            for (size_t i = 0; i < n; ++i) same_vals &= g[i] == (host_bit(v, i) ? (int8_t)(uint8_t)((uint8_t)p[i] + (uint8_t)q[i]) : 0);
            CHECK(same_vals);
            mnr_buf_free(o); mnr_bits_free(om);
            mnr_agg a;
            OK(mnr_reduce_stats(c, P, V, &a));
            int64_t sum = 0, mn = INT8_MAX, mx = INT8_MIN; uint64_t cnt = 0;
            for (size_t i = 0; i < n; ++i) if (host_bit(v, i)) { sum += p[i]; mn = std::min<int64_t>(mn, p[i]); mx = std::max<int64_t>(mx, p[i]); ++cnt; }
            CHECK(a.sum.i64 == sum && a.min.i64 == mn && a.max.i64 == mx && a.count == cnt);
        }
        mnr_buf_free(H); mnr_buf_free(P); mnr_buf_free(Q); mnr_bits_free(V);
    }
    {   // 8/16-bit division through the f32 pipe and float % without the libm loop, against the machine divide / libm fmod:
        // the whole 8-bit domain (zero divisors nulled), 16-bit pairs around every multiple of a few divisors, f64 pairs
        std::vector<int8_t> a8, b8; std::vector<uint16_t> a16, b16; std::vector<int16_t> s16a, s16b;
        for (int l = -128; l < 128; ++l) for (int r = -128; r < 128; ++r) { a8.push_back((int8_t)l); b8.push_back((int8_t)r); }
        for (uint32_t d : {1u, 2u, 3u, 7u, 255u, 256u, 257u, 1000u, 32767u, 32768u, 65535u})
            for (uint32_t k = 0; k * d <= 65535u; k += (65535u / d > 4096 ? 97 : 1))
                for (int e = -1; e <= 1; ++e) {
                    const int64_t l = (int64_t)k * d + e;
                    if (l < 0 || l > 65535) continue;
                    a16.push_back((uint16_t)l); b16.push_back((uint16_t)d);
                    if (l <= 32768 && d <= 32768)
                        for (int sg = 0; sg < 4; ++sg) {
                            const int64_t sl = (sg & 1) ? -l : l, sd = (sg & 2) ? -(int64_t)d : (int64_t)d;
                            if (sl > 32767 || sd > 32767) continue;
                            s16a.push_back((int16_t)sl); s16b.push_back((int16_t)sd);
                        }
                }
        auto run_int = [&](auto tag, mnr_dtype dt, const auto& l, const auto& r) {
            using T = decltype(tag);
            const size_t n = l.size();
            std::vector<uint8_t> ones((n + 7) / 8, 0xFF);
            mnr_buf *L = nullptr, *R = nullptr; mnr_bits* V = nullptr;
            OK(mnr_buf_upload(c, dt, l.data(), n, &L)); OK(mnr_buf_upload(c, dt, r.data(), n, &R)); OK(mnr_bits_upload(c, ones.data(), n, &V));
            for (mnr_op op : {MNR_DIV, MNR_REM, MNR_FLOORDIV}) {
                mnr_buf* o = nullptr; mnr_bits* om = nullptr;
                OK(mnr_ew_binary(c, op, L, R, V, nullptr, MNR_MASK_AND, &o, &om));
                std::vector<T> g(n); std::vector<uint8_t> gm((n + 7) / 8);
                OK(mnr_buf_download(c, o, g.data())); OK(mnr_bits_download(c, om, gm.data()));
                bool same_vals = true;
                for (size_t i = 0; i < n; ++i) {
                    const int x = l[i], y = r[i];
                    if (y == 0) { same_vals &= g[i] == 0 && !host_bit(gm, i); continue; }
                    const int qq = x / y, m = x % y;
                    const int e = op == MNR_DIV ? qq : op == MNR_REM ? m : ((m != 0 && ((x ^ y) < 0)) ? qq - 1 : qq);
                    same_vals &= g[i] == (T)e && host_bit(gm, i);
                }
                CHECK(same_vals);
                mnr_buf_free(o); mnr_bits_free(om);
            }
            mnr_buf_free(L); mnr_buf_free(R); mnr_bits_free(V);
        };
        run_int(int8_t{}, MNR_I8, a8, b8);
        run_int(uint16_t{}, MNR_U16, a16, b16);
        run_int(int16_t{}, MNR_I16, s16a, s16b);
        const size_t n = 50021;
        std::vector<double> x(n), y(n);
        for (size_t i = 0; i < n; ++i) {
            x[i] = std::ldexp((double)(int32_t)lcg() / 65536.0, (int)(lcg() % 80) - 40);
            y[i] = std::ldexp((double)(int32_t)lcg() / 65536.0 + 0.5, (int)(lcg() % 80) - 40);
            if (i % 11 == 0) x[i] = y[i] * (double)(lcg() % 100000);             // exact multiples
            if (i % 13 == 0) x[i] = std::nextafter(y[i] * (double)(lcg() % 1000), 0.0);
        }
        x[0] = 0.0; x[1] = -0.0; x[2] = INFINITY; y[3] = 0.0; y[4] = INFINITY; x[5] = NAN; x[6] = 5e-324; y[7] = 5e-324;
        auto f = apply_float_f64(x, y, Op::Remainder, nullptr);
        bool same_vals = true;
        for (size_t i = 0; i < n; ++i) {
            const double e = std::fmod(x[i], y[i]);
            same_vals &= std::isnan(e) ? std::isnan(f.data[i]) : (std::memcmp(&e, &f.data[i], 8) == 0);
        }
        CHECK(same_vals);
    }
    mnr_bits_free(A); mnr_bits_free(B);
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
        {   // datetime delegation (arithmetic/mod.rs:418-506: datetime_add, datetime_all_ops, datetime_masked_and_empty,
            // datetime_len_mismatch_panics) — the integer kernels with the two arrays' masks merged
            using DA = DatetimeArray<int64_t>;
            DA l = DA::from_slice({1000, 2000, 3000}), r = DA::from_slice({10, 20, 30});
            auto out = apply_datetime_i64({l, 0, l.len()}, {r, 0, r.len()}, Op::Add);
            CHECK(same<int64_t>(out.data, {1010, 2020, 3030}) && !out.null_mask);
            DA a = DA::from_slice({10, 20, 30, 40}), b = DA::from_slice({1, 2, 3, 4});
            auto run = [&](Op op) { return apply_datetime_i64({a, 0, 4}, {b, 0, 4}, op); };
            CHECK(same<int64_t>(run(Op::Add).data, {11, 22, 33, 44}));
            CHECK(same<int64_t>(run(Op::Subtract).data, {9, 18, 27, 36}));
            CHECK(same<int64_t>(run(Op::Multiply).data, {10, 40, 90, 160}));
            CHECK(same<int64_t>(run(Op::Divide).data, {10, 10, 10, 10}));
            CHECK(same<int64_t>(run(Op::Remainder).data, {0, 0, 0, 0}));
            CHECK(same<int64_t>(run(Op::Power).data, {10, 400, 27000, 2560000}));
            DA am = a;
            am.null_mask = Bitmask::from_bools({true, false, true, true});
            auto mo = apply_datetime_i64({am, 0, 4}, {b, 0, 4}, Op::Add);
            CHECK(same<int64_t>(mo.data, {11, 0, 33, 44}) && mask_is(mo.null_mask, {true, false, true, true}));
            DA bm = b;
            bm.null_mask = Bitmask::from_bools({true, true, false, true});
            auto both = apply_datetime_i64({am, 0, 4}, {bm, 0, 4}, Op::Add);               // both masks: per-row AND
            CHECK(same<int64_t>(both.data, {11, 0, 0, 44}) && mask_is(both.null_mask, {true, false, false, true}));
            auto win = apply_datetime_i64({am, 1, 2}, {b, 2, 2}, Op::Add);                 // windows: data offset, masks from bit 0 (dispatch.rs:321-322)
            CHECK(same<int64_t>(win.data, {23, 0}) && mask_is(win.null_mask, {true, false}));
            DA e = DA::from_slice({});
            CHECK(apply_datetime_i64({e, 0, 0}, {e, 0, 0}, Op::Add).is_empty());
            DA s2 = DA::from_slice({1000, 2000}), s1 = DA::from_slice({10});
            bool lm = false;
            try { apply_datetime_i64({s2, 0, 2}, {s1, 0, 1}, Op::Add); } catch (const KernelError& ex) { lm = ex.kind == "LengthMismatch"; }
            CHECK(lm);
            DA z = DA::from_slice({1, 0, 3, 4});
            CHECK(panics([&] { apply_datetime_i64({a, 0, 4}, {z, 0, 4}, Op::Divide); }));   // dense integer kernel: division by zero
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
        device_path_suite(ctx);
        std::printf("%d checks, %d failed, %llu kernel launches\n", g_checks, g_failed, (unsigned long long)ctx.launch_count());
    } catch (const std::exception& e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 99;
    }
    return g_failed;
}
