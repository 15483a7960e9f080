// Multi-GPU surface of the C ABI from ONE process (include/minarrow_b200.h "sharding"): mnr_shard_* chunk -> rank map,
// mnr_group_* (one context + mailbox per device, peer access), the batched fused reduction + exchange and the
// shard-local element-wise fan-out, on a SuperTable shaped like BASELINE configs[4] (i32 / i64 / f32 / f64 columns,
// chunks distributed over the ranks; reference routes: src/kernels/broadcast/super_table.rs:38-73, table.rs:31-62,
// super_array.rs:180-249; sum order benches/benchmark_parallel_simd.rs:81-97).  Expectations are computed here on the host
// with plain loops.  With fewer than 2 GPUs the ranks are "virtual": several contexts on device 0, same kernels, same
// mailbox protocol.
//
// Usage: test_shard_group [world]     run (exit code = number of failed checks)
//        test_shard_group --link      only prove that the binary links and the library loads (no GPU needed)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "minarrow_b200.h"

static int g_failed = 0, g_checks = 0;
#define CHECK(cond)                                                                     \
    do {                                                                                \
        ++g_checks;                                                                     \
        if (!(cond)) { ++g_failed; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)
#define OK(call)                                                                                         \
    do {                                                                                                 \
        const int rc__ = (call);                                                                         \
        ++g_checks;                                                                                      \
        if (rc__ != 0) { ++g_failed; std::printf("FAIL %s:%d  %s -> %d (%s)\n", __FILE__, __LINE__, #call, rc__, mnr_last_error()); } \
    } while (0)

static uint64_t g_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { g_state ^= g_state << 13; g_state ^= g_state >> 7; g_state ^= g_state << 17; return g_state; }

template <class T> struct Col {
    std::vector<std::vector<T>> data;          // per chunk
    std::vector<std::vector<uint8_t>> valid;   // per chunk, Arrow bytes
};
static bool bit(const std::vector<uint8_t>& m, size_t i) { return (m[i >> 3] >> (i & 7)) & 1; }

template <class T> static Col<T> make_col(size_t chunks, size_t rows) {
    Col<T> c;
    for (size_t k = 0; k < chunks; ++k) {
        const size_t n = rows + k * 3;   // ragged chunk lengths
        std::vector<T> d(n);
        std::vector<uint8_t> v((n + 7) / 8, 0);
        for (size_t i = 0; i < n; ++i) {
            if (std::is_floating_point<T>::value) d[i] = (T)((double)(int64_t)(rnd() % 2000001) / 1000.0 - 1000.0);
            else d[i] = (T)(int64_t)(rnd() % 200001) - (T)100000;
            if (rnd() % 10 != 0) v[i >> 3] |= (uint8_t)(1u << (i & 7));
        }
        c.data.push_back(std::move(d));
        c.valid.push_back(std::move(v));
    }
    return c;
}

template <class T> static mnr_dtype code();
template <> mnr_dtype code<int32_t>() { return MNR_I32; }
template <> mnr_dtype code<int64_t>() { return MNR_I64; }
template <> mnr_dtype code<float>() { return MNR_F32; }
template <> mnr_dtype code<double>() { return MNR_F64; }

template <class T> static void upload(mnr_group* g, const Col<T>& c, std::vector<mnr_buf*>& bufs, std::vector<mnr_bits*>& vals) {
    const size_t n = c.data.size();
    std::vector<const void*> hp(n);
    std::vector<size_t> lens(n);
    std::vector<const uint8_t*> vp(n);
    for (size_t k = 0; k < n; ++k) { hp[k] = c.data[k].data(); lens[k] = c.data[k].size(); vp[k] = c.valid[k].data(); }
    bufs.assign(n, nullptr); vals.assign(n, nullptr);
    OK(mnr_group_upload(g, code<T>(), n, hp.data(), lens.data(), vp.data(), bufs.data(), vals.data()));
}

// host expectation of a column's aggregate (integers wrap in 64 bits; floats accumulate in double)
template <class T> static void expect(const Col<T>& c, const mnr_agg& a, const char* name) {
    uint64_t cnt = 0;
    if constexpr (std::is_floating_point<T>::value) {
        long double s = 0, sabs = 0;
        double mn = NAN, mx = NAN;
        for (size_t k = 0; k < c.data.size(); ++k)
            for (size_t i = 0; i < c.data[k].size(); ++i)
                if (bit(c.valid[k], i)) {
                    const double x = (double)c.data[k][i];
                    s += x; sabs += std::fabs(x); ++cnt;
                    mn = (mn != mn || x < mn) ? x : mn;
                    mx = (mx != mx || x > mx) ? x : mx;
                }
        CHECK(a.count == cnt);
        CHECK(std::fabs((long double)a.sum.f64 - s) <= 1e-12L * sabs);
        CHECK(a.min.f64 == mn && a.max.f64 == mx);
    } else {
        uint64_t s = 0;
        int64_t mn = INT64_MAX, mx = INT64_MIN;
        for (size_t k = 0; k < c.data.size(); ++k)
            for (size_t i = 0; i < c.data[k].size(); ++i)
                if (bit(c.valid[k], i)) {
                    const int64_t x = (int64_t)c.data[k][i];
                    s += (uint64_t)x; ++cnt;
                    if (x < mn) mn = x;
                    if (x > mx) mx = x;
                }
        CHECK(a.count == cnt);
        CHECK(a.sum.u64 == s);
        CHECK(a.min.i64 == mn && a.max.i64 == mx);
    }
    std::printf("  column %-4s count %llu ok\n", name, (unsigned long long)cnt);
}

static mnr_group* g_group = nullptr;

template <class T> static void check_mul(const Col<T>& l, const Col<T>& r, size_t k, mnr_buf* ob, mnr_bits* om) {
    const size_t n = l.data[k].size();
    std::vector<T> got(n);
    std::vector<uint8_t> gm((n + 7) / 8);
    CHECK(ob && om && mnr_buf_len(ob) == n && mnr_bits_len(om) == n);
    if (!ob || !om) return;
    // a chunk lives on its owner's context; download through that context
    const int owner = mnr_shard_owner(k, l.data.size(), mnr_group_world(g_group));
    OK(mnr_buf_download(mnr_group_ctx(g_group, owner), ob, got.data()));
    OK(mnr_bits_download(mnr_group_ctx(g_group, owner), om, gm.data()));
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
        const bool v = bit(l.valid[k], i) || bit(r.valid[k], i);   // SuperArray route: OR-union (super_array.rs:214-230)
        T e;
        if constexpr (std::is_floating_point<T>::value) e = v ? l.data[k][i] * r.data[k][i] : (T)0;
        else e = v ? (T)((uint64_t)l.data[k][i] * (uint64_t)r.data[k][i]) : (T)0;
        if (std::memcmp(&e, &got[i], sizeof(T)) != 0 || bit(gm, i) != v) ++bad;
    }
    CHECK(bad == 0);
}
int main(int argc, char** argv) {
    if (argc > 1 && !std::strcmp(argv[1], "--link")) {
        std::printf("abi %d, %d CUDA device(s)\n", mnr_abi_version(), mnr_device_count());
        return 0;
    }
    // ---- chunk -> rank map ------------------------------------------------------------------------------------------------
    for (size_t n : {1u, 2u, 7u, 8u, 64u, 65u})
        for (int w : {1, 2, 4, 8}) {
            size_t covered = 0;
            for (int r = 0; r < w; ++r) {
                size_t lo = 0, hi = 0;
                OK(mnr_shard_chunk_range(n, w, r, &lo, &hi));
                CHECK(lo == covered || lo == hi);
                for (size_t i = lo; i < hi; ++i) CHECK(mnr_shard_owner(i, n, w) == r);
                if (hi > lo) covered = hi;
            }
            CHECK(covered == n);
        }
    {
        size_t pos = 0;
        for (int r = 0; r < 8; ++r) {
            size_t off = 0, len = 0;
            OK(mnr_shard_row_range(1000000007ull, 8, r, 64, &off, &len));
            CHECK(off == pos && off % 64 == 0);
            pos += len;
        }
        CHECK(pos == 1000000007ull);
    }
    CHECK(mnr_shard_owner(5, 5, 2) == MNR_ERR_OUT_OF_BOUNDS);

    // ---- the group ------------------------------------------------------------------------------------------------------------
    const int ndev = mnr_device_count();
    if (ndev < 1) { std::printf("no CUDA device: minarrow_b200 has no CPU fallback\n"); return 1; }
    int world = argc > 1 ? std::atoi(argv[1]) : (ndev >= 2 ? ndev : 3);
    if (world > 16) world = 16;
    std::vector<int> devices(world);
    for (int r = 0; r < world; ++r) devices[r] = r % ndev;
    mnr_group* g = nullptr;
    OK(mnr_group_create(world, devices.data(), &g));
    if (!g) return 1;
    g_group = g;
    CHECK(mnr_group_world(g) == world);
    std::printf("group: %d rank(s) on %d device(s)%s\n", world, ndev, ndev < world ? " (virtual ranks)" : "");

    const size_t chunks = 11, rows = 50021;
    Col<int32_t> a0 = make_col<int32_t>(chunks, rows), b0 = make_col<int32_t>(chunks, rows);
    Col<int64_t> a1 = make_col<int64_t>(chunks, rows), b1 = make_col<int64_t>(chunks, rows);
    Col<float> a2 = make_col<float>(chunks, rows), b2 = make_col<float>(chunks, rows);
    Col<double> a3 = make_col<double>(chunks, rows), b3 = make_col<double>(chunks, rows);
    std::vector<mnr_buf*> A[4], B[4];
    std::vector<mnr_bits*> AV[4], BV[4];
    upload(g, a0, A[0], AV[0]); upload(g, a1, A[1], AV[1]); upload(g, a2, A[2], AV[2]); upload(g, a3, A[3], AV[3]);
    upload(g, b0, B[0], BV[0]); upload(g, b1, B[1], BV[1]); upload(g, b2, B[2], BV[2]); upload(g, b3, B[3], BV[3]);

    // per-column sum / min / max / count of the whole SuperTable: ONE call, batched kernels + fused exchange
    std::vector<const mnr_buf*> bufs;
    std::vector<const mnr_bits*> vals;
    std::vector<uint32_t> col;
    for (uint32_t c = 0; c < 4; ++c)
        for (size_t k = 0; k < chunks; ++k) { bufs.push_back(A[c][k]); vals.push_back(AV[c][k]); col.push_back(c); }
    const mnr_dtype dts[4] = {MNR_I32, MNR_I64, MNR_F32, MNR_F64};
    mnr_agg agg[4];
    for (int rep = 0; rep < 3; ++rep) {   // epochs alternate mailbox parity
        std::memset(agg, 0, sizeof agg);
        OK(mnr_group_reduce_stats(g, bufs.size(), bufs.data(), vals.data(), 1, 4, col.data(), dts, agg));
    }
    expect(a0, agg[0], "i32"); expect(a1, agg[1], "i64"); expect(a2, agg[2], "f32"); expect(a3, agg[3], "f64");

    // a column whose chunks all sit on rank 0 (fewer chunks than ranks): the other ranks join with identity aggregates
    {
        std::vector<const mnr_buf*> one{A[1][0]};
        std::vector<const mnr_bits*> onev{AV[1][0]};
        const uint32_t c0 = 0;
        const mnr_dtype d0 = MNR_I64;
        mnr_agg r{};
        OK(mnr_group_reduce_stats(g, 1, one.data(), onev.data(), 1, 1, &c0, &d0, &r));
        Col<int64_t> sub; sub.data.push_back(a1.data[0]); sub.valid.push_back(a1.valid[0]);
        expect(sub, r, "i64 (1 chunk)");
    }

    // table * table, shard-local, OR-union validity (route_super_array_broadcast semantics per chunk)
    for (int c = 0; c < 4; ++c) {
        std::vector<mnr_buf*> ob(chunks, nullptr);
        std::vector<mnr_bits*> om(chunks, nullptr);
        std::vector<const mnr_buf*> l(A[c].begin(), A[c].end()), r(B[c].begin(), B[c].end());
        std::vector<const mnr_bits*> lm(AV[c].begin(), AV[c].end()), rm(BV[c].begin(), BV[c].end());
        OK(mnr_group_ew_binary(g, MNR_MUL, chunks, l.data(), r.data(), lm.data(), rm.data(), MNR_MASK_OR, ob.data(), om.data()));
        OK(mnr_group_synchronize(g));
        for (size_t k = 0; k < chunks; k += 5) {
            if (c == 0) check_mul(a0, b0, k, ob[k], om[k]);
            if (c == 1) check_mul(a1, b1, k, ob[k], om[k]);
            if (c == 2) check_mul(a2, b2, k, ob[k], om[k]);
            if (c == 3) check_mul(a3, b3, k, ob[k], om[k]);
        }
        for (size_t k = 0; k < chunks; ++k) { mnr_buf_free(ob[k]); mnr_bits_free(om[k]); }
    }
    // typed scalar broadcast per column: col_i64 + 3
    {
        std::vector<mnr_buf*> ob(chunks, nullptr);
        std::vector<mnr_bits*> om(chunks, nullptr);
        std::vector<const mnr_buf*> l(A[1].begin(), A[1].end());
        std::vector<const mnr_bits*> lm(AV[1].begin(), AV[1].end());
        const int64_t three = 3;
        std::vector<const void*> sc(chunks, &three);
        OK(mnr_group_ew_scalar(g, MNR_ADD, chunks, l.data(), sc.data(), 0, lm.data(), ob.data(), om.data()));
        const size_t k = chunks - 1, n = a1.data[k].size();
        std::vector<int64_t> got(n);
        OK(mnr_buf_download(mnr_group_ctx(g, mnr_shard_owner(k, chunks, world)), ob[k], got.data()));
        size_t bad = 0;
        for (size_t i = 0; i < n; ++i) bad += got[i] != (bit(a1.valid[k], i) ? a1.data[k][i] + 3 : 0);
        CHECK(bad == 0);
        for (size_t q = 0; q < chunks; ++q) { mnr_buf_free(ob[q]); mnr_bits_free(om[q]); }
    }
    // operands on different ranks are rejected (element-wise work is shard-local)
    if (world > 1) {
        const mnr_buf* l = A[0][0];
        const mnr_buf* r = B[0][chunks - 1];
        mnr_buf* ob = nullptr; mnr_bits* om = nullptr;
        CHECK(mnr_group_ew_binary(g, MNR_ADD, 1, &l, &r, nullptr, nullptr, MNR_MASK_AND, &ob, &om) == MNR_ERR_INVALID_ARGUMENTS);
    }
    uint64_t launches = 0;
    for (int r = 0; r < world; ++r) launches += mnr_ctx_launch_count(mnr_group_ctx(g, r));
    for (int c = 0; c < 4; ++c)
        for (size_t k = 0; k < chunks; ++k) { mnr_buf_free(A[c][k]); mnr_bits_free(AV[c][k]); mnr_buf_free(B[c][k]); mnr_bits_free(BV[c][k]); }
    mnr_group_destroy(g);
    std::printf("%d checks, %d failed, %llu kernel launches over %d rank(s)\n", g_checks, g_failed, (unsigned long long)launches, world);
    return g_failed;
}
