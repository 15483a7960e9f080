// CPU check of the division-by-invariant-scalar arithmetic the element-wise kernels use (minarrow_b200/csrc/divmagic.h)
// against the machine's own `/`: every edge divisor x every edge dividend, plus 20 M random pairs per width.
// The same header is compiled into the CUDA kernels; tests/test_gpu_* compare those with the oracle.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <vector>

#include "../../minarrow_b200/csrc/divmagic.h"

using namespace mnr;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <typename W> static W rnd_skewed() {   // mixes magnitudes: full range, small, near powers of two
    const uint64_t r = rnd();
    const int sh = (int)(rnd() % (sizeof(W) * 8));
    switch (rnd() % 4) {
        case 0: return (W)r;
        case 1: return (W)(r >> sh);
        case 2: return (W)(((uint64_t)1 << sh) + (int64_t)(rnd() % 5) - 2);
        default: return (W)((int64_t)(rnd() % 2001) - 1000);
    }
}

template <typename W> static std::vector<W> edges() {
    using L = std::numeric_limits<W>;
    std::vector<W> v = {0, 1, 2, 3, 5, 7, 10, 100, 1000, L::max(), (W)(L::max() - 1), (W)(L::max() / 2), (W)(L::max() / 2 + 1), (W)(L::max() / 3)};
    for (unsigned b = 1; b < sizeof(W) * 8 - 1; ++b) { v.push_back((W)((W)1 << b)); v.push_back((W)(((W)1 << b) - 1)); v.push_back((W)(((W)1 << b) + 1)); }
    if (L::is_signed) {
        const size_t n = v.size();
        for (size_t i = 0; i < n; ++i) v.push_back((W)(0 - (uint64_t)v[i]));
        v.push_back(L::min()); v.push_back((W)(L::min() + 1));
    }
    return v;
}

template <typename W> static W ref_div(W n, W d) {
    if (std::numeric_limits<W>::is_signed && n == std::numeric_limits<W>::min() && d == (W)-1) return n;   // wrapping
    return (W)(n / d);
}

template <typename W> static long check(const char* name) {
    constexpr int N = sizeof(W) * 8;
    long bad = 0, done = 0;
    auto one = [&](W n, W d) {
        if (d == 0) return;
        const DivMagic k = std::numeric_limits<W>::is_signed ? div_magic_signed((int64_t)d, N) : div_magic_unsigned((uint64_t)d, N);
        const W q = div_by_magic<W>(n, d, k), e = ref_div<W>(n, d);
        ++done;
        if (q != e && bad++ < 5) printf("  %s: %" PRId64 " / %" PRId64 " = %" PRId64 ", magic gives %" PRId64 "\n", name, (int64_t)n, (int64_t)d, (int64_t)e, (int64_t)q);
    };
    const auto ev = edges<W>();
    for (W d : ev) for (W n : ev) one(n, d);
    for (int i = 0; i < 20000000; ++i) one(rnd_skewed<W>(), rnd_skewed<W>());
    // a fixed divisor against a dense run of dividends (what a column looks like)
    for (W d : {(W)3, (W)7, (W)1000, (W)86400, (W)(std::numeric_limits<W>::max() / 5)})
        for (int64_t n = -70000; n <= 70000; ++n) { one((W)n, d); if (std::numeric_limits<W>::is_signed) one((W)n, (W)(0 - (uint64_t)d)); }
    printf("%s: %ld checks, %ld failed\n", name, done, bad);
    return bad;
}

int main() {
    long bad = 0;
    bad += check<uint32_t>("u32");
    bad += check<int32_t>("i32");
    bad += check<uint64_t>("u64");
    bad += check<int64_t>("i64");
    // 8/16-bit columns are widened to 32 bits on the device: exhaustive over the 16-bit domain for a few divisors
    long narrow = 0, nbad = 0;
    for (int d = -300; d <= 300; ++d) {
        if (!d) continue;
        const DivMagic ks = div_magic_signed(d, 32);
        for (int n = -32768; n <= 32767; ++n, ++narrow) if (div_by_magic<int32_t>(n, d, ks) != n / d) ++nbad;
        if (d > 0) {
            const DivMagic ku = div_magic_unsigned((uint64_t)d, 32);
            for (unsigned n = 0; n <= 65535; ++n, ++narrow) if (div_by_magic<uint32_t>(n, (uint32_t)d, ku) != n / (unsigned)d) ++nbad;
        }
    }
    printf("narrow (widened to 32 bits): %ld checks, %ld failed\n", narrow, nbad);
    bad += nbad;
    return bad ? 1 : 0;
}
