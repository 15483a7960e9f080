// CPU check of the float-remainder fast path the element-wise kernels use (minarrow_b200/csrc/fastmod.h) against libm's
// fmod / fmodf, bit for bit: edge x edge (zeros, subnormals, powers of two, Inf, NaN, extremes), exact multiples and
// their neighbours (the case where the rounded quotient lands one too high), and random pairs at every exponent
// distance.  IEEE division and fma are the same operations on the host and on the GPU; the device-only pieces (the
// add-round-toward-zero truncation) are covered by the GPU parity tests against the oracle.
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#include "../../minarrow_b200/csrc/fastmod.h"

static uint64_t rng_state = 0x1234567887654321ull;
static uint64_t rnd() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <typename F> struct Bits;
template <> struct Bits<double> { using U = uint64_t; static constexpr int MANT = 52, EXPBITS = 11; };
template <> struct Bits<float> { using U = uint32_t; static constexpr int MANT = 23, EXPBITS = 8; };

template <typename F> static F from_bits(typename Bits<F>::U u) { F f; memcpy(&f, &u, sizeof f); return f; }
template <typename F> static typename Bits<F>::U to_bits(F f) { typename Bits<F>::U u; memcpy(&u, &f, sizeof f); return u; }

template <typename F> static bool same(F x, F y) {
    if (std::isnan(x) || std::isnan(y)) return std::isnan(x) && std::isnan(y);
    return to_bits(x) == to_bits(y);
}

template <typename F> static F rnd_float(int exp_lo, int exp_hi) {   // biased exponents, random mantissa and sign
    using U = typename Bits<F>::U;
    const U e = (U)(exp_lo + (int)(rnd() % (uint64_t)(exp_hi - exp_lo + 1)));
    const U m = (U)rnd() & (((U)1 << Bits<F>::MANT) - 1);
    const U s = (U)(rnd() & 1) << (sizeof(F) * 8 - 1);
    return from_bits<F>(s | (e << Bits<F>::MANT) | m);
}

template <typename F> static int run(const char* name) {
    using L = std::numeric_limits<F>;
    long checked = 0, failed = 0;
    auto check = [&](F a, F b) {
        const F got = mnr::fast_fmod<F>(a, b), exp = mnr::fmod_lib(a, b);
        ++checked;
        if (!same(got, exp)) {
            if (++failed <= 10) printf("  MISMATCH %s: fmod(%a, %a) = %a, expected %a\n", name, (double)a, (double)b, (double)got, (double)exp);
        }
    };
    std::vector<F> e = {(F)0, (F)1, (F)2, (F)3, (F)0.5, (F)0.1, (F)1.5, (F)10, (F)1e10, (F)1e-10, L::min(), L::denorm_min(), (F)(L::denorm_min() * 3),
                        (F)(L::min() / 2), L::max(), (F)(L::max() / 2), L::epsilon(), (F)(1 + L::epsilon()), (F)(1 - L::epsilon() / 2),
                        L::infinity(), L::quiet_NaN(), (F)4503599627370496.0, (F)8388608.0, (F)16777216.0, (F)9007199254740992.0};
    const size_t n0 = e.size();
    for (size_t i = 0; i < n0; ++i) e.push_back(-e[i]);
    for (F a : e) for (F b : e) check(a, b);
    const int emax = (1 << Bits<F>::EXPBITS) - 2;
    // random pairs: same exponent neighbourhood (the ordinary case), any exponents, subnormal divisors
    for (long i = 0; i < 3000000; ++i) {
        const int ea = 1 + (int)(rnd() % (uint64_t)emax);
        const int d = (int)(rnd() % 70);   // exponent distance up to beyond the mantissa width: both paths
        const int eb = ea - d < 0 ? 0 : ea - d;
        const F a = rnd_float<F>(ea, ea), b = rnd_float<F>(eb, eb);
        check(a, b); check(b, a);
    }
    for (long i = 0; i < 3000000; ++i) check(rnd_float<F>(0, emax), rnd_float<F>(0, emax));
    // exact multiples and their neighbours: a = k * b (exact when it fits) and one ulp either side
    for (long i = 0; i < 3000000; ++i) {
        const F b = rnd_float<F>(emax / 2 - 20, emax / 2 + 20);
        const F k = (F)(rnd() % (1ull << (int)(rnd() % (Bits<F>::MANT + 2))));
        const F a = k * b;
        check(a, b); check(std::nextafter(a, (F)0), b); check(std::nextafter(a, L::infinity()), b);
        check(std::nextafter(a, -L::infinity()), b);
    }
    // small integers (the reference's own vectors use integer-valued floats)
    for (int a = -300; a <= 300; ++a) for (int b = -40; b <= 40; ++b) check((F)a, (F)b), check((F)a / 8, (F)b / 16);
    printf("%s: %ld checked, %ld failed\n", name, checked, failed);
    return failed != 0;
}

int main() {
    int bad = 0;
    bad |= run<double>("f64");
    bad |= run<float>("f32");
    return bad;
}
