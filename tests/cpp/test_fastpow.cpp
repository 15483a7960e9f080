// minarrow_b200/csrc/fastpow.h (float64 Power without the two libm calls) against the reference's expression
// exp(b * ln a) evaluated by glibc — the oracle's own arithmetic — and against the long-double value of the same
// expression.  Host-only; the device build of the same header is compared with the oracle in tests/test_gpu_parity.py.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <random>

#include "../../minarrow_b200/csrc/fastpow.h"

static int g_failed = 0;
static void run(const char* name, double lo_a, double hi_a, double max_x, double bar, uint64_t seed, int n) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> ua(lo_a, hi_a), ux(-max_x, max_x);
    double worst = 0, worst_true = 0;
    for (int i = 0; i < n; ++i) {
        const double a = std::exp(ua(rng));
        const double la = std::log(a);
        double b = std::fabs(la) > 1e-3 ? ux(rng) / la : ux(rng);
        const double ref = std::exp(b * la);                               // the oracle's expression
        const long double tru = expl((long double)(b * la));
        const double got = mnr::fast_pow_f64(a, b);
        if (!std::isfinite(ref) || std::fabs(ref) < 2.3e-308) continue;   // subnormal results: compared in the specials loop (absolute)
        const double e1 = std::fabs(got - ref) / std::fabs(ref);
        const double e2 = (double)(fabsl((long double)got - tru) / fabsl(tru));
        if (e1 > worst) worst = e1;
        if (e2 > worst_true) worst_true = e2;
    }
    const bool ok = worst <= bar;
    if (!ok) ++g_failed;
    std::printf("%-34s max rel. diff vs glibc exp(b*log a) %.3e, vs long double %.3e  (bar %.0e) %s\n", name, worst, worst_true, bar, ok ? "ok" : "FAIL");
}

int main() {
    run("a in e^[-9,9], |b ln a| <= 74", -9, 9, 74, 1e-13, 1, 4000000);
    run("a in e^[-0.01,0.01], |b ln a| <= 74", -0.01, 0.01, 74, 1e-13, 2, 2000000);
    run("a in e^[-700,700], |b ln a| <= 100", -700, 700, 100, 1e-13, 3, 4000000);
    run("a in e^[-30,30], |b ln a| <= 699", -30, 30, 699, 1e-12, 4, 4000000);
    run("a in e^[-30,30], |b ln a| in [690, 745]", -30, 30, 745, 1e-12, 5, 2000000);
    run("subnormal a, |b ln a| <= 74", -744.4, -708.5, 74, 1e-13, 6, 1000000);
    // specials are folded in branch-free: they must give what exp(b * log(a)) gives
    const double sp[] = {0.0, -0.0, -1.5, INFINITY, -INFINITY, NAN, 4.9e-324, 2.2e-308, 1.0, 2.0, 0.5, 1e308, -1e308, 1e-300, 700.0, -800.0};
    int bad = 0;
    for (double a : sp)
        for (double b : sp) {
            const double r = std::exp(b * std::log(a)), g = mnr::fast_pow_f64(a, b);
            if (!((r != r && g != g) || r == g || std::fabs(g - r) <= 1e-12 * std::fabs(r) + 1e-320)) { ++bad; std::printf("special a=%g b=%g: %g vs %g\n", a, b, g, r); }
        }
    if (bad) ++g_failed;
    std::printf("specials: %d mismatches\n%d failed\n", bad, g_failed);
    return g_failed;
}
