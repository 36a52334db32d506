// Pins the exact mode of the device libm (csrc/device_libm_glibc.cuh) to the host's glibc: the header's host
// rendition (MR_LIBM_HOST: fma() for __fma_rn, plain loads for __ldg) must return the bits of sin/exp/log for every
// argument tried.  Test infrastructure; run by tests/test_glibc_libm.py.
//
//   gcc -O2 -mfma -ffp-contract=off -fno-builtin-sin -fno-builtin-exp -fno-builtin-log -o glibc_libm_check glibc_libm_check.c -lm
//   ./glibc_libm_check [samples per range, default 10000000]
//
// Prints one line per (function, range): samples, differing results, and the first few differences; exits 1 if any.
#define MR_LIBM_HOST 1
#ifndef FAST_MODE            /* -DFAST_MODE: the default libm; its exp and log (the public names) must be exact too */
#define MR_LIBM_GLIBC 1
#endif
#include "../maray_b200/csrc/device_libm.cuh"

#include <stdio.h>
#include <stdlib.h>

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rng(void) {
    uint64_t x = rng_state;
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    return rng_state = x;
}
static double uniform01(void) { return (double)(rng() >> 11) * 0x1p-53; }
static uint64_t bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }

typedef double (*fn1)(double);
static double (*volatile libm_sin)(double) = sin;   // through pointers: the compiler must not fold or substitute
static double (*volatile libm_exp)(double) = exp;
static double (*volatile libm_log)(double) = log;

static long total_bad = 0;
static int same(double a, double b) { return bits(a) == bits(b) || (a != a && b != b); }
static void report(const char* name, const char* range, long n, long bad) {
    printf("%-4s %-34s %10ld samples  %ld differ\n", name, range, n, bad);
    total_bad += bad;
}
static long check_one(const char* name, fn1 mine, fn1 ref, double x, long bad) {
    const double a = mine(x), b = ref(x);
    if (same(a, b)) return 0;
    if (bad < 4) printf("    %s(%a) = %a, glibc %a\n", name, x, a, b);
    return 1;
}
// uniform in [lo, hi), random sign when `both`
static void run_uniform(const char* name, fn1 mine, fn1 ref, double lo, double hi, int both, long n) {
    long bad = 0;
    for (long i = 0; i < n; i++) {
        double x = lo + (hi - lo) * uniform01();
        if (both && (rng() & 1)) x = -x;
        bad += check_one(name, mine, ref, x, bad);
    }
    char r[64];
    snprintf(r, sizeof r, "%s[%g, %g)", both ? "+-" : "", lo, hi);
    report(name, r, n, bad);
}
// uniform over bit patterns between two positive doubles (log-uniform in value), random sign when `both`
static void run_bits(const char* name, fn1 mine, fn1 ref, double lo, double hi, int both, long n) {
    long bad = 0;
    const uint64_t a = bits(lo), span = bits(hi) - bits(lo);
    for (long i = 0; i < n; i++) {
        double x = from_bits(a + rng() % span);
        if (both && (rng() & 1)) x = -x;
        bad += check_one(name, mine, ref, x, bad);
    }
    char r[64];
    snprintf(r, sizeof r, "%sbits[%g, %g)", both ? "+-" : "", lo, hi);
    report(name, r, n, bad);
}
// every double within `w` steps of each given value, both signs
static void run_around(const char* name, fn1 mine, fn1 ref, const double* v, int nv, int w) {
    long bad = 0, n = 0;
    for (int k = 0; k < nv; k++)
        for (int s = 0; s < 2; s++)
            for (int d = -w; d <= w; d++) {
                const uint64_t u = bits(v[k]);
                if (d < 0 && u < (uint64_t)(-d)) continue;
                double x = from_bits(u + (uint64_t)(int64_t)d);
                if (s) x = -x;
                bad += check_one(name, mine, ref, x, bad);
                n++;
            }
    report(name, "neighbourhoods of thresholds", n, bad);
}

static double my_sin(double x) { return mr_sin_g(x); }
static double my_exp(double x) { return mr_exp(x); }       /* the names the kernels call, in the mode compiled */
static double my_log(double x) { return mr_log(x); }
static double ref_sin(double x) { return libm_sin(x); }
static double ref_exp(double x) { return libm_exp(x); }
static double ref_log(double x) { return libm_log(x); }

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 10000000L;
    const double inf = 1.0 / 0.0, nan_ = inf - inf;

    // sin: the four ranges of __sin below __branred's, their boundaries, the table points k/128, tiny, specials.
    run_uniform("sin", my_sin, ref_sin, 0x1p-26, 0.126, 1, n);
    run_uniform("sin", my_sin, ref_sin, 0.126, 0.855469, 1, n);
    run_uniform("sin", my_sin, ref_sin, 0.855469, 2.426265, 1, n);
    run_uniform("sin", my_sin, ref_sin, 2.426265, 64.0, 1, n);
    run_uniform("sin", my_sin, ref_sin, 64.0, 1e5, 1, n);
    run_uniform("sin", my_sin, ref_sin, 1e5, 105414350.0, 1, n);
    run_bits("sin", my_sin, ref_sin, 0x1p-1074, 105414350.0, 1, n);
    {
        double v[512];
        int nv = 0;
        const double th[] = {0x1p-26, 0.126, 0.855469, 2.426265, 105414350.0, 0x1.921fb54442d18p+0, 0x1.921fb54442d18p+1,
                             0x1.2d97c7f3321d2p+1, 0x1.921fb54442d18p+2, 0x1p-1022, 1.0, 0.5, 710.0};
        for (unsigned i = 0; i < sizeof th / sizeof th[0]; i++) v[nv++] = th[i];
        v[nv++] = from_bits(0x3e500000ull << 32); v[nv++] = from_bits(0x3feb6000ull << 32);
        v[nv++] = from_bits(0x400368fdull << 32); v[nv++] = from_bits(0x419921fbull << 32);
        for (int k = 1; k <= 112; k++) v[nv++] = k / 128.0;
        for (int k = 1; k <= 112; k++) v[nv++] = (k + 0.5) / 128.0;
        for (int k = 1; k <= 64; k++) v[nv++] = k * 0x1.921fb54442d18p+0;     // multiples of pi/2: deep cancellation
        run_around("sin", my_sin, ref_sin, v, nv, 300);
        const double sp[] = {0.0, -0.0, inf, -inf, nan_, 0x1p-1074, 0x1p-27};
        long bad = 0;
        for (unsigned i = 0; i < sizeof sp / sizeof sp[0]; i++) bad += check_one("sin", my_sin, ref_sin, sp[i], bad);
        report("sin", "specials", (long)(sizeof sp / sizeof sp[0]), bad);
    }

    // exp
    run_uniform("exp", my_exp, ref_exp, 0.0, 1.0, 1, n);
    run_uniform("exp", my_exp, ref_exp, 0.0, 64.0, 1, n);
    run_uniform("exp", my_exp, ref_exp, 0.0, 512.0, 1, n);
    run_uniform("exp", my_exp, ref_exp, 512.0, 760.0, 1, n / 4);
    run_bits("exp", my_exp, ref_exp, 0x1p-1074, 2000.0, 1, n);
    {
        const double th[] = {0x1p-54, 512.0, 1024.0, 709.782712893384, 708.3964185322641, 745.1332191019411, 0.0054152123481245725,
                             1.0, 0x1.62e42fefa39efp-1, 0x1.62e42fefa39efp-8, 0x1p-1022};
        run_around("exp", my_exp, ref_exp, th, (int)(sizeof th / sizeof th[0]), 300);
        const double sp[] = {0.0, -0.0, inf, -inf, nan_, 0x1p-1074, 1e308, -1e308};
        long bad = 0;
        for (unsigned i = 0; i < sizeof sp / sizeof sp[0]; i++) bad += check_one("exp", my_exp, ref_exp, sp[i], bad);
        report("exp", "specials", (long)(sizeof sp / sizeof sp[0]), bad);
    }

    // log
    run_uniform("log", my_log, ref_log, 0x1p-20, 2.0, 0, n);
    run_uniform("log", my_log, ref_log, 0.9375, 1.0647, 0, n);
    run_uniform("log", my_log, ref_log, 1.0, 65.0, 0, n);
    run_uniform("log", my_log, ref_log, 1.0, 1e6, 0, n);
    run_bits("log", my_log, ref_log, 0x1p-1074, 0x1.fffffffffffffp+1023, 0, n);
    run_bits("log", my_log, ref_log, 0x1p-1074, 0x1p-1022, 0, n / 10);
    {
        const double th[] = {1.0, 0.9375, 1.0 + 0x1.09p-4, 0x1p-1022, 0.5, 2.0, 0x1.6p-1, 0x1.6p+0, 0x1.fffffffffffffp+1023};
        run_around("log", my_log, ref_log, th, (int)(sizeof th / sizeof th[0]), 300);
        const double sp[] = {0.0, -0.0, inf, -inf, nan_, 0x1p-1074, -1.0, -0x1p-1074};
        long bad = 0;
        for (unsigned i = 0; i < sizeof sp / sizeof sp[0]; i++) bad += check_one("log", my_log, ref_log, sp[i], bad);
        report("log", "specials", (long)(sizeof sp / sizeof sp[0]), bad);
    }
    // sign of the sine (mr_sin_ge0: what step(sin(u)) evaluates): equal to (sin(x) >= 0) of glibc's sine AND of the fast
    // sine it replaces, everywhere -- multiples of pi/2 included, where the reduced argument nearly cancels.
    {
        long bad = 0, cnt = 0;
        for (int pass = 0; pass < 3; pass++) {
            const double hi_[] = {8.0, 4194304.0, 4.6e18};
            for (long i = 0; i < n; i++) {
                double x = pass < 2 ? hi_[pass] * uniform01() : from_bits(rng() % bits(hi_[pass]));
                if (rng() & 1) x = -x;
                const int a = mr_sin_ge0(x), b = libm_sin(x) >= 0.0;
                const int c = mr_sin_inrange_f(x) ? (mr_sin_fast_f(x) >= 0.0) : b;
                if (a != b || a != c) { if (bad < 4) printf("    sin_ge0(%a) = %d, glibc %d, fast %d\n", x, a, b, c); bad++; }
                cnt++;
            }
        }
        for (int k = -4000; k <= 4000; k++)
            for (int d = -64; d <= 64; d++) {
                const double m = k * 0x1.921fb54442d18p+0;
                const double x = from_bits(bits(m < 0 ? -m : m) + (uint64_t)(int64_t)d) * (m < 0 ? -1.0 : 1.0);
                if (k == 0 && d < 0) continue;
                const int a = mr_sin_ge0(x), b = libm_sin(x) >= 0.0, c = mr_sin_fast_f(x) >= 0.0;
                if (a != b || a != c) { if (bad < 4) printf("    sin_ge0(%a) = %d, glibc %d, fast %d\n", x, a, b, c); bad++; }
                cnt++;
            }
        for (long i = 0; i < n / 16; i++) {                       // doubles next to random multiples of pi/2 up to 2^22
            const double m = (double)(rng() % 2670176u) * 0x1.921fb54442d18p+0;
            const double x = from_bits(bits(m) + (uint64_t)(rng() % 33u)) - 0.0;
            const double xs = (rng() & 1) ? -x : x;
            const int a = mr_sin_ge0(xs), b = libm_sin(xs) >= 0.0;
            const int c = mr_sin_inrange_f(xs) ? (mr_sin_fast_f(xs) >= 0.0) : b;
            if (a != b || a != c) { if (bad < 4) printf("    sin_ge0(%a) = %d, glibc %d, fast %d\n", xs, a, b, c); bad++; }
            cnt++;
        }
        const double sp[] = {0.0, -0.0, inf, -inf, nan_, 0x1p-1074, -0x1p-1074, 0x1p-1022, 4194304.0, -4194304.0, 1e300};
        for (unsigned i = 0; i < sizeof sp / sizeof sp[0]; i++, cnt++)
            if (mr_sin_ge0(sp[i]) != (libm_sin(sp[i]) >= 0.0)) { printf("    sin_ge0(%a)\n", sp[i]); bad++; }
        report("sgn", "sign of the sine, all ranges", cnt, bad);
    }
    printf("total differing: %ld\n", total_bad);
    return total_bad ? 1 : 0;
}
