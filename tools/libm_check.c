/* Accuracy of the device sin/exp/log (csrc/device_libm.cuh) measured on the host: the header is
 * compiled with MR_LIBM_HOST (fma() for __fma_rn, same operations in the same order) and compared
 * with the 80-bit long double libm.  Also reports how often the result differs from glibc's double
 * routines (what the reference calls).
 * Build: gcc -O2 -ffp-contract=off -mfma -DMR_LIBM_HOST -I maray_b200/csrc -o /tmp/libm_check tools/libm_check.c -lm */
#include <stdio.h>
#include <stdlib.h>
#include "device_libm.cuh"
static uint64_t rs = 88172645463325252ull;
static double urand(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (rs >> 11) * (1.0 / 9007199254740992.0); }
static uint64_t d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static double ulp_err(double got, long double want) {
    if (want == 0) return got == 0 ? 0 : 1e9;
    int ex; frexpl(want, &ex);
    long double ulp = ldexpl(1.0L, ex - 53);
    return (double)fabsl(((long double)got - want) / ulp);
}
typedef double (*f1)(double); typedef long double (*fl)(long double);
/* mode 0: uniform in [lo,hi]; mode 1: log-spaced magnitudes in [lo,hi], positive; mode 2: same, random sign */
static void run(const char* name, f1 mine, f1 libm, fl ref, double lo, double hi, int mode, int n) {
    double maxe = 0, maxe_libm = 0, sum = 0, worst_x = 0; int mism = 0;
    for (int i = 0; i < n; i++) {
        double u = urand();
        double x = mode ? exp(log(lo) + (log(hi) - log(lo)) * u) : lo + (hi - lo) * u;
        if (mode == 2 && (i & 1)) x = -x;
        double g = mine(x), l = libm(x); long double w = ref((long double)x);
        double e = ulp_err(g, w), e2 = ulp_err(l, w);
        if (e > maxe) { maxe = e; worst_x = x; }
        if (e2 > maxe_libm) maxe_libm = e2;
        sum += e; if (d2u(g) != d2u(l)) mism++;
    }
    printf("%-4s %-28s max %.3f ULP (glibc %.3f)  mean %.3f  differs from glibc %.2f%%  worst x=%.17g\n",
           name, mode ? "log-spaced" : "uniform", maxe, maxe_libm, sum / n, 100.0 * mism / n, worst_x);
    printf("     range [%g, %g]\n", lo, hi);
}
int main(void) {
    int n = 2000000, bad = 0;
    run("sin", mr_sin, sin, sinl, -3.2, 3.2, 0, n);
    run("sin", mr_sin, sin, sinl, -100, 100, 0, n);
    run("sin", mr_sin, sin, sinl, 1e-300, 1e-5, 2, n);
    run("sin", mr_sin, sin, sinl, 100, 4194303.0, 2, n);
    run("exp", mr_exp, exp, expl, -1, 1, 0, n);
    run("exp", mr_exp, exp, expl, -40, 40, 0, n);
    run("exp", mr_exp, exp, expl, -707.9, 707.9, 0, n);
    run("log", mr_log, log, logl, 0.5, 2.0, 0, n);
    run("log", mr_log, log, logl, 0.99, 1.01, 0, n);
    run("log", mr_log, log, logl, 1e-300, 1e300, 1, n);
    run("log", mr_log, log, logl, 1.0, 3.0, 0, n);
    /* special values must equal glibc exactly (they take the fall-back branch or are exact) */
    double sp[] = {0.0, -0.0, INFINITY, -INFINITY, NAN, 1.0, 4.9e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
                   709.0, 709.78, 710.0, -745.0, -746.0, 1e10, 4194304.0, -4194304.0, 1e-310, -1.0, -1e-320};
    for (unsigned i = 0; i < sizeof sp / sizeof sp[0]; i++) {
        double x = sp[i];
        double a[3] = {mr_sin(x), mr_exp(x), mr_log(x)}, b[3] = {sin(x), exp(x), log(x)};
        for (int k = 0; k < 3; k++) {
            int same = d2u(a[k]) == d2u(b[k]) || (isnan(a[k]) && isnan(b[k]));
            int fast_path_ok = ulp_err(a[k], k == 0 ? sinl(x) : k == 1 ? expl(x) : logl(x)) < 1.0;
            if (!same && !(isfinite(b[k]) && fast_path_ok)) { printf("SPECIAL MISMATCH fn %d x=%g: %a vs %a\n", k, x, a[k], b[k]); bad++; }
        }
    }
    /* sin(-0) keeps its sign */
    if (!signbit(mr_sin(-0.0)) || signbit(mr_sin(0.0))) { printf("sin(+-0) sign wrong\n"); bad++; }
    printf(bad ? "FAILED\n" : "special values ok\n");
    return bad != 0;
}
