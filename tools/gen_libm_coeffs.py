"""Generates the polynomial coefficients of the device sin/exp/log in csrc/device_sem.cuh.

Near-minimax polynomials from Chebyshev interpolation in 60-digit arithmetic (mpmath), converted to
the monomial basis and rounded to double.  Prints C hex-float literals and the approximation error of
the ROUNDED polynomial (evaluated exactly), relative to the function value.

  sin:  sin(r) = r + r*s*S(s),       s = r*r,  |r| <= pi/4 (+ slack)          S = (sin(r)/r - 1)/s
  cos:  cos(r) = 1 + s*C(s)                                                    C = (cos(r) - 1)/s
  exp:  exp(r) = 1 + r*(1 + r*P(r)), |r| <= ln2/2 (+ slack)                    P = (exp(r) - 1 - r)/r^2
  log:  log(m) = u + u^3*Q(w),       u = 2(m-1)/(m+1), w = u*u, m in [sqrt(1/2), sqrt(2)]   Q = (2*atanh(u/2) - u)/u^3
"""
import sys
from mpmath import mp, mpf, cos, pi, sin, exp, log, atanh, sqrt, matrix, lu_solve

mp.dps = 60


def _fit(fn, a, b, powers, weight=None):
    """Least-squares combination of x^p (p in powers) over 12x as many Chebyshev-distributed points of
    [a, b] (dense Chebyshev sampling makes least squares close to minimax)."""
    from mpmath import qr_solve
    n = len(powers)
    m = 12 * n + 7
    xs = [(a + b) / 2 + (b - a) / 2 * cos(pi * (2 * k + 1) / (2 * m)) for k in range(m)]
    A = matrix(m, n)
    y = matrix(m, 1)
    for i, x in enumerate(xs):
        wgt = weight(x) if weight else mpf(1)
        for j, p in enumerate(powers):
            A[i, j] = wgt * x ** p
        y[i] = wgt * fn(x)
    return qr_solve(A, y)[0]


def cheb_fit(fn, a, b, deg, weight=None):
    """Coefficients rounded to double one at a time, lowest order first: after fixing c0..ck the
    remaining powers are refitted to fn - sum_{j<=k} cj x^j, so later coefficients absorb the rounding
    of earlier ones."""
    fixed = []
    for k in range(deg + 1):
        def resid(x, fixed=tuple(fixed)):
            acc = fn(x)
            for j, cj in enumerate(fixed):
                acc -= cj * x ** j
            return acc
        c = _fit(resid, a, b, list(range(k, deg + 1)), weight)
        fixed.append(mpf(float(c[0])))
    return fixed


def max_rel_err(fn_exact, fn_approx, a, b, samples=4001):
    worst = mpf(0)
    for k in range(samples):
        x = a + (b - a) * mpf(k) / (samples - 1)
        e = fn_exact(x)
        if e == 0:
            continue
        worst = max(worst, abs((fn_approx(x) - e) / e))
    return worst


def horner(c, x):
    acc = mpf(0)
    for v in reversed(c):
        acc = acc * x + v
    return acc


def show(name, coeffs):
    print(f"// {name}")
    for i, c in enumerate(coeffs):
        print(f"    {float(c).hex()},   // [{i}] {float(c):.17g}")


if __name__ == "__main__":
    # ---- sin / cos on [-pi/4, pi/4] ----
    from mpmath import factorial
    rmax = pi / 4 * mpf("1.001")
    def G(s):      # (sin(r)/r - 1)/s = sum_{k>=1} (-1)^k s^(k-1) / (2k+1)!
        return sum((-1) ** k * s ** (k - 1) / factorial(2 * k + 1) for k in range(1, 30))
    def C(s):      # (cos(r) - 1)/s = sum_{k>=1} (-1)^k s^(k-1) / (2k)!
        return sum((-1) ** k * s ** (k - 1) / factorial(2 * k) for k in range(1, 30))
    for deg in (5, 6):
        c = cheb_fit(G, mpf(0), rmax ** 2, deg)
        err = max_rel_err(lambda r: sin(r), lambda r: r + r * r * r * horner(c, r * r), mpf("1e-3"), rmax)
        print(f"sin: degree {deg} in s: max rel err {float(err):.3e} (2^-53 = 1.11e-16)")
        if err < mpf(2) ** -57:
            show(f"sin S(s), degree {deg}", c)
            break
    for deg in (5, 6, 7):
        c = cheb_fit(C, mpf(0), rmax ** 2, deg)
        err = max_rel_err(lambda r: cos(r), lambda r: 1 + r * r * horner(c, r * r), mpf(0), rmax)
        print(f"cos: degree {deg} in s: max rel err {float(err):.3e}")
        if err < mpf(2) ** -57:
            show(f"cos C(s), degree {deg}", c)
            break
    # ---- exp ----
    rmax = log(2) / 2 * mpf("1.001")
    def P(r):      # (exp(r) - 1 - r)/r^2 = sum_{k>=2} r^(k-2)/k!
        return sum(r ** (k - 2) / factorial(k) for k in range(2, 40))
    for deg in (8, 9, 10):
        c = cheb_fit(P, -rmax, rmax, deg)
        err = max_rel_err(lambda r: exp(r), lambda r: 1 + r * (1 + r * horner(c, r)), -rmax, rmax)
        print(f"exp: degree {deg}: max rel err {float(err):.3e}")
        if err < mpf(2) ** -57:
            show(f"exp P(r), degree {deg}", c)
            break
    # ---- log ----
    umax = 2 * (sqrt(2) - 1) / (sqrt(2) + 1) * mpf("1.001")
    def Q(w):      # (2 atanh(u/2) - u)/u^3 = sum_{k>=1} w^(k-1) / ((2k+1) 4^k)
        return sum(w ** (k - 1) / ((2 * k + 1) * mpf(4) ** k) for k in range(1, 60))
    for deg in (5, 6, 7, 8):
        c = cheb_fit(Q, mpf(0), umax ** 2, deg)
        err = max_rel_err(lambda u: 2 * atanh(u / 2), lambda u: u + u ** 3 * horner(c, u * u), mpf("1e-4"), umax)
        print(f"log: degree {deg} in w: max rel err {float(err):.3e}")
        if err < mpf(2) ** -57:
            show(f"log Q(w), degree {deg}", c)
            break
    # ---- constants ----
    def split3(v):
        hi = mpf(float(v)); mid = mpf(float(v - hi)); lo = mpf(float(v - hi - mid))
        return hi, mid, lo
    for nm, v in (("pi/2", pi / 2), ("ln2", log(2))):
        hi, mid, lo = split3(v)
        print(f"// {nm}: hi {float(hi).hex()}  mid {float(mid).hex()}  lo {float(lo).hex()}")
    print(f"// 2/pi {float(2 / pi).hex()}   log2(e) {float(1 / log(2)).hex()}")
    # pi/2 split whose last part is positive (keeps sin(-0) = -0 through the three FMAs)
    import math
    hi = mpf(float(pi / 2)); mid = math.nextafter(float(pi / 2 - hi), 0.0); lo = pi / 2 - hi - mpf(mid)
    print(f"// pi/2 (positive tail): hi {float(hi).hex()}  mid {mid.hex()}  lo {float(lo).hex()}")
