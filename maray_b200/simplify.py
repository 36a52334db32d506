"""Content-time rewrites of the reference, restated: `Expr::simplify` and `constant_reduction`.

These run when a scene is AUTHORED (reference examples/chess.rs:42), never during a render, and they
change floating-point results (they are algebraic rewrites on rational constants), so a scene must go
through them to be "the same scene" as one the reference's tooling produced.  Restated rule for rule,
in the reference's order:

  Expr::simplify            reference src/lib.rs:601-604  (constant_reduction at the node, then run)
  simplify::run / Case      reference src/simplify.rs:5-327
  constant_reduction::run   reference src/constant_reduction.rs:9-186
  get_* pattern helpers     reference src/lib.rs:402-536

Pinned by the reference's own unit tests (src/lib.rs:1287-1515, 1693-1719), ported in
tests/test_simplify.py.  Expressions here are hash-consed (maray_b200/expr.py), so
structural equality is identity and results are memoised per node; `Arc` nodes do not occur
(the Python builder never creates them) and natural numbers are unbounded integers -- the
reference's u64 arithmetic would wrap in release builds, which no scene of interest reaches.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

from . import expr as E
from .expr import Expr

# ---- pattern helpers (reference src/lib.rs:402-536) ------------------------------------------------


def get_nat(e: Expr) -> Optional[int]:
    return e.n if e.tag == E.NAT else None


def get_neg(e: Expr) -> Optional[Expr]:
    return e.a if e.tag == E.NEG else None


def get_recip(e: Expr) -> Optional[Expr]:
    return e.a if e.tag == E.RECIP else None


def get_add(e: Expr) -> Optional[Tuple[Expr, Expr]]:
    return (e.a, e.b) if e.tag == E.ADD else None


def get_sub(e: Expr) -> Optional[Tuple[Expr, Expr]]:
    if e.tag == E.ADD and e.b.tag == E.NEG:
        return e.a, e.b.a
    return None


def get_mul(e: Expr) -> Optional[Tuple[Expr, Expr]]:
    return (e.a, e.b) if e.tag == E.MUL else None


def get_div(e: Expr) -> Optional[Tuple[Expr, Expr]]:
    if e.tag == E.MUL and e.b.tag == E.RECIP:
        return e.a, e.b.a
    return None


def get_square(e: Expr) -> Optional[Expr]:
    return e.a if e.tag == E.MUL and e.a is e.b else None


# ---- Case (reference src/simplify.rs:5-127) --------------------------------------------------------
# A case is (sign, kind, p, q): kind "none" | "nat" (p) | "div" (p/q); sign True = non-negative.

_PRIMES_ONCE = (2, 3, 5, 7, 11, 13, 17)


def case_from_expr(e: Expr):
    t = e.tag
    if t == E.NAT:
        return True, "nat", e.n, 0
    if t == E.NEG:
        a = e.a
        if a.tag == E.NAT:
            return False, "nat", a.n, 0
        if a.tag == E.RECIP:
            return (False, "div", 1, a.a.n) if a.a.tag == E.NAT else (False, "none", 0, 0)
        if a.tag == E.MUL:
            if a.a.tag == E.NAT and a.b.tag == E.RECIP and a.b.a.tag == E.NAT:
                return False, "div", a.a.n, a.b.a.n          # NB: not reduced under a negation (:39-47)
            return False, "none", 0, 0
        return False, "none", 0, 0
    if t == E.RECIP:
        return (True, "div", 1, e.a.n) if e.a.tag == E.NAT else (True, "none", 0, 0)
    if t == E.MUL:
        if e.a.tag == E.NAT and e.b.tag == E.RECIP:
            if e.b.a.tag == E.NAT:
                a, b = e.a.n, e.b.a.n
                if a == b and b != 0:
                    return True, "nat", 1, 0
                for p in _PRIMES_ONCE:                       # each prime cancelled at most once (:66-71)
                    if a % p == 0 and b % p == 0:
                        a //= p
                        b //= p
                return True, "div", a, b
            return True, "none", 0, 0
        return True, "none", 0, 0
    return True, "none", 0, 0


def case_to_expr(case) -> Optional[Expr]:
    sign, kind, p, q = case
    if kind == "none":
        return None
    if kind == "nat":
        n = E.nat(p)
    elif p == 1:
        n = E.recip(E.nat(q))
    elif q == 1:
        n = E.nat(p)
    else:
        n = E.div(E.nat(p), E.nat(q))
    return n if sign else E.neg(n)


def normalize(e: Expr) -> Expr:
    out = case_to_expr(case_from_expr(e))
    if out is None:
        raise ValueError("Case::normalize on a non-constant (the reference unwraps None here)")
    return out


def _add_nat(a, b) -> Expr:
    (sa, va), (sb, vb) = a, b
    if sa and sb:
        return E.nat(va + vb)
    if not sa and not sb:
        return E.neg(E.nat(va + vb))
    pos, negv = (va, vb) if sa else (vb, va)
    return E.nat(pos - negv) if pos >= negv else E.neg(E.nat(negv - pos))


def _add_div(a, b) -> Expr:
    (sa, a0, a1), (sb, b0, b1) = a, b
    if sa and sb:
        if a1 == b1:
            return simplify(E.div(E.nat(a0 + b0), E.nat(a1)))
        return simplify(E.div(E.nat(a0 * b1 + a1 * b0), E.nat(a1 * b1)))
    if not sa and not sb:
        return simplify(E.neg(E.add(E.div(E.nat(a0), E.nat(a1)), E.div(E.nat(b0), E.nat(b1)))))
    # mixed signs: (a0, a1) names the positive operand, (b0, b1) the negative one (:104)
    if not sa:
        (a0, a1), (b0, b1) = (b0, b1), (a0, a1)
    if a1 == b1:
        if a0 >= b0:
            return simplify(E.div(E.nat(a0 - b0), E.nat(a1)))
        return simplify(E.neg(E.div(E.nat(b0 - a0), E.nat(a1))))
    d1, d2 = a0 * b1, a1 * b0
    if d1 >= d2:
        return simplify(E.div(E.nat(d1 - d2), E.nat(a1 * b1)))
    return simplify(E.neg(E.div(E.nat(d2 - d1), E.nat(a1 * b1))))


def _mul_nat(a, b) -> Expr:
    (sa, va), (sb, vb) = a, b
    return E.nat(va * vb) if sa == sb else E.neg(E.nat(va * vb))


def _mul_div(a, b) -> Expr:
    (sa, a0, a1), (sb, b0, b1) = a, b
    d = E.div(E.nat(a0 * b0), E.nat(a1 * b1))
    return d if sa == sb else E.neg(d)


# ---- constant_reduction (reference src/constant_reduction.rs:9-186) -------------------------------
_CR_FACTORS = (2,) * 6 + (3,) * 4 + (5,) * 3 + (7,) * 3 + (11,) * 2 + (13, 17)


def _cancel(values):
    """Repeated trial division of all `values` by the common factors 2^6 3^4 5^3 7^3 11^2 13 17."""
    vals = list(values)
    for p in _CR_FACTORS:
        if all(v % p == 0 for v in vals):
            vals = [v // p for v in vals]
    return vals


def constant_reduction(e: Expr) -> Expr:
    """Structure-preserving cancellation of constants at THIS node (the reference mutates in place;
    the patterns are tried in its order, each on the result of the previous one)."""
    if e.tag != E.MUL:
        return e
    a, b = e.a, e.b
    # `a1/k * k`                                                                            (:12-28)
    da = get_div(a)
    if da is not None and b.tag == E.NAT and da[1].tag == E.NAT:
        a2, bb = _cancel([da[1].n, b.n])
        a, b = E.div(da[0], E.nat(a2)), E.nat(bb)
    # `(k * a2) / k` and `(a1 * k) / k`                                                     (:29-62)
    ma = get_mul(a)
    if ma is not None and b.tag == E.RECIP:
        a1, a2 = ma
        rb = b.a
        if a1.tag == E.NAT and rb.tag == E.NAT:
            n1, nb = _cancel([a1.n, rb.n])
            a1, rb = E.nat(n1), E.nat(nb)
        if a2.tag == E.NAT and rb.tag == E.NAT:
            n2, nb = _cancel([a2.n, rb.n])
            a2, rb = E.nat(n2), E.nat(nb)
        a, b = E.mul(a1, a2), E.recip(rb)
    # `k * (b11/k - b21/k)`                                                                 (:63-83)
    sb = get_sub(b)
    if a.tag == E.NAT and sb is not None:
        d1, d2 = get_div(sb[0]), get_div(sb[1])
        if d1 is not None and d2 is not None and d1[1].tag == E.NAT and d2[1].tag == E.NAT:
            k, n1, n2 = _cancel([a.n, d1[1].n, d2[1].n])
            a, b = E.nat(k), E.sub(E.div(d1[0], E.nat(n1)), E.div(d2[0], E.nat(n2)))
    # `k * (b11/k + b21/k)`                                                                 (:84-104)
    ab = get_add(b)
    if a.tag == E.NAT and ab is not None:
        d1, d2 = get_div(ab[0]), get_div(ab[1])
        if d1 is not None and d2 is not None and d1[1].tag == E.NAT and d2[1].tag == E.NAT:
            k, n1, n2 = _cancel([a.n, d1[1].n, d2[1].n])
            a, b = E.nat(k), E.add(E.div(d1[0], E.nat(n1)), E.div(d2[0], E.nat(n2)))
    # `k*(b11*(b1211/k - b1221/k) - b21/k)`                                                 (:105-136)
    sb = get_sub(b)
    if a.tag == E.NAT and sb is not None:
        m1, d2 = get_mul(sb[0]), get_div(sb[1])
        if m1 is not None and d2 is not None and d2[1].tag == E.NAT:
            inner = get_sub(m1[1])
            if inner is not None:
                i1, i2 = get_div(inner[0]), get_div(inner[1])
                if i1 is not None and i2 is not None and i1[1].tag == E.NAT and i2[1].tag == E.NAT:
                    k, n1, n2, n3 = _cancel([a.n, i1[1].n, i2[1].n, d2[1].n])
                    a = E.nat(k)
                    b = E.sub(E.mul(m1[0], E.sub(E.div(i1[0], E.nat(n1)), E.div(i2[0], E.nat(n2)))),
                              E.div(d2[0], E.nat(n3)))
    # `k*(b11*(b1211/k - b1221/k) - b21*(b2211/k - b2221/k))`                               (:137-183)
    sb = get_sub(b)
    if a.tag == E.NAT and sb is not None:
        m1, m2 = get_mul(sb[0]), get_mul(sb[1])
        if m1 is not None and m2 is not None:
            s1, s2 = get_sub(m1[1]), get_sub(m2[1])
            if s1 is not None and s2 is not None:
                ds = [get_div(s1[0]), get_div(s1[1]), get_div(s2[0]), get_div(s2[1])]
                if all(d is not None and d[1].tag == E.NAT for d in ds):
                    k, n1, n2, n3, n4 = _cancel([a.n] + [d[1].n for d in ds])
                    a = E.nat(k)
                    b = E.sub(E.mul(m1[0], E.sub(E.div(ds[0][0], E.nat(n1)), E.div(ds[1][0], E.nat(n2)))),
                              E.mul(m2[0], E.sub(E.div(ds[2][0], E.nat(n3)), E.div(ds[3][0], E.nat(n4)))))
    return E.mul(a, b)


# ---- simplify (reference src/lib.rs:601-604, src/simplify.rs:129-327) -----------------------------
_MEMO: Dict[int, Expr] = {}
_KEEP = []      # keeps memoised operands alive so ids are not recycled
_ACTIVE = set() # expressions whose simplification is in progress (cycle detection)


class SimplifyDiverges(RecursionError):
    """`Expr::simplify` is a pure function of the expression, so re-entering it on an expression whose
    simplification is still in progress means the reference recurses until its stack overflows.
    HEAD does that on `(p/n)/m` with a non-constant `p`: the rule `(a0/a1)*b -> (a0*b)/a1`
    (reference src/simplify.rs:286) turns `(p/n)*(1/m)` into `(p*(1/m))/n`, whose first operand is
    again a division, and back.  examples/chess.rs reaches that shape (tests/test_simplify.py)."""


def simplify(e: Expr) -> Expr:
    """`Expr::simplify`: constant reduction at the node, then the rewrite rules (children first)."""
    hit = _MEMO.get(id(e))
    if hit is not None:
        return hit
    if id(e) in _ACTIVE:
        raise SimplifyDiverges("the reference's simplify does not terminate on this expression")
    _ACTIVE.add(id(e))
    try:
        out = _run(constant_reduction(e))
    finally:
        _ACTIVE.discard(id(e))
    _MEMO[id(e)] = out
    _KEEP.append(e)
    return out


def clear_memo() -> None:
    _MEMO.clear()
    _KEEP.clear()
    _ACTIVE.clear()


def _run(e: Expr) -> Expr:
    t = e.tag
    if t in (E.X, E.Y, E.TAU, E.E, E.NAT, E.VAR, E.LET):
        return e
    if t == E.ARC:
        raise NotImplementedError("Arc nodes do not occur in the Python builder")
    if t == E.NEG:
        a = simplify(e.a)
        inner = get_neg(a)
        if inner is not None:
            return simplify(inner)
        if get_nat(a) == 0:
            return E.nat(0)
        s = get_sub(a)
        if s is not None:
            return E.sub(s[1], s[0])
        return E.neg(a)
    if t == E.ABS:
        return E.abs(simplify(e.a))
    if t == E.RECIP:
        a = simplify(e.a)
        d = get_div(a)
        if d is not None:
            return simplify(E.div(d[1], d[0]))
        inner = get_neg(a)
        if inner is not None:
            return simplify(E.neg(E.recip(inner)))
        inner = get_recip(a)
        if inner is not None:
            return simplify(inner)
        if get_nat(a) == 1:
            return E.nat(1)
        return E.recip(a)
    if t == E.SQRT:
        return E.sqrt(simplify(e.a))
    if t == E.STEP:
        a = simplify(e.a)
        if get_nat(a) is not None:
            return E.nat(1)
        inner = get_neg(a)
        if inner is not None and get_nat(inner) is not None:
            return E.nat(1) if get_nat(inner) == 0 else E.nat(0)
        sign, kind, _, _ = case_from_expr(a)
        if kind == "div":
            return E.nat(1) if sign else E.nat(0)
        return E.step(a)
    if t == E.SIN:
        a = simplify(e.a)
        if a.tag == E.TAU:
            return E.nat(0)
        ad = get_add(a)
        if ad is not None:
            if ad[0].tag == E.TAU:
                return E.sin(ad[1])
            if ad[1].tag == E.TAU:
                return E.sin(ad[0])
        return E.sin(a)
    if t == E.EXP:
        a = simplify(e.a)
        if get_nat(a) == 0:
            return E.nat(1)
        if get_nat(a) == 1:
            return E.e()
        return E.exp(a)
    if t == E.LN:
        return E.ln(simplify(e.a))
    if t == E.ADD:
        a, b = simplify(e.a), simplify(e.b)
        ca, cb = case_from_expr(a), case_from_expr(b)
        if ca[1] == "nat" and ca[2] == 0:
            return b
        if cb[1] == "nat" and cb[2] == 0:
            return a
        if ca[1] != "none" and cb[1] != "none":
            if ca[1] == "nat" and cb[1] == "nat":
                return normalize(_add_nat((ca[0], ca[2]), (cb[0], cb[2])))
            if ca[1] == "div" and cb[1] == "div":
                return normalize(_add_div((ca[0], ca[2], ca[3]), (cb[0], cb[2], cb[3])))
            if ca[1] == "nat":
                return normalize(_add_div((ca[0], ca[2] * cb[3], cb[3]), (cb[0], cb[2], cb[3])))
            return normalize(_add_div((ca[0], ca[2], ca[3]), (cb[0], ca[3] * cb[2], ca[3])))
        na, nb = get_neg(a), get_neg(b)
        if na is not None and nb is not None:
            return simplify(E.neg(E.add(na, nb)))
        if na is not None:
            return simplify(E.sub(b, na))
        sa = get_sub(a)
        if sa is not None:
            ab = get_add(b)
            if ab is not None:
                if sa[1] is ab[0]:
                    return E.add(sa[0], ab[1])
                if sa[1] is ab[1]:
                    return E.add(sa[0], ab[0])
        sb = get_sub(b)
        if sb is not None:
            if sb[1] is a or sb[1] is b:
                return sb[0]
        return E.add(a, b)
    if t == E.MUL:
        a, b = simplify(e.a), simplify(e.b)
        ca, cb = case_from_expr(a), case_from_expr(b)
        if ca[1] == "nat" and ca[2] == 0:
            return E.nat(0)
        if cb[1] == "nat" and cb[2] == 0:
            return E.nat(0)
        if ca[0] and ca[1] == "nat" and ca[2] == 1:
            return b
        if cb[0] and cb[1] == "nat" and cb[2] == 1:
            return a
        if not ca[0] and ca[1] == "nat" and ca[2] == 1:
            return simplify(E.neg(b))
        if not cb[0] and cb[1] == "nat" and cb[2] == 1:
            return simplify(E.neg(a))
        if ca[1] != "none" and cb[1] != "none":
            if ca[1] == "nat" and cb[1] == "nat":
                return normalize(_mul_nat((ca[0], ca[2]), (cb[0], cb[2])))
            if ca[1] == "div" and cb[1] == "div":
                return normalize(_mul_div((ca[0], ca[2], ca[3]), (cb[0], cb[2], cb[3])))
            if ca[1] == "nat":
                return normalize(_mul_div((ca[0], ca[2], 1), (cb[0], cb[2], cb[3])))
            return normalize(_mul_div((cb[0], cb[2], 1), (ca[0], ca[2], ca[3])))
        na, nb = get_neg(a), get_neg(b)
        if na is not None and nb is not None:
            return simplify(E.mul(na, nb))
        if na is not None:
            return simplify(E.neg(E.mul(na, b)))
        if nb is not None:
            return simplify(E.neg(E.mul(a, nb)))
        ra, rb = get_recip(a), get_recip(b)
        if ra is not None and rb is not None:
            return simplify(E.recip(E.mul(ra, rb)))
        if ra is not None:
            return simplify(E.div(b, ra))
        da, db = get_div(a), get_div(b)
        if da is not None and db is not None:
            return simplify(E.div(E.mul(da[0], db[0]), E.mul(da[1], db[1])))
        if da is not None:
            return simplify(E.div(E.mul(da[0], b), da[1]))
        if db is not None:
            return simplify(E.div(E.mul(a, db[0]), db[1]))
        rb = get_recip(b)
        if rb is not None and get_nat(rb) is not None:
            ma = get_mul(a)
            if ma is not None and get_nat(ma[1]) is not None:
                return simplify(E.mul(E.div(E.nat(ma[1].n), E.nat(rb.n)), ma[0]))
        ma = get_mul(a)
        if ma is not None and get_nat(ma[0]) is not None and get_nat(b) is not None:
            return simplify(E.mul(E.nat(ma[0].n * b.n), ma[1]))
        return E.mul(a, b)
    if t in (E.MAX, E.MIN):
        a, b = simplify(e.a), simplify(e.b)
        if get_nat(a) is not None and get_nat(b) is not None:
            if t == E.MAX:
                return E.nat(a.n if a.n >= b.n else b.n)
            return E.nat(a.n if a.n <= b.n else b.n)
        return E.max(a, b) if t == E.MAX else E.min(a, b)
    if t == E.DECOR:
        return E.decor(simplify(e.a))
    if t == E.APP:
        return E.app(e.n, simplify(e.a), simplify(e.b))
    raise ValueError(f"unknown tag {t}")
