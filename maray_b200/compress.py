"""Content-time tooling of the reference, restated: `Display for Expr`, `compressor::flatten` and
`Expr::compress` (the automatic common-sub-expression compressor that produces the `Let` of a `.maray`
file).  Like maray_b200/simplify.py this runs when a scene is AUTHORED, never during a render.

  Display for Expr            reference src/lib.rs:196-367  (`fmt` below; compress measures benefit in
                              characters of this text, so it has to be exact)
  Expr::needs_parens          reference src/lib.rs:390-401
  compressor::is_simple_expr  reference src/compressor.rs:108-119
  Compressor::count_expr      reference src/compressor.rs:31-79   (terms in first-seen pre-order)
  last_max_benefit            reference src/compressor.rs:82-105  (the LAST term with the largest benefit)
  is_compressed               reference src/compressor.rs:140-151
  compression_benefit         reference src/compressor.rs:154-164
  flatten                     reference src/compressor.rs:167-212
  compress                    reference src/compressor.rs:215-236
  Expr::rewrite               reference src/lib.rs:560-598

Pinned by the shipped scene itself: flattening the `Let` of data/chess.maray and compressing the result
again reproduces the file byte for byte (tests/test_compress.py) -- 859 definitions, same order, same
formulas.  The reference has no unit test for the compressor.

Expressions are hash-consed (maray_b200/expr.py): structural equality is identity, a sub-term's
occurrence count in the TREE is computed on the DAG by propagating multiplicities, and `Arc` nodes
(the reference's own sharing device, transparent to every function here) do not occur.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

from . import expr as E
from .expr import Expr
from .simplify import get_div, get_mul, get_neg, get_recip, get_square, get_sub

_NO_PARENS = (E.X, E.Y, E.TAU, E.E, E.VAR, E.NAT, E.ABS, E.SIN, E.STEP, E.SQRT, E.EXP, E.LN, E.MIN, E.MAX)


def needs_parens(e: Expr) -> bool:
    return e.tag not in _NO_PARENS


class Formatter:
    """`format!("{}", expr)`, memoised per node (only lengths are needed by the compressor; texts of
    big terms are built from their operands' texts once)."""

    def __init__(self):
        self._text: Dict[int, str] = {}
        self._keep: List[Expr] = []

    def fmt(self, e: Expr) -> str:
        out = self._text.get(id(e))
        if out is None:
            for n in E.dag_nodes([e]):                  # operands first: no deep recursion
                if id(n) not in self._text:
                    self._text[id(n)] = self._one(n)
                    self._keep.append(n)
            out = self._text[id(e)]
        return out

    def length(self, e: Expr) -> int:
        return len(self.fmt(e))                          # chars().count(): Python strings are code points

    def _par(self, e: Expr) -> str:
        t = self._text[id(e)]
        return "(" + t + ")" if needs_parens(e) else t

    def _one(self, e: Expr) -> str:
        t = e.tag
        T = self._text
        if t == E.X: return "x"
        if t == E.Y: return "y"
        if t == E.TAU: return "τ"
        if t == E.E: return "\U0001d41e"
        if t == E.VAR: return f"${e.n}"
        if t == E.NAT: return str(e.n)
        if t == E.NEG: return "-" + self._par(e.a)
        if t == E.ABS: return f"abs({T[id(e.a)]})"
        if t == E.RECIP: return "1/" + self._par(e.a)
        if t == E.SQRT: return f"sqrt({T[id(e.a)]})"
        if t == E.STEP: return f"step({T[id(e.a)]})"
        if t == E.SIN: return f"sin({T[id(e.a)]})"
        if t == E.EXP: return f"\U0001d41e^({T[id(e.a)]})"
        if t == E.LN: return f"ln({T[id(e.a)]})"
        if t == E.ADD:
            a, b = e.a, e.b

            def light(z: Expr) -> bool:   # forms that bind tighter than + and are printed bare
                return get_recip(z) is not None or get_div(z) is not None or get_square(z) is not None

            nb = get_neg(b)
            if nb is not None:            # a - b
                bare_a = (not needs_parens(a)) or light(a) or get_sub(a) is not None or get_mul(a) is not None
                bare_b = (not needs_parens(nb)) or light(nb)
                return (T[id(a)] if bare_a else "(" + T[id(a)] + ")") + "-" + (T[id(nb)] if bare_b else "(" + T[id(nb)] + ")")
            bare_a = (not needs_parens(a)) or light(a) or get_sub(a) is not None
            bare_b = (not needs_parens(b)) or light(b)
            return (T[id(a)] if bare_a else "(" + T[id(a)] + ")") + "+" + (T[id(b)] if bare_b else "(" + T[id(b)] + ")")
        if t == E.MUL:
            rb = get_recip(e.b)
            if rb is not None:
                return self._par(e.a) + "/" + self._par(rb)
            if e.a is e.b:
                return self._par(e.a) + "^2"
            return self._par(e.a) + "*" + self._par(e.b)
        if t == E.MAX: return f"max({T[id(e.a)]},{T[id(e.b)]})"
        if t == E.MIN: return f"min({T[id(e.a)]},{T[id(e.b)]})"
        if t == E.LET:
            return T[id(e.a)] + "\nwhere\n" + "".join(f"  ${i} = {T[id(d)]}\n" for i, d in e.vars)
        if t == E.DECOR:
            return self._par(e.a) + " : "          # the Python builder only makes empty token lists
        if t == E.APP: return f"app({e.n},{T[id(e.a)]},{T[id(e.b)]})"
        raise ValueError(f"unknown tag {t}")


def fmt(e: Expr) -> str:
    return Formatter().fmt(e)


def is_simple_expr(e: Expr, level: int = 2) -> bool:
    t = e.tag
    if t in (E.X, E.Y, E.TAU, E.E, E.NAT, E.VAR):
        return True
    if t == E.RECIP and level >= 1:
        return is_simple_expr(e.a, level - 1)
    if t == E.MUL and level >= 2:
        return is_simple_expr(e.a, level - 1) and is_simple_expr(e.b, level - 1)
    return False


_SIMPLE2: Dict[int, bool] = {}      # is_simple_expr(e, 2) per node (nodes are interned and immutable)
_SIMPLE2_KEEP: List[Expr] = []


def _simple2(e: Expr) -> bool:
    hit = _SIMPLE2.get(id(e))
    if hit is None:
        hit = is_simple_expr(e, 2)
        _SIMPLE2[id(e)] = hit
        _SIMPLE2_KEEP.append(e)
    return hit


def _count_children(e: Expr) -> Sequence[Expr]:
    """Operands `count_expr` descends into (it does not look inside a `Let`)."""
    if e.tag == E.LET:
        return ()
    if e.b is not None:
        return (e.a, e.b)
    if e.a is not None:
        return (e.a,)
    return ()


def count_terms(root: Expr) -> Tuple[List[Expr], Dict[int, int]]:
    """The compressor's term table for `root`: every sub-term that is not "simple", in the order a
    pre-order walk of the tree first meets it, with its number of occurrences in the tree."""
    order: List[Expr] = []      # first-seen pre-order
    post: List[Expr] = []       # post-order of the same walk: operands before users
    seen = set()
    stack = [(root, False)]
    while stack:                                   # left operand first; a repeated node brings nothing
        n, done = stack.pop()                      # new (everything below it was seen the first time)
        if done:
            post.append(n)
            continue
        if id(n) in seen or _simple2(n):
            continue
        seen.add(id(n))
        order.append(n)
        stack.append((n, True))
        for c in reversed(_count_children(n)):
            stack.append((c, False))
    # occurrences in the tree: multiplicities flow from users to operands, users first
    occ: Dict[int, int] = {id(n): 0 for n in order}
    if order and order[0] is root:
        occ[id(root)] = 1
    for n in reversed(post):
        k = occ[id(n)]
        for c in _count_children(n):
            if id(c) in occ:
                occ[id(c)] += k
    return order, occ


def compression_benefit(length: int, count: int, var_len: int) -> int:
    cost = var_len + 3 + length + 3               # ` = `, and the line shift + two spaces of a definition
    if var_len > length:
        return 0
    main = (length - var_len) * count
    if cost > main:
        return 0
    return main - cost


class _Compressor:
    def __init__(self):
        self.F = Formatter()
        self._compressed: Dict[int, bool] = {}
        self._keep: List[Expr] = []

    def is_compressed(self, e: Expr) -> bool:
        hit = self._compressed.get(id(e))
        if hit is None:
            order, occ = count_terms(e)
            hit = True
            for t in order:
                c = occ[id(t)]
                if c > 1 and compression_benefit(self.F.length(t), c, 3) != 0:
                    hit = False
                    break
            self._compressed[id(e)] = hit
            self._keep.append(e)
        return hit

    def last_max_benefit(self, order: List[Expr], occ: Dict[int, int], min_count: int, var_len: int):
        best = None
        for t in order:
            c = occ[id(t)]
            if c < min_count:
                continue
            if not self.is_compressed(t):
                continue
            b = compression_benefit(self.F.length(t), c, var_len)
            if b == 0:
                continue
            if best is None or b >= best[1]:
                best = (t, b)
        return best


def rewrite(e: Expr, formulas: Dict[int, int]) -> Expr:
    """`Expr::rewrite`: every occurrence of a definition's formula becomes its variable (checked at a
    node before its operands are looked at); `Let` nodes are left alone.  Nodes without a formula below
    them are returned as they are."""
    out: Dict[int, Expr] = {}
    for n in E.dag_nodes([e]):
        vid = formulas.get(id(n))
        if vid is not None:
            out[id(n)] = E.var_id(vid)
        elif n.a is None or n.tag == E.LET:
            out[id(n)] = n
        else:
            a = out[id(n.a)]
            b = out[id(n.b)] if n.b is not None else None
            out[id(n)] = n if (a is n.a and b is n.b) else E._mk(n.tag, a, b, n.n)
    return out[id(e)]


def _local_nodes(root: Expr) -> List[Expr]:
    """Nodes below `root`, operands first, treating `Let` and `Var` as leaves."""
    out: List[Expr] = []
    seen = set()
    stack = [(root, False)]
    while stack:
        n, done = stack.pop()
        if done:
            out.append(n)
            continue
        if id(n) in seen:
            continue
        seen.add(id(n))
        stack.append((n, True))
        if n.tag in (E.LET, E.VAR):
            continue
        if n.b is not None:
            stack.append((n.b, False))
        if n.a is not None:
            stack.append((n.a, False))
    return out


def flatten(e: Expr, ctx: Optional[Dict[int, Expr]] = None) -> Expr:
    """`compressor::flatten`: every `Let` is dissolved, every variable replaced by its (flattened)
    definition, looked up in the context of the innermost `Let` only (the reference replaces, not
    extends, the context); `Decor` is dropped.  An unbound variable is an error (the reference panics)."""
    import sys
    memo: Dict[Tuple[int, int], Expr] = {}
    keep = []

    def go(root: Expr, env: Optional[Dict[int, Expr]]) -> Expr:
        key_env = id(env) if env is not None else 0
        if (id(root), key_env) in memo:
            return memo[(id(root), key_env)]
        for n in _local_nodes(root):
            k = (id(n), key_env)
            if k in memo:
                continue
            t = n.tag
            if t in (E.X, E.Y, E.TAU, E.E, E.NAT):
                memo[k] = n
            elif t == E.VAR:
                if env is None or n.n not in env:
                    raise ValueError("Could not find variable")
                memo[k] = go(env[n.n], env)
            elif t == E.LET:
                inner = {vid: d for vid, d in n.vars}
                keep.append(inner)                      # keeps id(inner) unique for the memo key
                memo[k] = go(n.a, inner)
            elif t == E.DECOR:
                memo[k] = memo[(id(n.a), key_env)]
            else:
                a = memo[(id(n.a), key_env)]
                b = memo[(id(n.b), key_env)] if n.b is not None else None
                memo[k] = E._mk(t, a, b, n.n)
        return memo[(id(root), key_env)]

    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 100000))             # one Python frame pair per level of variable nesting
    try:
        return go(e, ctx)
    finally:
        sys.setrecursionlimit(old)


def compress(e: Expr, log=None) -> Expr:
    """`Expr::compress`: flatten, then repeatedly name the sub-term whose extraction shortens the
    printed formula most (the last one among equals), until nothing pays for its own definition."""
    res = flatten(e, None)
    comp = _Compressor()
    ctx: List[Tuple[int, Expr]] = []
    formulas: Dict[int, int] = {}
    var_len = 2
    while True:
        order, occ = count_terms(res)
        pick = comp.last_max_benefit(order, occ, 2, var_len)
        if pick is None:
            break
        formula, benefit = pick
        vid = len(ctx)
        if log is not None:
            log(vid + 1, benefit, formula)
        var_len = len(f"a{vid}")
        ctx.append((vid, formula))
        formulas[id(formula)] = vid
        # Only the new formula can still occur: every earlier one was replaced when it was defined, and a
        # later rewrite cannot re-create it (it would have to contain a variable that did not exist yet).
        res = rewrite(res, {id(formula): vid})
    return E.let_(ctx, res) if ctx else res
