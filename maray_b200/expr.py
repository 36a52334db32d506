"""Scene authoring for tests and benchmarks: a mirror of Maray's `Expr` builder and wire format.

This is host-side content tooling (it builds `.maray` byte strings); nothing here evaluates an
expression.  Names, argument order and the shape of every generated tree follow the reference's
builder functions so that scenes written against `maray::*` read the same here:

  * `Expr` variants and their order      -- reference src/lib.rs:101-149
  * builder functions (`x`, `nat`, `step_at`, `lerp`, `chess`, `p2_*`, barycentrics ...)
                                          -- reference src/lib.rs:837-1151
  * `subst2` / `translate` / `scale` / `rotate`  -- reference src/lib.rs:709-735, 798-826
  * `save` / `open` wire format (bincode 1.3.3)  -- reference src/lib.rs:1216-1235, SURVEY.md App. A
  * texture function ids                  -- reference src/textures.rs:14-23

Nodes are hash-consed (structurally equal sub-trees are one Python object), which keeps big scenes
small in memory; on the wire sharing is expressed with one top-level `Let` per channel
(`share_let`), the canonical shape `Expr::compress` emits (reference src/compressor.rs:215-236).
"""
from __future__ import annotations

import builtins as _b
import struct
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

# Variant order of the reference enum (HEAD layout).  Legacy files have no `Arc` and every tag is
# one lower.
ARC, X, Y, TAU, E, VAR, NAT, NEG, ABS, RECIP, SQRT, STEP, SIN, EXP, LN, ADD, MUL, MAX, MIN, LET, DECOR, APP = range(22)
TAG_NAMES = ["Arc", "X", "Y", "Tau", "E", "Var", "Nat", "Neg", "Abs", "Recip", "Sqrt", "Step", "Sin",
             "Exp", "Ln", "Add", "Mul", "Max", "Min", "Let", "Decor", "App"]
UNARY = (NEG, ABS, RECIP, SQRT, STEP, SIN, EXP, LN)
BINARY = (ADD, MUL, MAX, MIN)

_INTERN: Dict[tuple, "Expr"] = {}
_range = _b.range   # this module defines its own `range`, `min`, `max`, `abs` (reference names)


class Expr:
    """One node.  `a`/`b` are operands, `n` is the Var id / Nat value / App id, `vars` the Let context."""

    __slots__ = ("tag", "a", "b", "n", "vars", "seq", "__weakref__")
    _count = 0

    def __init__(self, tag: int, a: Optional["Expr"] = None, b: Optional["Expr"] = None, n: int = 0,
                 vars: Optional[Tuple[Tuple[int, "Expr"], ...]] = None):
        self.tag, self.a, self.b, self.n, self.vars = tag, a, b, n, vars
        Expr._count += 1
        self.seq = Expr._count          # creation order: operands always have a smaller seq

    # operator sugar, as the reference's impl Add/Sub/Mul/Div/Neg (src/lib.rs:151-194)
    def __add__(self, o): return add(self, _lift(o))
    def __sub__(self, o): return sub(self, _lift(o))
    def __mul__(self, o): return mul(self, _lift(o))
    def __truediv__(self, o): return div(self, _lift(o))
    def __neg__(self): return neg(self)

    def subst2(self, p: Sequence["Expr"]) -> "Expr":
        """Substitute X and Y (reference src/lib.rs:709-735; `Let` and `Var` are left untouched)."""
        memo: Dict[int, Expr] = {}
        return _subst2(self, p[0], p[1], memo)

    def translate(self, off): return self.subst2(p2_sub([x(), y()], off))
    def scale(self, s): return self.subst2(p2_div([x(), y()], s))
    def scale_at(self, off, s): return self.translate(p2_neg(off)).scale(s).translate(off)

    def rotate(self, rad):
        s, c = sin(rad), cos(rad)
        ident = [x(), y()]
        return self.subst2([p2_dot([c, neg(s)], ident), p2_dot([s, c], ident)])

    def rotate_at(self, off, rad): return self.translate(p2_neg(off)).rotate(rad).translate(off)

    def __repr__(self):
        return f"<Expr {TAG_NAMES[self.tag]}>"


def _mk(tag, a=None, b=None, n=0, vars=None) -> Expr:
    key = (tag, id(a), id(b), n, None if vars is None else tuple((i, id(e)) for i, e in vars))
    e = _INTERN.get(key)
    if e is None:
        e = Expr(tag, a, b, n, vars)
        _INTERN[key] = e
    return e


def clear_intern_pool() -> None:
    """Drop the hash-cons table (scenes built afterwards no longer share nodes with earlier ones)."""
    _INTERN.clear()


def _lift(o) -> Expr:
    return nat(o) if isinstance(o, int) else o


def _subst2(e: Expr, px: Expr, py: Expr, memo: Dict[int, Expr]) -> Expr:
    # iterative post-order so deep scenes do not hit the Python recursion limit
    stack = [(e, False)]
    while stack:
        node, done = stack.pop()
        if id(node) in memo:
            continue
        t = node.tag
        if t == X: memo[id(node)] = px; continue
        if t == Y: memo[id(node)] = py; continue
        if t in (TAU, E, VAR, NAT, LET): memo[id(node)] = node; continue
        if not done:
            stack.append((node, True))
            stack.append((node.a, False))
            if node.b is not None: stack.append((node.b, False))
            continue
        a = memo[id(node.a)]
        b = memo[id(node.b)] if node.b is not None else None
        memo[id(node)] = _mk(t, a, b, node.n)
    return memo[id(e)]


# ---- leaves and operators (reference src/lib.rs:837-957) ------------------------------------------
def app(id: int, a: Expr, b: Expr) -> Expr: return _mk(APP, a, b, id)
def x() -> Expr: return _mk(X)
def y() -> Expr: return _mk(Y)
def var_id(id: int) -> Expr: return _mk(VAR, n=id)
def tau() -> Expr: return _mk(TAU)
def pi() -> Expr: return div(tau(), nat(2))
def rad_45() -> Expr: return div(tau(), nat(8))
def rad_90() -> Expr: return div(tau(), nat(4))
def e() -> Expr: return _mk(E)
def nat(a: int) -> Expr:
    assert 0 <= a < 1 << 64
    return _mk(NAT, n=a)
def half() -> Expr: return div(nat(1), nat(2))
def neg(a): return _mk(NEG, a)
def abs(a): return _mk(ABS, a)  # noqa: A001 - mirrors the reference name
def recip(a): return _mk(RECIP, a)
def sqrt(a): return _mk(SQRT, a)
def step(a): return _mk(STEP, a)
def step_at(a, x_): return step(sub(x_, a))
def step_pos(a): return set_inv(step(neg(a)))
def step_pos_at(a, x_): return step_pos(sub(x_, a))
def pos(cond, a, b): return lerp(b, a, step_pos(cond))
def range(a, b, x_): return mul(step_at(a, x_), set_inv(step_at(b, x_)))  # noqa: A001
def range_incl(a, b, x_): return mul(step_at(a, x_), set_inv(step_pos_at(b, x_)))
def clamp(a, b, x_): return pos(sub(x_, a), pos(sub(x_, b), b, x_), a)
def clamp_unit(x_): return clamp(nat(0), nat(1), x_)
def clamp_u8(x_): return clamp(nat(0), nat(255), x_)
def ge(a, b): return step(sub(a, b))
def gt(a, b): return step_pos(sub(a, b))
def le(a, b): return set_inv(gt(a, b))
def lt(a, b): return set_inv(ge(a, b))
def eq(a, b): return set_and(ge(a, b), le(a, b))
def set_inv(a): return sub(nat(1), a)
def set_and(a, b): return min(a, b)
def set_or(a, b): return max(a, b)
def set_xor(a, b): return set_or(set_and(a, set_inv(b)), set_and(b, set_inv(a)))
def sin(a): return _mk(SIN, a)
def cos(a): return sin(add(a, rad_90()))
def exp(a): return _mk(EXP, a)
def ln(a): return _mk(LN, a)
def max(a, b): return _mk(MAX, a, b)  # noqa: A001
def min(a, b): return _mk(MIN, a, b)  # noqa: A001
def add(a, b): return _mk(ADD, a, b)
def sub(a, b): return add(a, neg(b))
def mul(a, b): return _mk(MUL, a, b)
def div(a, b): return mul(a, recip(b))
def square(a): return mul(a, a)
def lerp(a, b, t): return add(a, mul(sub(b, a), t))
def unit_to_rad(a): return mul(a, tau())
def rad_to_unit(a): return div(a, tau())


def let_(vars: Iterable[Tuple[int, Expr]], body: Expr) -> Expr:
    return _mk(LET, body, None, 0, tuple(vars))


def decor(a: Expr) -> Expr:
    """Decor with an empty token list (transparent; reference src/lib.rs:663)."""
    return _mk(DECOR, a)


def chess(n: int) -> Expr:
    """Reference src/lib.rs:969-973."""
    sx = step(sin(mul(mul(div(nat(n), nat(2)), tau()), x())))
    sy = step(sin(mul(mul(div(nat(n), nat(2)), tau()), y())))
    return set_xor(sx, sy)


def set_unit_square(f: Expr) -> Expr:
    """Reference src/lib.rs:975-980."""
    return set_and(set_and(range(nat(0), nat(1), x()), range(nat(0), nat(1), y())), f)


# ---- 2-D vector helpers (reference src/lib.rs:982-1060) ------------------------------------------
def p2_pos(cond, a, b): return p2_lerp(b, a, step_pos(cond))
def p2_abs(a): return [abs(a[0]), abs(a[1])]
def p2_neg(a): return [neg(a[0]), neg(a[1])]
def p2_add(a, b): return [add(a[0], b[0]), add(a[1], b[1])]
def p2_sub(a, b): return [sub(a[0], b[0]), sub(a[1], b[1])]
def p2_mul(a, b): return [mul(a[0], b[0]), mul(a[1], b[1])]
def p2_div(a, b): return [div(a[0], b[0]), div(a[1], b[1])]
def p2_max(a, b): return [max(a[0], b[0]), max(a[1], b[1])]
def p2_scale(a, b): return p2_mul(a, [b, b])
def p2_circle(ang): return [cos(ang), sin(ang)]
def p2_spiral(ang): return p2_scale(p2_circle(ang), rad_to_unit(ang))
def p2_dot(a, b): return add(mul(a[0], b[0]), mul(a[1], b[1]))
def p2_len(a): return sqrt(p2_dot(a, a))
def p2_lerp(a, b, t): return [lerp(a[0], b[0], t), lerp(a[1], b[1], t)]
def p2_qbez(a, b, c, t): return p2_lerp(p2_lerp(a, b, t), p2_lerp(b, c, t), t)
def p2_cbez(a, b, c, d, t): return p2_lerp(p2_qbez(a, b, c, t), p2_qbez(b, c, d, t), t)
def p2_subst(p, off): return [p[0].subst2(off), p[1].subst2(off)]
def p4_same(v): return [v, v, v, v]
def p4_xy(p): return [p[0], p[1]]
def p4_zw(p): return [p[2], p[3]]


def quad_to_tri(quad, uv):
    """Reference src/lib.rs:1073-1080."""
    q0, q1, q2, q3 = quad
    uv0, uv1, uv2, uv3 = uv
    return [([q0, q1, q2], [uv0, uv1, uv2]), ([q1, q2, q3], [uv1, uv2, uv3])]


def quad_pos(quad, uv):
    """Reference src/lib.rs:1083-1091."""
    q0, q1, q2, q3 = quad
    return p2_lerp(p2_lerp(q0, q1, uv[0]), p2_lerp(q2, q3, uv[0]), uv[1])


def to_barycentric(triangle, pos_):
    """Reference src/lib.rs:1109-1124."""
    px, py = pos_
    (x1, y1), (x2, y2), (x3, y3) = triangle
    den = add(mul(sub(y2, y3), sub(x1, x3)), mul(sub(x3, x2), sub(y1, y3)))
    l1 = div(add(mul(sub(y2, y3), sub(px, x3)), mul(sub(x3, x2), sub(py, y3))), den)
    l2 = div(add(mul(sub(y3, y1), sub(px, x3)), mul(sub(x1, x3), sub(py, y3))), den)
    l3 = sub(sub(nat(1), l1), l2)
    return [l1, l2, l3]


def from_barycentric(triangle, lam):
    (x1, y1), (x2, y2), (x3, y3) = triangle
    return [add(add(mul(lam[0], x1), mul(lam[1], x2)), mul(lam[2], x3)),
            add(add(mul(lam[0], y1), mul(lam[1], y2)), mul(lam[2], y3))]


def inside_triangle(triangle, pos_):
    """Reference src/lib.rs:1094-1097."""
    b0, b1, b2 = to_barycentric(triangle, pos_)
    return set_and(set_and(step(b0), step(b1)), step(b2))


def to_uv(triangle, uv, pos_):
    """Reference src/lib.rs:1100-1106."""
    b0, b1, b2 = to_barycentric(triangle, pos_)
    return p2_add(p2_add(p2_scale(uv[0], b0), p2_scale(uv[1], b1)), p2_scale(uv[2], b2))


# ---- signed distance functions (reference src/sd.rs:27-59) ---------------------------------------
def sd_circle(r: Expr) -> Expr:
    return sub(p2_len([x(), y()]), r)


def sd_box(b) -> Expr:
    d = p2_sub(p2_abs([x(), y()]), b)
    return add(p2_len(p2_max(d, [nat(0), nat(0)])), min(max(d[0], d[1]), nat(0)))


def sd_rounded_box(b, r) -> Expr:
    rxy = p2_pos(x(), p4_xy(r), p4_zw(r))
    rx = pos(y(), rxy[0], rxy[1])
    q = p2_add(p2_sub(p2_abs([x(), y()]), b), [rx, rx])
    return sub(add(min(max(q[0], q[1]), nat(0)), p2_len(p2_max(q, [nat(0), nat(0)]))), rx)


def sd_inside(sd: Expr) -> Expr: return step(neg(sd))
def sd_outside(sd: Expr) -> Expr: return step(sd)


def grid_cell(size, pos_, quad):
    """`Grid2(size).cell(pos, quad)` -> (positions, uvs).  Reference src/grid.rs:10-32."""
    w, h = nat(size[0]), nat(size[1])
    fx, fy = div(nat(pos_[0]), w), div(nat(pos_[1]), h)
    gx, gy = div(nat(pos_[0] + 1), w), div(nat(pos_[1] + 1), h)
    uv0, uv1, uv2, uv3 = [fx, fy], [gx, fy], [fx, gy], [gx, gy]
    return ([quad_pos(quad, uv0), quad_pos(quad, uv1), quad_pos(quad, uv2), quad_pos(quad, uv3)],
            [uv0, uv1, uv2, uv3])


# ---- texture function ids (reference src/textures.rs:14-23) --------------------------------------
TEXTURE_ALIGN = 5
def channel(img: int, ch: int) -> int: return img * TEXTURE_ALIGN + ch
def image_width(img: int) -> int: return img * TEXTURE_ALIGN + 3
def image_height(img: int) -> int: return img * TEXTURE_ALIGN + 4


# ---- traversal helpers ----------------------------------------------------------------------------
def _children(e: Expr):
    if e.tag == LET:
        for _, d in e.vars:
            yield d
    if e.a is not None:
        yield e.a
    if e.b is not None:
        yield e.b


def dag_nodes(roots: Sequence[Expr]) -> List[Expr]:
    """All distinct nodes reachable from `roots`, children before parents."""
    out: List[Expr] = []
    seen = set()
    stack = [(r, False) for r in reversed(list(roots))]
    while stack:
        node, done = stack.pop()
        if done:
            out.append(node)
            continue
        if id(node) in seen:
            continue
        seen.add(id(node))
        stack.append((node, True))
        for c in reversed(list(_children(node))):
            if id(c) not in seen:
                stack.append((c, False))
    return out


def tree_size(e: Expr) -> int:
    """Node count of the expression written out as a tree (what the wire format stores)."""
    size: Dict[int, int] = {}
    for n in dag_nodes([e]):
        size[id(n)] = 1 + sum(size[id(c)] for c in _children(n))
    return size[id(e)]


def share_let(color: Sequence[Expr], min_tree_size: int = 3, bind: Sequence[Expr] = ()) -> List[Expr]:
    """Express DAG sharing on the wire: one `Let` with ids 0..n (definitions in dependency order,
    definition i may use $j for j<i) placed on top of every channel, bodies rewritten to `Var`s.

    This is the canonical shape `Expr::compress` produces (reference src/compressor.rs:215-236) and
    the shape for which var_fixer is the identity: all three channels carry the SAME
    context.  The choice of which sub-terms become variables is ours (every node used more than once
    whose tree has at least `min_tree_size` nodes, plus every node listed in `bind` -- an author may
    name any intermediate result); values are unchanged by construction.  Channels must be Let-free.
    """
    # definitions follow the order in which the scene's author built the values (operands first)
    order = sorted(dag_nodes(color), key=lambda n: n.seq)
    refs: Dict[int, int] = {}
    for n in order:
        assert n.tag != LET, "share_let expects Let-free channels"
        for c in _children(n):
            refs[id(c)] = refs.get(id(c), 0) + 1
    for r in color:
        refs[id(r)] = refs.get(id(r), 0) + 1
    forced = {id(n) for n in bind}
    tsize: Dict[int, int] = {}
    new: Dict[int, Expr] = {}          # node -> rewritten node (shared nodes replaced by Var)
    defs: List[Tuple[int, Expr]] = []
    for n in order:
        kids = list(_children(n))
        tsize[id(n)] = 1 + sum(tsize[id(c)] for c in kids)
        if n.tag in (X, Y, TAU, E, NAT, VAR):
            rew = n
        else:
            rew = _mk(n.tag, new[id(n.a)] if n.a is not None else None,
                      new[id(n.b)] if n.b is not None else None, n.n)
        shared = refs.get(id(n), 0) > 1 and tsize[id(n)] >= min_tree_size
        if (shared or id(n) in forced) and n.tag not in (X, Y, TAU, E, NAT, VAR):
            vid = len(defs)
            defs.append((vid, rew))
            rew = var_id(vid)
        new[id(n)] = rew
    if not defs:
        return list(color)
    return [let_(defs, new[id(c)]) for c in color]


# ---- wire format: maray::save / maray::open (reference src/lib.rs:1216-1235) ---------------------
def to_bytes(size: Sequence[int], color: Sequence[Expr], legacy: bool = False) -> bytes:
    """bincode(([u32;2],[Expr;3])).  `legacy=True` writes the pre-`Arc` numbering of data/chess.maray."""
    assert len(color) == 3
    out = bytearray(struct.pack("<II", size[0], size[1]))
    shift = 1 if legacy else 0
    pu32, pu64 = struct.Struct("<I").pack, struct.Struct("<Q").pack
    for root in color:
        stack: list = [root]
        while stack:
            it = stack.pop()
            if isinstance(it, bytes):
                out += it
                continue
            t = it.tag
            if legacy and t == ARC:
                stack.append(it.a)
                continue
            out += pu32(t - shift)
            if t in (VAR, NAT):
                out += pu64(it.n)
            elif t == APP:
                out += pu32(it.n)
                stack.append(it.b); stack.append(it.a)
            elif t == LET:
                out += pu64(len(it.vars))
                stack.append(it.a)
                for vid, d in reversed(it.vars):
                    stack.append(d)
                    stack.append(pu64(vid))
            elif t == DECOR:
                stack.append(pu64(0))          # empty Vec<Token>
                stack.append(it.a)
            else:
                if it.b is not None: stack.append(it.b)
                if it.a is not None: stack.append(it.a)
    return bytes(out)


class WireError(ValueError):
    pass


def _parse(buf: bytes, legacy: bool):
    u32, u64 = struct.Struct("<I").unpack_from, struct.Struct("<Q").unpack_from
    pos_ = 8
    n = len(buf)
    if n < 8:
        raise WireError("short file")
    size = list(struct.unpack_from("<II", buf, 0))
    shift = 1 if legacy else 0
    color = []
    for _ in _range(3):
        # explicit stacks: `work` holds pending actions, `vals` finished sub-expressions
        work: list = [("expr",)]
        vals: list = []
        while work:
            act = work.pop()
            kind = act[0]
            if kind == "expr":
                if pos_ + 4 > n: raise WireError("eof")
                t = u32(buf, pos_)[0] + shift; pos_ += 4
                if t >= 22 or (legacy and t == ARC): raise WireError(f"bad variant {t - shift}")
                if t in (X, Y, TAU, E):
                    vals.append(_mk(t))
                elif t in (VAR, NAT):
                    if pos_ + 8 > n: raise WireError("eof")
                    vals.append(_mk(t, n=u64(buf, pos_)[0])); pos_ += 8
                elif t == ARC:
                    work.append(("expr",))           # transparent
                elif t in UNARY:
                    work.append(("un", t)); work.append(("expr",))
                elif t in BINARY:
                    work.append(("bin", t, 0)); work.append(("expr",)); work.append(("expr",))
                elif t == APP:
                    if pos_ + 4 > n: raise WireError("eof")
                    fid = u32(buf, pos_)[0]; pos_ += 4
                    work.append(("bin", APP, fid)); work.append(("expr",)); work.append(("expr",))
                elif t == LET:
                    if pos_ + 8 > n: raise WireError("eof")
                    cnt = u64(buf, pos_)[0]; pos_ += 8
                    if cnt > (n - pos_) // 12: raise WireError("bad let length")
                    work.append(("let", cnt))
                    work.append(("expr",))           # body (runs after all definitions)
                    for _i in _range(cnt):
                        work.append(("expr",)); work.append(("id",))
                elif t == DECOR:
                    work.append(("decor",)); work.append(("expr",))
            elif kind == "id":
                if pos_ + 8 > n: raise WireError("eof")
                vals.append(u64(buf, pos_)[0]); pos_ += 8
            elif kind == "un":
                a = vals.pop(); vals.append(_mk(act[1], a))
            elif kind == "bin":
                b = vals.pop(); a = vals.pop(); vals.append(_mk(act[1], a, b, act[2]))
            elif kind == "let":
                body = vals.pop()
                cnt = act[1]
                flat = vals[len(vals) - 2 * cnt:] if cnt else []
                if cnt: del vals[len(vals) - 2 * cnt:]
                vs = tuple((flat[2 * i], flat[2 * i + 1]) for i in _range(cnt))
                vals.append(let_(vs, body))
            elif kind == "decor":
                if pos_ + 8 > n: raise WireError("eof")
                ntok = u64(buf, pos_)[0]; pos_ += 8
                if ntok: raise WireError("Decor tokens are not supported by the Python reader")
                a = vals.pop(); vals.append(_mk(DECOR, a))
        color.append(vals.pop())
    if pos_ != n:
        raise WireError("trailing bytes")
    if not _vars_bound(color):
        raise WireError("unbound variable")
    return size, color


def _vars_bound(color) -> bool:
    """Every Var is bound by an enclosing Let (layout plausibility check, see from_bytes)."""
    for root in color:
        stack = [(root, frozenset())]
        seen = set()
        while stack:
            node, bound = stack.pop()
            if (id(node), bound) in seen: continue
            seen.add((id(node), bound))
            if node.tag == VAR:
                if node.n not in bound: return False
            elif node.tag == LET:
                inner = bound | frozenset(i for i, _ in node.vars)
                for _, d in node.vars: stack.append((d, inner))
                stack.append((node.a, inner))
            else:
                if node.a is not None: stack.append((node.a, bound))
                if node.b is not None: stack.append((node.b, bound))
    return True



def from_bytes(buf: bytes):
    """Returns (size, color, legacy).  Tries HEAD numbering, then legacy; the accepted layout is the
    one that consumes the whole buffer with every variable bound (a legacy file can
    decode by accident under HEAD numbering, but then Nat reads as Var and is unbound)."""
    try:
        s, c = _parse(buf, False)
        return s, c, False
    except WireError:
        s, c = _parse(buf, True)
        return s, c, True


def save(file: str, data) -> None:
    size, color = data
    with open(file, "wb") as f:
        f.write(to_bytes(size, color))


def open_(file: str):
    with open(file, "rb") as f:
        size, color, _ = from_bytes(f.read())
    return size, color


def map_xy(color: Sequence[Expr], px: Expr, py: Expr) -> List[Expr]:
    """Replace X and Y everywhere, INCLUDING inside `Let` definitions (unlike `subst2`, which stops at
    a `Let`).  Used to rescale a stored scene to another resolution."""
    memo: Dict[int, Expr] = {}

    def go(root: Expr) -> Expr:
        for n in dag_nodes([root]):
            if id(n) in memo: continue
            t = n.tag
            if t == X: memo[id(n)] = px
            elif t == Y: memo[id(n)] = py
            elif t in (TAU, E, VAR, NAT): memo[id(n)] = n
            elif t == LET:
                memo[id(n)] = let_(tuple((i, memo[id(d)]) for i, d in n.vars), memo[id(n.a)])
            else:
                memo[id(n)] = _mk(t, memo[id(n.a)], memo[id(n.b)] if n.b is not None else None, n.n)
        return memo[id(root)]

    return [go(c) for c in color]
