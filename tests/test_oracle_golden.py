"""Pins the CPU oracle (oracle/maray_oracle.c) to everything the reference holds for this path:
the known-answer assertions of `it_works` (reference src/lib.rs:1241-1285), the shipped scene
data/chess.maray (legacy layout, exact EOF) and the shipped render images/chess.png (approximate
golden: SURVEY.md F4)."""
import hashlib
import os

import numpy as np
import pytest
from PIL import Image

from maray_b200 import expr as E
from oracle.oracle import OracleScene

from conftest import GOLDEN


def _eval(e, x, legacy=False):
    return OracleScene(E.to_bytes([4, 4], [e, e, e], legacy=legacy)).eval(0, x)


X = E.x()
KNOWN = [  # (expr, x, expected) -- reference src/lib.rs:1243-1284
    (E.mul(X, X), 2.0, 4.0),
    (E.neg(E.nat(1)), 0.0, -1.0),
    (E.div(E.nat(1), E.nat(2)), 0.0, 0.5),
    (E.pi(), 0.0, 3.141592653589793),
    (E.lerp(E.neg(E.nat(1)), E.nat(1), X), 0.0, -1.0),
    (E.lerp(E.neg(E.nat(1)), E.nat(1), X), 1.0, 1.0),
    (E.cos(X), 0.0, 1.0),
    (E.step(X), -1.0, 0.0), (E.step(X), 0.0, 1.0), (E.step(X), 1.0, 1.0),
    (E.step_at(E.nat(2), X), 1.0, 0.0), (E.step_at(E.nat(2), X), 2.0, 1.0),
    (E.range(E.nat(1), E.nat(2), X), 0.5, 0.0), (E.range(E.nat(1), E.nat(2), X), 1.5, 1.0),
    (E.range(E.nat(1), E.nat(2), X), 2.5, 0.0),
    (E.p2_len([X, X]), 0.0, 0.0), (E.p2_len([X, X]), 1.0, 2.0 ** 0.5),
    (E.clamp(E.nat(1), E.nat(5), X), 0.0, 1.0), (E.clamp(E.nat(1), E.nat(5), X), 1.0, 1.0),
    (E.clamp(E.nat(1), E.nat(5), X), 5.0, 5.0), (E.clamp(E.nat(1), E.nat(5), X), 6.0, 5.0),
]


@pytest.mark.parametrize("legacy", [False, True])
def test_it_works_known_answers(legacy):
    for e, x, want in KNOWN:
        assert _eval(e, x, legacy) == want


def test_step_min_max_cast_corner_cases():
    # Step: NaN -> 0, -0.0 -> 1 (reference src/lib.rs:644-647)
    nan = E.add(E.recip(E.nat(0)), E.neg(E.recip(E.nat(0))))       # inf - inf
    negzero = E.neg(E.nat(0))
    assert np.isnan(_eval(nan, 0.0))
    assert _eval(E.step(nan), 0.0) == 0.0
    assert _eval(E.step(negzero), 0.0) == 1.0
    # f64::max/min ignore NaN (reference src/lib.rs:655-658)
    assert _eval(E.max(nan, E.nat(3)), 0.0) == 3.0 and _eval(E.max(E.nat(3), nan), 0.0) == 3.0
    assert _eval(E.min(nan, E.nat(3)), 0.0) == 3.0 and _eval(E.min(E.nat(3), nan), 0.0) == 3.0
    # +0/-0 tie returns the first operand (x86-64 lowering; DESIGN.md "Semantics"): observable through 1/x
    assert _eval(E.recip(E.max(negzero, E.nat(0))), 0.0) == -np.inf
    assert _eval(E.recip(E.max(E.nat(0), negzero)), 0.0) == np.inf
    assert _eval(E.recip(E.min(E.nat(0), negzero)), 0.0) == np.inf
    # `as u8` (reference src/render.rs:26-28): saturating, truncating, NaN -> 0
    cases = [(E.nat(300), 255), (E.neg(E.nat(5)), 0), (nan, 0), (E.div(E.nat(511), E.nat(2)), 255),
             (E.div(E.nat(509), E.nat(2)), 254), (E.div(E.nat(1), E.nat(2)), 0), (E.recip(E.nat(0)), 255)]
    for e, want in cases:
        img = OracleScene(E.to_bytes([2, 2], [e, e, e])).render()
        assert (img == want).all()


def test_unbound_variable_is_nan_in_reference_semantics():
    # Cache::val falls through to NaN (reference src/cache.rs:40).  A file with a stray Var is
    # rejected by the layout check, so bind the name in an outer Let that the inner Let hides:
    # `Let` REPLACES the context (reference src/lib.rs:659-662), the inner body cannot see $7.
    e = E.let_([(0, E.nat(5))], E.let_([(1, E.nat(1))], E.var_id(0)))
    assert np.isnan(_eval(e, 0.0))
    # ... unless the cache already holds it (the cache is keyed by name only, reference src/cache.rs:29)
    e2 = E.let_([(0, E.nat(5))], E.add(E.var_id(0), E.let_([(1, E.nat(1))], E.var_id(0))))
    assert _eval(e2, 0.0) == 10.0
    # var_fixer renames the binding but not a reference hidden behind an inner Let
    # (reference src/var_fixer.rs:49-66): with non-canonical ids even that is lost.
    e3 = E.let_([(7, E.nat(5))], E.add(E.var_id(7), E.let_([(8, E.nat(1))], E.var_id(7))))
    assert np.isnan(_eval(e3, 0.0))


def test_chess_maray_layout_and_size(chess_bytes):
    assert hashlib.sha256(chess_bytes).hexdigest().startswith("b1ad82f4")
    s = OracleScene(chess_bytes)
    assert s.legacy and s.size == (1024, 1024)
    assert [s.tree_nodes(c) for c in range(3)] == [29314] * 3


def test_chess_rows_against_reference_png_and_golden(chess_bytes):
    """images/chess.png is an approximate golden (rendered from another revision of the expression):
    it differs from a faithful f64 evaluation only on rows 512 and 704 (100 + 56 pixels).
    tests/golden/chess_oracle_1024.png is this oracle's own full render (sha256 4d2ca7dd..., the anchor
    an independent numpy evaluator produced during the survey)."""
    ref_png = np.array(Image.open(os.path.join(GOLDEN, "chess_reference.png")).convert("RGB"))
    gold = np.array(Image.open(os.path.join(GOLDEN, "chess_oracle_1024.png")).convert("RGB"))
    assert hashlib.sha256(gold.tobytes()).hexdigest() == "4d2ca7dd8c0b5f48922ae7d20363f966f4faba6691d17218f38ca56852743bba"
    assert hashlib.sha256(ref_png.tobytes()).hexdigest() == "b6f0efcf4632279bfa6846224b75469cfaff971af8a910975437f7dd732ccac2"
    d = (gold != ref_png).any(axis=2)
    ys, counts = np.unique(np.nonzero(d)[0], return_counts=True)
    assert ys.tolist() == [512, 704] and counts.tolist() == [100, 56]

    rows = [0, 300, 511, 512, 513, 600, 703, 704, 705, 818, 819, 1023]
    got = OracleScene(chess_bytes).render_rows(rows)
    for i, y in enumerate(rows):
        assert np.array_equal(got[i], gold[y]), f"row {y} differs from the committed oracle render"
        if y not in (512, 704):
            assert np.array_equal(got[i], ref_png[y]), f"row {y} differs from images/chess.png"


def test_var_fixer_reference_assertions():
    """The reference's own `test_var_fixer` (src/lib.rs:1508-1691), run against the oracle's restatement of
    `var_fixer::fix_color` (src/var_fixer.rs:25-82): the scene is opened (which fixes it), written back, and
    compared with the tree the reference expects."""
    from maray_b200.expr import add, let_, nat, sub, to_bytes, var_id, x, y

    def fixed(color):
        return OracleScene(to_bytes([1, 1], color)).fixed_bytes()

    def nest(i0, d0, i1, d1, ref):
        return let_([(i0, d0)], let_([(i1, d1)], var_id(ref)))

    # one channel three times: {0: x} {0: y} $0  ->  {0: x} {1: y} $1 on every channel (:1520-1571)
    a = nest(0, x(), 0, y(), 0)
    b = nest(0, x(), 1, y(), 1)
    assert fixed([a, a, a]) == to_bytes([1, 1], [b, b, b])
    # inner definitions differ per channel: fresh ids 1, 2, 3; the shared outer x keeps id 0 (:1573-1631)
    a2 = [nest(0, x(), 0, add(y(), nat(k)), 0) for k in (1, 2, 3)]
    b2 = [nest(0, x(), k, add(y(), nat(k)), k) for k in (1, 2, 3)]
    assert fixed(a2) == to_bytes([1, 1], b2)
    # both levels differ per channel: ids are handed out in visiting order 0..5 (:1633-1690)
    a3 = [nest(0, sub(x(), nat(k)), 0, add(y(), nat(k)), 0) for k in (1, 2, 3)]
    b3 = [nest(2 * k - 2, sub(x(), nat(k)), 2 * k - 1, add(y(), nat(k)), 2 * k - 1) for k in (1, 2, 3)]
    assert fixed(a3) == to_bytes([1, 1], b3)
    # the canonical compress shape (one Let, ids 0..n, the same on every channel) is a fixed point
    from maray_b200 import scenes
    raw = scenes.chess_1k()
    size, color, _legacy = E.from_bytes(raw)
    assert OracleScene(raw).fixed_bytes() == to_bytes(size, color)


def test_jit_standin_cpu_baseline_equals_the_oracle():
    """bench.py's second CPU baseline (oracle/jit_standin.py: the generated straight-line program built
    for the host, the stand-in for the reference's WASM JIT) must render what the interpreter restatement
    renders: same libm, same operations -> identical bytes, on the shipped scene, an SDF scene and a
    textured one."""
    from maray_b200 import scenes
    from oracle.jit_standin import JitStandIn

    tex = scenes.synthetic_textures(2, 64)
    x, y = E.x(), E.y()
    textured = E.to_bytes([96, 40], [E.app(E.channel(0, 0), x, y), E.app(E.channel(1, 2), E.sub(x, E.nat(20)), y),
                                     E.mul(E.app(E.image_width(1), x, y), E.nat(2))])
    for scene, textures, w, rows in ((scenes.chess_1k(), [], 1024, [0, 512, 704, 1023]),
                                     (scenes.sdf(320, 200, 12, seed=3), [], 320, [0, 99, 199]),
                                     (textured, tex, 96, [0, 39])):
        js = JitStandIn(scene, textures)
        got = js.render_rows(rows, w, threads=2)
        want = OracleScene(scene, textures).render_rows(rows, w)
        assert np.array_equal(got, want)
