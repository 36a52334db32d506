"""Scene authoring (maray_b200/expr.py): wire format round trips, both layouts, builder shapes."""
import numpy as np
import pytest

from maray_b200 import expr as E
from maray_b200 import scenes
from oracle.oracle import OracleScene


def test_chess_roundtrip_both_layouts(chess_bytes):
    size, color, legacy = E.from_bytes(chess_bytes)
    assert legacy and size == [1024, 1024]
    assert E.tree_size(color[0]) == 29314
    assert E.to_bytes(size, color, legacy=True) == chess_bytes
    head = E.to_bytes(size, color)
    s2, c2, l2 = E.from_bytes(head)
    assert not l2 and s2 == size and all(a is b for a, b in zip(color, c2))   # hash-consing: identical objects


def test_builder_shapes_match_reference_definitions():
    x, y = E.x(), E.y()
    # sub = add(a, neg(b)); div = mul(a, recip(b))  (reference src/lib.rs:934-949)
    s = E.sub(x, y)
    assert s.tag == E.ADD and s.b.tag == E.NEG and s.b.a is y
    d = E.div(x, y)
    assert d.tag == E.MUL and d.b.tag == E.RECIP
    # cos(a) = sin(a + tau/4)  (reference src/lib.rs:921)
    c = E.cos(x)
    assert c.tag == E.SIN and c.a.tag == E.ADD and c.a.b is E.rad_90()
    # texture ids (reference src/textures.rs:14-23)
    assert (E.channel(2, 1), E.image_width(2), E.image_height(2)) == (11, 13, 14)
    # operator sugar lifts integers to Nat (reference src/lib.rs:151-194)
    assert (x * 3) is E.mul(x, E.nat(3)) and (x - 1) is E.sub(x, E.nat(1)) and (-x) is E.neg(x)


def test_share_let_preserves_values():
    b = scenes.deep(64, 64, n_values=400, seed=9)
    size, color, _ = E.from_bytes(b)
    assert all(c.tag == E.LET for c in color)
    assert color[0].vars == color[1].vars == color[2].vars          # canonical shape
    ids = [i for i, _ in color[0].vars]
    assert ids == list(range(len(ids)))
    # inline the Let again in Python and compare oracle values of both forms at a few points
    flat = []
    for c in color:
        env = {}
        for i, d in c.vars:
            env[i] = _inline(d, env)
        flat.append(_inline(c.a, env))
    a, bb = OracleScene(b), OracleScene(E.to_bytes(size, flat))
    for (px, py) in [(0, 0), (13, 57), (63, 1)]:
        for ch in range(3):
            assert a.eval(ch, px, py) == bb.eval(ch, px, py)


def _inline(e, env):
    memo = {}
    for n in E.dag_nodes([e]):
        if n.tag == E.VAR: memo[id(n)] = env[n.n]
        elif n.a is None: memo[id(n)] = n
        else: memo[id(n)] = E._mk(n.tag, memo[id(n.a)], memo[id(n.b)] if n.b is not None else None, n.n)
    return memo[id(e)]


def test_scene_generators_are_deterministic_and_sized():
    assert scenes.sdf(64, 48, 6) == scenes.sdf(64, 48, 6)
    assert scenes.deep(32, 32, 300) == scenes.deep(32, 32, 300)
    assert E.from_bytes(scenes.sdf())[0] == [1920, 1080]
    assert E.from_bytes(scenes.textured())[0] == [3840, 2160]
    tex = scenes.synthetic_textures(2, 64)
    assert tex[1][5, 9, 2] == (9 * 7 + 5 * 13 + 2 * 31 + 101 + ((9 ^ 5) & 0xFF)) & 0xFF


def test_ambiguous_legacy_bytes_are_resolved_by_binding_check():
    # mul(nat 1, recip(nat 2)) in the legacy numbering also decodes under HEAD numbering (as
    # add(var 1, abs(var 2))) -- but then its variables are unbound, so the legacy reading wins.
    e = E.div(E.nat(1), E.nat(2))
    b = E.to_bytes([1, 1], [e, e, e], legacy=True)
    _, color, legacy = E.from_bytes(b)
    assert legacy and color[0] is e
    assert OracleScene(b).legacy
