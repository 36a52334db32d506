"""`Expr::compress` / `compressor::flatten` / `Display for Expr` restated (maray_b200/compress.py).

The reference has no unit test for its compressor; the pin is the shipped scene: dissolving the `Let`
of data/chess.maray (flatten) and compressing the result again must give back the file, byte for byte --
the same 859 definitions in the same order with the same formulas.  Every choice the compressor makes
depends on the printed length of terms (Display), on the first-seen order of terms, on tree occurrence
counts and on the "last of the best" tie rule, so this exercises all of it 859 times."""
import pytest

from maray_b200 import compress as C
from maray_b200 import expr as E
from maray_b200 import scenes
from maray_b200.expr import add, div, let_, max as emax, min as emin, mul, nat, neg, recip, sin, sqrt, step, sub, tau, var_id, x, y


def test_display_forms():
    """`impl fmt::Display for Expr` (reference src/lib.rs:196-367): parenthesisation rules."""
    f = C.fmt
    assert f(sub(div(y(), nat(1024)), div(nat(61), nat(80)))) == "y/1024-61/80"
    assert f(sub(div(mul(nat(6), sub(div(x(), nat(1024)), div(nat(59), nat(80)))), nat(160)),
                 div(mul(nat(2), var_id(0)), nat(160)))) == "(6*(x/1024-59/80))/160-(2*$0)/160"
    assert f(neg(add(x(), nat(1)))) == "-(x+1)" and f(neg(x())) == "-x"
    assert f(recip(add(x(), y()))) == "1/(x+y)" and f(recip(nat(3))) == "1/3"
    assert f(mul(add(x(), nat(1)), add(x(), nat(1)))) == "(x+1)^2" and f(mul(x(), x())) == "x^2"
    assert f(add(mul(x(), y()), nat(1))) == "(x*y)+1"                 # a product is parenthesised in a sum ...
    assert f(sub(mul(x(), y()), nat(1))) == "x*y-1"                   # ... but not as the minuend of a difference
    assert f(add(mul(x(), x()), div(y(), nat(2)))) == "x^2+y/2"       # squares and quotients are printed bare
    assert f(add(sub(x(), y()), nat(1))) == "x-y+1" and f(sub(x(), sub(y(), nat(1)))) == "x-(y-1)"
    assert f(step(sin(mul(mul(nat(4), tau()), var_id(246))))) == "step(sin((4*τ)*$246))"
    assert f(emax(emin(x(), y()), sqrt(x()))) == "max(min(x,y),sqrt(x))"
    assert f(let_([(0, add(x(), nat(1)))], mul(var_id(0), var_id(0)))) == "$0^2\nwhere\n  $0 = x+1\n"


def test_is_simple_expr_and_benefit():
    # reference src/compressor.rs:108-119
    assert C.is_simple_expr(x()) and C.is_simple_expr(recip(nat(3))) and C.is_simple_expr(mul(x(), recip(y())))
    assert not C.is_simple_expr(recip(mul(x(), y()))) and not C.is_simple_expr(mul(mul(x(), y()), x()))
    assert not C.is_simple_expr(add(x(), y())) and not C.is_simple_expr(neg(x()))
    # reference src/compressor.rs:154-164: len 10 seen 3 times with a 2-character name
    assert C.compression_benefit(10, 3, 2) == (10 - 2) * 3 - (2 + 3 + 10 + 3)
    assert C.compression_benefit(10, 2, 2) == 0 and C.compression_benefit(1, 50, 2) == 0


def test_small_compress_and_flatten_round_trip():
    t = add(mul(x(), nat(3)), sqrt(add(y(), nat(7))))                  # "(x*3)+sqrt(y+7)": long enough to pay
    e = emax(mul(t, t), add(t, emin(t, sin(t))))
    out = C.compress(e)
    assert out.tag == E.LET and out.vars == ((0, t),)
    assert out.a is emax(mul(var_id(0), var_id(0)), add(var_id(0), emin(var_id(0), sin(var_id(0)))))
    assert C.flatten(out) is e
    assert C.compress(add(x(), y())) is add(x(), y())                  # nothing repeats: no Let
    with pytest.raises(ValueError):
        C.flatten(add(var_id(3), x()))                                 # the reference panics: "Could not find variable"
    # the context of a Let REPLACES the outer one (reference src/compressor.rs:198)
    with pytest.raises(ValueError):
        C.flatten(let_([(0, x())], let_([(1, y())], add(var_id(0), var_id(1)))))


def test_compress_reproduces_the_shipped_scene():
    raw = scenes.chess_1k()
    size, color, legacy = E.from_bytes(raw)
    assert legacy and size == [1024, 1024]
    shipped = color[0]
    assert shipped.tag == E.MUL and shipped.a.tag == E.LET and shipped.b is nat(255)
    assert color[1] is shipped and color[2] is shipped
    flat = C.flatten(shipped.a)
    assert len(C.fmt(flat)) == 303059                                  # the formula examples/chess.rs would print, uncompressed
    again = C.compress(flat)
    assert again.tag == E.LET and len(again.vars) == 859
    assert again is shipped.a                                          # hash-consed: same definitions, same body
    channel = mul(again, nat(255))
    assert E.to_bytes(size, [channel, channel, channel], legacy=True) == raw
