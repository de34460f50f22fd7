"""Port of the reference optimizer, pinned by the five tree equalities of optimizer.rs:450-590 and by the
shapes SURVEY.md §8(a) derives for the named configs."""
import math

import numpy as np

from tuun_b200.optimizer import first_root, optimize
from tuun_b200.waveform import (Alt, Append, BinaryPointOp, Const, Fin, Fixed, Marked, Operator, Sine, Time, add,
                                div, f32, merge, mul, sub)


def sine1():
    return Sine(Const(1.0), Const(0.0))


def test_reference_equalities():  # optimizer.rs:450-590
    w1 = add(add(Const(1.0), add(Const(2.0), Const(3.0))), Const(4.0))
    assert optimize(w1) == Const(10.0)
    w2 = add(add(Const(2.0), add(Const(3.0), sine1())), Const(5.0))
    assert optimize(w2) == add(sine1(), Const(10.0))
    w3 = mul(mul(Const(2.0), mul(Const(3.0), sine1())), Const(5.0))
    assert optimize(w3) == mul(sine1(), Const(30.0))
    w4 = mul(add(Const(2.0), mul(Const(3.0), sine1())), Const(5.0))
    assert optimize(w4) == add(mul(sine1(), Const(15.0)), Const(10.0))
    w5 = mul(Fin(add(Time(), Const(-2.0)), Const(3.0)), Fin(add(Time(), Const(-1.5)), Const(5.0)))
    assert optimize(w5) == Fin(add(Time(), Const(-1.5)), Const(15.0))


def test_cfg1_shape():
    """`$440 * Qw`: the Fin is pulled out of the product and the `* 1` dropped (optimizer.rs:278,336-343)."""
    tau440 = f32(f32(2.0) * f32(3.14159265) * f32(440.0))
    qw = Fin(sub(Time(), Const(0.5)), Const(1.0))
    w = mul(Sine(Const(tau440), Const(0.0)), qw)
    assert optimize(w) == Fin(add(Time(), Const(-0.5)), Sine(Const(tau440), Const(0.0)))


def test_true_fm_shape():
    """fm-variations.tuunp:2 — commute, distribute, re-associate (optimizer.rs:148-152,282-311)."""
    two_pi = f32(np.float32(2.0) * np.float32(3.14159265))
    mod = Sine(Const(f32(np.float32(two_pi) * np.float32(220.0))), Const(f32(np.float32(3.14159265) / np.float32(2.0))))
    inner = add(Const(440.0), mul(Const(f32(np.float32(6.0) * np.float32(220.0))), mod))
    w = Sine(mul(Const(two_pi), inner), Const(0.0))
    got = optimize(w)
    assert isinstance(got, Sine) and got.phase == Const(0.0)
    f = got.frequency
    assert isinstance(f, BinaryPointOp) and f.op == Operator.Add and isinstance(f.b, Const)
    assert f.b.value == f32(np.float32(440.0) * np.float32(two_pi))
    assert f.a == mul(mod, Const(f32(np.float32(1320.0) * np.float32(two_pi))))


def test_misc_rules():
    e = Fixed([])
    assert optimize(Fin(Const(0.0), sine1())) == e
    assert optimize(Fin(Time(), sine1())) == e  # appendix A11
    assert optimize(Append(e, sine1())) == sine1()
    assert optimize(Append(Fixed([1.0]), Fixed([2.0, 3.0]))) == Fixed([1.0, 2.0, 3.0])
    assert optimize(Sine(Const(0.0), Const(1.0))) == Const(f32(np.sin(np.float32(1.0))))
    assert optimize(sub(sine1(), Const(-0.0))) == sine1()  # x - (-0) -> x + 0 -> x
    assert optimize(div(sine1(), Const(4.0))) == mul(sine1(), Const(0.25))
    assert optimize(Alt(Const(-1.0), Const(2.0), Const(3.0))) == Const(3.0)
    assert optimize(Alt(Const(0.0), Const(2.0), Const(3.0))) == Const(2.0)
    assert optimize(merge(e, sine1())) == sine1()
    assert optimize(merge(Time(), Const(0.0))) == Time()
    assert first_root(sub(Time(), Const(0.5))) == Const(0.5)
    assert first_root(Const(1.0)) is None
    # w | fin(t) | seq(t): Merge(Fin, Append(Fin, c)) with equal roots folds into the Append's head
    note = Fin(add(Time(), Const(-1.0)), sine1())
    rest = Append(Fin(add(Time(), Const(-1.0)), Const(0.0)), Fixed([7.0]))
    got = optimize(merge(note, rest))
    assert got == Append(Fin(add(Time(), Const(-1.0)), merge(sine1(), Const(0.0))), Fixed([7.0]))
    got = optimize(merge(Marked(3, note), rest))
    assert got == Append(Marked(3, Fin(add(Time(), Const(-1.0)), merge(sine1(), Const(0.0)))), Fixed([7.0]))
    # nested Fin keeps the shorter
    assert optimize(Fin(sub(Time(), Const(2.0)), Fin(sub(Time(), Const(1.0)), sine1()))) == \
        Fin(add(Time(), Const(-1.0)), sine1())
