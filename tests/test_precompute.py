"""Generator::precompute (generator.rs:868-1229): the decision table on CPU with the oracle as the
renderer, and the reference's `run_tests` third pass (precomputed trees reproduce the vectors)."""
import numpy as np
import pytest

from oracle.binding import OracleProgram
from tests.golden_cases import cases
from tuun_b200.optimizer import optimize
from tuun_b200.precompute import precompute
from tuun_b200.waveform import (Alt, Append, BinaryPointOp, Const, Filter, Fin, Fixed, Marked, Operator, Reset, Sine,
                                Time, add, mul, sub)


def oracle_render(sample_rate):
    return lambda w, n: OracleProgram(w, sample_rate).render(n, block=1024)


def pre(w, sr=1):
    return precompute(w, sr, render=oracle_render(sr))


def test_decision_table():
    fin4 = Fin(sub(Time(), Const(4.0)), Time())
    # finite and static: baked whole (test_append :1614-1621 expects Fixed)
    assert pre(Append(Fixed([1, 1, 1]), Fixed([2, 2, 2]))) == Fixed([1, 1, 1, 2, 2, 2])
    assert pre(fin4) == Fixed([0, 1, 2, 3])
    # infinite stays
    assert pre(Time()) == Time() and pre(add(Time(), Const(1.0))) == add(Time(), Const(1.0))
    # Multiply / Divide by an infinite operand is finite (min length) and bakes; Add does not lengthen either
    assert pre(mul(fin4, Const(2.0))) == Fixed([0, 2, 4, 6])
    assert pre(mul(Const(2.0), fin4)) == Fixed([0, 2, 4, 6])
    # Add with an infinite side: the finite side is baked, the node stays (generator.rs:1104-1111)
    assert pre(add(fin4, Time())) == add(Fixed([0, 1, 2, 3]), Time())
    # dynamic: Marked blocks its ancestors, its static children are still baked
    got = pre(mul(Marked(7, fin4), Const(2.0)))
    assert got == mul(Marked(7, Fixed([0, 1, 2, 3])), Const(2.0))
    got = pre(Fin(sub(Time(), Marked(1, Const(4.0))), Const(1.0)))
    assert isinstance(got, Fin) and isinstance(got.length.b, Marked)
    # Filter: Const coefficients count as infinite operands (generator.rs:1128-1151), so the node
    # stays and its finite input is baked; finite coefficient waveforms are baked too
    assert pre(Filter(fin4, [Const(2.0), Const(2.0)], [])) == Filter(Fixed([0, 1, 2, 3]), [Const(2.0), Const(2.0)], [])
    assert pre(Filter(fin4, [Fixed([2.0] * 4), Fixed([2.0] * 4)], [])) == Fixed([2, 6, 10, 6])
    f = pre(Filter(Time(), [Const(1.0), Fin(sub(Time(), Const(2.0)), Const(3.0))], []))
    assert f == Filter(Time(), [Const(1.0), Fixed([3, 3])], [])
    # Alt with two static branches and an infinite trigger
    a = pre(Alt(Sine(Const(1.0), Const(0.0)), fin4, Fixed([9, 9])))
    assert a == Alt(Sine(Const(1.0), Const(0.0)), Fixed([0, 1, 2, 3]), Fixed([9, 9]))


def test_ten_second_cap():
    # an infinite waveform under a Fin that never ends is cut at sample_rate * 10 (generator.rs:917)
    got = pre(Fin(Const(-1.0), Const(0.5)), sr=100)
    assert isinstance(got, Fixed) and len(got.samples) == 1000


@pytest.mark.parametrize("name,w,expected", cases(), ids=[c[0] for c in cases()])
def test_run_tests_precomputed_pass(name, w, expected):
    # generator.rs:1326-1350: precompute(optimize(w)) must reproduce the same vector, chunks 1/2/4/8
    p = pre(optimize(w))
    for size in (1, 2, 4, 8):
        o = OracleProgram(p, 1)
        got = o.render(len(expected), block=size) if len(expected) else np.zeros(0, np.float32)
        np.testing.assert_array_equal(got[:len(expected)], expected, err_msg=f"{name} chunk {size}: {p}")


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,expected", cases(), ids=[c[0] for c in cases()])
def test_gpu_run_tests_optimized_and_precomputed(name, w, expected):
    """The reference's run_tests passes 2 and 3 on the device: optimized, then precomputed ON the
    GPU (tb_render bakes the Fixed buffers), both reproduce the known-answer vectors exactly."""
    from tuun_b200.generator import Generator
    g = Generator(1)
    opt = optimize(w)
    baked = precompute(opt, 1)
    for tree in (opt, baked):
        assert g.length(g.initialize_state(tree), len(expected)) == len(expected)
        for size in (1, 2, 4, 8):
            p = g.initialize_state(tree)
            out = np.full(len(expected), np.inf, dtype=np.float32)
            for n in range(len(out) // size + 1):
                end = min(len(out), (n + 1) * size)
                assert g.generate(p, out[n * size:end]) == end - n * size
            np.testing.assert_array_equal(out, expected, err_msg=f"{name} chunk {size}: {tree}")


@pytest.mark.gpu
def test_gpu_precompute_harmonica_note():
    """A whole harmonica note is finite and static: precompute bakes it to one Fixed of 22,050
    samples, within tolerance of the oracle's render of the original tree."""
    from tuun_b200 import workloads as W
    from tuun_b200.builder import Std, to_waveform
    s = Std()
    note = optimize(to_waveform(s.harmonica(s.Q, 440)))
    baked = precompute(note, 44100)
    assert isinstance(baked, Fixed) and len(baked.samples) == 22050
    ref = OracleProgram(note, 44100).render(30000, block=1024)
    assert np.max(np.abs(baked.samples - ref)) <= 1e-4
