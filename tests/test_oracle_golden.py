"""Pins the CPU oracle against the reference's own known-answer vectors
(generator.rs:1353-1925), the way `run_tests` does: chunk sizes 1/2/4/8, buffers pre-filled
with +inf, `length()` agreeing with the generated length."""
import numpy as np
import pytest

from oracle.binding import OracleProgram
from tests.golden_cases import cases, length_cases, sine_cases
from tuun_b200.waveform import (Append, BinaryPointOp, Const, Fin, Marked, Operator, Time)

CASES = cases()


def run_chunks(prog, expected, size):
    out = np.full(len(expected), np.inf, dtype=np.float32)
    for n in range(len(out) // size + 1):  # generator.rs:1294-1298
        end = min(len(out), (n + 1) * size)
        got = prog.generate(out[n * size:end])
        assert got == end - n * size
    return out


@pytest.mark.parametrize("name,w,expected", CASES, ids=[c[0] for c in CASES])
def test_run_tests(name, w, expected):
    # check_length(g, waveform, 0, expected.len(), expected.len())  generator.rs:1290
    assert OracleProgram(w, 1).length(len(expected)) == len(expected)
    for size in (1, 2, 4, 8):
        out = run_chunks(OracleProgram(w, 1), expected, size)
        np.testing.assert_array_equal(out, expected, err_msg=f"{name} chunk {size}")


@pytest.mark.parametrize("name,w,expected", sine_cases(), ids=[c[0] for c in sine_cases()])
def test_sine(name, w, expected):
    out = np.zeros(len(expected), dtype=np.float32)
    OracleProgram(w, 44100).generate(out)
    assert np.max(np.abs(out - expected)) < 1e-5  # generator.rs:1487


@pytest.mark.parametrize("name,w,position,expected,max_", length_cases(), ids=[c[0] for c in length_cases()])
def test_check_length(name, w, position, expected, max_):
    p = OracleProgram(w, 1)
    p.generate(np.zeros(position, dtype=np.float32))
    assert p.length(max_) == expected


def test_fixed_exhausted():  # generator.rs:1364-1371
    from tuun_b200.waveform import Fixed
    p = OracleProgram(Fixed([1, 2, 3, 4, 5]), 1)
    out = np.zeros(6, dtype=np.float32)
    assert p.generate(out) == 5
    assert p.generate(out) == 0


def test_fin_substitute():  # generator.rs:1398-1463
    w = Append(Fin(BinaryPointOp(Operator.Subtract, Time(), Marked(7, Const(2.0))), Const(1.0)), Const(0.5))
    p = OracleProgram(w, 1)
    out = np.zeros(12, dtype=np.float32)
    assert p.generate(out[:6]) == 6
    np.testing.assert_array_equal(out[:6], [1, 1, .5, .5, .5, .5])
    assert p.substitute_const(7, 8.0) == 1
    assert p.generate(out[6:]) == 6
    np.testing.assert_array_equal(out, [1, 1] + [.5] * 10)

    w = Append(Fin(BinaryPointOp(Operator.Subtract, Time(), Marked(7, Const(3.0))), Time()), Const(0.5))
    p = OracleProgram(w, 1)
    out = np.zeros(12, dtype=np.float32)
    assert p.generate(out[:6]) == 6
    np.testing.assert_array_equal(out[:6], [0, 1, 2, .5, .5, .5])
    p.substitute_const(7, 9.0)
    assert p.generate(out[6:]) == 6
    np.testing.assert_array_equal(out, [0, 1, 2] + [.5] * 9)


def test_greater_or_equals_at():  # generator.rs:1906-1925
    w2 = Fin(BinaryPointOp(Operator.Add, Time(), Const(-5.0)), Time())
    out = np.zeros(10, dtype=np.float32)
    n = OracleProgram(w2, 1).generate(out)
    assert n == 5
    np.testing.assert_array_equal(out[:5], np.arange(5, dtype=np.float32))
