"""Parity of the CUDA path against the reference's own known-answer vectors
(generator.rs:1353-1925) and against the CPU oracle, through the C ABI.  Same harness shape as the
reference's `run_tests`: sample_rate 1, chunk sizes 1/2/4/8, output pre-filled with +inf."""
import numpy as np
import pytest

from tests.golden_cases import cases, length_cases, sine_cases

pytestmark = pytest.mark.gpu

CASES = cases()


def _gen(sample_rate=1):
    from tuun_b200.generator import Generator
    return Generator(sample_rate)


def run_chunks(g, w, expected, size):
    p = g.initialize_state(w)
    out = np.full(len(expected), np.inf, dtype=np.float32)
    for n in range(len(out) // size + 1):  # generator.rs:1294-1298
        end = min(len(out), (n + 1) * size)
        got = g.generate(p, out[n * size:end])
        assert got == end - n * size, f"chunk {n} of size {size}: generated {got}"
    return out


@pytest.mark.parametrize("name,w,expected", CASES, ids=[c[0] for c in CASES])
def test_run_tests(name, w, expected):
    g = _gen(1)
    assert g.length(g.initialize_state(w), len(expected)) == len(expected)  # check_length, :1290
    for size in (1, 2, 4, 8):
        out = run_chunks(g, w, expected, size)
        np.testing.assert_array_equal(out, expected, err_msg=f"{name} chunk {size}")


@pytest.mark.parametrize("name,w,expected", sine_cases(), ids=[c[0] for c in sine_cases()])
def test_sine(name, w, expected):
    g = _gen(44100)
    out = np.zeros(len(expected), dtype=np.float32)
    g.generate(g.initialize_state(w), out)
    assert np.max(np.abs(out - expected)) < 1e-5  # generator.rs:1487


@pytest.mark.parametrize("name,w,position,expected,max_", length_cases(), ids=[c[0] for c in length_cases()])
def test_check_length(name, w, position, expected, max_):
    g = _gen(1)
    p = g.initialize_state(w)
    g.generate(p, np.zeros(position, dtype=np.float32))
    assert g.length(p, max_) == expected


def test_fixed_exhausted():  # generator.rs:1364-1371
    from tuun_b200.waveform import Fixed
    g = _gen(1)
    p = g.initialize_state(Fixed([1, 2, 3, 4, 5]))
    out = np.zeros(6, dtype=np.float32)
    assert g.generate(p, out) == 5
    assert g.generate(p, out) == 0


def test_reinitialize_generates_same_samples():  # generator.rs:83-85
    from tuun_b200.waveform import Reset, Time
    from tests.golden_cases import sin_waveform
    g = _gen(1)
    p = g.initialize_state(Reset(sin_waveform(0.25, 0.0), Time()))
    a = np.zeros(8, dtype=np.float32)
    b = np.zeros(8, dtype=np.float32)
    g.generate(p, a)
    p.reset()
    g.generate(p, b)
    np.testing.assert_array_equal(a, b)
