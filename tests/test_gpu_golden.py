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


def test_fin_substitute_mid_stream():  # generator.rs:1398-1463, both scenarios, on the device
    """`waveform::substitute` between two generate calls: Fin has advanced BOTH of its children to the end of
    the block (generator.rs:141-167), so a new length 'picks up where it would have been' — and a Fin that has
    already ended (the first part of an Append) is not revisited."""
    from tuun_b200.waveform import Append, BinaryPointOp, Const, Fin, Marked, Operator, Time
    g = _gen(1)
    w = Append(Fin(BinaryPointOp(Operator.Subtract, Time(), Marked(7, Const(2.0))), Const(1.0)), Const(0.5))
    p = g.initialize_state(w)
    out = np.zeros(12, dtype=np.float32)
    assert g.generate(p, out[:6]) == 6
    np.testing.assert_array_equal(out[:6], [1, 1, .5, .5, .5, .5])
    assert p.substitute(7, 8.0) == 1
    assert p.substitute(99, 1.0) == 0          # "zero or more parts" (waveform.rs:393-395)
    assert g.generate(p, out[6:]) == 6
    np.testing.assert_array_equal(out, [1, 1] + [.5] * 10)

    w = Append(Fin(BinaryPointOp(Operator.Subtract, Time(), Marked(7, Const(3.0))), Time()), Const(0.5))
    p = g.initialize_state(w)
    out = np.zeros(12, dtype=np.float32)
    assert g.generate(p, out[:6]) == 6
    np.testing.assert_array_equal(out[:6], [0, 1, 2, .5, .5, .5])
    assert p.substitute(7, 9.0) == 1
    assert g.generate(p, out[6:]) == 6
    np.testing.assert_array_equal(out, [0, 1, 2] + [.5] * 9)


def test_substitute_lengthens_a_note_that_is_still_sounding():
    """The case the two reference scenarios bracket: the Fin is still open when its length is replaced — the
    Time child of the length kept counting, so the note ends where the NEW length says (oracle: same calls)."""
    from oracle.binding import OracleProgram
    from tuun_b200.waveform import Append, BinaryPointOp, Const, Fin, Marked, Operator, Sine, Time
    w = Append(Fin(BinaryPointOp(Operator.Subtract, Time(), Marked(3, Const(6.0))), Sine(Const(0.7), Const(0.0))),
               Const(0.25))
    g = _gen(1)
    p = g.initialize_state(w)
    o = OracleProgram(w, 1)
    got = np.zeros(16, dtype=np.float32)
    ref = np.zeros(16, dtype=np.float32)
    assert g.generate(p, got[:4]) == 4 and o.generate(ref[:4]) == 4
    assert p.substitute(3, 9.0) == 1 and o.substitute_const(3, 9.0) == 1
    assert g.generate(p, got[4:]) == 12 and o.generate(ref[4:]) == 12
    assert np.all(ref[9:] == 0.25) and ref[8] != 0.25      # the note now lasts 9 samples
    np.testing.assert_allclose(got, ref, atol=1e-6)
    np.testing.assert_array_equal(got[9:], ref[9:])


def test_substitute_a_slider_in_a_batch():
    """player.rs:110 — slider values are Marked constants replaced by constants; here mid-stream on a batch of
    parameter-swept voices, against the oracle doing the same per voice."""
    from oracle.binding import OracleProgram
    from tuun_b200.generator import Program
    from tuun_b200.waveform import BinaryPointOp, Const, Marked, Operator, Sine
    w = BinaryPointOp(Operator.Multiply, Sine(Const(1.0, param=0), Const(0.0)), Marked(5, Const(0.5)))
    params = (2 * np.pi * np.array([[220.0], [330.0], [440.0]])).astype(np.float32)
    p = Program(w, 44100)
    a = np.zeros((3, 1000), dtype=np.float32)
    b = np.zeros((3, 1000), dtype=np.float32)
    p.render(a, params=params)
    assert p.substitute(5, 0.125) == 1
    p.render(b, params=params)
    for v in range(3):
        o = OracleProgram(w, 44100)
        o.set_params(params[v])
        ra = o.render(1000)
        o.substitute_const(5, 0.125)
        rb = o.render(1000)
        assert np.abs(a[v] - ra).max() <= 1e-6 and np.abs(b[v] - rb).max() <= 1e-6
    assert np.abs(b).max() <= 0.125 + 1e-6 < np.abs(a).max()


def test_substitute_refuses_what_it_cannot_do():
    from tuun_b200._abi import TB_ERR_UNSUPPORTED, TuunB200Error
    from tuun_b200.waveform import BinaryPointOp, Const, Marked, Operator, Time
    g = _gen(1)
    p = g.initialize_state(BinaryPointOp(Operator.Add, Marked(1, Time()), Const(1.0)))
    with pytest.raises(TuunB200Error) as e:
        p.substitute(1, 2.0)
    assert e.value.status == TB_ERR_UNSUPPORTED
