"""A root sequence — what `<[a, b, c]>`, `a \\ b` and Player::beats_waveform evaluate to: a tree of Appends over Fins
of analytic length (builtins.rs:208-299, optimizer.rs:212-229) — is held as one program per part (lower.h
sequence_parts, abi.cpp launch_sequence): the second arm of an Append starts from Initial state
(generator.rs:169-188), so every part is a stream of its own that begins at a known sample, and the parts a call
overlaps render side by side.  Same samples as the tree rendered as one program and as the oracle, for any
blocking of the stream."""
import numpy as np
import pytest

from oracle.binding import OracleProgram
from tuun_b200.waveform import Append, BinaryPointOp, Const, Fin, Marked, Noise, Operator, Sine, Time, add, f32, mul

pytestmark = pytest.mark.gpu
SR = 44100


def gpu(w, n, blocks=None, params=None, voices=1, seed=None, env=None, monkeypatch=None, length_at=None):
    from tuun_b200.generator import Program
    for k, v in (env or {}).items():
        monkeypatch.setenv(k, v)
    p = Program(w, SR)
    if seed is not None:
        p.seed_noise(seed)
    out = np.full((voices, n), np.inf, dtype=np.float32)
    done = np.zeros(voices, dtype=np.int64)
    a = 0
    for c in (blocks or [n]):
        c = min(c, n - a)
        if c <= 0:
            break
        if length_at is not None and a == length_at[0]:      # Generator::length in the middle of the stream
            lens = p.lengths(voices, length_at[1], params=params)
            done += np.minimum(lens.astype(np.int64), length_at[1])
            out[:, a:a + length_at[1]] = np.nan
            a += length_at[1]
            continue
        blk = np.full((voices, c), np.inf, dtype=np.float32)
        lens = p.render(blk, params=params)
        out[:, a:a + c] = blk
        done += lens.astype(np.int64)
        a += c
    info = p.info
    for k in (env or {}):
        monkeypatch.delenv(k)
    return out, done, info


def oracle_rows(w, n, params=None, voices=1, seed=None):
    o = OracleProgram(w, SR)
    rows = np.zeros((voices, n), dtype=np.float32)
    lens = np.zeros(voices, dtype=np.int64)
    for v in range(voices):
        o.initialize_state()
        if seed is not None:
            o.seed_noise(seed, v)
        if params is not None:
            o.set_params(params[v])
        r = o.render(n)
        rows[v, :len(r)] = r
        lens[v] = len(r)
    return rows, lens


def close(got, ref, lens, tol=1e-4):
    for v in range(len(lens)):
        assert np.abs(got[v, :lens[v]] - ref[v, :lens[v]]).max() <= tol, v


def test_harmonica_sequence_part_by_part(monkeypatch):
    from tuun_b200 import workloads as W
    w = W.cfg2_harmonica(4)
    n = 100000
    ref, rl = oracle_rows(w, n)
    assert rl[0] == 88200
    one, l1, i1 = gpu(w, n, env={"TUUN_B200_SEQ": "0"}, monkeypatch=monkeypatch)
    assert i1.sequence_parts == 0 and l1[0] == 88200
    got, l2, i2 = gpu(w, n)
    assert i2.sequence_parts == 4 and i2.sequence_renders == 1 and l2[0] == 88200
    close(got, ref, rl)
    assert np.abs(got[0, :88200] - one[0, :88200]).max() <= 1e-5
    # any blocking of the stream: boundaries inside blocks, on block ends, one sample either side
    for blocks in ([1024] * 100, [22049, 2, 22050, 44099, 5000, 5000], [22050] * 5, [7, 50000, 50000]):
        g, l, _ = gpu(w, n, blocks=blocks)
        assert l[0] == 88200, blocks
        assert np.abs(g[0, :88200] - got[0, :88200]).max() <= 1e-5, blocks


def test_left_nested_appends_under_marks(monkeypatch):
    """benches/tracker_benches.rs:92-117 (marks_4_40): `marks = Append(marks, one)` — the chain nests to the LEFT and
    every `one` is a Marked sequence of its own; the leaves are what counts."""
    from tuun_b200 import workloads as W
    name, w, blocks = W.tracker_benches()[3]
    assert name == "marks_4_40"
    n = 12 * 22050 + 300
    ref, rl = oracle_rows(w, n)
    got, l, info = gpu(w, n)
    assert info.sequence_parts == 160 and l[0] == rl[0] == n
    close(got, ref, rl, tol=0.0)


def test_noise_streams_keep_their_numbers(monkeypatch):
    """A Noise node's stream is numbered by its node in the WHOLE tree (tb_seed_noise): a part keeps those numbers."""
    note = lambda sec, g: Fin(add(Time(), Const(-f32(sec))), BinaryPointOp(Operator.Merge, mul(Noise(), Const(g)), Const(0.0)))
    w = Append(note(0.05, 0.5), Append(note(0.03, 0.25), mul(Noise(), Const(0.125))))
    n = 6000
    ref, rl = oracle_rows(w, n, seed=1234)
    got, l, info = gpu(w, n, seed=1234)
    assert info.sequence_parts == 3 and l[0] == n
    np.testing.assert_array_equal(got[0], ref[0])
    g2, l2, _ = gpu(w, n, seed=1234, blocks=[2000, 205, 1323, 3000])
    np.testing.assert_array_equal(g2[0], ref[0])


def test_batch_with_per_voice_notes_and_a_finite_tail(monkeypatch):
    """Per-voice pitch (parameters), voice-independent lengths, a last part that ENDS (out_len short), host rows cut in
    time by a small staging buffer."""
    tone = lambda sec: Fin(add(Time(), Const(-f32(sec))), BinaryPointOp(Operator.Merge, Sine(Const(1.0, param=0), Const(0.0)), Const(0.0)))
    w = Append(tone(0.25), Append(tone(0.125), Fin(add(Time(), Const(-f32(0.2))), Sine(Const(1.0, param=1), Const(0.0)))))
    V, n = 37, 30000
    rng = np.random.default_rng(3)
    params = (2 * np.pi * rng.uniform(100, 2000, (V, 2))).astype(np.float32)
    ref, rl = oracle_rows(w, n, params=params, voices=V)
    assert (rl == 11025 + 5513 + 8820).all()
    got, l, info = gpu(w, n, params=params, voices=V)
    assert info.sequence_parts == 3 and (l == rl).all()
    close(got, ref, rl, tol=2e-6)
    g2, l2, _ = gpu(w, n, params=params, voices=V, blocks=[11000, 30, 5600, 13370],
                    env={"TUUN_B200_STAGE_MB": "1"}, monkeypatch=monkeypatch)
    assert (l2 == rl).all()
    close(g2, ref, rl, tol=2e-6)


def test_length_in_the_middle_of_a_sequence(monkeypatch):
    """Generator::length(Append(a, b), max) advances a, then b by what is left (generator.rs:705-722): part by part.  The
    oracle makes the same three calls (length() moves positions, not phases: what follows is what the reference gives)."""
    tone = lambda sec, f: Fin(add(Time(), Const(-f32(sec))), BinaryPointOp(Operator.Merge, Sine(Const(f32(f)), Const(0.0)), Const(0.0)))
    w = Append(tone(0.1, 3000.0), Append(tone(0.1, 5000.0), Sine(Const(f32(7000.0)), Const(0.0))))
    o = OracleProgram(w, SR)
    r1 = o.render(3000)
    assert o.length(4000) == 4000       # crosses the first boundary (4410), not the second (8820)
    r2 = o.render(5000)
    got, l, info = gpu(w, 12000, blocks=[3000, 4000, 5000], length_at=(3000, 4000))
    assert info.sequence_parts == 3 and l[0] == 12000
    assert np.abs(got[0, :3000] - r1).max() <= 2e-6
    assert np.abs(got[0, 7000:] - r2).max() <= 2e-6
    # the same calls on the tree as ONE program
    one, l1, i1 = gpu(w, 12000, blocks=[3000, 4000, 5000], length_at=(3000, 4000), env={"TUUN_B200_SEQ": "0"}, monkeypatch=monkeypatch)
    assert i1.sequence_parts == 0 and np.abs(one[0, 7000:] - r2).max() <= 2e-6


def test_a_tune_of_three_hundred_notes(monkeypatch):
    """As one program such a sequence needs a control stack deeper than the interpreter has (about 120 notes:
    TB_ERR_UNSUPPORTED); part by part it is 300 small programs."""
    from tuun_b200._abi import TB_ERR_UNSUPPORTED, TuunB200Error
    from tuun_b200.generator import Program, lower_check
    rng = np.random.default_rng(9)
    notes = [Fin(add(Time(), Const(-f32(0.01))), BinaryPointOp(Operator.Merge, Sine(Const(f32(2 * np.pi * f)), Const(0.0)), Const(0.0)))
             for f in rng.uniform(200, 2000, 300)]
    w = notes[-1]
    for x in reversed(notes[:-1]):
        w = Append(x, w)
    assert lower_check(w).sequence_parts == 300
    n = 300 * 441
    ref, rl = oracle_rows(w, n + 50)
    assert rl[0] == n
    got, l, info = gpu(w, n + 50)
    assert info.sequence_parts == 300 and l[0] == n
    close(got, ref, rl, tol=2e-6)
    monkeypatch.setenv("TUUN_B200_SEQ", "0")
    with pytest.raises(TuunB200Error) as e:
        Program(w, SR)
    assert e.value.status == TB_ERR_UNSUPPORTED
