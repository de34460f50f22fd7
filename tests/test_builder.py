"""The expression-value layer (tuun_b200/builder.py mirrors builtins.rs + lib/v0/std.tuun) and
the named workloads built with it: tree shapes against the hand-derived forms of SURVEY 8(a),
and the oracle's lengths for them (segment boundaries are integer facts)."""
import numpy as np

from oracle.binding import OracleProgram
from tuun_b200 import workloads as W
from tuun_b200.builder import Seq, Std, followed_by, sequence, times, to_waveform
from tuun_b200.optimizer import optimize
from tuun_b200.waveform import (Alt, Append, BinaryPointOp, Const, Filter, Fin, Operator, Reset, Sine, Time, add,
                                flatten, mul)

SR = 44100
F = np.float32


def test_cfg1_matches_hand_derived_tree():
    # `$440 * Qw`: the optimizer pulls the Fin out and drops the `* 1` (optimizer.rs:278,336-343)
    assert W.cfg1_from_source() == W.cfg1_sine()
    assert W.cfg1_from_source() == Fin(add(Time(), Const(-0.5)), Sine(Const(F(2) * F(3.14159265) * F(440)), Const(0.0)))


def test_true_fm_and_pm_shapes():
    progs = dict(W.cfg3_fm_variations())
    assert len(progs) == 12
    # commute, distribute, re-associate (optimizer.rs:148-152,282-311)
    assert progs["true-fm"] == Sine(add(mul(Sine(Const(1382.3008), Const(F(3.14159265) / F(2))), Const(8293.805)),
                                        Const(2764.6016)), Const(0.0))
    assert progs["pm"] == Sine(Const(2764.6016), mul(Sine(Const(1382.3008), Const(0.0)), Const(6.0)))
    assert isinstance(progs["square-fm"].frequency.a.a, Alt)
    assert isinstance(progs["pulse-pm"].phase.a.trigger.a.a, Reset)  # pulse = alt(sawtooth - w), sawtooth = reset


def test_lpf_coefficients_are_f32_scalars():
    s = Std()
    w = s.lpf(0.5, 1900)(Time())
    assert isinstance(w, Filter)
    got = [c.value for c in w.feed_forward + w.feedback]
    # lpf(0.5, 1900) @ 44.1 kHz
    np.testing.assert_allclose(got, [0.014366774, 0.028733548, 0.014366774, -1.5205542, 0.5780213], rtol=2e-7)
    b0, b1, b2, a1, a2 = W.lpf_coefficients(0.5, 1900)
    assert got == [float(b0), float(b1), float(b2), float(a1), float(a2)]


def test_seq_offsets_add_up():
    # `a \\ b` with two seqs: offsets add through first_root (builtins.rs:179-206)
    s = Std()
    q = s.note(s.Q)
    both = followed_by(q, q)
    assert isinstance(both, Seq)
    assert optimize(both.offset) == add(Time(), Const(-1.0))
    w = optimize(to_waveform(sequence([q, q, q, q])))
    assert OracleProgram(w, SR).render(SR * 4).shape[0] == 4 * 22050


def test_harmonica_note_lengths():
    # harmonica(Q, 440) at tempo 120: one note is 22,050 samples; the envelope pieces are
    # a = 0.13 s, d = r = 0.33 s, s = 0; the 4-note sequence is 88,200 samples.
    s = Std()
    note = optimize(to_waveform(s.harmonica(s.Q, 440)))
    assert isinstance(note, Fin) and note.length == add(Time(), Const(-0.5))
    env = note.waveform.b
    assert isinstance(env, Append) and env.a.length == add(Time(), Const(-F(0.13)))
    o = OracleProgram(note, SR)
    assert o.length(10 ** 9) == 22050
    seq4 = W.cfg2_harmonica(4)
    out = OracleProgram(seq4, SR).render(200000, block=1024)
    assert out.shape[0] == 88200
    assert np.isfinite(out).all() and 0.05 < float(np.max(np.abs(out))) < 2.0
    # every note repeats the first one exactly (fresh state per Append arm, generator.rs:169-188)
    np.testing.assert_array_equal(out[:22050], out[22050:44100])


def test_cfg4_sources():
    progs = dict(W.cfg4_filters(noise_seconds=0.5))
    assert set(progs) == {"square220-lpf", "noise-lpf", "square-cascade", "pulse-filter_4_3"}
    n = OracleProgram(progs["noise-lpf"], SR).render(SR)
    assert n.shape[0] == SR // 2  # a filter's output is as long as its input (generator.rs:382-515)
    assert len(flatten(progs["square-cascade"]).nodes) > 10


def test_every_named_config_lowers_on_the_host():
    # tb_lower_check: the host half of tb_program_create, no device needed
    from tuun_b200.generator import lower_check
    trees = [("cfg1", W.cfg1_from_source()), ("cfg2", W.cfg2_harmonica(4)), ("cfg2x16", W.cfg2_harmonica(16)),
             ("cfg5", W.fm_filter_voice())] + W.cfg3_fm_variations() + W.cfg4_filters(0.1)
    for name, w in trees:
        info = lower_check(w)
        assert info.smem_bytes <= 220 * 1024 and info.threads in (32, 64, 128, 256), name
        if name == "cfg2":  # length programs fold their constant subtrees (lower.cpp len_trivial): 1,681 words before
            assert info.n_code_words // 4 < 1000, info.n_code_words
    # infinite, window-free trees take the 512-sample steady tiles; finite ones the general 256
    assert lower_check(W.fm_filter_voice()).tile == 512
    assert lower_check(W.cfg1_from_source()).tile == 256
