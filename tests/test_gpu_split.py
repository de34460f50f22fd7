"""Time-axis split (tuun_b200/csrc/split.cu, abi.cpp render_split): one voice rendered as S segments side by
side must be the stream the serial walk produces (generator.rs:76-85: "pick up where this one left off").

Partition invariance, S = 1 against many: sines — whose state is a sum of phase increments, kept as an exact
u64 — come out bit-identical; filters, whose histories come out of an f64 scan of affine maps instead of
the serial f32 recurrence, agree to the recurrence's own round-off noise (<= 1e-6 for the plain low-passes
here, a few 1e-6 for resonant ones: 1e-7 x the noise gain of workloads.biquad_noise_gain).  Every case is also
held against the CPU oracle at north_star's 1e-4."""
import math

import numpy as np
import pytest

from oracle.binding import OracleProgram
from tuun_b200.waveform import Alt, Const, Filter, Noise, Sine, Time, add, f32, mul

pytestmark = pytest.mark.gpu
SR = 44100
TAU = f32(2 * math.pi)


def render(w, n, monkeypatch, split, params=None, voices=1, calls=None, seed=None):
    """Rows [voices, n] through tb_render with TUUN_B200_SPLIT=split ("0": serial); `calls` cuts the render
    into several calls."""
    from tuun_b200.generator import Program
    monkeypatch.setenv("TUUN_B200_SPLIT", str(split))
    p = Program(w, SR)
    if seed is not None:
        p.seed_noise(seed)
    out = np.zeros((voices, n), dtype=np.float32)
    a = 0
    for c in (calls or [n]):
        blk = np.zeros((voices, c), dtype=np.float32)
        lens = p.render(blk, params=params)
        assert (lens == c).all()
        out[:, a:a + c] = blk
        a += c
    assert a == n
    return out, p.info


def oracle(w, n, params=None, voices=1, seed=None):
    o = OracleProgram(w, SR)
    rows = np.zeros((voices, n), dtype=np.float32)
    for v in range(voices):
        o.initialize_state()
        if seed is not None:
            o.seed_noise(seed, v)
        if params is not None:
            o.set_params(params[v])
        rows[v] = o.render(n)
    return rows


def lpf(x, q, fc):
    from tuun_b200.workloads import lpf as _lpf
    return _lpf(x, q, fc)


def test_sine_is_bit_identical_for_any_partition(monkeypatch):
    w = Sine(Const(TAU * f32(440.0)), Const(0.25))
    n = 256 + 37 * 512 + 99
    serial, i0 = render(w, n, monkeypatch, 0)
    assert i0.split_passes == 1 and i0.split_rounds == 0
    for s in (2, 8, 32):
        got, info = render(w, n, monkeypatch, s)
        assert info.split_rounds >= 1 and info.split_segments >= 2
        assert np.array_equal(got, serial), (s, np.abs(got - serial).max())
    assert np.abs(serial - oracle(w, n)).max() <= 1e-6


def test_fm_phase_sum_is_exact_across_segments(monkeypatch):
    """A frequency-modulated carrier: the accumulator at a segment's start is the u64 sum of all earlier
    increments, from per-segment sums of a first pass (two passes)."""
    from tuun_b200.workloads import fm_filter_params, fm_filter_sample_ids, fm_pair_voice
    w = fm_pair_voice()
    params = fm_filter_params(fm_filter_sample_ids(6))
    n = 256 + 64 * 512
    serial, i0 = render(w, n, monkeypatch, 0, params=params, voices=6)
    assert i0.split_passes == 2
    for s in (4, 64):
        got, info = render(w, n, monkeypatch, s, params=params, voices=6)
        assert info.split_rounds >= 1
        assert np.array_equal(got, serial), (s, np.abs(got - serial).max())
    assert np.abs(serial - oracle(w, n, params=params, voices=6)).max() <= 1e-5


def test_filters_agree_to_their_round_off_noise(monkeypatch):
    """square | lpf (config 4), a three-biquad cascade (one more pass per filter), noise | lpf, the
    tracker_benches shapes filter_1_1 / filter_4_3 over a clock."""
    sq = Alt(Sine(Const(TAU * f32(220.0)), Const(0.0)), Const(1.0), Const(-1.0))
    cases = [
        ("square-lpf", lpf(sq, 0.707, 2000.0), 2, 1e-6, None),
        ("cascade", lpf(lpf(lpf(sq, 4.0, 800.0), 2.0, 1600.0), 1.0, 3200.0), 4, 1e-5, None),  # Q = 4: noise gain 19
        ("noise-lpf", lpf(mul(Noise(), Const(0.1)), 0.7, 2000.0), 2, 1e-6, 77),
        ("filter_4_3", Filter(sq, [Const(0.00107949), Const(0.00323847), Const(0.00323847), Const(0.00107949)],
                              [Const(-2.5610316), Const(2.2132402), Const(-0.6435727)]), 2, 1e-6, None),
        ("fir-5", Filter(sq, [Const(0.2)] * 5, []), 2, 0.0, None),
    ]
    n = 256 + 48 * 512 + 300
    for name, w, passes, tol, seed in cases:
        serial, i0 = render(w, n, monkeypatch, 0, seed=seed)
        assert i0.split_passes == passes, (name, i0.split_passes)
        ref = oracle(w, n, seed=seed)
        assert np.abs(serial - ref).max() <= 1e-4, name
        for s in (2, 16):
            got, info = render(w, n, monkeypatch, s, seed=seed)
            assert info.split_rounds >= 1, name
            d = float(np.abs(got - serial).max())
            assert d <= tol, (name, s, d)
            assert np.abs(got - ref).max() <= 1e-4, (name, s)


def test_clock_under_a_filter(monkeypatch):
    """filter_1_1 of benches/tracker_benches.rs:19-34: a one-pole filter over Time — the clock's position is
    analytic, the ramp's running average is an affine map like any other."""
    w = Filter(Time(), [Const(0.5)], [Const(-0.5)])
    n = 43 * 1024
    serial, _ = render(w, n, monkeypatch, 0)
    got, info = render(w, n, monkeypatch, 8)
    assert info.split_rounds >= 1
    assert np.abs(got - serial).max() <= 1e-6 * max(1.0, float(np.abs(serial).max()))
    assert np.abs(got - oracle(w, n)).max() <= 1e-4


def test_streams_continue_across_split_calls(monkeypatch):
    """The last segment's final state is the voice's state: a split call, a serial call and another split call
    are one stream."""
    from tuun_b200.workloads import fm_filter_params, fm_filter_voice
    w = fm_filter_voice()
    params = fm_filter_params([49157, 34061, 16389 + 256 * 9])   # cutoffs >= 632 Hz: noise gain < 40
    n = 3 * 20000
    serial, _ = render(w, n, monkeypatch, 0, params=params, voices=3)
    got, info = render(w, n, monkeypatch, 8, params=params, voices=3, calls=[20000, 1000, 19000, 20000])
    assert info.split_passes == 3 and info.split_rounds >= 3
    assert np.abs(got - serial).max() <= 1e-5   # the 503 Hz low-pass of the third voice has noise gain 38
    assert np.abs(got - oracle(w, n, params=params, voices=3)).max() <= 1e-4


def test_long_single_voice_splits_by_itself(monkeypatch):
    """Config 4's shape: one voice x 60 s.  No knob: the call is split because one warp would render it alone."""
    from tuun_b200.generator import Program
    monkeypatch.delenv("TUUN_B200_SPLIT", raising=False)
    sq = Alt(Sine(Const(TAU * f32(220.0)), Const(0.0)), Const(1.0), Const(-1.0))
    w = lpf(sq, 0.707, 2000.0)
    n = 60 * SR
    p = Program(w, SR)
    out = np.zeros((1, n), dtype=np.float32)
    lens = p.render(out)
    assert lens[0] == n
    info = p.info
    assert info.split_rounds >= 1 and info.split_segments >= 256
    ref = oracle(w, n)
    e = np.abs(out - ref)
    assert e.max() <= 1e-4
    assert e[:, -SR:].max() <= 2 * e[:, :SR].max() + 1e-6   # no drift over the minute


def test_reset_oscillators_split(monkeypatch):
    """sawtooth / pulse / triangle of lib/v0/std.tuun are a Reset over a clock (Appendix B of SURVEY.md): a segment
    needs the class of the trigger's last sample and the run's local clock at its start — a copy and a "last set"
    scan over the segments (split.cu SP_RESET_SIGN / SP_CLK).  Edges must land on the same samples."""
    from tuun_b200.builder import Std, pipe, to_waveform, filter_
    from tuun_b200.optimizer import optimize
    s = Std()
    f43 = filter_([0.2, 0.3, 0.2, 0.1], [-0.5, 0.2, -0.1])
    cases = [("sawtooth", s.sawtooth(220), 2), ("pulse", s.pulse(0.3, 110), 2), ("triangle", s.triangle(55), 2),
             ("pulse-filter_4_3", pipe(s.pulse(0.5, 110), f43), 3)]
    n = 256 + 64 * 512 + 77
    for name, v, passes in cases:
        w = optimize(to_waveform(v))
        serial, i0 = render(w, n, monkeypatch, 0)
        assert i0.split_passes == passes, (name, i0.split_passes)
        ref = oracle(w, n)
        assert np.abs(serial - ref).max() <= 1e-4, name
        for sgm in (4, 64):
            got, info = render(w, n, monkeypatch, sgm)
            assert info.split_rounds >= 1, name
            bad = np.abs(got - ref) > 1e-4
            assert not bad.any(), (name, sgm, int(bad.sum()), np.nonzero(bad[0])[0][:5])
            assert np.abs(got - serial).max() <= 2e-6, (name, sgm)


def test_pulse_modulated_fm_splits_by_itself(monkeypatch):
    """fm-variations.tuunp:22 (config 3): a pulse — Reset, Alt — driving the frequency of a carrier: the Reset's
    clock, then the carrier's phase sum, then the samples (three passes), 10 s of one voice."""
    from tuun_b200 import workloads as W
    from tuun_b200.generator import Program
    monkeypatch.delenv("TUUN_B200_SPLIT", raising=False)
    name, w = W.cfg3_fm_variations()[10]
    assert name == "pulse-fm"
    n = 441000
    p = Program(w, SR)
    out = np.zeros((1, n), dtype=np.float32)
    assert p.render(out)[0] == n
    assert p.info.split_rounds >= 1 and p.info.split_passes == 3
    e = np.abs(out - oracle(w, n))
    assert e.max() <= 1e-4 and e[:, -SR:].max() <= 2 * e[:, :SR].max() + 1e-6


def test_fm_batch_split_with_filter_warm_up(monkeypatch):
    """The form strong scaling takes (65,536 voices over 8 GPUs leave 8,192 each: too few to fill the lane kernel):
    every voice in S segments, the carrier's phase sums from a pass that computes nothing else, the biquad's history
    from a warm-up before each segment (abi.cpp render_split_fm).  Forced on a small batch here; voices cover every
    filter shape, including the slowest-forgetting one (200 Hz, Q = 2: 3,500 samples of warm-up)."""
    import torch
    from tuun_b200.generator import Program
    from tuun_b200.workloads import fm_filter_cover_ids, fm_filter_params, fm_filter_tolerance, fm_filter_voice
    w = fm_filter_voice()
    ids = fm_filter_cover_ids(1)[::2]            # 128 voices
    params = fm_filter_params(ids)
    V, n = len(ids), 256 + 4 * 40000 + 100
    monkeypatch.setenv("TUUN_B200_LANE_MIN_VOICES", "1")
    monkeypatch.setenv("TUUN_B200_SPLIT_FM", "0")
    serial = np.zeros((V, n), dtype=np.float32)
    Program(w, SR).render(serial, params=params)
    monkeypatch.setenv("TUUN_B200_SPLIT_FM", "4")
    p = Program(w, SR)
    got = torch.zeros((V, n), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()  # (the fill runs on torch's stream, the renders on the program's own)
    lens = np.zeros(V, dtype=np.uint64)
    p.render(got, params=params, out_len=lens)
    info = p.info
    assert (lens == n).all() and info.split_rounds == 1 and info.split_segments == 4 and info.split_seg_samples == 40000
    assert info.split_fm_rounds == 1
    got = got.cpu().numpy()
    ref, _, _, _ = OracleProgram(w, SR).render_batch(params, V, n, threads=8)
    tol = fm_filter_tolerance(params, 1e-4)
    assert (np.abs(got - ref).max(axis=1) <= tol).all()
    assert (np.abs(serial - ref).max(axis=1) <= tol).all()
    d = np.abs(got - serial).max(axis=1)
    from tuun_b200.workloads import biquad_noise_gain
    g = biquad_noise_gain(params[:, 6], params[:, 7])
    assert (d <= 6e-7 * np.maximum(g, 5.0)).all(), float((d / np.maximum(g, 5.0)).max())   # the filter's own round-off noise
    assert float(np.median(d)) <= 2e-6
    # and the stream continues: the next (serial) call of both programs agrees the same way
    monkeypatch.setenv("TUUN_B200_SPLIT_FM", "0")
    nxt = np.zeros((V, 3000), dtype=np.float32)
    p.render(nxt, params=params)
    o = OracleProgram(w, SR)
    ref2, _, _, _ = o.render_batch(params, V, n + 3000, threads=8)
    assert (np.abs(nxt - ref2[:, n:]).max(axis=1) <= tol).all()
