"""The named synthetic workloads of BASELINE.json / SURVEY.md §8(d), as optimized Waveform trees.

The trees are what the reference's parse -> evaluate -> optimize pipeline produces for the given
Tuun source (hand-derived from src/lib/optimizer.rs rules and lib/v0/std.tuun);
all scalar folding is done in f32 like the reference evaluator (expr.rs:155).
"""
from __future__ import annotations

import numpy as np

from .waveform import (Alt, Append, BinaryPointOp, Const, Filter, Fin, Fixed, Operator, Sine, Time, Waveform, add, mul)

F = np.float32
PI = F(3.14159265)  # lib/v0/std.tuun:7
TWO_PI = F(2.0) * PI


def lpf_coefficients(q, fc, sample_rate=44100):
    """lib/v0/std.tuun:118-129 (RBJ low-pass), evaluated with f32 scalars like builtins.rs:352,385."""
    q = np.asarray(q, dtype=F)
    fc = np.asarray(fc, dtype=F)
    w0 = TWO_PI * fc / F(sample_rate)
    alpha = np.sin(w0, dtype=F) / (F(2.0) * q)
    cosw = np.cos(w0, dtype=F)
    a0 = F(1.0) + alpha
    b1 = (F(1.0) - cosw) / a0
    b0 = b1 / F(2.0)
    a1 = (F(-2.0) * cosw) / a0
    a2 = (F(1.0) - alpha) / a0
    return b0.astype(F), b1.astype(F), b0.astype(F), a1.astype(F), a2.astype(F)


def lpf(x: Waveform, q: float, fc: float, sample_rate=44100) -> Waveform:
    b0, b1, b2, a1, a2 = lpf_coefficients(q, fc, sample_rate)
    return Filter(x, [Const(b0), Const(b1), Const(b2)], [Const(a1), Const(a2)])


def cfg1_sine(q_seconds=0.5) -> Waveform:
    """`$440 * Qw` (config 1): Fin{Time + (-Q), Sine{2*pi*440, 0}} after the optimizer pulls the Fin
    out of the product and drops the `* 1` (optimizer.rs:278,336-343)."""
    return Fin(add(Time(), Const(-F(q_seconds))), Sine(Const(TWO_PI * F(440.0)), Const(0.0)))


def fm_filter_voice() -> Waveform:
    """Config 5, one shared op list: sine(2*pi*(fc + I*fm*sine(2*pi*fm, pi/2)), 0) | lpf(Qv, cut)
    with the eight swept constants as per-voice parameters:
      p0 = 2*pi*fm, p1 = I*fm*2*pi, p2 = 2*pi*fc, p3..p5 = b0,b1,b2, p6,p7 = a1,a2."""
    mod = Sine(Const(1.0, param=0), Const(PI / F(2.0)))
    car = Sine(add(mul(mod, Const(1.0, param=1)), Const(1.0, param=2)), Const(0.0))
    return Filter(car, [Const(1.0, param=3), Const(1.0, param=4), Const(1.0, param=5)],
                  [Const(0.0, param=6), Const(0.0, param=7)])


def fm_pair_voice() -> Waveform:
    """The carrier of config 5 without its filter (same parameter table): against the oracle its error is the
    phase error itself."""
    mod = Sine(Const(1.0, param=0), Const(PI / F(2.0)))
    return Sine(add(mul(mod, Const(1.0, param=1)), Const(1.0, param=2)), Const(0.0))


def fm_filter_params(voice_ids, sample_rate=44100) -> np.ndarray:
    """The sweep of SURVEY §8(d) config 5, all in f32: fc = 55*2^(5*(v%256)/256),
    I = 10*((v>>8)%16)/16, D in {0.5,1,2,3}[(v>>12)%4], fm = D/2*fc, cut = 200*40^(((v>>14)%4)/4),
    Qv = 0.5 + 0.25*(v%7)."""
    v = np.asarray(voice_ids, dtype=np.int64)
    fc = F(55.0) * np.power(F(2.0), F(5.0) * (v % 256).astype(F) / F(256.0), dtype=F)
    idx = F(10.0) * ((v >> 8) % 16).astype(F) / F(16.0)
    d = np.array([0.5, 1.0, 2.0, 3.0], dtype=F)[(v >> 12) % 4]
    fm = d / F(2.0) * fc
    cut = F(200.0) * np.power(F(40.0), ((v >> 14) % 4).astype(F) / F(4.0), dtype=F)
    q = F(0.5) + F(0.25) * (v % 7).astype(F)
    b0, b1, b2, a1, a2 = lpf_coefficients(q, cut, sample_rate)
    p = np.stack([TWO_PI * fm, (idx * fm) * TWO_PI, TWO_PI * fc, b0, b1, b2, a1, a2], axis=1)
    return np.ascontiguousarray(p, dtype=F)


def biquad_noise_gain(a1, a2, n=16384) -> np.ndarray:
    """Round-off noise gain of the reference's direct-form recurrence y -= a1*y[n-1]; y -= a2*y[n-2]
    (generator.rs:500-502): the l2 norm of the impulse response of 1 / (1 + a1 z^-1 + a2 z^-2).  Every f32
    rounding inside the recurrence (about 6e-8 * |y| each) reaches the output multiplied by it."""
    a1 = np.atleast_1d(np.asarray(a1, dtype=np.float64))
    a2 = np.atleast_1d(np.asarray(a2, dtype=np.float64))
    y1 = np.zeros_like(a1)
    y2 = np.zeros_like(a1)
    acc = np.zeros_like(a1)
    for i in range(n):
        y = (1.0 if i == 0 else 0.0) - a1 * y1 - a2 * y2
        acc += y * y
        y2, y1 = y1, y
    return np.sqrt(acc)


def fm_filter_tolerance(params, base=1e-4) -> np.ndarray:
    """Per-voice max-abs tolerance of config 5 against the reference's f32 render: `base` (BASELINE.json: 1e-4)
    for 26 of the 28 filter shapes of the sweep, 1.5 x base for the two whose round-off noise gain is >= 190 —
    the 200 Hz low-passes with Q = 1.75 and Q = 2 (gain 196 and 209; 4,680 of the 65,536 voices).  Two f32
    evaluations of the reference's own recurrence whose inputs differ in the last bits (libm sin against any
    other sine) decorrelate into two realisations of its round-off noise (tests/test_cfg5_noise_floor.py shows
    that with the oracle alone).  Measured on a B200 over ALL 16,384 voices of the 200 Hz class x 441,000 samples
    (tests/diag/cfg5_highgain_sweep.py): Q <= 1.5: max 9.1e-5; Q = 1.75: 3 voices of 2,340 above 1e-4 (max
    1.10e-4); Q = 2: 42 of 2,340 (max 1.29e-4); every other cutoff class stays below 2.7e-5."""
    params = np.asarray(params)
    uniq, inv = np.unique(params[:, 6:8], axis=0, return_inverse=True)
    g = biquad_noise_gain(uniq[:, 0], uniq[:, 1])
    return (base * np.where(g >= 190.0, 1.5, 1.0))[inv.reshape(-1)]


def square(freq_rad: Waveform) -> Waveform:
    """lib/v0/std.tuun:21."""
    return Alt(Sine(freq_rad, Const(0.0)), Const(1.0), Const(-1.0))


# ---------------------------------------------------------------------------------------------
# Configs 1-4 written the way their Tuun source reads (tuun_b200.builder mirrors builtins.rs and
# lib/v0/std.tuun), then optimized like the reference pipeline does before generating.
# ---------------------------------------------------------------------------------------------
def _std(tempo=120.0, sample_rate=44100):
    from .builder import Std
    return Std(tempo=tempo, sample_rate=sample_rate)


def _finish(value) -> Waveform:
    from .builder import to_waveform
    from .optimizer import optimize
    return optimize(to_waveform(value))


def cfg1_from_source(tempo=120.0) -> Waveform:
    """`$440 * Qw` with `open std` (config 1)."""
    from .builder import times
    s = _std(tempo)
    return _finish(times(s.hz(440), s.note(s.Q)))


def cfg2_harmonica(n_notes=4, tempo=120.0, freq=440.0) -> Waveform:
    """`let h = harmonica(Q, 440) in <[h, h, h, h]>` (config 2): envelopes, seq / time-shift combinators."""
    from .builder import sequence
    s = _std(tempo)
    return _finish(sequence([s.harmonica(s.Q, freq) for _ in range(n_notes)]))


def cfg3_fm_variations():
    """The 12 uncommented programs of fm-variations.tuunp (lines 2,4,6,7,9,10,11,12,15,19,22,24),
    `capture(..)` stripped.  Returns [(name, optimized waveform)]; each is rendered for 10 s."""
    from .builder import divide, plus, sine, times
    s = _std()
    pi = s.pi
    fc, I, D = F(440), F(6), F(1)
    fm = times(divide(D, 2), fc)
    two_pi = times(2, pi)
    half_pi = divide(pi, 2)
    w_fm = times(two_pi, fm)   # 2*pi*fm
    w_fc = times(two_pi, fc)
    i_fm = times(I, fm)        # I * fm

    def true_fm(mod):  # sine(2*pi*(fc + (I * fm * mod)), 0)
        return sine(times(two_pi, plus(fc, times(i_fm, mod))), 0)

    def pm(mod):       # sine(2*pi*fc, I * mod)
        return sine(w_fc, times(I, mod))

    sweep = times(s.linear(0, 0.25), half_pi)  # linear(0,0.25)*pi/2  ((linear * pi) / 2 by precedence)
    sweep = divide(times(s.linear(0, 0.25), pi), 2)
    progs = [
        ("true-fm", true_fm(sine(w_fm, half_pi))),
        ("pm", pm(sine(w_fm, 0))),
        ("true-fm-mod-sweep", true_fm(sine(w_fm, plus(half_pi, sweep)))),
        ("pm-mod-sweep", pm(sine(w_fm, sweep))),
        ("true-fm-freq", times(two_pi, plus(fc, times(i_fm, sine(w_fm, half_pi))))),
        ("true-fm-mod-only", times(two_pi, times(i_fm, sine(w_fm, half_pi)))),
        ("square-fm", true_fm(s.square(w_fm))),   # passes 2*pi*fm to square as written
        ("square-pm", pm(s.square(w_fm))),
        ("cos-fm", true_fm(sine(w_fm, 0))),
        ("cos-pm", pm(sine(w_fm, half_pi))),
        ("pulse-fm", true_fm(s.pulse(0.5, fm))),
        ("pulse-pm", pm(s.pulse(0.5, fm))),
    ]
    return [(name, _finish(v)) for name, v in progs]


def cfg4_filters(noise_seconds=60.0, sample_rate=44100):
    """Config 4: resonant IIR chains (docs/filter.md style) over pulse / noise sources, 60 s each.
    The reference's Noise draws from an unseeded RNG, so the noise source is a Fixed buffer
    (Philox, seed 0x7475756E) as SURVEY 8(d) prescribes.  Returns [(name, optimized waveform)]."""
    from .builder import filter_, pipe, times
    s = _std(sample_rate=sample_rate)
    n = int(round(noise_seconds * sample_rate))
    rng = np.random.Generator(np.random.Philox(0x7475756E))
    noise = Fixed((rng.random(n, dtype=np.float32) * F(2) - F(1)).astype(F))
    f43 = filter_([0.2, 0.3, 0.2, 0.1], [-0.5, 0.2, -0.1])  # benches/tracker_benches.rs:74-80 shape
    progs = [
        ("square220-lpf", pipe(s.square(220), s.lpf(0.707, 2000))),
        ("noise-lpf", pipe(times(noise, 0.1), s.lpf(0.7, 2000))),
        ("square-cascade", pipe(s.square(220), s.lpf(4, 800), s.lpf(2, 1600), s.lpf(1, 3200))),
        ("pulse-filter_4_3", pipe(s.pulse(0.5, 110), f43)),
    ]
    return [(name, _finish(v)) for name, v in progs]


def tracker_benches(sample_rate=44100):
    """The five shapes of the reference's own criterion benches (benches/tracker_benches.rs), each
    driven there as N blocks of 1024 samples.  Returns [(name, waveform, n_blocks)]."""
    from .builder import Std, fin, mark, minus, pipe, plus, seq, sequence, times, to_waveform
    from .optimizer import optimize
    from .waveform import Marked, Noise
    t = Time()
    filter_1_1 = Filter(Time(), [Const(0.5)], [Const(-0.5)])                       # :20-34
    filter_1_1_linear = Filter(Time(), [add(mul(Time(), Const(-0.5)), Const(0.5))],  # :36-67
                               [add(mul(Time(), Const(0.5)), Const(-0.5))])
    filter_4_3 = Filter(Time(), [Const(0.00107949), Const(0.00323847), Const(0.00323847), Const(0.00107949)],
                        [Const(-2.5610316), Const(2.2132402), Const(-0.6435727)])  # :69-89
    # marks_4_40 (:92-117): 40 x Player::beats_waveform (player.rs:232-260) appended; tempo 120, 4 beats
    spb = F(0.5)
    beats = sequence([pipe(0, fin(minus(Time(), spb)), seq(minus(Time(), spb)), mark(i + 1)) for i in range(4)])
    one = Marked(0, optimize(to_waveform(beats)))
    marks = one
    for _ in range(39):
        marks = Append(marks, one)
    # large_440 (:119-165): `triangle(55) + (noise * 0.2) | R(1.0, 1.0)`, evaluated but NOT optimized
    s = Std(sample_rate=sample_rate)
    large = times(plus(s.triangle(55), times(Noise(), 0.2)), s.Rw(1.0, 1.0))
    return [("filter_1_1", filter_1_1, 43), ("filter_1_1_linear", filter_1_1_linear, 43),
            ("filter_4_3", filter_4_3, 43), ("marks_4_40", marks, 3438), ("large_440", to_waveform(large), 43)]


def fm_filter_sample_ids(n, first=0) -> np.ndarray:
    """`n` voices of the config 5 sweep picked by an odd multiplier (a permutation of 0..65535 whose
    neighbours differ in every bit field: carrier, modulation index, ratio, cutoff and Q all change from
    one pick to the next).  Pick 0 is voice 49230, a constant-rate carrier (index 0) under the 200 Hz, Q = 2
    low-pass.  Use this — not a small stride — wherever a subset of the sweep stands for the sweep."""
    k = np.arange(first, first + n, dtype=np.int64)
    return (k * 40503 + 49230) % 65536


def fm_filter_cover_ids(per_class=2) -> np.ndarray:
    """Voices covering every (modulation index, ratio, cutoff) class of the sweep — the 256 values of
    bits 8..15 — `per_class` times with different carriers; Q = 0.5 + 0.25 (v mod 7) takes all 7 values for
    every cutoff along the way (28 filter shapes)."""
    hi = np.repeat(np.arange(256, dtype=np.int64), per_class)
    k = np.tile(np.arange(per_class, dtype=np.int64), 256)
    lo = (hi * 89 + k * 131 + 17) % 256
    return (hi << 8) | lo
