"""The named synthetic workloads of BASELINE.json / SURVEY.md §8(d), as optimized Waveform trees.

The trees are what the reference's parse -> evaluate -> optimize pipeline produces for the given
Tuun source (hand-derived from src/lib/optimizer.rs rules and lib/v0/std.tuun);
all scalar folding is done in f32 like the reference evaluator (expr.rs:155).
"""
from __future__ import annotations

import numpy as np

from .waveform import (Alt, BinaryPointOp, Const, Filter, Fin, Operator, Sine, Time, Waveform, add, mul)

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


def square(freq_rad: Waveform) -> Waveform:
    """lib/v0/std.tuun:21."""
    return Alt(Sine(freq_rad, Const(0.0)), Const(1.0), Const(-1.0))
