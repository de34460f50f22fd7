"""Expression-value layer: what the reference evaluator's built-ins do to VALUES, mirrored in Python.

The reference turns Tuun source into a `Waveform` tree with `eval.rs` + `builtins.rs` and the
library `lib/v0/std.tuun`.  This module mirrors the value semantics of that layer — not the
parser — so the named workloads (BASELINE.json configs 1-4) can be written the way the Tuun
source reads and come out as the same un-optimized trees; `tuun_b200.optimizer.optimize` then
produces the shapes the generator sees.  All scalars are f32 (expr.rs:155).

  builtins.rs:35-98    binary_op  (Float/Waveform/Seq dispatch)       -> plus, minus, times, ...
  builtins.rs:179-206  add_offsets                                     -> _add_offsets
  builtins.rs:208-299  followed_by (`\\`)                               -> followed_by
  builtins.rs:344-398  sine, cos                                       -> sine, cos
  builtins.rs:600-727  curry, fin, seq, unseq                          -> fin, seq, unseq
  builtins.rs:780-842  filter                                          -> filter_
  builtins.rs:844-898  reset, alt                                      -> reset, alt
  builtins.rs:921-973  chord `{..}`, sequence `<..>`                   -> chord, sequence
  lib/v0/std.tuun                                                      -> class Std
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Union

import numpy as np

from . import optimizer
from .waveform import (Alt, Append, BinaryPointOp, Captured, Const, Filter, Fin, Fixed, Operator, Reset, Sine, Time,
                       Waveform)

F = np.float32
Value = Union[np.float32, Waveform, "Seq", list, Callable]


@dataclass
class Seq:
    """Expr::Seq { offset, waveform } as a value (both waveforms), expr.rs."""
    offset: Waveform
    waveform: Waveform


class TuunError(Exception):
    """Expr::Error."""


def _is_float(x) -> bool:
    return isinstance(x, (float, int, np.floating, np.integer)) and not isinstance(x, bool)


def _wf(x) -> Waveform:
    if isinstance(x, Waveform):
        return x
    if _is_float(x):
        return Const(float(F(x)))
    raise TuunError(f"expected a waveform or float, got {x!r}")


def _binary_op(name: str, float_op, op: Operator):
    """builtins.rs:35-98."""

    def apply(a, b):
        if _is_float(a) and _is_float(b):
            with np.errstate(all="ignore"):
                return F(float_op(F(a), F(b)))
        if isinstance(a, Seq) and isinstance(b, Seq):
            raise TuunError(f"Invalid arguments for {name}")
        if isinstance(a, Seq):
            return Seq(a.offset, BinaryPointOp(op, a.waveform, _wf(b)))
        if isinstance(b, Seq):
            return Seq(b.offset, BinaryPointOp(op, _wf(a), b.waveform))
        return BinaryPointOp(op, _wf(a), _wf(b))

    return apply


plus = _binary_op("+", lambda a, b: a + b, Operator.Add)
_minus2 = _binary_op("-", lambda a, b: a - b, Operator.Subtract)
times = _binary_op("*", lambda a, b: a * b, Operator.Multiply)
divide = _binary_op("/", lambda a, b: a / b, Operator.Divide)
power = _binary_op("pow", lambda a, b: np.power(a, b, dtype=F), Operator.Power)


def minus(a, b=None):
    """Binary `-`, or unary `-x` = Const(-1) * x for waveforms (builtins.rs:113-132)."""
    if b is None:
        if _is_float(a):
            return F(-F(a))
        return BinaryPointOp(Operator.Multiply, Const(-1.0), _wf(a))
    return _minus2(a, b)


def merge(a, b):
    """`&` (builtins.rs:156-177): two floats are promoted to constants."""
    if _is_float(a) and _is_float(b):
        return BinaryPointOp(Operator.Merge, Const(float(F(a))), Const(float(F(b))))
    return _binary_op("&", None, Operator.Merge)(a, b)


def sine(freq, phase):
    """builtins.rs:344-376: radians per second, radians; frequency 0 with a float phase folds to f32 sin."""
    if _is_float(freq) and _is_float(phase) and F(freq) == 0.0:
        return F(np.sin(F(phase), dtype=F))
    return Sine(_wf(freq), _wf(phase))


def cos(x):
    """builtins.rs:379-398."""
    if _is_float(x):
        return F(np.cos(F(x), dtype=F))
    return Sine(Const(0.0), BinaryPointOp(Operator.Add, _wf(x), Const(float(F(np.pi / 2)))))


def alt(trigger, positive, negative):
    """builtins.rs:869-898."""
    return Alt(_wf(trigger), _wf(positive), _wf(negative))


def reset(trigger, waveform):
    """builtins.rs:844-867."""
    if not isinstance(trigger, Waveform):
        raise TuunError("First argument must be a waveform")
    return Reset(trigger, _wf(waveform))


def _curry(f: Callable[[Waveform], Waveform]):
    """builtins.rs:600-641: a waveform -> waveform function that also maps over floats and seqs."""

    def apply(x):
        if isinstance(x, Seq):
            return Seq(x.offset, f(x.waveform))
        return f(_wf(x))

    return apply


def fin(length):
    """builtins.rs:643-678 (a float is a constant waveform, not a duration)."""
    length = _wf(length)
    return _curry(lambda w: Fin(length, w))


def seq(offset):
    """builtins.rs:680-727."""
    offset = _wf(offset)

    def apply(x):
        return Seq(offset, _wf(x))

    return apply


def unseq():
    def apply(x):
        if not isinstance(x, Seq):
            raise TuunError("Expected seq as argument to unseq")
        return x.waveform

    return apply


def mark(mark_id: int):
    """`mark(n)`: tag a waveform so the tracker reports when it starts (builtins.rs, curried)."""
    from .waveform import Marked
    return _curry(lambda w: Marked(int(mark_id), w))


def capture(stem: str):
    return _curry(lambda w: Captured(stem, w))


def filter_(feed_forward: list, feedback: list):
    """builtins.rs:780-842."""
    if not feed_forward:
        raise TuunError("Filter requires at least one feed-forward coefficient")
    ff = [_wf(c) for c in feed_forward]
    fb = [_wf(c) for c in feedback]
    return _curry(lambda w: Filter(w, list(ff), list(fb)))


def _add_offsets(a: Waveform, b: Waveform) -> Waveform:
    """builtins.rs:179-206: Time + optimize((root_a + root_b) * -1)."""
    ra, rb = optimizer.first_root(a), optimizer.first_root(b)
    if ra is None or rb is None:
        raise TuunError("Cannot add offsets that are not linear functions of Time")
    s = optimizer.optimize(BinaryPointOp(Operator.Multiply, BinaryPointOp(Operator.Add, ra, rb), Const(-1.0)))
    return BinaryPointOp(Operator.Add, Time(), s)


def followed_by(a, b):
    """`a \\ b` (builtins.rs:208-299): Merge(a, Append(Fin(offset_a, 0), b))."""
    if not isinstance(a, Seq):
        raise TuunError("Expected seq as first argument to \\")
    body = lambda bw: BinaryPointOp(Operator.Merge, a.waveform, Append(Fin(a.offset, Const(0.0)), bw))
    if isinstance(b, Seq):
        return Seq(_add_offsets(a.offset, b.offset), body(b.waveform))
    return body(_wf(b))


def chord(xs: list) -> Waveform:
    """`{[..]}` (builtins.rs:921-944)."""
    result: Waveform = Fin(Const(0.0), Const(0.0))
    for x in reversed(xs):
        result = BinaryPointOp(Operator.Merge, _wf(x), result)
    return result


def sequence(xs: list):
    """`<[..]>` (builtins.rs:946-973): right fold with `\\`."""
    if not xs:
        return Fixed([])
    if len(xs) == 1:
        return _wf(xs[0])
    result = xs[-1]
    for x in reversed(xs[:-1]):
        result = followed_by(x, result)
    return result


def pipe(x, *fs):
    """`x | f | g` — reverse application."""
    for f in fs:
        x = f(x)
    return x


def to_waveform(x) -> Waveform:
    """What a player does with a program's value: a Seq plays its waveform (player.rs / wasm.rs:259)."""
    if isinstance(x, Seq):
        return x.waveform
    return _wf(x)


class Std:
    """lib/v0/std.tuun with `tempo` and `sample_rate` bound (wasm.rs:184-266 binds both as floats)."""

    def __init__(self, tempo: float = 120.0, sample_rate: int = 44100):
        self.tempo = F(tempo)
        self.sample_rate = F(sample_rate)
        self.pi = F(3.14159265)  # std.tuun:7
        self.time = Time
        # std.tuun:152-168
        self.beats_per_measure = F(4)
        self.W = self.duration_from_beats(self.beats_per_measure)
        self.H = divide(self.W, 2)
        self.Q = divide(self.W, 4)
        self.E = divide(self.W, 8)

    # -- math (std.tuun:5-6)
    def min(self, x, y):
        return alt(minus(x, y), y, x)

    def max(self, x, y):
        return alt(minus(x, y), x, y)

    # -- waves (std.tuun:14-37)
    def sin(self, phase):
        return sine(0, phase)

    def hz(self, freq_hz):
        """`$f`."""
        return sine(times(times(2, self.pi), freq_hz), 0)

    def sawtooth(self, freq_hz):
        return times(plus(reset(self.hz(freq_hz), times(minus(freq_hz), Time())), 0.5), 2)

    def square(self, freq_hz):
        return alt(self.hz(freq_hz), 1, -1)

    def triangle(self, freq_hz):
        slope = times(4, freq_hz)
        a = minus(times(Time(), slope), 1)
        b = plus(times(Time(), minus(slope)), 3)
        return alt(self.hz(freq_hz), reset(self.hz(freq_hz), a), reset(self.hz(freq_hz), b))

    def pulse(self, width, freq_hz):
        return alt(minus(self.sawtooth(freq_hz), width), 1, -1)

    # -- helpers (std.tuun:59-73)
    def db2amp(self, db):
        return power(10.0, divide(db, 20.0))

    def linear(self, initial, slope):
        return plus(initial, times(Time(), slope))

    def midi(self, m):
        """`@m`."""
        return times(power(2, divide(minus(m, 69), 12)), 440)

    # -- envelopes (std.tuun:76-101)
    def Aw(self, dur):
        return pipe(self.linear(0.0, divide(1.0, dur)), fin(minus(Time(), dur)), seq(minus(Time(), dur)))

    def Dw(self, dur, a):
        return pipe(self.linear(1.0, divide(minus(a, 1.0), dur)), fin(minus(Time(), dur)), seq(minus(Time(), dur)))

    def Sw(self, dur, a):
        return pipe(a, fin(minus(Time(), dur)), seq(minus(Time(), dur)))

    def Rw(self, dur, level):
        return pipe(self.linear(level, divide(minus(level), dur)), fin(minus(Time(), dur)))

    def ADSR(self, attack_dur, decay_dur, sustain_level, sustain_dur, release_dur):
        return lambda w: times(w, sequence([self.Aw(attack_dur), self.Dw(decay_dur, sustain_level),
                                            self.Sw(sustain_dur, sustain_level),
                                            self.Rw(release_dur, sustain_level)]))

    def add_semitones(self, freq, n):
        return times(power(2, divide(n, 12)), freq)

    def add_cents(self, freq, n):
        return times(power(2, divide(n, 1200)), freq)

    # -- filters (std.tuun:112-129)
    def moving_average(self, n):
        c = divide(1, plus(n, 1))
        return filter_([c] * (int(n) + 1), [])

    def lpf(self, Q, fc):
        w0 = divide(times(times(2, self.pi), fc), self.sample_rate)
        alpha = divide(self.sin(w0), times(2, Q))
        b0 = divide(minus(1, cos(w0)), 2)
        b1 = minus(1, cos(w0))
        b2 = divide(minus(1, cos(w0)), 2)
        a0 = plus(1, alpha)
        a1 = times(-2, cos(w0))
        a2 = minus(1, alpha)
        return filter_([divide(b0, a0), divide(b1, a0), divide(b2, a0)], [divide(a1, a0), divide(a2, a0)])

    # -- instruments (std.tuun:134-150)
    def harmonica(self, dur, freq):
        osc1 = lambda: self.pulse(plus(0.93, times(0.05, self.hz(1.6))), freq)
        osc2 = reset(osc1(), self.pulse(0.7, self.add_cents(self.add_semitones(freq, 8), 7)))
        osc = plus(times(0.375, osc1()), times(0.5, osc2))
        a, r = F(0.13), F(0.33)
        d = self.max(0.33, minus(dur, plus(a, r)))
        s = self.max(minus(dur, plus(plus(a, d), r)), 0)
        return pipe(osc, self.lpf(0.5, 1900), self.ADSR(a, d, 0.5, s, r), fin(minus(Time(), dur)),
                    seq(minus(Time(), dur)))

    # -- time (std.tuun:153-176)
    def duration_from_beats(self, beats):
        return times(beats, divide(60, self.tempo))

    def note(self, dur):
        """`Ww`, `Hw`, `Qw`, `Ew`: 1 | fin(time - dur) | seq(time - dur)."""
        return pipe(1, fin(minus(Time(), dur)), seq(minus(Time(), dur)))

    def rest(self, dur):
        """`Wrw`, ...: 0 | fin(0) | seq(time - dur)."""
        return pipe(0, fin(0), seq(minus(Time(), dur)))
