"""Port of the reference optimizer (src/lib/optimizer.rs:9-442) over the Python Waveform mirror.

The reference-side host (Rust) runs `optimizer::optimize` before a tree crosses the C ABI; this port
exists so the named workloads can be built here, without a Rust toolchain, in exactly the shape the
reference's generator would see.  Match arms are kept in the reference's order because the first match
wins; float patterns (`Const(0.0)`, `Const(1.0)`) compare with `==`, so `-0.0` matches `0.0`
(SURVEY appendix A11).  All scalar arithmetic is f32.  Pinned by tests/test_optimizer.py against the five
tree equalities of optimizer.rs:450-590.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .waveform import (Alt, Append, BinaryPointOp, Captured, Const, Filter, Fin, Fixed, Marked, Noise,
                       Operator, Reset, Sine, Time, Waveform)

F = np.float32


def _c(x) -> Const:
    return Const(float(F(x)))


def _empty() -> Fixed:
    return Fixed(np.zeros(0, dtype=np.float32))


def _is_empty_fixed(w) -> bool:
    return isinstance(w, Fixed) and len(w.samples) == 0


def _is_const(w, value=None) -> bool:
    return isinstance(w, Const) and (value is None or w.value == value)


def _bin(op, a, b) -> BinaryPointOp:
    return BinaryPointOp(op, a, b)


def first_root(w: Waveform) -> Optional[Waveform]:
    """optimizer.rs:9-43."""
    if isinstance(w, Const):
        return _c(0.0) if w.value == 0.0 else None
    if isinstance(w, Time):
        return _c(0.0)
    if isinstance(w, BinaryPointOp) and w.op == Operator.Add:
        if isinstance(w.a, Time):
            return optimize(_bin(Operator.Multiply, w.b, _c(-1.0)))
        if isinstance(w.b, Time):
            return optimize(_bin(Operator.Multiply, w.a, _c(-1.0)))
        return None
    if isinstance(w, BinaryPointOp) and w.op == Operator.Subtract:
        return first_root(_bin(Operator.Add, w.a, optimize(_bin(Operator.Multiply, w.b, _c(-1.0)))))
    return None


def optimize(w: Waveform) -> Waveform:
    """optimizer.rs:52-442."""
    if isinstance(w, (Const, Time, Noise, Fixed)):
        return w
    if isinstance(w, Fin):  # :60-104
        length = optimize(w.length)
        if isinstance(length, Const) and length.value >= 0.0:
            return _empty()
        if isinstance(length, Fixed) and len(length.samples) > 0 and length.samples[0] >= 0.0:
            return _empty()
        if isinstance(length, Time):
            return _empty()
        inner = optimize(w.waveform)
        if isinstance(inner, Fin):
            ra, rb = first_root(length), first_root(inner.length)
            if isinstance(ra, Const) and isinstance(rb, Const):
                m = float(min(F(ra.value), F(rb.value)))
                return Fin(optimize(_bin(Operator.Subtract, Time(), _c(m))), inner.waveform)
            return Fin(length, Fin(inner.length, inner.waveform))
        return Fin(length, inner)
    if isinstance(w, Append):  # :105-114
        a, b = optimize(w.a), optimize(w.b)
        if _is_empty_fixed(a):
            return b
        if _is_empty_fixed(b):
            return a
        if isinstance(a, Fixed) and isinstance(b, Fixed):
            return Fixed(np.concatenate([a.samples, b.samples]))
        return Append(a, b)
    if isinstance(w, Sine):  # :116-135
        f, p = optimize(w.frequency), optimize(w.phase)
        if _is_const(f, 0.0) and isinstance(p, Const):
            return _c(np.sin(F(p.value), dtype=F))
        if _is_const(f, 0.0) and isinstance(p, Fixed):
            return Fixed(np.sin(p.samples, dtype=F))
        return Sine(f, p)
    if isinstance(w, Filter):  # :136-146
        return Filter(optimize(w.waveform), [optimize(c) for c in w.feed_forward], [optimize(c) for c in w.feedback])
    if isinstance(w, BinaryPointOp):
        op = w.op
        if op == Operator.Add:  # :139-184
            a, b = optimize(w.a), optimize(w.b)
            if _is_empty_fixed(a) or _is_empty_fixed(b):
                return _empty()
            if isinstance(a, Const) and isinstance(b, Const):
                return _c(F(a.value) + F(b.value))
            if _is_const(b, 0.0):
                return a
            if isinstance(a, Const):
                return optimize(_bin(Operator.Add, b, a))
            if isinstance(a, BinaryPointOp) and a.op == Operator.Add and isinstance(b, Const):
                return _bin(Operator.Add, a.a, optimize(_bin(Operator.Add, a.b, b)))
            if isinstance(a, Fin) and isinstance(b, Fin) and first_root(a.length) == first_root(b.length):
                return Fin(a.length, optimize(_bin(Operator.Add, a.waveform, b.waveform)))
            return _bin(Operator.Add, a, b)
        if op == Operator.Subtract:  # :185-193
            return optimize(_bin(Operator.Add, w.a, optimize(_bin(Operator.Multiply, w.b, _c(-1.0)))))
        if op == Operator.Merge:  # :194-274
            a, b = optimize(w.a), optimize(w.b)
            if _is_empty_fixed(a):
                return b
            if _is_empty_fixed(b):
                return a
            if isinstance(a, Const) and isinstance(b, Const):
                return _c(F(a.value) + F(b.value))
            if isinstance(a, (Time, Noise)) and _is_const(b, 0.0):
                return a
            if isinstance(a, Const):
                return optimize(_bin(Operator.Merge, b, a))
            if isinstance(a, Fin) and isinstance(b, Append):
                bb = b.a
                if isinstance(bb, Fin) and first_root(a.length) == first_root(bb.length):
                    return optimize(Append(Fin(a.length, _bin(Operator.Merge, a.waveform, bb.waveform)), b.b))
                return _bin(Operator.Merge, Fin(a.length, a.waveform), Append(bb, b.b))
            if isinstance(a, Marked) and isinstance(b, Append):
                aa, bb = a.waveform, b.a
                if isinstance(aa, Fin) and isinstance(bb, Fin) and first_root(aa.length) == first_root(bb.length):
                    return optimize(Append(Marked(a.id, Fin(aa.length, _bin(Operator.Merge, aa.waveform, bb.waveform))), b.b))
                return _bin(Operator.Merge, Marked(a.id, aa), Append(bb, b.b))
            return _bin(Operator.Merge, a, b)
        if op == Operator.Multiply:  # :275-347
            a, b = optimize(w.a), optimize(w.b)
            if _is_empty_fixed(a) or _is_empty_fixed(b):
                return _empty()
            if _is_const(b, 1.0):
                return a
            if isinstance(a, Const) and isinstance(b, Const):
                return _c(F(a.value) * F(b.value))
            if isinstance(a, Fixed) and isinstance(b, Const):
                return Fixed(a.samples * F(b.value))
            if isinstance(a, Const):
                return optimize(_bin(Operator.Multiply, b, a))
            if isinstance(a, BinaryPointOp) and isinstance(b, Const):
                if a.op == Operator.Multiply:
                    return _bin(Operator.Multiply, a.a, optimize(_bin(Operator.Multiply, a.b, b)))
                if a.op == Operator.Add:
                    return _bin(Operator.Add, optimize(_bin(Operator.Multiply, a.a, b)),
                                optimize(_bin(Operator.Multiply, a.b, b)))
                if a.op == Operator.Divide:
                    return _bin(Operator.Divide, optimize(_bin(Operator.Multiply, a.a, b)), a.b)
            if isinstance(a, Fin):
                return optimize(Fin(a.length, optimize(_bin(Operator.Multiply, a.waveform, b))))
            if isinstance(b, Fin):
                return optimize(Fin(b.length, optimize(_bin(Operator.Multiply, a, b.waveform))))
            return _bin(Operator.Multiply, a, b)
        if op == Operator.Divide:  # :348-393
            a, b = optimize(w.a), optimize(w.b)
            if _is_empty_fixed(b):
                return _empty()
            if isinstance(b, Const):
                with np.errstate(divide="ignore"):
                    return optimize(_bin(Operator.Multiply, a, _c(F(1.0) / F(b.value))))
            if isinstance(a, BinaryPointOp) and a.op == Operator.Divide:
                return _bin(Operator.Divide, a.a, optimize(_bin(Operator.Multiply, a.b, b)))
            if isinstance(b, BinaryPointOp) and b.op == Operator.Divide:
                return _bin(Operator.Divide, optimize(_bin(Operator.Multiply, a, b.b)), b.a)
            if isinstance(a, Fin):
                return optimize(Fin(a.length, optimize(_bin(Operator.Divide, a.waveform, b))))
            if isinstance(b, Fin):
                return optimize(Fin(b.length, optimize(_bin(Operator.Divide, a, b.waveform))))
            return _bin(Operator.Divide, a, b)
        if op == Operator.Power:  # :394-405
            a, b = optimize(w.a), optimize(w.b)
            if _is_empty_fixed(a) or _is_empty_fixed(b):
                return _empty()
            if isinstance(a, Const) and _is_const(b, 0.0):
                return _c(1.0)
            if _is_const(b, 1.0):
                return a
            if isinstance(a, Const) and isinstance(b, Const):
                with np.errstate(all="ignore"):
                    return _c(np.power(F(a.value), F(b.value), dtype=F))
            if isinstance(a, Fixed) and isinstance(b, Const):
                with np.errstate(all="ignore"):
                    return Fixed(np.power(a.samples, F(b.value), dtype=F))
            return _bin(Operator.Power, a, b)
        raise ValueError(op)
    if isinstance(w, Reset):  # :406-414
        return Reset(optimize(w.trigger), optimize(w.waveform))
    if isinstance(w, Alt):  # :415-431
        t, p, n = optimize(w.trigger), optimize(w.positive_waveform), optimize(w.negative_waveform)
        if isinstance(t, Const) and t.value >= 0.0:
            return p
        if isinstance(t, Const) and t.value < 0.0:
            return n
        return Alt(t, p, n)
    if isinstance(w, Marked):
        return Marked(w.id, optimize(w.waveform))
    if isinstance(w, Captured):
        return Captured(w.file_stem, optimize(w.waveform))
    raise TypeError(type(w))
