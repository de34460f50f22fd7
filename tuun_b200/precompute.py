"""`Generator::precompute` (src/lib/generator.rs:868-1229) with the baking done on the GPU.

The reference replaces every part of a tree that is finite and static with an equivalent `Fixed`
(rendered ahead of time, at most 10 s), leaving infinite parts (Const, Time, Noise and what is built
only from them) and dynamic parts (Marked, Captured and what contains them) in place.  The decision
table below is the reference's; `generate_fixed` renders through the C ABI (tb_render) instead of
the CPU generator — the natural first caller of the batch renderer.

`render` may be injected (tests use the CPU oracle there to check the tree logic without a device).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from .waveform import (Alt, Append, BinaryPointOp, Captured, Const, Filter, Fin, Fixed, Marked, Noise, Operator, Reset,
                       Sine, Time, Waveform)

INFINITE = "infinite"  # Reason::Infinite (:876)
DYNAMIC = "dynamic"    # Reason::Dynamic (:878)


@dataclass
class _Result:
    w: Waveform
    why: Optional[str] = None  # None = Pc (pre-computable)

    @property
    def pc(self) -> bool:
        return self.why is None


def _resolve(a: str, b: str) -> str:  # resolve_reason (:980-986)
    return INFINITE if (a == INFINITE and b == INFINITE) else DYNAMIC


def gpu_render(sample_rate: int, device: int = -1) -> Callable[[Waveform, int], np.ndarray]:
    def render(w: Waveform, max_len: int) -> np.ndarray:
        from .generator import Program
        p = Program(w, sample_rate, device=device)
        out = np.zeros((1, max_len), dtype=np.float32)
        n = int(p.render(out)[0])
        p.close()
        return out[0, :n].copy()

    return render


def precompute(waveform: Waveform, sample_rate: int, device: int = -1,
               render: Optional[Callable[[Waveform, int], np.ndarray]] = None, log=None) -> Waveform:
    render = render or gpu_render(sample_rate, device)
    max_len = int(sample_rate) * 10  # :917
    say = log or (lambda *_: None)

    def generate_fixed(w: Waveform) -> Waveform:  # :896-930
        if isinstance(w, (Fixed, Const)):
            return w
        out = render(w, max_len)
        if len(out) == max_len:
            say(f"Warning: precompute generated max samples (maybe not finite?): {w}")
        return Fixed(out)

    def do_two(a, b, wf):  # :957-977
        ra, rb = go(a), go(b)
        if ra.pc and rb.pc:
            return _Result(wf(ra.w, rb.w))
        if ra.pc:
            return _Result(wf(generate_fixed(ra.w), rb.w), rb.why)
        if rb.pc:
            return _Result(wf(ra.w, generate_fixed(rb.w)), ra.why)
        return _Result(wf(ra.w, rb.w), _resolve(ra.why, rb.why))

    def go(w: Waveform) -> _Result:
        if isinstance(w, (Const, Time, Noise)):
            return _Result(w, INFINITE)
        if isinstance(w, Fixed):
            return _Result(w)
        if isinstance(w, Fin):  # :1041-1081
            rl, rw = go(w.length), go(w.waveform)
            if rw.why == DYNAMIC or rl.why == DYNAMIC:
                return _Result(Fin(rl.w, rw.w), DYNAMIC)
            return _Result(Fin(rl.w, rw.w))
        if isinstance(w, Append):
            return do_two(w.a, w.b, lambda a, b: Append(a, b))
        if isinstance(w, Sine):
            return do_two(w.frequency, w.phase, lambda f, p: Sine(f, p))
        if isinstance(w, Reset):
            return do_two(w.trigger, w.waveform, lambda t, x: Reset(t, x))
        if isinstance(w, BinaryPointOp):  # :1094-1121
            ra, rb = go(w.a), go(w.b)
            op = w.op
            if ra.pc and rb.pc:
                return _Result(BinaryPointOp(op, ra.w, rb.w))
            if op in (Operator.Multiply, Operator.Divide) and (
                    (ra.why == INFINITE and rb.pc) or (ra.pc and rb.why == INFINITE)):
                return _Result(BinaryPointOp(op, ra.w, rb.w))  # finite because min-length
            if ra.pc:
                return _Result(BinaryPointOp(op, generate_fixed(ra.w), rb.w), rb.why)
            if rb.pc:
                return _Result(BinaryPointOp(op, ra.w, generate_fixed(rb.w)), ra.why)
            return _Result(BinaryPointOp(op, ra.w, rb.w), _resolve(ra.why, rb.why))
        if isinstance(w, Filter):  # :1122-1181
            inner = go(w.waveform)
            ff = [go(c) for c in w.feed_forward]
            fb = [go(c) for c in w.feedback]
            reason = None
            for r in [inner] + ff + fb:
                if not r.pc:
                    reason = r.why if reason is None else _resolve(reason, r.why)
            extract = lambda r: generate_fixed(r.w) if (r.pc and reason is not None) else r.w
            return _Result(Filter(extract(inner), [extract(r) for r in ff], [extract(r) for r in fb]), reason)
        if isinstance(w, Alt):  # do_three (:988-1035)
            rs = [go(w.trigger), go(w.positive_waveform), go(w.negative_waveform)]
            if all(r.pc for r in rs):
                return _Result(Alt(rs[0].w, rs[1].w, rs[2].w))
            why = None
            for r in rs:
                if not r.pc:
                    why = r.why if why is None else _resolve(why, r.why)
            parts = [generate_fixed(r.w) if r.pc else r.w for r in rs]
            return _Result(Alt(*parts), why)
        if isinstance(w, (Marked, Captured)):  # do_one_dynamic (:941-953)
            r = go(w.waveform)
            inner = generate_fixed(r.w) if r.pc else r.w
            return _Result(Marked(w.id, inner) if isinstance(w, Marked) else Captured(w.file_stem, inner), DYNAMIC)
        raise TypeError(f"not a Waveform: {type(w)}")

    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 20000))
    try:
        r = go(waveform)
    finally:
        sys.setrecursionlimit(old)
    return generate_fixed(r.w) if r.pc else r.w
