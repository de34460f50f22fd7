"""Waveform IR — Python mirror of the reference's `enum Waveform<MarkId, State>` and
`enum Operator` (src/lib/waveform.rs:5-100), plus the flattening into the `tb_node` op list
that crosses the C ABI (include/tuun_b200.h).

Same variant names and field meaning as the reference so tests read like the reference's own
(`generator.rs:1353-1925`).  A tree is a tree, not a DAG: a Python object that occurs twice is
flattened twice, exactly as the reference clones shared sub-expressions.
"""
from __future__ import annotations

import ctypes
import enum
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np


class Operator(enum.IntEnum):  # waveform.rs:5-19
    Add = 0
    Subtract = 1
    Multiply = 2
    Divide = 3
    Merge = 4
    Power = 5


class Kind(enum.IntEnum):  # waveform.rs:23-100, same order
    Const = 0
    Time = 1
    Noise = 2
    Fixed = 3
    Fin = 4
    Append = 5
    Sine = 6
    Filter = 7
    BinaryPointOp = 8
    Reset = 9
    Alt = 10
    Marked = 11
    Captured = 12


class Waveform:
    """Base class; subclasses are the 13 variants."""

    def __repr__(self):  # compact, like the reference's Display impl (waveform.rs:102-176)
        return show(self)


def f32(x) -> float:
    """Round to f32 the way every reference scalar is (expr.rs:155)."""
    return float(np.float32(x))


@dataclass(eq=True, repr=False)
class Const(Waveform):
    value: float
    param: int = -1  # >= 0: column of the per-voice parameter table (tb_node.param_slot)

    def __post_init__(self):
        self.value = f32(self.value)


@dataclass(eq=True, repr=False)
class Time(Waveform):
    pass


@dataclass(eq=True, repr=False)
class Noise(Waveform):
    pass


@dataclass(eq=False, repr=False)
class Fixed(Waveform):
    samples: Sequence[float]

    def __post_init__(self):
        self.samples = np.ascontiguousarray(self.samples, dtype=np.float32)

    def __eq__(self, other):
        return isinstance(other, Fixed) and np.array_equal(self.samples, other.samples)


@dataclass(eq=True, repr=False)
class Fin(Waveform):
    length: Waveform
    waveform: Waveform


@dataclass(eq=True, repr=False)
class Append(Waveform):
    a: Waveform
    b: Waveform


@dataclass(eq=True, repr=False)
class Sine(Waveform):
    frequency: Waveform
    phase: Waveform


@dataclass(eq=True, repr=False)
class Filter(Waveform):
    waveform: Waveform
    feed_forward: List[Waveform]
    feedback: List[Waveform] = field(default_factory=list)


@dataclass(eq=True, repr=False)
class BinaryPointOp(Waveform):
    op: Operator
    a: Waveform
    b: Waveform


@dataclass(eq=True, repr=False)
class Reset(Waveform):
    trigger: Waveform
    waveform: Waveform


@dataclass(eq=True, repr=False)
class Alt(Waveform):
    trigger: Waveform
    positive_waveform: Waveform
    negative_waveform: Waveform


@dataclass(eq=True, repr=False)
class Marked(Waveform):
    id: int
    waveform: Waveform


@dataclass(eq=True, repr=False)
class Captured(Waveform):
    file_stem: str
    waveform: Waveform


def show(w: Waveform) -> str:
    if isinstance(w, Const):
        return f"Const({w.value!r}{'@p%d' % w.param if w.param >= 0 else ''})"
    if isinstance(w, Time):
        return "Time"
    if isinstance(w, Noise):
        return "Noise"
    if isinstance(w, Fixed):
        s = w.samples
        return f"Fixed({list(s)!r})" if len(s) <= 10 else f"Fixed([...], len={len(s)})"
    if isinstance(w, Fin):
        return f"Fin({show(w.length)}, {show(w.waveform)})"
    if isinstance(w, Append):
        return f"Append({show(w.a)}, {show(w.b)})"
    if isinstance(w, Sine):
        return f"Sine({show(w.frequency)}, {show(w.phase)})"
    if isinstance(w, Filter):
        return "Filter(%s, [%s], [%s])" % (
            show(w.waveform),
            ", ".join(map(show, w.feed_forward)),
            ", ".join(map(show, w.feedback)),
        )
    if isinstance(w, BinaryPointOp):
        return f"{Operator(w.op).name}({show(w.a)}, {show(w.b)})"
    if isinstance(w, Reset):
        return f"Reset({show(w.trigger)}, {show(w.waveform)})"
    if isinstance(w, Alt):
        return f"Alt({show(w.trigger)}, {show(w.positive_waveform)}, {show(w.negative_waveform)})"
    if isinstance(w, Marked):
        return f"Marked({w.id}, {show(w.waveform)})"
    if isinstance(w, Captured):
        return f"Captured({w.file_stem}, {show(w.waveform)})"
    raise TypeError(type(w))


class TbNode(ctypes.Structure):
    """`struct tb_node` of include/tuun_b200.h (64 bytes)."""

    _fields_ = [
        ("kind", ctypes.c_uint32),
        ("op", ctypes.c_uint32),
        ("a", ctypes.c_int32),
        ("b", ctypes.c_int32),
        ("c", ctypes.c_int32),
        ("value", ctypes.c_float),
        ("param_slot", ctypes.c_int32),
        ("list_off", ctypes.c_uint32),
        ("ff_count", ctypes.c_uint32),
        ("fb_count", ctypes.c_uint32),
        ("mark_id", ctypes.c_uint32),
        ("reserved", ctypes.c_uint32),
        ("fixed_off", ctypes.c_uint64),
        ("fixed_len", ctypes.c_uint64),
    ]


assert ctypes.sizeof(TbNode) == 64


@dataclass
class OpList:
    """The flat program: nodes in topological order (children first, root last)."""

    nodes: ctypes.Array
    lists: np.ndarray  # int32
    fixed_pool: np.ndarray  # float32
    n_params: int

    @property
    def n_nodes(self) -> int:
        return len(self.nodes)


def flatten(root: Waveform) -> OpList:
    nodes: List[TbNode] = []
    lists: List[int] = []
    pool: List[np.ndarray] = []
    pool_len = 0
    n_params = 0
    stems = {}

    def emit(**kw) -> int:
        n = TbNode(kind=0, op=0, a=-1, b=-1, c=-1, value=0.0, param_slot=-1)
        for k, v in kw.items():
            setattr(n, k, v)
        nodes.append(n)
        return len(nodes) - 1

    def go(w: Waveform) -> int:
        nonlocal pool_len, n_params
        if isinstance(w, Const):
            n_params = max(n_params, w.param + 1)
            return emit(kind=Kind.Const, value=w.value, param_slot=w.param)
        if isinstance(w, Time):
            return emit(kind=Kind.Time)
        if isinstance(w, Noise):
            return emit(kind=Kind.Noise)
        if isinstance(w, Fixed):
            off = pool_len
            pool.append(w.samples)
            pool_len += len(w.samples)
            return emit(kind=Kind.Fixed, fixed_off=off, fixed_len=len(w.samples))
        if isinstance(w, Fin):
            a = go(w.length)
            b = go(w.waveform)
            return emit(kind=Kind.Fin, a=a, b=b)
        if isinstance(w, Append):
            a = go(w.a)
            b = go(w.b)
            return emit(kind=Kind.Append, a=a, b=b)
        if isinstance(w, Sine):
            a = go(w.frequency)
            b = go(w.phase)
            return emit(kind=Kind.Sine, a=a, b=b)
        if isinstance(w, Filter):
            if len(w.feed_forward) < 1:
                raise ValueError("Filter needs at least one feed-forward coefficient (generator.rs:233)")
            a = go(w.waveform)
            idx = [go(c) for c in w.feed_forward] + [go(c) for c in w.feedback]
            off = len(lists)
            lists.extend(idx)
            return emit(kind=Kind.Filter, a=a, list_off=off, ff_count=len(w.feed_forward),
                        fb_count=len(w.feedback))
        if isinstance(w, BinaryPointOp):
            a = go(w.a)
            b = go(w.b)
            return emit(kind=Kind.BinaryPointOp, op=int(w.op), a=a, b=b)
        if isinstance(w, Reset):
            a = go(w.trigger)
            b = go(w.waveform)
            return emit(kind=Kind.Reset, a=a, b=b)
        if isinstance(w, Alt):
            a = go(w.trigger)
            b = go(w.positive_waveform)
            c = go(w.negative_waveform)
            return emit(kind=Kind.Alt, a=a, b=b, c=c)
        if isinstance(w, Marked):
            a = go(w.waveform)
            return emit(kind=Kind.Marked, a=a, mark_id=int(w.id))
        if isinstance(w, Captured):
            a = go(w.waveform)
            sid = stems.setdefault(w.file_stem, len(stems))
            return emit(kind=Kind.Captured, a=a, mark_id=sid)
        raise TypeError(f"not a Waveform: {type(w)}")

    import sys

    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 20000))
    try:
        go(root)
    finally:
        sys.setrecursionlimit(old)
    arr = (TbNode * len(nodes))(*nodes)
    return OpList(
        nodes=arr,
        lists=np.asarray(lists, dtype=np.int32),
        fixed_pool=(np.concatenate(pool).astype(np.float32) if pool else np.zeros(0, np.float32)),
        n_params=n_params,
    )


# Convenience constructors used by the config builders and tests.
def add(a, b):
    return BinaryPointOp(Operator.Add, a, b)


def sub(a, b):
    return BinaryPointOp(Operator.Subtract, a, b)


def mul(a, b):
    return BinaryPointOp(Operator.Multiply, a, b)


def div(a, b):
    return BinaryPointOp(Operator.Divide, a, b)


def merge(a, b):
    return BinaryPointOp(Operator.Merge, a, b)


def power(a, b):
    return BinaryPointOp(Operator.Power, a, b)
