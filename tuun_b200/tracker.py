"""Offline tracker: the scheduling and mix loop of `Tracker::generate` (src/lib/tracker.rs:484-644)
over the GPU renderer, with 32-bit-float mono WAV output (tracker.rs:207-226 / hound's WavSpec).

The reference tracker is an audio callback: every buffer it promotes pending waveforms whose start
time has come, splits the buffer into segments at sample-accurate start times, asks each active
waveform for the segment (`Generator::generate`), adds the results into the pre-zeroed buffer in
activation order, and retires waveforms that return short.  Its clock is the wall clock
(`Instant::now()`, tracker.rs:326) — batch mode (`--ui=false`, main.rs:91-174) spins the same
callback — so the only deterministic part is `generate(buffer_start, out)`; this module mirrors that
function on a virtual clock that advances by exactly one buffer per callback.

What runs on the GPU: every waveform is rendered WHOLE when it is activated (`Generator::length` for
its extent, then one tb_render of exactly that many samples; one program per waveform object, so a
repeating waveform is re-rendered, not re-lowered; streaming it segment by segment would be one tiny
launch per waveform per segment).  The reference
generator is block-size invariant for the trees the reference itself tests that way (chunks of
1/2/4/8, generator.rs:1284-1351), so handing out consecutive slices of that render is what the
per-segment `generate` calls would have produced — with two exceptions the reference's code has and its
tests do not reach: a Fin whose length is a rendered, non-monotonic waveform is re-polled per call
(generator.rs:672-687), and a finite input under a Filter; such trees come out as
ONE-call renders here.  Discarding a prefix is the late-start catch-up (tracker.rs:509-531), except that
the reference still writes the discarded prefix of a Captured node to its file.  A waveform that has not
ended after `max_seconds` is cut there with a RuntimeWarning.  The adds stay in
activation order, each rounded, exactly like `out[filled + j] += tmp[j]` (tracker.rs:617-619).

Times are integer nanoseconds with Rust's `Duration` conversions (from_secs_f32 rounds to nearest,
as_secs_f32 = secs + nanos / 1e9 in f32), so segment boundaries are computed in the same arithmetic.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np

from .waveform import Captured, Marked, Waveform

F = np.float32
NANOS = 1_000_000_000


def from_secs_f32(x) -> int:
    """Duration::from_secs_f32 (rounds to the nearest nanosecond)."""
    x = float(F(x))
    if not (x >= 0.0):
        raise ValueError("Duration::from_secs_f32: negative or NaN")
    return int(round(x * NANOS))


def as_secs_f32(ns: int) -> np.float32:
    """Duration::as_secs_f32."""
    return F(F(ns // NANOS) + F(ns % NANOS) / F(NANOS))


@dataclass
class _Pending:
    id: object
    waveform: Waveform
    start: int  # ns
    repeat_every: Optional[int] = None


@dataclass
class _Active:
    id: object
    start: int
    samples: np.ndarray  # the whole render (host)
    cursor: int = 0
    captures: Dict[str, np.ndarray] = field(default_factory=dict)


def _captured_subtrees(w: Waveform, out: Dict[str, Waveform]):
    """process_captured (tracker.rs:147-226): every Captured node, duplicate stems are an error."""
    import dataclasses
    if isinstance(w, Captured):
        if w.file_stem in out:
            raise ValueError(f"Captured waveform with duplicate file stem: {w.file_stem}")
        out[w.file_stem] = w.waveform
    if dataclasses.is_dataclass(w):
        for f in dataclasses.fields(w):
            v = getattr(w, f.name)
            if isinstance(v, Waveform):
                _captured_subtrees(v, out)
            elif isinstance(v, list):
                for c in v:
                    if isinstance(c, Waveform):
                        _captured_subtrees(c, out)


class OfflineTracker:
    """`Tracker::new` + `Command::Play` + the audio callback, on a virtual clock."""

    def __init__(self, sample_rate: int = 44100, buffer_size: int = 1024, device: int = -1,
                 max_seconds: float = 600.0, render: Optional[Callable[[Waveform, int], np.ndarray]] = None):
        self.sample_rate = int(sample_rate)
        self.buffer_size = int(buffer_size)
        self.max_samples = int(max_seconds * sample_rate)
        self._render = render or self._gpu_render(device)
        self.pending: List[_Pending] = []   # sorted by start (tracker.rs:126)
        self.active: List[_Active] = []
        self.now = 0                        # ns: start of the next buffer
        self.captured: Dict[str, List[np.ndarray]] = {}
        self.launches = 0
        self._keep: List[Waveform] = []

    def _gpu_render(self, device):
        programs: Dict[int, object] = {}   # one program per waveform object: a repeating waveform is re-rendered, not re-lowered

        def render(w: Waveform, n: int) -> np.ndarray:
            from .generator import Program
            p = programs.get(id(w))
            if p is None:
                p = programs[id(w)] = Program(w, self.sample_rate, device=device)
                self._keep.append(w)  # id() stays unique while the waveform is alive
            before = int(p.info.kernel_launches)
            p.reset()
            length = int(p.lengths(1, n)[0])   # Generator::length: how much there is (at most n), no samples
            p.reset()
            out = np.zeros((1, max(length, 1)), dtype=np.float32)
            got = int(p.render(out[:, :length])[0]) if length else 0
            self.launches += int(p.info.kernel_launches) - before
            return out[0, :got]

        return render

    # -- Command::Play (tracker.rs:376-420) ---------------------------------------------------------
    def play(self, waveform: Waveform, start: float = 0.0, repeat_every: Optional[float] = None, id=None):
        """Schedule `waveform` to start `start` seconds after time zero (optionally repeating)."""
        p = _Pending(id if id is not None else len(self.pending) + len(self.active), waveform,
                     from_secs_f32(start), None if repeat_every is None else from_secs_f32(repeat_every))
        self.pending.append(p)
        self.pending.sort(key=lambda w: w.start)  # stable, like sort_by_key

    def _activate(self, pending: _Pending, segment_start: int) -> _Active:
        samples = self._render(pending.waveform, self.max_samples)
        if len(samples) >= self.max_samples:
            import warnings
            warnings.warn(f"waveform {pending.id} has not ended after {self.max_samples} samples (max_seconds): cut there",
                          RuntimeWarning, stacklevel=2)
        a = _Active(pending.id, pending.start, samples)
        caps: Dict[str, Waveform] = {}
        _captured_subtrees(pending.waveform, caps)
        root = pending.waveform
        while isinstance(root, Marked):
            root = root.waveform
        for stem, sub in caps.items():
            # The reference writes what flows through the Captured node (generator.rs:354-378); for
            # a capture at the top of a program (`... | capture("stem")`) that is the program's own
            # output.  A nested capture is rendered standalone for as long as the waveform runs.
            if isinstance(root, Captured) and root.file_stem == stem:
                a.captures[stem] = samples
            else:
                a.captures[stem] = self._render(sub, max(1, len(samples)))[:len(samples)]
        if pending.start < segment_start:  # late start: generate and discard (tracker.rs:509-531)
            # f32::round (tracker.rs:517) rounds halves away from zero; np.round would round them to even
            delta = int(np.floor(np.float64(F(as_secs_f32(segment_start - pending.start) * F(self.sample_rate))) + 0.5))
            a.cursor = min(delta, len(samples))
        return a

    # -- Tracker::generate (tracker.rs:484-644) -----------------------------------------------------
    def generate(self, buffer_start: int, out: np.ndarray) -> List[_Active]:
        sr = self.sample_rate
        segment_start = buffer_start
        segment_length = len(out)
        finished: List[_Active] = []
        out[:] = 0.0
        filled = 0
        while filled < len(out):
            while self.pending:
                if self.pending[0].start <= segment_start:
                    pending = self.pending.pop(0)
                    self.active.append(self._activate(pending, segment_start))
                    if pending.repeat_every is not None:
                        pending.start += pending.repeat_every
                        while pending.start <= segment_start:  # missed repetitions
                            pending.start += pending.repeat_every
                        self.pending.append(pending)
                        self.pending.sort(key=lambda w: w.start)
                else:
                    gap = F(as_secs_f32(self.pending[0].start - segment_start) * F(sr))
                    segment_length = min(segment_length, int(np.ceil(gap)))
                    break
            if self.active:
                i = 0
                while i < len(self.active):
                    a = self.active[i]
                    take = a.samples[a.cursor:a.cursor + segment_length]
                    n = len(take)
                    out[filled:filled + n] += take  # f32 adds in activation order (tracker.rs:617-619)
                    for stem, rows in a.captures.items():
                        self.captured.setdefault(stem, []).append(rows[a.cursor:a.cursor + n])
                    a.cursor += n
                    if n < segment_length:
                        finished.append(self.active.pop(i))
                    else:
                        i += 1
            filled += segment_length
            segment_start += from_secs_f32(F(segment_length) / F(sr))
            segment_length = len(out) - filled
        return finished

    def callback(self, out: np.ndarray):
        """One audio callback on the virtual clock: the buffer starts where the last one ended."""
        done = self.generate(self.now, out)
        self.now += from_secs_f32(F(len(out)) / F(self.sample_rate))
        return done

    def render_all(self, max_seconds: Optional[float] = None, max_samples: Optional[int] = None) -> np.ndarray:
        """Batch mode (main.rs:159-174): call the callback until nothing is active or pending
        (repeating waveforms never finish: bound them with `max_seconds` / `max_samples`)."""
        limit = max_samples if max_samples is not None else (
            None if max_seconds is None else int(round(max_seconds * self.sample_rate)))
        chunks = []
        total = 0
        while (self.active or self.pending) and (limit is None or total < limit):
            out = np.zeros(self.buffer_size, dtype=np.float32)
            self.callback(out)
            chunks.append(out)
            total += len(out)
        mix = np.concatenate(chunks) if chunks else np.zeros(0, np.float32)
        return mix if limit is None else mix[:limit]

    def captured_output(self) -> Dict[str, np.ndarray]:
        return {k: (np.concatenate(v) if v else np.zeros(0, np.float32)) for k, v in self.captured.items()}


def write_wav(path: str, samples: np.ndarray, sample_rate: int):
    """Mono 32-bit IEEE-float WAV — hound's WavSpec{channels: 1, bits_per_sample: 32,
    sample_format: Float} (tracker.rs:212-217).  The data chunk is the f32 samples, little endian;
    the header is the canonical 16-byte `fmt ` (format tag 3) plus a `fact` chunk."""
    data = np.ascontiguousarray(samples, dtype="<f4").tobytes()
    fmt = struct.pack("<HHIIHH", 3, 1, sample_rate, sample_rate * 4, 4, 32)
    fact = struct.pack("<I", len(data) // 4)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"fact" + struct.pack("<I", 4) + fact \
        + b"data" + struct.pack("<I", len(data)) + data
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)


def read_wav(path: str):
    """Inverse of write_wav (float32 mono): returns (samples, sample_rate)."""
    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE"
    pos, rate, samples = 12, None, None
    while pos + 8 <= len(raw):
        tag, size = raw[pos:pos + 4], struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        body = raw[pos + 8:pos + 8 + size]
        if tag == b"fmt ":
            fmt_tag, ch, rate, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            assert fmt_tag == 3 and ch == 1 and bits == 32
        elif tag == b"data":
            samples = np.frombuffer(body, dtype="<f4").astype(np.float32)
        pos += 8 + size + (size & 1)
    return samples, rate
