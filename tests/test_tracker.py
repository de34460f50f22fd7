"""tuun_b200/tracker.py against the reference's tracker loop written out literally
(tracker.rs:484-644: per-segment `generate` calls on streaming state), with the CPU oracle as the
generator on both sides; then the same schedule through the GPU renderer."""
import os

import numpy as np
import pytest

from oracle.binding import OracleProgram
from tuun_b200 import workloads as W
from tuun_b200.builder import Std, capture, to_waveform
from tuun_b200.optimizer import optimize
from tuun_b200.tracker import OfflineTracker, as_secs_f32, from_secs_f32, read_wav, write_wav
from tuun_b200.waveform import Const, Fin, Sine, Time, add, mul

SR = 44100
F = np.float32


def reference_tracker(schedule, n_buffers, buffer_size=1024):
    """Tracker::generate, literally: streaming generators, segment by segment."""
    pending = sorted([dict(w=w, start=from_secs_f32(s), rep=None if r is None else from_secs_f32(r))
                      for w, s, r in schedule], key=lambda p: p["start"])
    active = []
    now = 0
    chunks = []
    for _ in range(n_buffers):
        out = np.zeros(buffer_size, dtype=np.float32)
        segment_start, segment_length, filled = now, buffer_size, 0
        while filled < buffer_size:
            while pending:
                if pending[0]["start"] <= segment_start:
                    p = pending.pop(0)
                    o = OracleProgram(p["w"], SR)
                    if p["start"] < segment_start:
                        delta = int(np.round(F(as_secs_f32(segment_start - p["start"]) * F(SR))))
                        if delta > 0:
                            o.render(delta, block=delta)
                    active.append(o)
                    if p["rep"] is not None:
                        p["start"] += p["rep"]
                        while p["start"] <= segment_start:
                            p["start"] += p["rep"]
                        pending.append(p)
                        pending.sort(key=lambda q: q["start"])
                else:
                    segment_length = min(segment_length,
                                         int(np.ceil(F(as_secs_f32(pending[0]["start"] - segment_start) * F(SR)))))
                    break
            i = 0
            while i < len(active):
                tmp = active[i].render(segment_length, block=segment_length)
                out[filled:filled + len(tmp)] += tmp
                if len(tmp) < segment_length:
                    active.pop(i)
                else:
                    i += 1
            filled += segment_length
            segment_start += from_secs_f32(F(segment_length) / F(SR))
            segment_length = buffer_size - filled
        now += from_secs_f32(F(buffer_size) / F(SR))
        chunks.append(out)
    return np.concatenate(chunks)


def schedule():
    s = Std()
    blip = Fin(add(Time(), Const(-0.05)), mul(Sine(Const(5000.0), Const(0.0)), Const(0.25)))
    return [
        (W.cfg1_from_source(), 0.0, None),
        (optimize(to_waveform(s.harmonica(s.Q, 330))), 0.1234, None),  # starts mid-buffer
        (blip, 0.05, 0.3),                                              # repeats every 0.3 s
        (W.cfg1_from_source(), 0.7001, None),
    ]


def oracle_render(w, n):
    return OracleProgram(w, SR).render(n, block=1024)


def test_duration_arithmetic():
    assert from_secs_f32(0.5) == 500_000_000 and from_secs_f32(0.999e-9) == 1
    assert as_secs_f32(1_500_000_000) == F(1.5)
    assert from_secs_f32(F(1024) / F(44100)) == 23_219_954  # one buffer at 44.1 kHz


def test_schedule_matches_literal_tracker_loop():
    n_buffers = 60  # 1.39 s
    ref = reference_tracker(schedule(), n_buffers)
    t = OfflineTracker(SR, 1024, render=oracle_render, max_seconds=5)
    for w, s, r in schedule():
        t.play(w, s, r)
    got = t.render_all(max_samples=n_buffers * 1024)
    assert len(got) == n_buffers * 1024
    np.testing.assert_array_equal(got, ref)
    # the harmonica note starts exactly ceil(0.1234 s * 44100) = 5442 samples in (a mid-buffer split)
    t2 = OfflineTracker(SR, 1024, render=oracle_render, max_seconds=5)
    t2.play(schedule()[1][0], 0.1234)
    alone = t2.render_all()
    first = int(np.ceil(F(as_secs_f32(from_secs_f32(0.1234)) * F(SR))))
    assert first == 5442 and not alone[:first].any()
    np.testing.assert_array_equal(alone[first:first + 22050], oracle_render(schedule()[1][0], 30000))


def test_late_start_discards_and_batch_mode_ends():
    t = OfflineTracker(SR, 1024, render=oracle_render, max_seconds=5)
    buf = np.zeros(1024, dtype=np.float32)
    for _ in range(5):
        t.callback(buf)                      # 5 silent buffers pass
    assert not buf.any()
    w = W.cfg1_from_source()
    t.play(w, 0.05)                          # should have started 0.0661 s ago
    t.callback(buf)
    late = int(np.round(F(as_secs_f32(t.now - from_secs_f32(F(1024) / F(SR)) - from_secs_f32(0.05)) * F(SR))))
    whole = oracle_render(w, 30000)
    np.testing.assert_array_equal(buf, whole[late:late + 1024])
    rest = t.render_all()
    assert len(whole) - late - 1024 <= len(rest) < len(whole) - late - 1024 + 1024
    assert not t.active and not t.pending


def test_capture_and_wav_roundtrip(tmp_path):
    t = OfflineTracker(SR, 1024, render=oracle_render, max_seconds=5)
    prog = dict(W.cfg3_fm_variations())["true-fm"]
    w = capture("true-fm")(Fin(add(Time(), Const(-0.2)), prog))
    t.play(w, 0.0)
    mix = t.render_all()
    cap = t.captured_output()["true-fm"]
    assert len(cap) == 8820 and np.array_equal(cap, mix[:8820]) and not mix[8820:].any()
    path = os.path.join(tmp_path, "true-fm.wav")
    write_wav(path, cap, SR)
    back, rate = read_wav(path)
    assert rate == SR and np.array_equal(back, cap)
    raw = open(path, "rb").read()
    assert raw[20:22] == bytes([3, 0]) and raw[22:24] == bytes([1, 0]) and raw[34:36] == bytes([32, 0])  # float, mono, 32 bit
    import scipy.io.wavfile
    r2, d2 = scipy.io.wavfile.read(path)
    assert r2 == SR and d2.dtype == np.float32 and np.array_equal(d2, cap)


@pytest.mark.gpu
def test_gpu_tracker_schedule_within_tolerance():
    n_buffers = 60
    ref = reference_tracker(schedule(), n_buffers)
    t = OfflineTracker(SR, 1024, max_seconds=5)
    for w, s, r in schedule():
        t.play(w, s, r)
    got = t.render_all(max_samples=n_buffers * 1024)
    assert len(got) == len(ref)
    assert np.max(np.abs(got - ref)) <= 4e-4  # up to four waveforms overlap, 1e-4 each
    assert t.launches >= 4
