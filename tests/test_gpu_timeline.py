"""Instruments of lib/v0/std.tuun as large batches on the lane-per-voice kernels (tuun_b200/csrc/lanes.cuh):

* a Reset nested in a Reset — hard sync, `reset(pulse(f1), pulse(f2))`: the inner clock restarts with the outer
  one (ST_RESET_CLK with an enclosing clock slot; generator.rs:273-318 and set_state, waveform.rs:322-392);
* a timeline — `Append(Fin{T - c0, e0}, Append(Fin{T - c1, e1}, ..))` under a root Fin, every length a literal:
  the ADSR envelopes (ST_SEG_CLK / ST_SEG_SEL, program.h; generator.rs:133-188).

Config 2 (the harmonica: both of the above, a biquad, a sequence of four notes) is the tree these were built for.
Everything is compared with the CPU oracle; the general interpreter renders the first tile and the last samples of
every call on the same state blocks, so streaming in odd block sizes checks the state the lane kernels leave behind
for it at every kind of position (inside a piece, on a piece's first sample, after the note's end)."""
import math

import numpy as np
import pytest

from oracle.binding import OracleProgram
from tuun_b200 import workloads as W
from tuun_b200.waveform import Alt, Append, Const, Fin, Reset, Sine, Time, add, f32, mul, sub

pytestmark = pytest.mark.gpu
SR = 44100
TOL = 1e-4  # north_star
TAU = f32(2 * math.pi)


def program(w, monkeypatch, lanes=True):
    from tuun_b200.generator import Program
    monkeypatch.setenv("TUUN_B200_SPLIT", "0")
    monkeypatch.setenv("TUUN_B200_LANES", "1" if lanes else "0")
    if lanes:
        monkeypatch.setenv("TUUN_B200_LANE_MIN_VOICES", "1")
    return Program(w, SR)


def oracle_rows(w, params, V, N):
    o = OracleProgram(w, SR)
    rows = np.zeros((V, N), dtype=np.float32)
    lens = np.zeros(V, dtype=np.int64)
    for v in range(V):
        o.initialize_state()
        if params is not None:
            o.set_params(params[v])
        r = o.render(N)
        rows[v, :len(r)] = r
        lens[v] = len(r)
    return rows, lens


def streamed(p, V, N, blocks, params=None):
    """The stream in blocks, the way the tracker asks for it; rows and per-voice lengths."""
    rows = np.zeros((V, N), dtype=np.float32)
    lens = np.zeros(V, dtype=np.int64)
    a = 0
    live = np.ones(V, dtype=bool)
    for n in blocks:
        n = min(n, N - a)
        if n <= 0:
            break
        blk = np.zeros((V, n), dtype=np.float32)
        got = np.asarray(p.render(blk, params=params)).astype(np.int64)
        for v in np.nonzero(live)[0]:
            rows[v, a:a + got[v]] = blk[v, :got[v]]
        lens[live] += got[live]
        live &= got == n
        a += n
    return rows, lens


def saw(fp, rp):
    """lib/v0/std.tuun sawtooth: 2 (reset($f, -f T) + 0.5), f a per-voice parameter."""
    return mul(add(Reset(Sine(Const(1.0, param=fp), Const(0.0)), mul(Time(), Const(1.0, param=rp))), Const(0.5)), Const(2.0))


def test_hard_sync_batch(monkeypatch):
    from tuun_b200.generator import lower_check
    V, N = 200, 256 + 16 * 400 + 9
    rng = np.random.default_rng(5)
    f1 = rng.uniform(60.0, 900.0, V).astype(np.float32)
    f2 = (f1 * rng.uniform(1.2, 3.7, V).astype(np.float32)).astype(np.float32)
    params = np.stack([TAU * f1, -f1, TAU * f2, -f2], axis=1).astype(np.float32)
    master = Alt(sub(saw(0, 1), Const(0.93)), Const(1.0), Const(-1.0))
    slave = Alt(add(saw(2, 3), Const(0.3)), Const(1.0), Const(-1.0))
    ramp = mul(saw(2, 3), Const(0.5))
    for name, w in (("sync pulse", Reset(master, slave)), ("sync saw", Reset(master, ramp)),
                    ("sync of a sine by a sine", Reset(Sine(Const(1.0, param=0), Const(0.0)),
                                                       mul(Sine(Const(1.0, param=2), Const(0.25)), add(Time(), Const(0.5)))))):
        info = lower_check(w)
        assert info.lane_smem_bytes > 0, name
        ref, _ = oracle_rows(w, params, V, N)
        p = program(w, monkeypatch)
        out = np.zeros((V, N), dtype=np.float32)
        lens = p.render(out, params=params)
        assert (np.asarray(lens) == N).all() and p.info.lane_launches == 1, name
        # a trigger within rounding of zero may move an edge by one sample (hard part 1); under a nested
        # Reset that moves the run behind it: count voices, not samples
        bad_voices = int(np.count_nonzero((np.abs(out - ref) > TOL).any(axis=1)))
        assert bad_voices <= 1, (name, bad_voices)
        q = program(w, monkeypatch)
        rows, _ = streamed(q, V, N, (1000, 272, 3001, 16, 4000), params)
        assert int(np.count_nonzero((np.abs(rows - ref) > TOL).any(axis=1))) <= 1, name
        assert q.info.lane_launches >= 3


def test_timeline_envelopes(monkeypatch):
    """Attack / decay / sustain as a timeline over a per-voice sine; the last piece infinite (a held level) or one more
    Fin; a tremolo inside a piece; the note ends inside the timeline."""
    from tuun_b200.generator import lower_check
    V, N = 150, 20000
    rng = np.random.default_rng(6)
    f = rng.uniform(100.0, 2000.0, V).astype(np.float32)
    params = (TAU * f).reshape(V, 1).astype(np.float32)
    tone = lambda: Sine(Const(1.0, param=0), Const(0.0))
    held = Append(Fin(add(Time(), Const(-0.1)), mul(Time(), Const(10.0))),
                  Append(Fin(add(Time(), Const(-0.15)), add(mul(Time(), Const(-2.0)), Const(1.0))), Const(0.7)))
    trem = Append(Fin(add(Time(), Const(-0.05)), mul(Time(), Const(20.0))),
                  Append(Fin(add(Time(), Const(-0.2)), add(mul(Sine(Const(f32(TAU * 6)), Const(0.0)), Const(0.2)), Const(0.8))),
                         Fin(add(Time(), Const(-0.3)), add(mul(Time(), Const(-1.0)), Const(0.8)))))
    for name, env, dur in (("held", held, 0.4), ("tremolo", trem, 0.41), ("short note", trem, 0.12)):
        w = Fin(add(Time(), Const(-f32(dur))), mul(tone(), env))
        info = lower_check(w)
        assert info.lane_smem_bytes > 0, name
        ref, rlens = oracle_rows(w, params, V, N)
        p = program(w, monkeypatch)
        out = np.zeros((V, N), dtype=np.float32)
        lens = np.asarray(p.render(out, params=params))
        assert (lens == rlens).all() and p.info.lane_launches == 1, (name, lens[:4], rlens[:4])
        for v in range(V):
            assert np.max(np.abs(out[v, :lens[v]] - ref[v, :lens[v]])) <= 2e-5, (name, v)
        # blocks that end inside every piece, on the first sample of one (0.05 s = 2,205; 0.25 s = 11,025), past the end
        for blocks in ((2205, 272, 8548, 1024, 1024, 1024, 1024, 1024, 1024, 1024, 1024, 1024, 1024),
                       (1000, 300, 1500, 5000, 3225, 4000, 4000, 4000)):
            q = program(w, monkeypatch)
            rows, slens = streamed(q, V, N, blocks, params)
            assert (slens == rlens).all(), (name, blocks[:3], slens[:4], rlens[:4])
            for v in range(V):
                assert np.max(np.abs(rows[v, :slens[v]] - ref[v, :slens[v]])) <= 2e-5, (name, v, blocks[:3])
    # a timeline that would end inside the note stays on the general interpreter (the zero tail of the product)
    w = Fin(add(Time(), Const(-1.5)), mul(tone(), trem))
    assert lower_check(w).lane_smem_bytes == 0


def test_harmonica_note_and_sequence(monkeypatch):
    """Config 2: one harmonica note as a batch through the lane kernels, then the four-note sequence (one program per
    part, each on the lane kernels)."""
    from tuun_b200.generator import lower_check
    V = 130
    seq = W.cfg2_harmonica(4)
    note = seq.a
    assert lower_check(note).lane_smem_bytes > 0
    ref = OracleProgram(note, SR).render(30000, block=1024)
    assert len(ref) == 22050
    p = program(note, monkeypatch)
    out = np.zeros((V, 30000), dtype=np.float32)
    lens = np.asarray(p.render(out))
    assert (lens == 22050).all() and p.info.lane_launches == 1
    err = np.abs(out[:, :22050] - ref[None, :])
    assert float(err.max()) <= TOL, float(err.max())
    assert np.array_equal(out[0, :22050].view(np.uint32), out[V - 1, :22050].view(np.uint32))
    # against the general interpreter on the same device
    g = program(note, monkeypatch, lanes=False)
    gen = np.zeros((V, 30000), dtype=np.float32)
    g.render(gen)
    assert g.info.lane_launches == 0
    assert float(np.max(np.abs(gen[:, :22050] - out[:, :22050]))) <= TOL
    # streamed: 5,733 and 20,287 are the first samples of the decay and of the sustain ramp
    for blocks in ((1024,) * 30, (5733, 14554, 1000, 763, 512), (300, 5000, 433, 16000, 317, 9000)):
        q = program(note, monkeypatch)
        rows, slens = streamed(q, V, 30000, blocks)
        assert (slens == 22050).all(), (blocks[:3], slens[:4])
        assert float(np.max(np.abs(rows[:, :22050] - ref[None, :]))) <= TOL, blocks[:3]
    # the sequence
    sref = OracleProgram(seq, SR).render(100000, block=1024)
    assert len(sref) == 88200
    s = program(seq, monkeypatch)
    sout = np.zeros((V, 100000), dtype=np.float32)
    slens = np.asarray(s.render(sout))
    assert (slens == 88200).all() and s.info.sequence_parts == 4 and s.info.lane_launches >= 4
    assert float(np.max(np.abs(sout[:, :88200] - sref[None, :]))) <= TOL


def test_random_timelines_streamed(monkeypatch):
    """Seeded random timelines — 2 to 5 pieces of random literal lengths, each a ramp, a level, a tremolo or a square
    LFO, the last one a Fin or a held level — over a per-voice tone that a hard-synced pulse may replace, under a note
    that ends inside the timeline; rendered whole and in random block sizes (blocks end inside pieces, on their first
    samples and past the note's end) against the oracle."""
    from tuun_b200.generator import lower_check
    rng = np.random.default_rng(20261019)
    V = 70
    f = rng.uniform(80.0, 1500.0, V).astype(np.float32)
    g = (f * rng.uniform(1.3, 2.9, V).astype(np.float32)).astype(np.float32)
    params = np.stack([TAU * f, -f, TAU * g, -g], axis=1).astype(np.float32)

    def piece_tree(kind, level):
        if kind == 0:
            return add(mul(Time(), Const(f32(rng.uniform(-3.0, 3.0)))), Const(f32(level)))
        if kind == 1:
            return Const(f32(level))
        if kind == 2:
            return add(mul(Sine(Const(f32(TAU * rng.uniform(3.0, 9.0))), Const(f32(rng.uniform(0.0, 1.0)))), Const(0.2)), Const(f32(level)))
        return Alt(Sine(Const(f32(TAU * rng.uniform(4.0, 12.0))), Const(0.0)), Const(f32(level)), mul(Time(), Const(2.0)))

    taken = 0
    for case in range(14):
        n_pieces = int(rng.integers(2, 6))
        lens = [float(np.float32(rng.uniform(0.004, 0.09))) for _ in range(n_pieces)]
        held = bool(rng.integers(0, 2))
        parts = [Fin(add(Time(), Const(-f32(d))), piece_tree(int(rng.integers(0, 4)), rng.uniform(0.2, 1.0))) for d in lens]
        env = Const(f32(0.6)) if held else parts.pop()
        for p in reversed(parts):
            env = Append(p, env)
        total = sum(lens[:len(parts)]) + (10.0 if held else lens[-1])
        dur = float(np.float32(rng.uniform(0.3, 0.95) * min(total, 0.3)))
        if rng.integers(0, 2):
            tone = Sine(Const(1.0, param=0), Const(0.0))
        else:
            tone = Reset(Alt(add(saw(0, 1), Const(-0.9)), Const(1.0), Const(-1.0)), mul(saw(2, 3), Const(0.5)))
        w = Fin(add(Time(), Const(-f32(dur))), mul(tone, env))
        N = int(dur * SR) + 700
        ref, rlens = oracle_rows(w, params, V, N)
        if lower_check(w).lane_smem_bytes == 0:
            continue  # (a note rounded past the end of its timeline)
        taken += 1
        p = program(w, monkeypatch)
        out = np.zeros((V, N), dtype=np.float32)
        got = np.asarray(p.render(out, params=params)).astype(np.int64)
        assert (got == rlens).all(), (case, got[:3], rlens[:3])
        assert p.info.lane_launches >= 1 or N < 272 + 16, case
        blocks = []
        while sum(blocks) < N:
            blocks.append(int(rng.choice([16, 100, 256, 272, 300, 1024, 1500, 4000])))
        q = program(w, monkeypatch)
        rows, slens = streamed(q, V, N, blocks, params)
        assert (slens == rlens).all(), (case, blocks[:4], slens[:3], rlens[:3])
        for name, x, xl in (("whole", out, got), ("streamed", rows, slens)):
            bad = 0
            for v in range(V):
                bad += int((np.abs(x[v, :xl[v]] - ref[v, :xl[v]]) > TOL).any())
            # an edge of a pulse or of a square LFO within rounding of zero moves by a sample (hard part 1)
            assert bad <= 1, (case, name, bad, blocks[:4])
    assert taken >= 10


def test_clocked_batch_at_the_default_threshold(monkeypatch):
    """Programs with clocked words take the lane kernels from 4,096 voices by default (no environment switches here):
    rows, and the mixdown without rows (tb_render_mix on the chip), of a pulse | lpf batch of 4,200 voices against the
    rows the general interpreter renders."""
    import os
    from tuun_b200.generator import Program
    from tuun_b200.workloads import lpf
    for k in ("TUUN_B200_LANE_MIN_VOICES", "TUUN_B200_LANES", "TUUN_B200_STEADY", "TUUN_B200_SPLIT"):
        monkeypatch.delenv(k, raising=False)
    V, N = 4200, 256 + 16 * 120 + 3
    rng = np.random.default_rng(11)
    f = rng.uniform(40.0, 1500.0, V).astype(np.float32)
    params = np.stack([TAU * f, -f], axis=1).astype(np.float32)
    w = lpf(Alt(sub(saw(0, 1), Const(0.3)), Const(1.0), Const(-1.0)), 0.707, 2000)
    p = Program(w, SR)
    assert p.info.lane_min_voices == 4096
    rows = np.zeros((V, N), dtype=np.float32)
    p.render(rows, params=params)
    assert p.info.lane_launches == 1
    monkeypatch.setenv("TUUN_B200_LANES", "0")
    g = Program(w, SR)
    monkeypatch.delenv("TUUN_B200_LANES")
    gen = np.zeros((V, N), dtype=np.float32)
    g.render(gen, params=params)
    assert g.info.lane_launches == 0
    bad = int(np.count_nonzero((np.abs(rows - gen) > TOL).any(axis=1)))
    assert bad <= 2, bad  # (an edge within rounding of zero may move by a sample on one of the two kernels)
    q = Program(w, SR)
    mix = np.zeros(N, dtype=np.float32)
    q.render_mix(mix, V, params=params)
    assert q.info.lane_launches >= 1
    want = rows.astype(np.float64).sum(axis=0)
    assert np.max(np.abs(mix - want)) <= 2e-3 * max(1.0, float(np.max(np.abs(want))))
