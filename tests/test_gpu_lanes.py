"""The lane-per-voice kernel (tuun_b200/csrc/lanes.cu): large batches of steady-state voices, one
thread per voice.  Forced on small batches here (TUUN_B200_LANE_MIN_VOICES=1) and compared with the
CPU oracle (generator.rs:86-515 restated), with the warp-per-voice kernel, and with itself across
different call partitions.  Filters run the reference's own recurrence in its own operation order
there, so filters over exactly representable inputs are bit-identical to the oracle."""
import math

import numpy as np
import pytest

from oracle.binding import OracleProgram
from tuun_b200.waveform import (Alt, BinaryPointOp, Const, Filter, Noise, Operator, Sine, Time, add, f32, mul, sub)

pytestmark = pytest.mark.gpu
SR = 44100
TOL = 1e-4  # north_star
TAU = f32(2 * math.pi)


def program(w, monkeypatch, lanes=True):
    from tuun_b200.generator import Program
    monkeypatch.setenv("TUUN_B200_SPLIT", "0")  # these tests count launches of the serial forms (split: test_gpu_split.py)
    if lanes:
        monkeypatch.setenv("TUUN_B200_LANES", "1")
        monkeypatch.setenv("TUUN_B200_LANE_MIN_VOICES", "1")
    else:
        monkeypatch.setenv("TUUN_B200_LANES", "0")
    return Program(w, SR)


def oracle_rows(w, params, V, N, seed=None):
    o = OracleProgram(w, SR)
    rows = np.zeros((V, N), dtype=np.float32)
    for v in range(V):
        o.initialize_state()
        if seed is not None:
            o.seed_noise(seed, v)
        if params is not None:
            o.set_params(params[v])
        r = o.render(N)
        assert len(r) == N
        rows[v] = r
    return rows


def cfg5(V):
    from tuun_b200.workloads import fm_filter_params, fm_filter_sample_ids, fm_filter_voice
    ids = fm_filter_sample_ids(V)  # every bit field of the sweep changes from one pick to the next
    return fm_filter_voice(), fm_filter_params(ids)


def within(out, ref, params):
    """Config 5 against the oracle: 1e-4; 1.5e-4 for the two filter shapes whose own round-off noise reaches
    it (workloads.fm_filter_tolerance: the 200 Hz low-passes with Q = 1.75 and 2)."""
    from tuun_b200.workloads import fm_filter_tolerance
    err = np.max(np.abs(out - ref), axis=1)
    tol = fm_filter_tolerance(params, TOL)
    bad = np.nonzero(err > tol)[0]
    assert len(bad) == 0, (bad[:5], err[bad[:5]], tol[bad[:5]])
    return float(err.max())


def test_cfg5_batch_against_oracle_and_warp_kernel(monkeypatch):
    V, N = 1000, 256 + 16 * 220 + 7  # ragged: a partial CTA, a partial warp, a tail shorter than a tile
    w, params = cfg5(V)
    p = program(w, monkeypatch)
    assert p.info.lane_smem_bytes > 0 and p.info.lane_min_voices == 1
    out = np.full((V, N), np.inf, dtype=np.float32)
    lens = p.render(out, params=params)
    # one launch: the fused-FM-voice kernel starts the stream itself and takes the 7 samples past the last tile
    assert p.info.lane_launches == 1 and p.info.kernel_launches == 1
    assert (lens == N).all()
    o = OracleProgram(w, SR)
    ref, olens, _, _ = o.render_batch(params, V, N)
    assert (olens == N).all()
    within(out, ref, params)
    q = program(w, monkeypatch, lanes=False)
    assert q.info.lane_smem_bytes == 0
    warp = np.zeros((V, N), dtype=np.float32)
    q.render(warp, params=params)
    assert q.info.lane_launches == 0
    within(out, warp, params)  # high-Q, low-cutoff voices: the round-off noise of the two feedback forms


def test_cfg5_one_second_drift(monkeypatch):
    """A whole second of the headline voices: the error against the oracle at the end of the render
    is the error at its start (the 64-bit phase sums are exact, nothing accumulates)."""
    V, N = 128, SR
    w, params = cfg5(V)
    out = np.zeros((V, N), dtype=np.float32)
    program(w, monkeypatch).render(out, params=params)
    ref, _, _, _ = OracleProgram(w, SR).render_batch(params, V, N)
    within(out, ref, params)
    d = np.abs(out - ref)
    assert d[:, -4410:].max() <= max(2.0 * d[:, :4410].max(), 2e-5)


def test_partition_invariance_and_device_rows(monkeypatch):
    """One call, or the same samples over several calls of odd sizes (each call: general head tile,
    lane tiles, general tail): the carried state makes them the same stream.  Device rows (torch)
    take the same kernels as host rows."""
    import torch
    V, N = 200, 6000
    w, params = cfg5(V)
    one = np.zeros((V, N), dtype=np.float32)
    program(w, monkeypatch).render(one, params=params)
    p = program(w, monkeypatch)
    parts = np.zeros((V, N), dtype=np.float32)
    a = 0
    for n in (1000, 273, 16, 2500, 2211):
        blk = np.zeros((V, n), dtype=np.float32)
        lens = p.render(blk, params=params)
        assert (lens == n).all()
        parts[:, a:a + n] = blk
        a += n
    assert a == N
    # Every call renders its first 256 samples and its last < 16 with the general tiles, whose FAST
    # sines round the same exact phase differently (32 top bits of 2^-44-turn sums there, 23 bits of
    # 2^-32-turn sums in lane tiles): inputs of the filters differ by ~1e-6, which for most voices is
    # what the outputs differ by; the high-Q 200 Hz low-passes (poles at radius 0.993) turn any
    # input change into another realisation of their round-off noise (6e-5).
    per_voice = np.max(np.abs(parts - one), axis=1)
    within(parts, one, params)
    assert np.median(per_voice) <= 5e-6
    dev = torch.zeros((V, N + 8), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()  # (the fill runs on torch's stream, the renders on the program's own)
    q = program(w, monkeypatch)
    q.render(dev[:, :N], params=torch.from_numpy(params).cuda())
    torch.cuda.synchronize()
    assert q.info.lane_launches == 1
    np.testing.assert_array_equal(dev[:, :N].cpu().numpy(), one)
    assert float(dev[:, N:].abs().max()) == 0.0


def test_filters_bit_exact_on_exact_inputs(monkeypatch):
    """Time, constants, Noise and Alt of them are exact on both sides, so every filter shape below
    must reproduce the oracle's f32 recurrence bit for bit (generator.rs:496-507)."""
    x = add(mul(Noise(), Const(0.25)), Alt(sub(mul(Time(), Const(3.0)), Const(0.05)), Const(0.5), Const(-0.5)))
    shapes = [
        ([0.2, 0.3, 0.2], [-1.2, 0.5]),                       # biquad
        ([0.5], [-0.5]),                                      # tracker_benches.rs filter_1_1
        ([0.4, 0.1], [0.3]),
        ([0.1, 0.2, 0.3, 0.4], [-0.3, 0.2, -0.1]),            # tracker_benches.rs filter_4_3
        ([0.25, 0.25, 0.25, 0.25], []),                       # pure FIR
        ([0.1, 0.05, 0.02, 0.3, 0.1, 0.2, 0.05, 0.1, 0.08], [-0.2, 0.1, -0.05, 0.02]),  # K = 9, J = 4
    ]
    V, N = 70, 256 + 16 * 40 + 3
    for ff, fb in shapes:
        w = Filter(x, [Const(f32(c)) for c in ff], [Const(f32(c)) for c in fb])
        p = program(w, monkeypatch)
        p.seed_noise(7, 0)
        out = np.zeros((V, N), dtype=np.float32)
        p.render(out)
        assert p.info.lane_launches == 1
        ref = oracle_rows(w, None, V, N, seed=7)
        np.testing.assert_array_equal(out, ref, err_msg=f"{ff} {fb}")
    # a cascade, per-voice coefficients
    w = Filter(Filter(x, [Const(1.0, param=0), Const(0.3)], [Const(0.0, param=1)]),
               [Const(0.2), Const(0.0, param=2), Const(0.2)], [Const(-1.0), Const(0.0, param=3)])
    rng = np.random.default_rng(3)
    params = np.stack([rng.uniform(0.1, 0.9, V), rng.uniform(-0.8, 0.8, V), rng.uniform(0.1, 0.5, V),
                       rng.uniform(0.1, 0.6, V)], axis=1).astype(np.float32)
    p = program(w, monkeypatch)
    p.seed_noise(7, 0)
    out = np.zeros((V, N), dtype=np.float32)
    p.render(out, params=params)
    np.testing.assert_array_equal(out, oracle_rows(w, params, V, N, seed=7))


def test_every_steady_operator(monkeypatch):
    """A tree touching every word of the lane program: all four Sine forms in both precision
    classes, Alt with waveform and constant branches, point operators with waveform and constant
    right-hand sides (incl. Divide and Power), Time, Noise, slots."""
    trig = Sine(Const(1.0, param=0), Const(0.3))                                   # CC, feeds a trigger: exact
    fm = Sine(add(mul(Time(), Const(900.0)), Const(1.0, param=1)),                 # AA: frequency and phase waveforms
              mul(Sine(Const(1.0, param=2), Const(0.5)), Const(2.0)))
    pm = Sine(Const(1.0, param=1), mul(trig, Const(3.0)))                          # CA
    ac = Sine(add(mul(fm, Const(500.0)), Const(2000.0)), Const(0.25))              # AC; makes fm exact-class
    alt = Alt(trig, add(ac, pm), mul(pm, Const(-0.5)))
    sq = Alt(Sine(Const(1.0, param=2), Const(0.0)), Const(1.0), Const(-1.0))
    body = add(alt, mul(sq, mul(Noise(), Const(0.1))))
    div = BinaryPointOp(Operator.Divide, body, add(Time(), Const(1.0)))
    pw = BinaryPointOp(Operator.Power, add(mul(div, Const(0.25)), Const(1.5)), Const(2.0))
    w = BinaryPointOp(Operator.Subtract, pw, BinaryPointOp(Operator.Divide, sq, Const(4.0)))
    V, N = 96, 256 + 16 * 64 + 11
    rng = np.random.default_rng(5)
    params = np.stack([TAU * rng.uniform(20, 300, V), TAU * rng.uniform(100, 1500, V), TAU * rng.uniform(30, 700, V)],
                      axis=1).astype(np.float32)
    p = program(w, monkeypatch)
    p.seed_noise(11, 0)
    out = np.zeros((V, N), dtype=np.float32)
    lens = p.render(out, params=params)
    assert p.info.lane_launches == 1 and (lens == N).all()
    ref = oracle_rows(w, params, V, N, seed=11)
    d = np.abs(out - ref)
    # Alt / square edges: a trigger within rounding of zero may flip one sample (hard part 1)
    bad = int(np.count_nonzero(d > TOL))
    assert bad <= 2, (bad, float(d.max()))
    q = program(w, monkeypatch, lanes=False)
    q.seed_noise(11, 0)
    warp = np.zeros((V, N), dtype=np.float32)
    q.render(warp, params=params)
    assert int(np.count_nonzero(np.abs(out - warp) > 2e-5)) <= 2


def test_default_threshold_keeps_small_batches_on_the_warp_kernel(monkeypatch):
    from tuun_b200.generator import Program
    monkeypatch.delenv("TUUN_B200_LANES", raising=False)
    monkeypatch.delenv("TUUN_B200_LANE_MIN_VOICES", raising=False)
    w, params = cfg5(64)
    p = Program(w, SR)
    assert p.info.lane_smem_bytes > 0 and p.info.lane_min_voices >= 148 * 64  # one 64-voice CTA per SM
    out = np.zeros((64, 2000), dtype=np.float32)
    p.render(out, params=params)
    assert p.info.lane_launches == 0


def test_mixdown_on_chip(monkeypatch):
    """tb_render_mix without rows on a large steady batch: the voices are summed on the chip in blocks — the 32
    voices of a warp in voice order, warps in order — and never written as rows (the fused-FM-voice kernel takes the
    whole call, from the stream's first sample to the samples that do not fill a tile).  Checked bit for bit against
    that order applied to the rows the same kernel renders, and against the oracle's serial mix within the
    per-voice tolerance."""
    V, N = 1000, 256 + 16 * 50
    w, params = cfg5(V)
    rows = np.zeros((V, N), dtype=np.float32)
    program(w, monkeypatch).render(rows, params=params)
    p = program(w, monkeypatch)
    mix = np.full(N, np.inf, dtype=np.float32)
    lens = p.render_mix(mix, V, params=params)
    assert p.info.lane_launches == 1 and p.info.kernel_launches == 2 and (lens == N).all()   # the lane kernel + one tb_mix_kernel
    padded = np.zeros((1024, N), dtype=np.float32)
    padded[:V] = rows
    want = None
    for wi in range(0, 1024, 32):
        part = padded[wi].copy()
        for v in range(1, 32):
            part += padded[wi + v]
        want = part if want is None else want + part
    np.testing.assert_array_equal(mix, want)
    # ragged length, device mix buffer, against the oracle
    import torch
    N2 = N + 5
    q = program(w, monkeypatch)
    dmix = torch.zeros(N2, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()  # (the fill runs on torch's stream, the renders on the program's own)
    q.render_mix(dmix, V, params=torch.from_numpy(params).cuda())
    torch.cuda.synchronize()
    _, _, omix, _ = OracleProgram(w, SR).render_batch(params, V, N2, mix=True, threads=1)
    assert np.max(np.abs(dmix.cpu().numpy() - omix)) <= TOL * V
    assert np.max(np.abs(dmix.cpu().numpy()[:N] - mix)) <= 1e-3


def test_streaming_blocks_skip_the_general_head(monkeypatch):
    """A caller that streams blocks (main.rs:42-43 uses 1024) pays one launch per block, and the
    stream is the one a single call renders (same tolerance as test_partition_invariance).  (Programs
    that run through the lane interpreter pay the general head tile once per stream:
    test_every_steady_operator's tree is checked for that below.)"""
    V, N = 160, 1024 * 5 + 777
    w, params = cfg5(V)
    one = np.zeros((V, N), dtype=np.float32)
    program(w, monkeypatch).render(one, params=params)
    p = program(w, monkeypatch)
    parts = np.zeros((V, N), dtype=np.float32)
    launches = []
    for a in range(0, N, 1024):
        b = min(N, a + 1024)
        k0 = p.info.kernel_launches
        blk = np.zeros((V, b - a), dtype=np.float32)
        assert (p.render(blk, params=params) == b - a).all()
        parts[:, a:b] = blk
        launches.append(int(p.info.kernel_launches - k0))
    assert launches == [1, 1, 1, 1, 1, 1]  # the fused-FM-voice kernel needs no general tile, not even for 777 = 48 * 16 + 9
    per_voice = np.max(np.abs(parts - one), axis=1)
    within(parts, one, params)
    assert np.median(per_voice) <= 5e-6
    ref, _, _, _ = OracleProgram(w, SR).render_batch(params, V, N)
    within(parts, ref, params)
    # length() moves node positions by different amounts: the next render takes the general head again
    p.lengths(V, 100, params=params)
    k0 = p.info.kernel_launches
    p.render(np.zeros((V, 1024), dtype=np.float32), params=params)
    assert p.info.kernel_launches - k0 == 2
    # a mixdown without rows continues a primed stream with the lane kernel alone
    q = program(w, monkeypatch)
    q.render_mix(np.zeros(1024, dtype=np.float32), V, params=params)
    k0 = q.info.kernel_launches
    mix = np.zeros(1024, dtype=np.float32)
    q.render_mix(mix, V, params=params)
    assert q.info.kernel_launches - k0 == 2  # lane kernel + the add of its partial rows
    assert np.max(np.abs(mix - one[:, 1024:2048].sum(axis=0, dtype=np.float64))) <= TOL * V
    # an interpreter program (two filters): general head tile on the first call only
    from tuun_b200.workloads import lpf
    r = program(lpf(w, 2.0, 1600), monkeypatch)
    counts = []
    for _ in range(3):
        k0 = r.info.kernel_launches
        r.render(np.zeros((V, 1024), dtype=np.float32), params=params)
        counts.append(int(r.info.kernel_launches - k0))
    assert counts == [2, 1, 1]


def test_work_queue_segments(monkeypatch):
    """When a batch has more 64-voice groups than the device holds CTAs, the lane launch is cut into
    time segments handed out through a work queue (forced here with TUUN_B200_LANE_QUEUE=1): same
    stream as the plain launch (the constant-rate sines re-derive their carried sin/cos pair from the
    exact accumulator at every segment start, hence 'within one f32 rounding' and not 'equal')."""
    V, N = 1000, 256 + 16 * 300
    w, params = cfg5(V)
    monkeypatch.setenv("TUUN_B200_LANE_QUEUE", "0")
    plain = np.zeros((V, N), dtype=np.float32)
    program(w, monkeypatch).render(plain, params=params)
    monkeypatch.setenv("TUUN_B200_LANE_QUEUE", "1")
    p = program(w, monkeypatch)
    q = np.full((V, N), np.inf, dtype=np.float32)
    lens = p.render(q, params=params)
    assert (lens == N).all() and p.info.lane_launches == 1
    per_voice = np.max(np.abs(q - plain), axis=1)
    within(q, plain, params)
    assert np.median(per_voice) <= 5e-6  # the plain render is the self-primed FM kernel, the queued one starts with a general tile
    ref, _, _, _ = OracleProgram(w, SR).render_batch(params, V, N)
    within(q, ref, params)
    # a tree that runs through the interpreter (two more biquads behind the fused voice), and the mixdown
    from tuun_b200.workloads import lpf
    w3 = lpf(lpf(w, 2.0, 1600), 1.0, 3200)
    a = np.zeros((V, N), dtype=np.float32)
    program(w3, monkeypatch).render(a, params=params)
    ref3, _, _, _ = OracleProgram(w3, SR).render_batch(params, V, N)
    within(a, ref3, params)
    mix = np.zeros(N, dtype=np.float32)
    program(w, monkeypatch).render_mix(mix, V, params=params)
    assert np.max(np.abs(mix - q.sum(axis=0, dtype=np.float64))) <= 1e-3


@pytest.mark.parametrize("fast_sines", ["0", "1"])
def test_sine_precision_modes(monkeypatch, fast_sines):
    """TUUN_B200_FAST_SINES: "0" makes every sine EXACT class (f64 polynomial), "1" evaluates the FAST
    class with the f32 polynomial instead of MUFU.SIN; the lane interpreter has its own code for both."""
    monkeypatch.setenv("TUUN_B200_FAST_SINES", fast_sines)
    V, N = 130, 256 + 16 * 90 + 4
    w, params = cfg5(V)
    p = program(w, monkeypatch)
    out = np.zeros((V, N), dtype=np.float32)
    p.render(out, params=params)
    assert p.info.lane_launches == 1 and p.info.lane_fm_capacity == 0  # the FM kernel is for MUFU carriers only
    ref, _, _, _ = OracleProgram(w, SR).render_batch(params, V, N)
    within(out, ref, params)
    # phase-modulated and doubly modulated sines through the same modes
    pm = Sine(Const(1.0, param=0), mul(Sine(add(mul(Sine(Const(1.0, param=1), Const(0.0)), Const(40.0)), Const(1.0, param=2)),
                                            Const(0.2)), Const(4.0)))
    rng = np.random.default_rng(9)
    pr = np.stack([TAU * rng.uniform(100, 900, V), TAU * rng.uniform(1, 9, V), TAU * rng.uniform(50, 400, V)], axis=1).astype(np.float32)
    q = program(pm, monkeypatch)
    out = np.zeros((V, N), dtype=np.float32)
    q.render(out, params=pr)
    assert q.info.lane_launches == 1
    assert np.max(np.abs(out - oracle_rows(pm, pr, V, N))) <= TOL


def test_notes_of_fixed_duration(monkeypatch):
    """Root Fin with an analytic length over a steady tree — `$f * Qw` (config 1), any note with a fixed
    duration: the lane kernel renders the inner tree and counts how much of it belongs to each voice.
    Lengths are bit-exact (per-voice durations), samples within tolerance up to each voice's end (the
    tail of a finished voice's row is undefined, generator.rs:76-95), later calls return 0."""
    from tuun_b200.waveform import Fin
    from tuun_b200.workloads import fm_filter_voice
    V, N = 700, 256 + 16 * 1200 + 9
    rng = np.random.default_rng(4)
    dur = rng.uniform(0.003, 0.6, V).astype(np.float32)   # 132 .. 26,460 samples: some end inside the head tile,
    dur[:5] = [0.0, 0.5, 1.0, 256 / SR, 19465 / SR]         # some never inside this call, some on tile edges
    note = Fin(add(Time(), Const(0.0, param=0)), Sine(Const(1.0, param=1), Const(0.0)))   # cfg1 with swept Q and f
    params = np.stack([-dur, TAU * rng.uniform(100, 2000, V).astype(np.float32)], axis=1).astype(np.float32)
    p = program(note, monkeypatch)
    out = np.full((V, N), np.inf, dtype=np.float32)
    lens = p.render(out, params=params)
    assert p.info.lane_launches == 1
    o = OracleProgram(note, SR)
    olens = np.zeros(V, dtype=np.int64)
    for v in range(V):
        o.initialize_state()
        o.set_params(params[v])
        ref = o.render(N)
        olens[v] = len(ref)
        assert lens[v] == len(ref), (v, dur[v])
        assert np.max(np.abs(out[v, :len(ref)] - ref), initial=0.0) <= 1e-6, v
    assert (olens == N).sum() > 50 and (olens < 256).sum() > 5
    # the stream continues: unfinished voices go on where they were, finished ones return 0 for good
    out2 = np.full((V, 4096), np.inf, dtype=np.float32)
    lens2 = p.render(out2, params=params)
    for v in range(V):
        o.initialize_state()
        o.set_params(params[v])
        full = o.render(N + 4096)
        assert lens2[v] == max(0, len(full) - N), v
        assert np.max(np.abs(out2[v, :lens2[v]] - full[N:]), initial=0.0) <= 1e-6
    # an FM + low-pass note: the fused voice kernel under a Fin
    fmnote = Fin(sub(Time(), Const(0.25)), fm_filter_voice())
    w, fp = cfg5(V)
    q = program(fmnote, monkeypatch)
    out = np.zeros((V, N), dtype=np.float32)
    lens = q.render(out, params=fp)
    assert (lens == 11025).all() and q.info.lane_launches == 1 and q.info.lane_fm_capacity > 0
    ref, _, _, _ = OracleProgram(w, SR).render_batch(fp, V, 11025)
    within(out[:, :11025], ref, fp)


def test_full_batch_of_65536_voices():
    """BASELINE.json config 5 at its full voice count (default kernel selection, device rows, 4096
    samples per voice): 512 voices spread over the sweep against the oracle, every length, the
    mixdown against the rows, and the stream continued by a second call against one longer call."""
    import torch
    from tuun_b200.generator import Program
    from tuun_b200.workloads import fm_filter_params, fm_filter_voice
    V, N = 65536, 4096
    w = fm_filter_voice()
    params = fm_filter_params(np.arange(V))
    pd = torch.from_numpy(params).cuda()
    p = Program(w, SR)
    assert p.info.lane_min_voices <= V and p.info.lane_fm_capacity * 64 >= V
    out = torch.zeros((V, 2 * N), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()  # (the fill runs on torch's stream, the renders on the program's own)
    lens = np.zeros(V, dtype=np.uint64)
    p.render(out[:, :N], params=pd, out_len=lens)
    assert (lens == N).all() and p.info.lane_launches == 1
    p.render(out[:, N:], params=pd, out_len=lens)       # primed stream: the lane kernel alone
    assert (lens == N).all() and p.info.lane_launches == 2 and p.info.kernel_launches == 2
    pick = (np.arange(512) * 127 + 5) % V
    got = out[torch.from_numpy(pick).cuda()].cpu().numpy()
    ref, _, _, _ = OracleProgram(w, SR).render_batch(params[pick], len(pick), 2 * N, threads=4)
    within(got, ref, params[pick])
    one = torch.zeros((V, 2 * N), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()  # (the fill runs on torch's stream, the renders on the program's own)
    Program(w, SR).render(one, params=pd)
    d = (one - out).abs().amax(dim=1)
    from tuun_b200.workloads import fm_filter_tolerance
    assert bool((d.cpu().numpy() <= fm_filter_tolerance(params, TOL)).all()) and float(d.median()) <= 5e-6
    assert bool(torch.isfinite(out).all())
    mix = torch.zeros(2 * N, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()  # (the fill runs on torch's stream, the renders on the program's own)
    Program(w, SR).render_mix(mix, V, params=pd)
    want = one.sum(dim=0, dtype=torch.float64)
    # 65,536 f32 terms summed in blocks against an f64 sum (the voices start in phase: |mix| reaches thousands)
    assert float((mix.double() - want).abs().max()) <= 0.05 + 2e-4 * float(want.abs().max())


def test_fm_voice_kernel_edges(monkeypatch):
    """The fused-FM-voice kernel at the edges of its loop: calls shorter than two tiles, lengths around
    tile multiples, fresh and continued streams, few voices, the voice without a filter, and a voice whose
    carrier frequency bound is beyond the exact range of the magic-number conversion (slow conversion)."""
    from tuun_b200.workloads import fm_filter_params, fm_filter_voice
    w = fm_filter_voice()
    for V in (1, 31, 33, 65):
        params = fm_filter_params((np.arange(V) * 40503 + 17) % 65536)
        o = OracleProgram(w, SR)
        for sizes in ((16,), (17, 15, 1, 31), (33, 16, 47), (5, 300), (255, 257, 16)):
            p = program(w, monkeypatch)
            total = sum(sizes)
            got = np.zeros((V, total), dtype=np.float32)
            a = 0
            for n in sizes:
                blk = np.full((V, n), np.inf, dtype=np.float32)
                lens = p.render(blk, params=params)
                assert (lens == n).all()
                got[:, a:a + n] = blk
                a += n
            assert p.info.lane_launches == sum(1 for n in sizes if n >= 16), sizes
            ref, _, _, _ = o.render_batch(params, V, total)
            within(got, ref, params)
    # no filter behind the pair
    fm = Sine(add(mul(Sine(Const(1.0, param=0), Const(f32(math.pi / 2))), Const(1.0, param=1)), Const(1.0, param=2)), Const(0.0))
    V = 40
    params = fm_filter_params(np.arange(V) * 997 % 65536)
    for n in (16, 23, 1000 + 9):
        p = program(fm, monkeypatch)
        out = np.zeros((V, n), dtype=np.float32)
        p.render(out, params=params)
        assert p.info.lane_launches == 1 and p.info.kernel_launches == 1
        assert np.max(np.abs(out - oracle_rows(fm, params, V, n))) <= 5e-6
    # one voice with an absurd modulation gain: its warp converts frequencies the slow, exact way
    big = params.copy()
    big[3, 1] = 6.0e7   # rad/s peak deviation, above 100 x 2 pi x 44100
    p = program(w, monkeypatch)
    out = np.zeros((V, 2000), dtype=np.float32)
    p.render(out, params=big)
    ref = oracle_rows(w, big, V, 2000)
    assert np.max(np.abs(np.delete(out - ref, 3, axis=0))) <= TOL
    # the wild voice itself (its frequency aliases hundreds of times per sample): same stream
    assert np.max(np.abs(out[3] - ref[3])) <= 1e-3


def test_fm_ws_matches_single_thread_kernel(monkeypatch):
    """The fused FM voice as a phase warp and a tone warp per 32 voices (lanes_fm_ws.cu, the default when the device
    holds all those CTAs at once) against the one-thread-a-voice form (lanes_fm.cu, TUUN_B200_FM_WS=0): the same
    operations in the same order on both sides of the hand-over, so rows, lengths, carried state (later calls of the
    same stream) and the on-chip mixdown are bit-identical — over ragged batch sizes (a last CTA with dead lanes),
    call lengths around tile pairs, a slow-conversion voice, and rows that are only 8-byte aligned."""
    import torch
    from tuun_b200.workloads import fm_filter_params, fm_filter_sample_ids, fm_filter_voice
    w = fm_filter_voice()

    def run(ws, V, calls, mix=False, wild=False):
        monkeypatch.setenv("TUUN_B200_FM_WS", "1" if ws else "0")
        p = program(w, monkeypatch)
        prm = fm_filter_params(fm_filter_sample_ids(V))
        if wild:
            prm[min(3, V - 1), 1] = 6.0e7
        params = torch.from_numpy(prm).cuda()
        outs = []
        for n in calls:
            if mix:
                m = torch.zeros((n,), dtype=torch.float32, device="cuda")
                torch.cuda.synchronize()  # (the fill runs on torch's stream, the renders on the program's own)
                p.render_mix(m, V, params=params)
                torch.cuda.synchronize()  # (the program renders on its own stream; no host lengths were asked for)
                outs.append(m.cpu().numpy())
            else:
                out = torch.full((V, n), float("inf"), dtype=torch.float32, device="cuda")
                torch.cuda.synchronize()  # (the fill runs on torch's stream, the render on the program's own)
                lens = p.render(out, params=params, out_len=np.zeros(V, dtype=np.uint64))
                assert (lens == n).all()
                outs.append(out.cpu().numpy())
        return outs, int(p.info.fm_ws_launches), int(p.info.lane_launches)

    for V, calls, wild in [(64, [4096, 1000], False), (100, [4101, 16, 31, 777], False), (33, [17, 15, 16, 48, 64], False),
                           (2049, [2048 + 13, 32], False), (40, [2000, 18], True), (12352, [4096 + 5], False)]:
        a, wa, la = run(False, V, calls, wild=wild)
        b, wb, lb = run(True, V, calls, wild=wild)
        assert wa == 0 and wb == lb == la == sum(1 for n in calls if n >= 16), (V, calls, wa, wb, la, lb)
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x.view(np.uint32), y.view(np.uint32))
    # the whole 65,536-voice batch, every SM holding its 14 CTAs: compared on the device
    def whole(ws):
        monkeypatch.setenv("TUUN_B200_FM_WS", "1" if ws else "0")
        p = program(w, monkeypatch)
        params = torch.from_numpy(fm_filter_params(np.arange(65536))).cuda()
        out = torch.empty((65536, 8192 + 16), dtype=torch.float32, device="cuda")
        lens = p.render(out, params=params, out_len=np.zeros(65536, dtype=np.uint64))
        assert (lens == out.shape[1]).all() and p.info.fm_ws_launches == (1 if ws else 0)
        return out
    a = whole(False)
    b = whole(True)
    assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    del a, b
    # a batch cut in time (abi.cpp render_split_fm): the warm-up and the samples pass run on the two-warp kernel over
    # virtual voices (lanes_fm_ws_split.cu), the phase-sum pass keeps its own kernel
    def cut(ws):
        monkeypatch.setenv("TUUN_B200_FM_WS", "1" if ws else "0")
        monkeypatch.setenv("TUUN_B200_SPLIT_FM", "4")  # (read when a call is planned)
        p = program(w, monkeypatch)
        params = torch.from_numpy(fm_filter_params(fm_filter_sample_ids(4096))).cuda()
        out = torch.empty((4096, 4 * 32768 + 24), dtype=torch.float32, device="cuda")
        lens = p.render(out, params=params, out_len=np.zeros(4096, dtype=np.uint64))
        monkeypatch.delenv("TUUN_B200_SPLIT_FM")
        assert (lens == out.shape[1]).all() and p.info.split_fm_rounds == 1
        return out, int(p.info.fm_ws_launches)
    a, wa = cut(False)
    b, wb = cut(True)
    assert wa == 0 and wb >= 2 and torch.equal(a.view(torch.int32), b.view(torch.int32))
    del a, b
    for V, calls in [(4096, [4096 + 7, 640]), (1000, [8000])]:
        a, _, _ = run(False, V, calls, mix=True)
        b, wb, _ = run(True, V, calls, mix=True)
        assert wb == len(calls)
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x.view(np.uint32), y.view(np.uint32))


def test_reset_oscillators(monkeypatch):
    """sawtooth, pulse (with a modulated width) and triangle of lib/v0/std.tuun — a Reset over a tree that is
    closed-form in the run's own clock — as large batches: the lane kernels render the trigger, turn it into a
    per-sample local clock (ST_RESET_CLK) and evaluate the inner tree against it.  Per-voice frequencies; a
    filter behind; streamed in blocks (the Reset's sign state and the clocked nodes' positions carry over)."""
    from tuun_b200.generator import lower_check
    from tuun_b200.waveform import Reset
    from tuun_b200.workloads import lpf
    V, N = 300, 256 + 16 * 500 + 6
    rng = np.random.default_rng(12)
    f = rng.uniform(30.0, 1800.0, V).astype(np.float32)
    params = np.stack([TAU * f, -f, f32(4) * f, f32(-4) * f], axis=1).astype(np.float32)
    trig = lambda: Sine(Const(1.0, param=0), Const(0.0))
    saw = mul(add(Reset(trig(), mul(Time(), Const(1.0, param=1))), Const(0.5)), Const(2.0))
    width = add(mul(Sine(Const(f32(TAU * 1.6)), Const(0.0)), Const(0.05)), Const(0.43))      # pulse-width modulation
    pulse = Alt(sub(saw, width), Const(1.0), Const(-1.0))
    tri = Alt(trig(), Reset(trig(), add(mul(Time(), Const(1.0, param=2)), Const(-1.0))),
              Reset(trig(), add(mul(Time(), Const(1.0, param=3)), Const(3.0))))
    ringing = Reset(trig(), mul(Sine(Const(f32(TAU * 3000)), Const(0.0)), add(mul(Time(), Const(-40.0)), Const(1.0))))
    for name, w, tol in (("saw", saw, 2e-5), ("pulse", pulse, TOL), ("triangle", tri, 2e-5),
                         ("saw|lpf", lpf(saw, 0.9, 1500), 2e-5), ("reset sine", ringing, 2e-5)):
        info = lower_check(w)
        assert info.lane_smem_bytes > 0 and info.tile == 256, name   # lane kernels yes, warp steady interpreter no
        p = program(w, monkeypatch)
        out = np.zeros((V, N), dtype=np.float32)
        lens = p.render(out, params=params)
        assert (lens == N).all() and p.info.lane_launches == 1, name
        ref = oracle_rows(w, params, V, N)
        d = np.abs(out - ref)
        # a trigger within rounding of zero may move an edge by one sample (hard part 1): count them
        bad = int(np.count_nonzero(d > tol))
        assert bad <= 3, (name, bad, float(d.max()))
        # the same stream in blocks
        q = program(w, monkeypatch)
        parts = np.zeros((V, N), dtype=np.float32)
        a = 0
        for n in (1000, 272, 4000, N - 5272):
            blk = np.zeros((V, n), dtype=np.float32)
            q.render(blk, params=params)
            parts[:, a:a + n] = blk
            a += n
        assert int(np.count_nonzero(np.abs(parts - ref) > tol)) <= 3, name
