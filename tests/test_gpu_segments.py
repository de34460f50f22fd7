"""Time-segment sharding over GPUs (include/tuun_b200.h tb_segments_*; tuun_b200/sharding.py TimeShard): here
two "ranks" are two programs on ONE device that exchange their segments' state blocks by plain copies — the
protocol a 2-GPU run performs with one NCCL all-gather per pass (bench.py --time-shard; the gloo test of the
exchange itself is tests/test_sharding_gloo.py).  Each rank renders its own half of the time axis of every
voice; together they must be the stream one serial program produces, and both must be able to continue it."""
import math

import numpy as np
import pytest

from oracle.binding import OracleProgram
from tuun_b200.waveform import Alt, Const, Sine, f32

pytestmark = pytest.mark.gpu
SR = 44100
TAU = f32(2 * math.pi)


def two_rank_render(w, params, V, S, seg, monkeypatch, tail=1000, fm_form=False):
    import torch
    from tuun_b200.generator import Program
    from tuun_b200.sharding import TimeShard, segment_range
    if fm_form:   # the cheaper form for fused FM voices (phase-sum pass, filter warm-up): same protocol for the caller
        monkeypatch.setenv("TUUN_B200_SPLIT_FM", "0")
        monkeypatch.setenv("TUUN_B200_LANE_MIN_VOICES", "1")
    else:
        monkeypatch.setenv("TUUN_B200_SPLIT", "0")  # the serial reference and the heads: no split of their own
    head = 256
    n = head + S * seg
    serial = np.zeros((V, n + tail), dtype=np.float32)
    Program(w, SR).render(serial, params=params)
    ranks = [Program(w, SR) for _ in range(2)]
    heads = []
    for p in ranks:                                  # every rank renders the first tile of the stream itself
        h = np.zeros((V, head), dtype=np.float32)
        p.render(h, params=params)
        heads.append(h)
    outs = [torch.zeros((V, S // 2 * seg), dtype=torch.float32, device="cuda") for _ in ranks]
    torch.cuda.synchronize()  # (the fill runs on torch's stream, the renders on the program's own)
    if fm_form:
        monkeypatch.setenv("TUUN_B200_SPLIT_FM", "1")
    shards = [TimeShard(p, V, S, seg, *segment_range(S, r, 2), params=params) for r, p in enumerate(ranks)]
    if fm_form:
        monkeypatch.setenv("TUUN_B200_SPLIT_FM", "0")
    assert shards[0].passes == shards[1].passes
    for k in range(1, shards[0].passes + 1):
        for ts, o in zip(shards, outs):
            ts.run_pass(k, o)
        torch.cuda.synchronize()
        a, b = shards[0].states, shards[1].states    # [V, S, words]; the all-gather, by hand
        assert a.shape == (V, S, a.shape[2])
        b[:, :S // 2] = a[:, :S // 2]
        a[:, S // 2:] = b[:, S // 2:]
        torch.cuda.synchronize()
        if k < shards[0].passes:
            for ts in shards:
                ts.fix(k)
    for ts in shards:
        ts.end()
    torch.cuda.synchronize()
    got = np.concatenate([heads[0], outs[0].cpu().numpy(), outs[1].cpu().numpy()], axis=1)
    tails = []
    for p in ranks:                                  # both ranks continue the same stream
        t = np.zeros((V, tail), dtype=np.float32)
        p.render(t, params=params)
        tails.append(t)
    assert (ranks[0].info.split_fm_rounds == 1) == fm_form
    return serial, got, tails, shards[0].passes


def test_two_ranks_render_one_stream_fm_filter(monkeypatch):
    from tuun_b200.workloads import fm_filter_params, fm_filter_voice
    w = fm_filter_voice()
    params = fm_filter_params([49157, 34061, 16389 + 256 * 9, 40000])
    V, S, seg = 4, 8, 2048
    serial, got, tails, passes = two_rank_render(w, params, V, S, seg, monkeypatch)
    assert passes == 3
    n = got.shape[1]
    assert np.abs(got - serial[:, :n]).max() <= 1e-5          # filters: their own round-off noise (gain <= 38)
    for t in tails:
        assert np.abs(t - serial[:, n:]).max() <= 1e-5
    o = OracleProgram(w, SR)
    for v in range(V):
        o.initialize_state()
        o.set_params(params[v])
        assert np.abs(got[v] - o.render(n)).max() <= 1e-4


def test_two_ranks_fm_form(monkeypatch):
    """The same protocol when the library takes the cheaper form for fused FM voices (what the time_shard record of
    bench.py runs: 64 voices x 512 segments per rank): pass 1 phase sums only — the snapshot for the warm-ups rides in
    the exchanged state blocks —, pass 2 the filters' warm-ups, local to each rank, pass 3 the samples."""
    from tuun_b200.workloads import biquad_noise_gain, fm_filter_cover_ids, fm_filter_params, fm_filter_tolerance, fm_filter_voice
    w = fm_filter_voice()
    ids = fm_filter_cover_ids(1)[::8]            # 32 voices, every cutoff and Q
    params = fm_filter_params(ids)
    V, S, seg = len(ids), 4, 16384               # the slowest filter needs 3,500 samples of warm-up: a quarter of a segment
    serial, got, tails, passes = two_rank_render(w, params, V, S, seg, monkeypatch, fm_form=True)
    assert passes == 3
    n = got.shape[1]
    g = np.maximum(biquad_noise_gain(params[:, 6], params[:, 7]), 5.0)
    assert (np.abs(got - serial[:, :n]).max(axis=1) <= 6e-7 * g).all()
    for t in tails:
        assert (np.abs(t - serial[:, n:]).max(axis=1) <= 6e-7 * g).all()
    o = OracleProgram(w, SR)
    ref, _, _, _ = o.render_batch(params, V, n, threads=8)
    assert (np.abs(got - ref).max(axis=1) <= fm_filter_tolerance(params, 1e-4)).all()


def test_two_ranks_sines_are_bit_identical(monkeypatch):
    from tuun_b200.workloads import fm_filter_params, fm_filter_sample_ids, fm_pair_voice
    w = fm_pair_voice()
    params = fm_filter_params(fm_filter_sample_ids(5))
    serial, got, tails, passes = two_rank_render(w, params, 5, 4, 1024, monkeypatch)
    assert passes == 2
    n = got.shape[1]
    # no filter: this program is steady from its first sample, so the serial render tiles the stream from 0 while the
    # ranks start behind a 256-sample head — the phase is the same u64, FAST sines of the two tile forms differ by 3e-7
    assert np.abs(got - serial[:, :n]).max() <= 1e-6
    assert np.array_equal(tails[0], tails[1])
    assert np.abs(tails[0] - serial[:, n:]).max() <= 1e-6


def test_segments_api_refuses_misuse():
    from tuun_b200._abi import TB_ERR_STATE, TB_ERR_UNSUPPORTED, TuunB200Error
    from tuun_b200.generator import Program
    from tuun_b200.sharding import TimeShard
    from tuun_b200.waveform import Fin, Time, add
    from tuun_b200.workloads import lpf
    sq = Alt(Sine(Const(TAU * f32(220.0)), Const(0.0)), Const(1.0), Const(-1.0))
    p = Program(lpf(sq, 0.7, 2000.0), SR)
    with pytest.raises(TuunB200Error) as e:      # the stream has not started: the filter has not read ahead yet
        TimeShard(p, 1, 4, 1024, 0, 4)
    assert e.value.status == TB_ERR_STATE
    q = Program(Fin(add(Time(), Const(-0.5)), sq), SR)
    q.render(np.zeros((1, 256), dtype=np.float32))
    with pytest.raises(TuunB200Error) as e:      # a note that ends: not a steady program
        TimeShard(q, 1, 4, 1024, 0, 4)
    assert e.value.status == TB_ERR_UNSUPPORTED
