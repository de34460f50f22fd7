"""The N>1 host logic on CPU: two processes over gloo (world_size 2) run the same sharding and
mixdown-reduce code the GPU ranks run (tuun_b200/sharding.py); the per-voice renders come from the
CPU oracle here, since this test is about the exchange step, not the kernel."""
import os
import socket

import numpy as np
import pytest

from tuun_b200.sharding import voice_range, weak_voice_range

SR = 44100


def test_voice_ranges_tile_the_batch():
    for n in (1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            parts = [voice_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    assert [weak_voice_range(100, r) for r in range(3)] == [(0, 100), (100, 200), (200, 300)]
    with pytest.raises(ValueError):
        voice_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_voices, n_samples, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from oracle.binding import OracleProgram
    from tuun_b200.sharding import gather_lengths, reduce_mix, voice_range
    from tuun_b200.waveform import Const, Fin, Time, add
    from tuun_b200.workloads import fm_filter_params, fm_filter_voice

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = voice_range(n_voices, rank, world)
        ids = (np.arange(lo, hi) * 211) % 65536
        o = OracleProgram(fm_filter_voice(), SR)
        rows, lens, _, _ = o.render_batch(fm_filter_params(ids), hi - lo, n_samples)
        partial = np.zeros(n_samples, dtype=np.float32)
        for v in range(hi - lo):  # the rank-local half of the tracker's serial mix
            partial += rows[v]
        t = torch.from_numpy(partial.copy())
        reduce_mix(t, dst=0)
        all_lens = gather_lengths(np.asarray(lens), n_voices, rank, world)
        if rank == 0:
            q.put((t.numpy().copy(), all_lens))
    finally:
        dist.destroy_process_group()


def test_two_rank_mixdown_over_gloo():
    import torch.multiprocessing as mp

    from oracle.binding import OracleProgram
    from tuun_b200.workloads import fm_filter_params, fm_filter_voice

    n_voices, n_samples, world = 13, 4096, 2  # an odd count: the ranges are 6 and 7 voices
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_voices, n_samples, q)) for r in range(world)]
    for p in procs:
        p.start()
    mix, lens = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process tracker order over all voices
    ids = (np.arange(n_voices) * 211) % 65536
    rows, ref_lens, _, _ = OracleProgram(fm_filter_voice(), SR).render_batch(fm_filter_params(ids), n_voices, n_samples)
    serial = np.zeros(n_samples, dtype=np.float32)
    for v in range(n_voices):
        serial += rows[v]
    assert (lens == np.asarray(ref_lens)).all() and len(lens) == n_voices
    # two partial sums instead of one running sum: f32 re-association, bounded by a few ulps of the mix
    assert np.max(np.abs(mix - serial)) <= 4e-6 * n_voices


# ---- time-segment sharding: the exchange step of tb_segments_* (sharding.exchange_segment_states) ----
def test_segment_plans():
    from tuun_b200.sharding import plan_segments, segment_range
    for world in (1, 2, 4, 8):
        for n in (600 * SR, 60 * SR, 10 * SR, 65536):
            S, seg = plan_segments(n, world)
            assert S % world == 0 and seg % 512 == 0 and seg >= 512 and S * seg <= n
            assert n - S * seg < S * 512                      # what is left for the serial tail
            parts = [segment_range(S, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == S
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert len({hi - lo for lo, hi in parts}) == 1    # equal shares: one plain all-gather
    assert plan_segments(16 * 2048 + 100, 8, per_rank=2) == (16, 2048)
    with pytest.raises(ValueError):
        segment_range(10, 0, 4)
    with pytest.raises(ValueError):
        plan_segments(1000, 8)


def _state_block(v, s, words):
    """What a rank writes for segment s of voice v in this test: a pattern that names the block."""
    return (np.arange(words, dtype=np.int64) * 7 + v * 1000003 + s * 101).astype(np.int32)


def _seg_worker(rank, world, port, V, S, words, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from tuun_b200.sharding import exchange_segment_states, segment_range

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = segment_range(S, rank, world)
        states = torch.full((V, S, words), -1, dtype=torch.int32)   # what other ranks own is stale here
        for v in range(V):
            for s in range(lo, hi):
                states[v, s] = torch.from_numpy(_state_block(v, s, words))
        exchange_segment_states(states, lo, hi)
        q.put((rank, states.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_segment_state_exchange_over_gloo():
    """After a pass every rank holds the final states of its own time range of every voice; the all-gather must
    leave EVERY rank with all [voice, segment] blocks in place (the scan over segments then runs redundantly
    and identically on every rank)."""
    import torch.multiprocessing as mp
    V, S, words, world = 3, 8, 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_seg_worker, args=(r, world, port, V, S, words, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([np.stack([_state_block(v, s, words) for s in range(S)]) for v in range(V)])
    for r in range(world):
        np.testing.assert_array_equal(got[r], want)
