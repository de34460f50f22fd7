"""Multi-GPU host logic: voices are independent (each `Command::Play` waveform owns its state,
tracker.rs:84-95), so a batch shards into contiguous voice ranges with NO data-path collective.
The one exchange step is the optional mixdown — the tracker's `out[j] += tmp[j]` (tracker.rs:617-619)
summed across ranks: every rank reduces its own voices on the device (tb_render_mix) and the
[n_samples] partial mixes are reduced over the process group (NCCL over NVLink on the GPU box;
gloo in the CPU tests of this logic).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def voice_range(n_voices: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous share [lo, hi) of `n_voices` for `rank`: ranges tile the batch exactly and differ
    in size by at most one voice."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return n_voices * rank // world, n_voices * (rank + 1) // world


def weak_voice_range(voices_per_gpu: int, rank: int) -> Tuple[int, int]:
    """Weak scaling: every rank renders `voices_per_gpu` voices; ids continue across ranks."""
    return voices_per_gpu * rank, voices_per_gpu * (rank + 1)


def reduce_mix(partial, dst: int = 0, group=None):
    """Sum the per-rank partial mixes into rank `dst` (in place on `dst`): the cross-rank half of
    the tracker's mix loop.  `partial` is a 1-D float32 torch tensor on the device the process
    group's backend works on.  No-op without an initialised process group (single process)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return partial
    dist.reduce(partial, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return partial


def gather_lengths(local_lens: np.ndarray, n_total: int, rank: int, world: int, group=None) -> Optional[np.ndarray]:
    """out_len of every voice on rank 0 (None elsewhere): ranks hold contiguous ranges."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return np.asarray(local_lens, dtype=np.uint64)
    sizes = [voice_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    mine = torch.zeros(width, dtype=torch.int64)
    mine[: len(local_lens)] = torch.from_numpy(np.asarray(local_lens, dtype=np.int64))
    bucket = [torch.zeros(width, dtype=torch.int64) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, bucket, dst=0, group=group)
    if rank != 0:
        return None
    return np.concatenate([bucket[r][: hi - lo].numpy() for r, (lo, hi) in enumerate(sizes)]).astype(np.uint64)


class ShardedBatch:
    """One rank's share of a batch render: program + its voice range; `render`/`render_mix` are the
    local calls, `mixdown` adds the cross-rank reduce."""

    def __init__(self, waveform, sample_rate: int, n_voices: int, rank: int, world: int, device: int = -1,
                 weak: bool = False):
        from .generator import Program
        self.rank, self.world = rank, world
        self.lo, self.hi = weak_voice_range(n_voices, rank) if weak else voice_range(n_voices, rank, world)
        self.n_total = n_voices * world if weak else n_voices
        self.program = Program(waveform, sample_rate, device=device)

    @property
    def n_local(self) -> int:
        return self.hi - self.lo

    def render(self, out, params=None, out_len=None):
        return self.program.render(out, params=params, out_len=out_len)

    def mixdown(self, mix, params=None, dst: int = 0, group=None, after_local=None):
        """Local tb_render_mix (rows never leave the device), then reduce_mix across ranks.
        `after_local` runs between the two (stream hand-over on the GPU box)."""
        self.program.render_mix(mix, self.n_local, params=params)
        if after_local is not None:
            after_local()
        return reduce_mix(mix, dst=dst, group=group)
