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


# ---------------------------------------------------------------------------------------------------
# Time-segment sharding: few voices, long renders (include/tuun_b200.h tb_segments_*).
# ---------------------------------------------------------------------------------------------------
def segment_range(n_segments: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous share [lo, hi) of a voice's segments for `rank` (every rank renders that time range of EVERY
    voice).  n_segments must be a multiple of the world size so that the exchange is one plain all-gather."""
    if n_segments % world != 0:
        raise ValueError(f"{n_segments} segments do not divide over {world} ranks")
    per = n_segments // world
    return per * rank, per * (rank + 1)


def plan_segments(n_samples: int, world: int, per_rank: int = 0, tile: int = 512) -> Tuple[int, int]:
    """(n_segments, seg_samples) for a call of n_samples over `world` ranks: `per_rank` segments each (0: as many
    as keep segments at >= 4 tiles, at most 64), whole 512-sample tiles; what the segments do not cover
    (n_samples - n_segments * seg_samples < n_segments * tile) is rendered serially behind them."""
    if per_rank <= 0:
        per_rank = max(1, min(64, n_samples // (world * 4 * tile)))
    n_segments = world * per_rank
    seg = n_samples // n_segments // tile * tile
    if seg == 0:
        raise ValueError("call too short to shard in time")
    return n_segments, seg


def exchange_segment_states(states, lo: int, hi: int, group=None):
    """The one exchange step of a pass: `states` is [n_voices, n_segments, words] (any integer dtype; on the
    device the process group's backend works on); every rank has just written the blocks of its own segments
    [lo, hi) and receives everybody else's — one all-gather of n_voices * (hi - lo) * words * 4 bytes per rank.
    In place.  No-op without a process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return states
    world = dist.get_world_size(group)
    if world == 1:
        return states
    V, S, W = states.shape
    per = hi - lo
    assert per * world == S, "segments must divide evenly over the ranks"
    mine = states[:, lo:hi, :].contiguous()
    everyone = torch.empty((world * V, per, W), dtype=states.dtype, device=states.device)  # rank-major
    dist.all_gather_into_tensor(everyone, mine, group=group)
    states.copy_(everyone.view(world, V, per, W).permute(1, 0, 2, 3).reshape(V, S, W))
    return states


class _DeviceWords:
    """A device buffer of uint32 words as a CUDA array, for torch.as_tensor (zero copy)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i4", "data": (ptr, False), "version": 3}


class TimeShard:
    """One rank's side of a time-sharded render (tb_segments_*): begin in the constructor, then for every pass
    k = 1..passes: run_pass(k, out_local); exchange `states`; fix(k) unless it was the last; finally end()."""

    def __init__(self, program, n_voices: int, n_segments: int, seg_samples: int, lo: int, hi: int, params=None,
                 device=None):
        import ctypes

        import torch

        from . import _abi
        self._abi, self._L, self.program = _abi, _abi.lib(), program
        self.lo, self.hi, self.seg_samples = lo, hi, seg_samples
        self.flags = _abi.TB_OUT_DEVICE
        pptr, n_params = None, 0
        if params is not None:
            if hasattr(params, "data_ptr"):
                pptr, n_params = ctypes.c_void_p(params.data_ptr()), params.shape[1]
                self.flags |= _abi.TB_PARAMS_DEVICE
            else:
                params = np.ascontiguousarray(params, dtype=np.float32)
                pptr, n_params = params.ctypes.data_as(ctypes.c_void_p), params.shape[1]
        self._params = params  # keep alive
        passes = ctypes.c_uint32(0)
        _abi.check(self._L.tb_segments_begin(program._h, pptr, n_params, n_voices, n_segments, seg_samples, self.flags,
                                             ctypes.byref(passes)))
        self.passes = int(passes.value)
        sp, sb = ctypes.c_void_p(), ctypes.c_uint64(0)
        _abi.check(self._L.tb_segments_states(program._h, ctypes.byref(sp), ctypes.byref(sb)))
        self.states = torch.as_tensor(_DeviceWords(sp.value, (n_voices, n_segments, sb.value // 4)),
                                      device=device if device is not None else "cuda")

    def run_pass(self, k: int, out_local):
        import ctypes
        assert out_local.shape[1] >= (self.hi - self.lo) * self.seg_samples and out_local.stride(1) == 1
        self._abi.check(self._L.tb_segments_pass(self.program._h, k, self.lo, self.hi,
                                                 ctypes.c_void_p(out_local.data_ptr()), out_local.stride(0), self.flags))

    def fix(self, k: int):
        self._abi.check(self._L.tb_segments_fix(self.program._h, k))

    def end(self):
        self._abi.check(self._L.tb_segments_end(self.program._h))


def render_time_sharded(program, out_local, n_voices: int, n_segments: int, seg_samples: int, rank: int, world: int,
                        params=None, group=None):
    """Rank `rank`'s share of the next n_segments * seg_samples samples of every voice: out_local is a CUDA
    tensor [n_voices, (n_segments / world) * seg_samples] that receives the rank's own time range.  Every
    rank calls this with the same program state (same tb_render history).  Returns the number of passes.
    The exchange between passes is exchange_segment_states over the group (NCCL on the GPU box)."""
    import torch
    lo, hi = segment_range(n_segments, rank, world)
    ts = TimeShard(program, n_voices, n_segments, seg_samples, lo, hi, params=params, device=out_local.device)
    mine = torch.cuda.ExternalStream(program.stream, device=out_local.device)
    for k in range(1, ts.passes + 1):
        ts.run_pass(k, out_local)
        if world > 1:
            # the pass runs on the program's stream, the collective on torch's: hand over both ways
            torch.cuda.current_stream().wait_stream(mine)
            exchange_segment_states(ts.states, lo, hi, group)
            mine.wait_stream(torch.cuda.current_stream())
        if k < ts.passes:
            ts.fix(k)
    ts.end()
    return ts.passes


def bind_to_gpu_numa(local_rank: int) -> dict:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (Linux sysfs), BEFORE it allocates pinned
    host buffers: pinned pages are then first-touched on that node and device-to-host copies do not cross the
    socket interconnect.  With 8 ranks draining rows over PCIe at once the host side is the bottleneck
    (bench.py e2e).  Returns what it found; a missing sysfs entry or a single-node box leaves the affinity alone."""
    import os
    info = {"numa_node": None, "cpus": None}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
    except Exception as e:  # noqa: BLE001 — best effort: the render does not depend on it
        info["error"] = str(e)[:80]
    return info
