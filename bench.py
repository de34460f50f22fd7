#!/usr/bin/env python
"""bench.py — rendered voice-samples/sec of the B200 renderer on BASELINE.json's config 5
(65,536 parameter-swept FM+filter voices x 10 s @ 44.1 kHz), next to the CPU oracle.

  python bench.py [--gpus N] [--steps K] [--warmup W]          the CUDA path (one rank per GPU)
  python bench.py --impl reference [...]                       the reference algorithm on host cores

A step renders the whole batch once from Initial state: every rank renders its own contiguous
range of voices into HBM (no data-path collective); the optional mixdown reduces the per-rank
partial mixes over NCCL.  Scaling is "weak" by default: every GPU renders --voices voices (voice
ids continue the parameter sweep, rank r takes [r*V, (r+1)*V)); --scaling strong divides the
--voices voices over the ranks instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLE_RATE = 44100
METRIC = "rendered voice-samples/sec"
UNIT = "voice-samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--voices", type=int, default=65536)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--e2e-voices", type=int, default=0, help="rows of the reused pinned host window (default 8192)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mix", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the secondary records: strong scaling, time sharding, exact sines, per-config table")
    ap.add_argument("--shard-voices", type=int, default=64, help="voices of the time-sharded render (x --shard-seconds)")
    ap.add_argument("--shard-seconds", type=float, default=600.0)
    return ap.parse_args()


def total_voices(args, world):
    return args.voices * world if args.scaling == "weak" else args.voices


def config(args, n_samples, world=None):
    world = world or args.gpus
    return {
        "workload": "cfg5: 65,536 parameter-swept FM+biquad voices x 10 s @ 44.1 kHz "
                    "(sine(2pi(fc+I*fm*sine(2pi*fm,pi/2)),0) | lpf(Q,cut)), one shared op list + [V x 8] f32 table",
        "voices": total_voices(args, world),
        "voices_per_gpu": total_voices(args, world) // world,
        "samples_per_voice": n_samples,
        "sample_rate": SAMPLE_RATE,
        "sharding": "contiguous voice ranges per GPU, no data-path collective"
                    + (" (weak: every GPU renders the full 65,536-voice sweep shape)" if args.scaling == "weak" else ""),
        "l2": "output rows (>= 14 GB per GPU) are far larger than L2; nothing is re-read between steps",
    }


class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples inside the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(args, n_samples, target_seconds, threads):
    """The CPU oracle (a port of generator.rs) timed on a bounded sample of the same workload:
    whole voices of the same length, as many as fit in about `target_seconds` of wall time."""
    from oracle.binding import OracleProgram
    from tuun_b200.workloads import fm_filter_params, fm_filter_voice
    o = OracleProgram(fm_filter_voice(), SAMPLE_RATE)
    probe_ids = (np.arange(threads) * 40503 + 49230) % args.voices
    t = time.perf_counter()
    o.render_batch(fm_filter_params(probe_ids), len(probe_ids), min(n_samples, 44100), keep=False, threads=threads)
    dt = time.perf_counter() - t
    rate = len(probe_ids) * min(n_samples, 44100) / max(dt, 1e-6)
    n_v = int(max(threads, min(args.voices, rate * target_seconds / n_samples)))
    n_v = max(threads, (n_v // threads) * threads)
    ids = (np.arange(n_v) * 40503 + 49230) % args.voices
    t = time.perf_counter()
    _, _, _, total = o.render_batch(fm_filter_params(ids), n_v, n_samples, keep=False, threads=threads)
    dt = time.perf_counter() - t
    return {"value": total / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_v} voices (ids (v*40503+49230) mod {args.voices}) x {n_samples} samples, 1024-sample blocks, "
                      f"{dt:.1f} s wall"}, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_samples = int(round(args.seconds * SAMPLE_RATE))
    threads = os.cpu_count() or 1
    per_step = max(2.0, 60.0 / max(1, args.steps + args.warmup))
    vals, times = [], []
    sample = ""
    for i in range(args.warmup + args.steps):
        cb, dt = cpu_baseline(args, n_samples, per_step, threads)
        if i >= args.warmup:
            vals.append(cb["value"]); times.append(dt)
        sample = cb["sample"]
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32 (f64 phase)",
            "data": "synthetic", "config": config(args, n_samples),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the Rust reference cannot be built here (no cargo/rustc); this arm times the C++ "
                    "restatement of generator.rs (oracle/) on all host threads"}
    print(json.dumps(line))


def timed_steps(step, steps, stream, barrier, world):
    """`steps` calls of step() bracketed by CUDA events on `stream`; device milliseconds, max over ranks."""
    import torch
    import torch.distributed as dist
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    for a, b in evs:
        a.record(stream)
        step()
        b.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = float(sum(a.elapsed_time(b) for a, b in evs))
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def strong_record(args, world, rank, local, n_samples, barrier, peak):
    """north_star's own shape: the 65,536-voice batch DIVIDED over the ranks (contiguous voice ranges, no
    data-path collective).  Reported next to the weak-scaling `value`; efficiency is value / (N x the N = 1 value)."""
    import torch
    from tuun_b200.generator import Program
    from tuun_b200.sharding import voice_range
    from tuun_b200.workloads import fm_filter_params, fm_filter_voice
    lo, hi = voice_range(args.voices, rank, world)
    params = torch.from_numpy(fm_filter_params(np.arange(lo, hi))).cuda()
    prog = Program(fm_filter_voice(), SAMPLE_RATE, device=local)
    stream = torch.cuda.ExternalStream(prog.stream, device=local)
    out = torch.empty((hi - lo, n_samples), dtype=torch.float32, device="cuda")

    def step():
        prog.reset()
        prog.render(out, params=params)

    for _ in range(3):
        step()
    steps = max(3, args.steps)
    ms = timed_steps(step, steps, stream, barrier, world)
    info = prog.info
    value = args.voices * n_samples * steps / (ms * 1e-3)
    kernel = (("tb_render_lanes_fm_ws_split_kernel" if info.fm_ws_launches else "tb_render_lanes_fm_split_kernel") +
              " after tb_render_lanes_fm_sums_kernel (every voice cut in time: phase-sum pass, "
              "filter warm-up, samples)" if info.lane_launches and info.lane_fm_capacity and info.split_rounds
              else "tb_render_lanes_fm_ws_kernel" if info.lane_launches and info.fm_ws_launches
              else "tb_render_lanes_fm_kernel" if info.lane_launches and info.lane_fm_capacity
              else "tb_render_lanes_kernel" if info.lane_launches else "tb_render_kernel")
    return {"value": value, "unit": UNIT, "ms_per_step": ms / steps, "voices": args.voices,
            "voices_per_gpu": hi - lo, "kernel": kernel, "split_segments": int(info.split_segments),
            "hbm_frac_per_gpu": 4.0 * (hi - lo) * n_samples / (ms / steps * 1e-3) / 1e9 / peak,
            "scaling": "strong: total work fixed at the 65,536-voice batch"}


def time_shard_record(args, world, rank, local, barrier, peak):
    """Few voices, a long render: --shard-voices FM + low-pass voices x --shard-seconds, every rank rendering its
    own TIME range of every voice (tb_segments_*: one all-gather of the segments' state blocks per pass over
    NCCL).  At N = 1 this is what tb_render does by itself for a small batch."""
    import torch
    import torch.distributed as dist
    from tuun_b200.generator import Program
    from tuun_b200.sharding import plan_segments, render_time_sharded
    from tuun_b200.workloads import fm_filter_params, fm_filter_sample_ids, fm_filter_voice
    V = args.shard_voices
    n = int(round(args.shard_seconds * SAMPLE_RATE))
    head = 256
    # segments of at least 16,384 samples (room for the filters' warm-up: the cheaper form of the split), at most 512 a rank
    S, seg = plan_segments(n - head, world, per_rank=max(32, min(512, (n - head) // (world * 16384))))
    params = torch.from_numpy(fm_filter_params(fm_filter_sample_ids(V))).cuda()
    prog = Program(fm_filter_voice(), SAMPLE_RATE, device=local)
    stream = torch.cuda.ExternalStream(prog.stream, device=local)
    head_out = torch.empty((V, head), dtype=torch.float32, device="cuda")
    out = torch.empty((V, S // world * seg), dtype=torch.float32, device="cuda")
    passes = [0]

    def step():
        prog.reset()
        prog.render(head_out, params=params)          # the first tile of the stream: every rank, serial
        passes[0] = render_time_sharded(prog, out, V, S, seg, rank, world, params=params)
        stream.wait_stream(torch.cuda.current_stream())

    for _ in range(2):
        step()
    steps = 3
    ms = timed_steps(step, steps, stream, barrier, world)
    samples = head + S * seg
    value = V * samples * steps / (ms * 1e-3)
    words = prog.info.state_words
    # what this rank rendered against the same voices rendered serially (no split of any kind) on this rank
    os.environ["TUUN_B200_SPLIT"] = "0"
    try:
        ser = Program(fm_filter_voice(), SAMPLE_RATE, device=local)
        full = torch.empty((V, samples), dtype=torch.float32, device="cuda")
        ser.render(full, params=params)
        torch.cuda.synchronize()
    finally:
        del os.environ["TUUN_B200_SPLIT"]
    lo = head + rank * (S // world) * seg
    diff = (full[:, lo:lo + (S // world) * seg] - out).abs().amax(dim=1)
    dmax = torch.tensor([float(diff.max()), float(diff.median())], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dmax, op=dist.ReduceOp.MAX)
    del full, ser
    return {"value": value, "unit": UNIT, "ms_per_step": ms / steps, "voices": V, "samples_per_voice": samples,
            "segments": S, "segment_samples": seg, "passes": passes[0],
            "exchange_bytes_per_pass_per_rank": int(V * (S // world) * words * 4),
            "max_abs_diff_vs_serial": float(dmax[0].item()), "median_voice_diff_vs_serial": float(dmax[1].item()),
            "diff_note": "every rank's time range against the same 64 voices rendered serially on that rank (max over ranks); "
                         "the difference is the biquads' round-off noise (affine scan vs serial recurrence)",
            "hbm_frac_per_gpu": 4.0 * V * samples / world / (ms / steps * 1e-3) / 1e9 / peak,
            "sharding": "time: rank r renders segments [r S/N, (r+1) S/N) of every voice; states all-gathered per pass"}


def streaming_record(args, local, n_local, n_samples, params_d):
    """The reference's own calling pattern (main.rs:42-43, tracker.rs:597-642: generate() once per block of 1024
    samples): the same batch rendered block by block into one reused [voices, 1024] device buffer, state carried by
    the program from call to call; CUDA events around the whole stream of calls."""
    import torch
    from tuun_b200.generator import Program
    from tuun_b200.workloads import fm_filter_voice
    block = 1024
    n = n_samples // block * block
    prog = Program(fm_filter_voice(), SAMPLE_RATE, device=local)
    stream = torch.cuda.ExternalStream(prog.stream, device=local)
    buf = torch.empty((n_local, block), dtype=torch.float32, device=f"cuda:{local}")
    best = None
    for _ in range(2):
        prog.reset()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(stream)
        for _ in range(n // block):
            prog.render(buf, params=params_d)
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    return {"value": n_local * n / (best * 1e-3), "unit": UNIT, "block_samples": block, "calls": n // block,
            "us_per_call": best * 1e3 / (n // block), "voices": n_local, "samples_per_voice": n,
            "note": "one tb_render per 1024-sample block (the reference's block size), one launch a call, rows in HBM"}


def exact_sines_record(args, local, n_local, n_samples, params_d, out, peak):
    """The same batch with every sine in the EXACT class (TUUN_B200_FAST_SINES=0: f64 polynomial on the 64-bit
    phase, the f32 the reference's libm sin rounds to on 99.98 % of samples) — what the FAST-class carrier
    (MUFU.SIN on 23 phase bits, the default) buys, and the error of both against the oracle on a sample."""
    import torch
    from oracle.binding import OracleProgram
    from tuun_b200.generator import Program
    from tuun_b200.workloads import fm_filter_params, fm_filter_cover_ids, fm_filter_voice
    os.environ["TUUN_B200_FAST_SINES"] = "0"
    try:
        prog = Program(fm_filter_voice(), SAMPLE_RATE, device=local)
    finally:
        del os.environ["TUUN_B200_FAST_SINES"]
    stream = torch.cuda.ExternalStream(prog.stream, device=local)

    def step():
        prog.reset()
        prog.render(out, params=params_d)

    step()
    step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(2):
        step()
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 2
    info = prog.info
    # error of both classes on 64 covering voices x 2 s (the full-length comparison is tests/test_gpu_cfg5_full.py)
    ids = fm_filter_cover_ids(1)[::4]
    pr = fm_filter_params(ids)
    n_err = 2 * SAMPLE_RATE
    ref, _, _, _ = OracleProgram(fm_filter_voice(), SAMPLE_RATE).render_batch(pr, len(ids), n_err, threads=os.cpu_count() or 1)
    errs = {}
    for name, env in (("exact", "0"), ("fast", None)):
        if env is not None:
            os.environ["TUUN_B200_FAST_SINES"] = env
        os.environ["TUUN_B200_LANE_MIN_VOICES"] = "1"
        try:
            q = Program(fm_filter_voice(), SAMPLE_RATE, device=local)
            got = np.zeros((len(ids), n_err), dtype=np.float32)
            q.render(got, params=pr)
        finally:
            os.environ.pop("TUUN_B200_FAST_SINES", None)
            os.environ.pop("TUUN_B200_LANE_MIN_VOICES", None)
        errs[name] = float(np.abs(got - ref).max())
    value = n_local * n_samples / (ms * 1e-3)
    return {"value": value, "unit": UNIT, "ms_per_step": ms, "hbm_frac": 4.0 * value / 1e9 / peak,
            "kernel": "tb_render_lanes_kernel" if info.lane_launches else "tb_render_kernel",
            "max_abs_err_vs_oracle": errs["exact"], "fast_class_max_abs_err_vs_oracle": errs["fast"],
            "err_sample": f"{len(ids)} voices covering every index / ratio / cutoff class x {n_err} samples, lane kernels",
            "note": "TUUN_B200_FAST_SINES=0: both sines of a voice by the f64 polynomial"}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist

    from tuun_b200.generator import Program
    from tuun_b200.workloads import fm_filter_params, fm_filter_voice

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the renderer has no CPU path")
    torch.cuda.set_device(local)
    from tuun_b200.sharding import bind_to_gpu_numa
    numa = bind_to_gpu_numa(local) if world > 1 else {"numa_node": None}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_samples = int(round(args.seconds * SAMPLE_RATE))
    from tuun_b200.sharding import reduce_mix, voice_range, weak_voice_range
    n_total = total_voices(args, world)
    lo, hi = weak_voice_range(args.voices, rank) if args.scaling == "weak" else voice_range(n_total, rank, world)
    n_local = hi - lo
    params_h = fm_filter_params(np.arange(lo, hi))
    params_d = torch.from_numpy(params_h).cuda()
    prog = Program(fm_filter_voice(), SAMPLE_RATE, device=local)
    stream = torch.cuda.ExternalStream(prog.stream, device=local)
    out = torch.empty((n_local, n_samples), dtype=torch.float32, device="cuda")
    lens = np.zeros(n_local, dtype=np.uint64)

    def step():
        prog.reset()
        prog.render(out, params=params_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = prog.info.kernel_launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t0 = time.time()
    w0 = time.perf_counter()
    for a, b in evs:
        a.record(stream)
        step()
        b.record(stream)
    torch.cuda.synchronize()
    barrier()
    w1 = time.perf_counter()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if sampler else None
    launches = prog.info.kernel_launches - launches0
    kern_ms = [a.elapsed_time(b) for a, b in evs]
    dev_ms = float(sum(kern_ms))
    tmax = torch.tensor([dev_ms, (w1 - w0) * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax[0].item())
    lens_ok = True
    prog.reset()
    chk = prog.render(out, params=params_d, out_len=lens)
    lens_ok = bool((chk == n_samples).all())
    value = n_total * n_samples * args.steps / (total_ms * 1e-3)

    # roofline of the one kernel: 4 B stored per voice-sample, per launch
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback"
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
    # The dominant kernel: tb_render_lanes_kernel when the batch takes the lane-per-voice path (its
    # own CUDA-event times on the launching stream, tb_lane_kernel_times; it renders all samples of
    # a call but the first 256-sample tile and the last < 16), else the one tb_render_kernel launch.
    info = prog.info
    lane_ms = prog.lane_kernel_times(args.steps) if info.lane_launches else np.zeros(0)
    if len(lane_ms):
        groups = (n_local + 63) // 64
        kernel = ("tb_render_lanes_fm_ws_kernel" if info.fm_ws_launches
                  else "tb_render_lanes_fm_kernel" if info.lane_fm_capacity and groups <= info.lane_fm_capacity
                  else "tb_render_lanes_queue_kernel" if groups > info.lane_capacity else "tb_render_lanes_kernel")
        # the fused-FM-voice kernel renders the whole call in one launch; the interpreter kernels leave the first
        # 256-sample tile and the < 16 samples past the last lane tile to the general kernel
        lane_samples = n_samples if launches == args.steps else (n_samples - 256) // 16 * 16
        alg_bytes = 4.0 * n_local * lane_samples
        avg_launch_s = float(np.mean(lane_ms)) * 1e-3
        dominant_share = float(np.sum(lane_ms)) / dev_ms if len(lane_ms) == args.steps else None
    else:
        kernel = "tb_render_kernel"
        lane_samples = n_samples
        alg_bytes = 4.0 * n_local * n_samples
        avg_launch_s = (dev_ms / max(1, launches)) * 1e-3
        dominant_share = 1.0
    achieved = alg_bytes / avg_launch_s / 1e9
    # DRAM traffic of the kernel from the committed ncu --set full capture (profiles/), which ran
    # the same kernel and per-voice work on a smaller launch: bytes per voice-sample x this launch.
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("kernel", "tb_render_kernel") == kernel:
            traffic = float(tj["dram_bytes_per_voice_sample"]) * n_local * lane_samples
            traffic_src = tj["source"]
    # pure-store stream over the same rows, same run: cudaMemsetAsync, best of 3, CUDA events
    store_gbs = 0.0
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out.zero_()
        b.record()
        torch.cuda.synchronize()
        store_gbs = max(store_gbs, out.numel() * 4 / (a.elapsed_time(b) * 1e-3) / 1e9)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "store_peak": store_gbs, "frac_of_store_peak": achieved / store_gbs if store_gbs else None,
                "store_peak_how": "cudaMemsetAsync over this run's output rows, best of 3 (a pure-store stream; `peak` is a "
                                  "read+write copy)",
                "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel,
                "share_of_step": dominant_share,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "avg_launch_ms": avg_launch_s * 1e3,
                "note": "bound by instruction issue and the 16-lane conversion/special-function unit, not by HBM: "
                        "see profiles/README.md (thread-instructions per voice-sample, pipe utilisation)"}

    # end to end through the C ABI with HOST buffers: H2D of the parameter table, D2H of every row.
    # All local voices are rendered; the pinned host window (e2e_group rows) is reused group after
    # group, the way a consumer draining the rows would.
    e2e = None
    if not args.no_e2e:
        # pinned host window per rank: 8192 rows (14.4 GB) on one GPU, shrunk with the rank count so that the
        # box's pinned memory stays the same when 8 ranks run side by side
        grp = min(n_local, args.e2e_voices or max(1024, 8192 // world))
        host = torch.empty((grp, n_samples), dtype=torch.float32, pin_memory=True)
        host_np = host.numpy()
        ph = params_h
        e2e_steps = max(1, min(args.steps, 2))
        prog_h = Program(fm_filter_voice(), SAMPLE_RATE, device=local)

        def e2e_step():
            for a in range(0, n_local, grp):
                b = min(n_local, a + grp)
                prog_h.reset()
                prog_h.render(host_np[: b - a], params=ph[a:b])

        prog_h.render(host_np, params=ph[:grp])  # warm: staging buffers, page touch
        barrier()
        e0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e1 = time.perf_counter()
        et = torch.tensor([e1 - e0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        # the ceiling of that path: the same rows device -> pinned host by plain copies, all ranks at once
        dev_win = out[:grp]
        barrier()
        c0 = time.perf_counter()
        for a in range(0, n_local, grp):
            host[: min(grp, n_local - a)].copy_(dev_win[: min(grp, n_local - a)], non_blocking=True)
        torch.cuda.synchronize()
        c1 = time.perf_counter()
        ct = torch.tensor([c1 - c0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ct, op=dist.ReduceOp.MAX)
        d2h_gbs = n_local * n_samples * 4 / float(ct.item()) / 1e9
        e2e_rate = n_total * n_samples * e2e_steps / float(et.item())
        e2e = {"value": n_total * n_samples * e2e_steps / float(et.item()), "unit": UNIT,
               "d2h_ceiling_gbs_per_gpu": d2h_gbs,
               "frac_of_d2h_ceiling": (e2e_rate / world * 4 / 1e9) / d2h_gbs if d2h_gbs else None,
               "d2h_ceiling_how": "the same rows copied device -> the same pinned window with cudaMemcpyAsync alone, all ranks "
                                  "at once, same run",
               "h2d_bytes_per_step": int(n_local * 8 * 4), "d2h_bytes_per_step": int(n_local * n_samples * 4 + n_local * 8),
               "voices": n_total, "steps": e2e_steps, "host_window_rows": grp, "numa": numa,
               "path": "tb_render with pinned host rows: voice groups rendered into 2 device staging buffers, "
                       "each group leaves with one cudaMemcpyAsync on a second stream while the next renders"}
        del host, host_np, prog_h

    # optional mixdown: per-GPU mix on the device, NCCL reduce of the [n_samples] partials over NVLink
    mixdown = None
    if not args.no_mix:
        mix_d = torch.empty(n_samples, dtype=torch.float32, device="cuda")
        prog_m = Program(fm_filter_voice(), SAMPLE_RATE, device=local)
        mstream = torch.cuda.ExternalStream(prog_m.stream, device=local)
        done_ev = torch.cuda.Event()

        def mix_step():
            prog_m.reset()
            prog_m.render_mix(mix_d, n_local, params=params_d)
            if world > 1:
                done_ev.record(mstream)
                torch.cuda.current_stream().wait_event(done_ev)
                reduce_mix(mix_d, dst=0)

        mix_step()
        barrier()
        m0 = time.perf_counter()
        msteps = max(1, min(args.steps, 2))
        for _ in range(msteps):
            mix_step()
        torch.cuda.synchronize()
        barrier()
        m1 = time.perf_counter()
        mt = torch.tensor([m1 - m0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(mt, op=dist.ReduceOp.MAX)
        # the same with a HOST consumer of the mix (the reference-shaped contract of tracker.rs:597-642: one mono
        # buffer leaves): H2D of the parameter table from pinned memory and D2H of [n_samples] floats inside the timed region
        mix_h = torch.empty(n_samples, dtype=torch.float32, pin_memory=True)
        params_pin = torch.from_numpy(params_h).pin_memory()
        params_dev = torch.empty_like(params_d)

        def mix_e2e_step():
            params_dev.copy_(params_pin, non_blocking=True)
            mstream.wait_stream(torch.cuda.current_stream())
            prog_m.reset()
            prog_m.render_mix(mix_d, n_local, params=params_dev)
            done_ev.record(mstream)
            torch.cuda.current_stream().wait_event(done_ev)
            if world > 1:
                reduce_mix(mix_d, dst=0)
            if rank == 0:
                mix_h.copy_(mix_d, non_blocking=True)
            torch.cuda.synchronize()

        mix_e2e_step()
        barrier()
        x0 = time.perf_counter()
        for _ in range(msteps):
            mix_e2e_step()
        barrier()
        x1 = time.perf_counter()
        xt = torch.tensor([x1 - x0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(xt, op=dist.ReduceOp.MAX)
        mixdown = {"value": n_total * n_samples * msteps / float(mt.item()), "unit": UNIT,
                   "e2e": {"value": n_total * n_samples * msteps / float(xt.item()), "unit": UNIT,
                           "h2d_bytes_per_step": int(n_local * 8 * 4), "d2h_bytes_per_step": int(n_samples * 4),
                           "path": "pinned parameter table -> device, tb_render_mix, NCCL reduce, mix -> pinned host"},
                   "ms_per_step": 1e3 * float(mt.item()) / msteps, "nccl_reduce_bytes": int(n_samples * 4) if world > 1 else 0,
                   "mode": "tb_render_mix(TB_NO_VOICE_OUT | TB_OUT_DEVICE) per rank (voices summed on the chip inside the lane "
                           "kernel, per-warp partial rows added in order), then ncclReduce(sum,f32) to rank 0"}
        del prog_m

    # secondary records (none of them inside `value`)
    extras = {}
    if not args.no_extras:
        if world == 1:
            extras["exact_sines"] = exact_sines_record(args, local, n_local, n_samples, params_d, out, peak)
            extras["streaming_1024"] = streaming_record(args, local, n_local, n_samples, params_d)
        del out
        torch.cuda.empty_cache()
        if world > 1 and args.scaling == "weak":
            extras["strong"] = strong_record(args, world, rank, local, n_samples, barrier, peak)
        extras["time_shard"] = time_shard_record(args, world, rank, local, barrier, peak)
        if rank == 0 and world == 1:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import config_bench
            extras["configs"] = config_bench.run(reps=3, hbm_gbs=peak)
            try:
                extras["config_batches"] = config_bench.run_batches(reps=2, hbm_gbs=peak)
            except Exception as e:  # a secondary record must not cost the line
                extras["config_batches"] = {"error": str(e)[:200]}
            extras["configs_note"] = ("one voice each (the reference's own bench shape), wall clock of tb_render with device rows "
                                      "against the CPU port on one core in 1024-sample blocks; see tools/config_bench.py")
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f32 samples, u64 fixed-point phase, f64 sine core (EXACT class), MUFU sine (FAST class)",
                "data": "synthetic", "config": config(args, n_samples, world), "clocks": clocks,
                "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e, "mixdown": mixdown,
                "wall_ms_per_step": float(tmax[1].item()) / args.steps, "lengths_ok": lens_ok}
        line.update(extras)
        if not args.no_cpu:
            cb, _ = cpu_baseline(args, n_samples, args.cpu_seconds, os.cpu_count() or 1)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
