"""Per-config throughput next to the CPU port: BASELINE.json's configs 1-4 and the five shapes of the
reference's own criterion benches (benches/tracker_benches.rs:19-166), each ONE voice — the shape the
reference runs — through tb_render with device rows, and through the oracle on one host core in
1024-sample blocks (tracker_benches.rs drives `generate` exactly so).

Imported by bench.py (key "configs" of the JSON line, N = 1) and runnable alone:
    python tools/config_bench.py [--reps 5]
GPU time is wall clock around tb_render + device synchronisation (launch overheads included: for a
44,032-sample render they are most of it), best of `reps` after one warm-up, state reset in between.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 44100


def shapes():
    """[(name, waveform, samples asked for)] — lengths of finite programs are found by the render."""
    from tuun_b200 import workloads as W
    out = [("cfg1 $440*Qw", W.cfg1_from_source(), 30000),
           ("cfg2 harmonica x4", W.cfg2_harmonica(4), 100000)]
    out += [("cfg3 " + name, w, 441000) for name, w in W.cfg3_fm_variations()]
    out += [("cfg4 " + name, w, 60 * SR) for name, w in W.cfg4_filters()]
    tb = W.tracker_benches()
    out += [("bench " + name, w, blocks * 1024) for name, w, blocks in tb]
    # the criterion shapes run for 1 s (43 blocks): a render that short is mostly launch latency on a GPU, so the
    # two constant-coefficient ones are also shown at config 4's length
    out += [("bench " + name + " x 60 s", w, 60 * SR) for name, w, _ in tb if name in ("filter_1_1", "filter_4_3")]
    return out


def time_gpu(w, n, reps):
    import torch
    from tuun_b200.generator import Program
    p = Program(w, SR)
    out = torch.empty((1, n), dtype=torch.float32, device="cuda")
    lens = np.zeros(1, dtype=np.uint64)
    p.render(out, out_len=lens)                      # warm-up; also the length
    length = int(lens[0])
    best = float("inf")
    for _ in range(reps):
        p.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        p.render(out)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    info = p.info
    return length, best, info


def time_cpu(w, n, min_seconds=0.25):
    from oracle.binding import OracleProgram
    o = OracleProgram(w, SR)
    reps, total, length = 0, 0.0, 0
    while total < min_seconds and reps < 50:
        o.initialize_state()
        t0 = time.perf_counter()
        length = len(o.render(n, block=1024))
        total += time.perf_counter() - t0
        reps += 1
    return length, total / reps


def run(reps=5, hbm_gbs=None, names=None):
    rows = []
    for name, w, n in shapes():
        if names and not any(k in name for k in names):
            continue
        glen, gdt, info = time_gpu(w, n, reps)
        clen, cdt = time_cpu(w, n)
        row = {"config": name, "samples": glen, "length_matches_cpu": glen == clen,
               "gpu_ms": gdt * 1e3, "gpu_value": glen / gdt, "cpu_1core_value": clen / cdt,
               "gpu_over_cpu_1core": (glen / gdt) / (clen / cdt),
               "split_segments": int(info.split_segments) if info.split_rounds else 0,
               "split_passes": int(info.split_passes)}
        if hbm_gbs:
            row["hbm_frac"] = 4.0 * glen / gdt / 1e9 / hbm_gbs
        rows.append(row)
    return rows


def run_batches(reps=3, hbm_gbs=None, voices=16384):
    """Config 2 as a batch (the four-note harmonica sequence, every voice the same tree): the default kernel
    selection — from 10,656 voices the lane-per-voice kernels, which take the notes' nested Resets and ADSR timeline
    (lanes.cuh ST_RESET_CLK / ST_SEG_*) — next to the general interpreter alone (TUUN_B200_LANES=0)."""
    import torch
    from tuun_b200 import workloads as W
    from tuun_b200.generator import Program
    w, n = W.cfg2_harmonica(4), 88200
    out = torch.empty((voices, n), dtype=torch.float32, device="cuda")
    rows = []
    keep = os.environ.get("TUUN_B200_LANES")
    try:
        for name, lanes in (("general interpreter", "0"), ("default selection", None)):
            if lanes is None:
                os.environ.pop("TUUN_B200_LANES", None)
            else:
                os.environ["TUUN_B200_LANES"] = lanes
            p = Program(w, SR)
            best = float("inf")
            for _ in range(reps + 1):
                p.reset()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                p.render(out)
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            info = p.info
            row = {"config": "cfg2 harmonica x4", "voices": voices, "samples_per_voice": n, "kernels": name,
                   "gpu_ms": best * 1e3, "gpu_value": voices * n / best, "kernel_launches_per_render": int(info.kernel_launches) // (reps + 1),
                   "lane_launches_per_render": int(info.lane_launches) // (reps + 1)}
            if hbm_gbs:
                row["hbm_frac"] = 4.0 * voices * n / best / 1e9 / hbm_gbs
            rows.append(row)
            del p
    finally:
        if keep is None:
            os.environ.pop("TUUN_B200_LANES", None)
        else:
            os.environ["TUUN_B200_LANES"] = keep
    del out
    torch.cuda.empty_cache()
    return rows


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", nargs="*")
    ap.add_argument("--json", action="store_true")
    args = ap.parse_args()
    peak = None
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peak = float(json.load(open(pp))["hbm_gbs"])
    rows = run(args.reps, peak, args.only)
    print(f"{'config':34s} {'samples':>9s} {'gpu ms':>9s} {'gpu v-s/s':>11s} {'cpu 1 core':>11s} {'ratio':>8s} {'split':>6s}")
    for r in rows:
        print(f"{r['config']:34s} {r['samples']:9d} {r['gpu_ms']:9.3f} {r['gpu_value']:11.3e} {r['cpu_1core_value']:11.3e} "
              f"{r['gpu_over_cpu_1core']:8.1f} {r['split_segments']:6d}" + ("" if r["length_matches_cpu"] else "  LENGTH MISMATCH"))
    batches = run_batches(args.reps, peak)
    for r in batches:
        print(f"{r['config']} x {r['voices']} voices, {r['kernels']}: {r['gpu_ms']:.2f} ms, {r['gpu_value']:.3e} voice-samples/s "
              f"({r['lane_launches_per_render']} lane launches of {r['kernel_launches_per_render']})")
    if args.json:
        print(json.dumps(rows + batches))


if __name__ == "__main__":
    main()
