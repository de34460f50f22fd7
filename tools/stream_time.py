"""The reference's calling pattern (main.rs:42-43, tracker.rs:597-642: generate() per block of 1024 samples) on the
65,536-voice FM + low-pass batch: device rows [V, block] reused every call, blocks of 1024 / 4096 / 16384 samples against one
call for the whole length.    python tools/stream_time.py [seconds]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tuun_b200.generator import Program
from tuun_b200.workloads import fm_filter_params, fm_filter_voice

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
V = 65536
N = int(round(secs * 44100)) // 16384 * 16384
params = torch.from_numpy(fm_filter_params(np.arange(V))).cuda()
for block in (1024, 4096, 16384, N):
    p = Program(fm_filter_voice(), 44100)
    out = torch.empty((V, block), dtype=torch.float32, device="cuda")
    best = 1e9
    for rep in range(3):
        p.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(N // block):
            p.render(out, params=params)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    calls = N // block
    print(f"block {block:7d}: {calls:4d} calls, {best * 1e3:8.2f} ms total, {best / calls * 1e6:8.1f} us a call, "
          f"{V * N / best:.3e} voice-samples/s (launches a call {p.info.kernel_launches / (3 * calls):.1f})")
