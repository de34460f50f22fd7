"""A few FM + low-pass voices for a long time through plain tb_render (device rows): which split form the library
picks and how long it takes.  python tools/few_voices_time.py [voices] [seconds]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tuun_b200.generator import Program
from tuun_b200.workloads import fm_filter_params, fm_filter_sample_ids, fm_filter_voice
V = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(float(sys.argv[2]) * 44100) if len(sys.argv) > 2 else 600 * 44100
p = Program(fm_filter_voice(), 44100)
params = torch.from_numpy(fm_filter_params(fm_filter_sample_ids(V))).cuda()
out = torch.empty((V, N), dtype=torch.float32, device="cuda")
best = 1e9
for _ in range(5):
    p.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p.render(out, params=params)
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
i = p.info
print(f"{V} voices x {N} samples: {best * 1e3:.2f} ms -> {V * N / best:.3e} voice-samples/s; segments {i.split_segments} x "
      f"{i.split_seg_samples}, lane launches per call {i.lane_launches // 5}")
