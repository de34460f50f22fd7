"""One GPU's share of the 65,536-voice batch under strong scaling: V = 65,536 / N voices x 10 s, device rows, wall clock
and CUDA events around tb_render.  python tools/strong_time.py [N ...]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tuun_b200.generator import Program
from tuun_b200.workloads import fm_filter_params, fm_filter_voice
N = 441000
for n_gpu in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    V = 65536 // n_gpu
    p = Program(fm_filter_voice(), 44100)
    params = torch.from_numpy(fm_filter_params(np.arange(V))).cuda()
    out = torch.empty((V, N), dtype=torch.float32, device="cuda")
    best = 1e9
    for _ in range(6):
        p.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        p.render(out, params=params)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    i = p.info
    print(f"N={n_gpu}: {V} voices: {best * 1e3:.2f} ms -> {V * N / best:.3e}/GPU, x{n_gpu} = {V * N * n_gpu / best:.3e}; "
          f"split segments {i.split_segments} x {i.split_seg_samples}, launches/call {i.kernel_launches // 6}, lane {i.lane_launches // 6}")
    del out
