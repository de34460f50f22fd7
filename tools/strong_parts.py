import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tuun_b200.generator import Program
from tuun_b200.workloads import fm_filter_params, fm_filter_voice
N = 441000
V = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
p = Program(fm_filter_voice(), 44100)
params = torch.from_numpy(fm_filter_params(np.arange(V))).cuda()
out = torch.empty((V, N), dtype=torch.float32, device="cuda")
for _ in range(4):
    p.reset()
    p.render(out, params=params)
torch.cuda.synchronize()
i = p.info
print("lane kernel ms of the last launches:", np.round(p.lane_kernel_times(6), 3), "segments", i.split_segments, i.split_seg_samples)
