"""Kernel time of the headline shape at a reduced length: 65,536 FM + low-pass voices x SECONDS (default 2 s),
device rows, the lane kernel's own CUDA-event times (tb_lane_kernel_times).  A/B aid for kernel edits:
    TUUN_B200_LIB=/path/to/other/libtuun_b200.so python tools/fm_time.py [seconds] [voices]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tuun_b200.generator import Program
from tuun_b200.workloads import fm_filter_params, fm_filter_voice

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
V = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
N = int(round(secs * 44100))
p = Program(fm_filter_voice(), 44100)
params = torch.from_numpy(fm_filter_params(np.arange(V))).cuda()
out = torch.empty((V, N), dtype=torch.float32, device="cuda")
for _ in range(8):
    p.reset()
    p.render(out, params=params)
torch.cuda.synchronize()
ms = p.lane_kernel_times(5)
print(f"{os.environ.get('TUUN_B200_LIB', 'default')}: V={V} N={N} lane kernel ms {np.round(ms, 3)} -> "
      f"{V * N / (np.median(ms) * 1e-3):.4e} voice-samples/s, launches {p.info.kernel_launches}")
