"""Where the time of tb_render_mix(TB_NO_VOICE_OUT) goes on the headline batch: wall clock of the call, the lane
kernel's own time (events), the rest = head tile + the two tb_mix_kernel launches."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tuun_b200.generator import Program
from tuun_b200.workloads import fm_filter_params, fm_filter_voice
V, N = 65536, 441000
p = Program(fm_filter_voice(), 44100)
params = torch.from_numpy(fm_filter_params(np.arange(V))).cuda()
mix = torch.empty(N, dtype=torch.float32, device="cuda")
for _ in range(3):
    p.reset(); p.render_mix(mix, V, params=params)
torch.cuda.synchronize()
ts = []
for _ in range(4):
    p.reset(); torch.cuda.synchronize(); t0 = time.perf_counter()
    p.render_mix(mix, V, params=params); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("call ms", np.round(np.array(ts) * 1e3, 2), "lane kernel ms", np.round(p.lane_kernel_times(4), 2), "launches/call", p.info.kernel_launches // 7)
