"""Diagnostic (GPU): EVERY voice of config 5's 200 Hz cutoff class (ids 0..16383: the six filter shapes with
round-off noise gain >= 100 live here) for the full 441,000 samples on the default kernel, against the
oracle.  Prints the error distribution per Q; decides whether a flat 1e-4 holds for the whole sweep."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle.binding import OracleProgram
from tuun_b200.generator import Program
from tuun_b200.workloads import fm_filter_params, fm_filter_voice

SR, N = 44100, 441000
lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 16384
ids = np.arange(lo, hi)
V = len(ids)
w = fm_filter_voice()
params = fm_filter_params(ids)
p = Program(w, SR)
out = torch.empty((V, N), dtype=torch.float32, device="cuda")
p.render(out, params=params)
torch.cuda.synchronize()
print("launches", p.info.lane_launches, p.info.kernel_launches, flush=True)
o = OracleProgram(w, SR)
err = np.zeros(V)
CH = 1024
t0 = time.time()
for a in range(0, V, CH):
    b = min(V, a + CH)
    ref, _, _, _ = o.render_batch(params[a:b], b - a, N, threads=os.cpu_count())
    got = out[a:b].cpu().numpy()
    err[a:b] = np.abs(got - ref).max(axis=1)
    print(a, f"{err[a:b].max():.3e}", f"{time.time() - t0:.0f}s", flush=True)
for q in range(7):
    m = (ids % 7) == q
    e = err[m]
    print(f"Q={0.5 + 0.25 * q:.2f}: voices {m.sum()}  max {e.max():.3e}  p99.9 {np.quantile(e, 0.999):.3e}  median {np.median(e):.3e}  >1e-4: {(e > 1e-4).sum()}")
worst = np.argsort(err)[-8:]
print("worst:", [(int(ids[i]), float(err[i])) for i in worst])
np.save("gpurun_out/r2_cfg5_highgain_err.npy", err)
