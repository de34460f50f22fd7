"""Wall-clock of the general interpreter on config 2 (harmonica notes: Reset, Alt, Filter, Append, Fin) over a batch."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tuun_b200 import workloads as W
from tuun_b200.generator import Program
V, N = 4096, 88200
w = W.cfg2_harmonica(4)
p = Program(w, 44100)
out = torch.empty((V, N), dtype=torch.float32, device="cuda")
for rep in range(3):
    p.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p.render(out)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"cfg2 x {V} voices: {dt * 1e3:.1f} ms, {V * N / dt:.3e} voice-samples/s")
