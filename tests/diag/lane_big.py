"""Diagnostics: one harmonica note (config 2) as batches of growing size on the lane kernels."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tuun_b200 import workloads as W
from tuun_b200.generator import Program
note = W.cfg2_harmonica(2).a
N = 22050
for V in [int(a) for a in sys.argv[1:]] or [4096, 9472, 9473, 16384]:
    for q in (None, "0", "1"):
        if q is None:
            os.environ.pop("TUUN_B200_LANE_QUEUE", None)
        else:
            os.environ["TUUN_B200_LANE_QUEUE"] = q
        try:
            p = Program(note, 44100)
            out = torch.empty((V, N), dtype=torch.float32, device="cuda")
            best = 1e9
            for rep in range(3):
                p.reset()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                p.render(out)
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            i = p.info
            print(f"V={V} queue={q}: {best*1e3:.2f} ms {V*N/best:.3e}/s lane launches {i.lane_launches} capacity {i.lane_capacity} smem {i.lane_smem_bytes}", flush=True)
        except Exception as e:
            print(f"V={V} queue={q}: {e}", flush=True)
