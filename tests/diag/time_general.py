"""Wall-clock of config 2 (harmonica notes: Reset in Reset, Alt, Filter, an ADSR timeline, a sequence of four notes) over a
batch: the general interpreter (TUUN_B200_LANES=0) against the lane-per-voice kernels (nested clocks and timelines in
the steady stream, lanes.cuh ST_RESET_CLK / ST_SEG_*).
    python tests/diag/time_general.py [voices ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tuun_b200 import workloads as W
from tuun_b200.generator import Program
N = 88200
w = W.cfg2_harmonica(4)
voices = [int(a) for a in sys.argv[1:]] or [4096]
for V in voices:
    out = torch.empty((V, N), dtype=torch.float32, device="cuda")
    rows = {}
    for lanes in ("0", "1"):
        os.environ["TUUN_B200_LANES"] = lanes
        p = Program(w, 44100)
        best = 1e9
        for rep in range(4):
            p.reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            p.render(out)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        i = p.info
        rows[lanes] = out[:: max(1, V // 8), :].cpu().numpy()
        print(f"cfg2 x {V} voices, lanes={lanes}: {best * 1e3:.2f} ms, {V * N / best:.3e} voice-samples/s "
              f"(kernel launches {i.kernel_launches}, lane launches {i.lane_launches}, lane_min_voices {i.lane_min_voices}, "
              f"lane smem {i.lane_smem_bytes})", flush=True)
        del p
    print(f"  max |lanes - general| = {float(np.max(np.abs(rows['0'] - rows['1']))):.3e}", flush=True)
    del out
