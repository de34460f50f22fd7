"""Repeat the time-cut render of 8,192 FM voices with both forms of the FM voice kernel and compare everything with the
first one-thread render: which form, if any, is not reproducible."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tuun_b200.generator import Program
from tuun_b200.workloads import fm_filter_params, fm_filter_sample_ids, fm_filter_voice

w = fm_filter_voice()
V, n = 8192, 176440
prm = fm_filter_params(fm_filter_sample_ids(V))
params = torch.from_numpy(prm).cuda()


def run(ws):
    os.environ["TUUN_B200_FM_WS"] = str(ws)
    p = Program(w, 44100)
    out = torch.zeros((V, n), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()  # the fill runs on torch's stream, the render on the program's own
    lens = p.render(out, params=params, out_len=np.zeros(V, dtype=np.uint64))
    assert (np.asarray(lens) == n).all()
    return out


ref = run(0)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10):
    for ws in (0, 1):
        out = run(ws)
        if not torch.equal(out.view(torch.int32), ref.view(torch.int32)):
            d = (out - ref).abs()
            rows = torch.nonzero(d.amax(dim=1) > 0).flatten().cpu().numpy()
            print(f"iteration {it} ws={ws}: {len(rows)} rows differ, max {float(d.max()):.3e}, rows {rows[:10]}")
            for r in rows[:4]:
                idx = torch.nonzero(d[r] > 0).flatten()
                print(f"   row {r}: m={prm[r, 1]:.6g} c={prm[r, 2]:.6g} samples {int(idx[0])}..{int(idx[-1])} ({len(idx)}), max {float(d[r].max()):.3e}")
        del out
print("done")
