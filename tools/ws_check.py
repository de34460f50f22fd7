"""A/B of the fused-FM-voice kernel's two forms (one thread a voice: lanes_fm.cu; a phase warp and a tone warp per
32 voices: lanes_fm_ws.cu): bit-identity of rows, lengths, carried state (a second call) and the on-chip mixdown over
ragged batch sizes and call lengths, then the kernel time of the headline shape.
    python tools/ws_check.py [seconds]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tuun_b200.generator import Program
from tuun_b200.workloads import fm_filter_params, fm_filter_voice, fm_filter_sample_ids

os.environ["TUUN_B200_LANE_MIN_VOICES"] = "1"
w = fm_filter_voice()


def render(ws, V, calls, mix=False):
    os.environ["TUUN_B200_FM_WS"] = "1" if ws else "0"
    p = Program(w, 44100)
    params = torch.from_numpy(fm_filter_params(fm_filter_sample_ids(V))).cuda()
    outs = []
    for n in calls:
        out = torch.zeros((V, n), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()  # (the fills run on torch's stream, the renders on the program's own)
        if mix:
            m = torch.zeros((n,), dtype=torch.float32, device="cuda")
            p.render_mix(m, V, params=params)
            torch.cuda.synchronize()
            outs.append(m.cpu().numpy())
        else:
            lens = p.render(out, params=params, out_len=np.zeros(V, dtype=np.uint64))
            assert (np.asarray(lens) == n).all(), (lens[:8], n)
            outs.append(out.cpu().numpy())
    info = p.info
    return outs, int(info.fm_ws_launches), int(info.lane_launches)


bad = 0
for V, calls in [(64, [4096, 1000]), (100, [4101, 16, 31, 777]), (12352, [8192 + 5, 4096]), (10007, [2048 + 13]),
                 (33, [17, 15, 16, 48])]:
    a, wa, la = render(False, V, calls)
    b, wb, lb = render(True, V, calls)
    same = all(np.array_equal(x.view(np.uint32), y.view(np.uint32)) for x, y in zip(a, b))
    print(f"V={V} calls={calls}: ws launches {wb}/{lb} (off: {wa}/{la}) bit-identical rows: {same}")
    if not same:
        bad += 1
        for i, (x, y) in enumerate(zip(a, b)):
            d = np.abs(x - y)
            v, s = np.unravel_index(np.argmax(d), d.shape)
            print(f"   call {i}: max diff {d.max():.3e} at voice {v} sample {s}; rows differing {(d.max(axis=1) > 0).sum()}")
try:
    for V, calls in [(4096, [4096 + 7, 640]), (1000, [8000])]:
        a, wa, la = render(False, V, calls, mix=True)
        b, wb, lb = render(True, V, calls, mix=True)
        same = all(np.array_equal(x.view(np.uint32), y.view(np.uint32)) for x, y in zip(a, b))
        print(f"mix V={V} calls={calls}: ws launches {wb}/{lb} bit-identical: {same}")
        bad += 0 if same else 1
except Exception as e:  # noqa: BLE001
    print("mix check skipped:", repr(e))
# a batch too small for the lane kernel is cut in time (abi.cpp render_split_fm): warm-up and samples pass on virtual voices
del os.environ["TUUN_B200_LANE_MIN_VOICES"]
for V, n in [(8192, 176400 + 40), (300, 441000)]:
    res = []
    for ws in (0, 1):
        os.environ["TUUN_B200_FM_WS"] = str(ws)
        p = Program(w, 44100)
        params = torch.from_numpy(fm_filter_params(fm_filter_sample_ids(V))).cuda()
        out = torch.zeros((V, n), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        lens = p.render(out, params=params, out_len=np.zeros(V, dtype=np.uint64))
        assert (np.asarray(lens) == n).all()
        res.append((out, int(p.info.fm_ws_launches), int(p.info.lane_launches), int(p.info.split_fm_rounds), int(p.info.split_segments)))
    same = torch.equal(res[0][0].view(torch.int32), res[1][0].view(torch.int32))
    print(f"split V={V} n={n}: ws launches {res[1][1]}/{res[1][2]} (off {res[0][1]}/{res[0][2]}), split-fm rounds {res[1][3]}, "
          f"segments {res[1][4]}, bit-identical rows: {same}")
    bad += 0 if same else 1
    if not same:
        d = (res[0][0] - res[1][0]).abs()
        rows = torch.nonzero(d.amax(dim=1) > 0).flatten().cpu().numpy()
        prm = fm_filter_params(fm_filter_sample_ids(V))
        print(f"   {len(rows)} rows differ, max diff {float(d.max()):.3e}; first rows {rows[:12]}")
        for r in rows[:6]:
            first = int(torch.nonzero(d[r] > 0).flatten()[0])
            print(f"   row {r}: m={prm[r, 1]:.6g} c={prm[r, 2]:.6g} first differing sample {first}, max {float(d[r].max()):.3e}")
    del res, out
os.environ["TUUN_B200_LANE_MIN_VOICES"] = "1"
print("PARITY", "OK" if bad == 0 else f"FAILED ({bad})")

del os.environ["TUUN_B200_LANE_MIN_VOICES"]
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
V, N = 65536, int(round(secs * 44100))
params = torch.from_numpy(fm_filter_params(np.arange(V))).cuda()
out = torch.empty((V, N), dtype=torch.float32, device="cuda")
for ws in (0, 1, 0, 1):
    os.environ["TUUN_B200_FM_WS"] = str(ws)
    p = Program(w, 44100)
    for _ in range(6):
        p.reset()
        p.render(out, params=params)
    torch.cuda.synchronize()
    ms = p.lane_kernel_times(4)
    print(f"ws={ws}: V={V} N={N} kernel ms {np.round(ms, 3)} -> {V * N / (np.median(ms) * 1e-3):.4e} voice-samples/s "
          f"(ws launches {p.info.fm_ws_launches} of {p.info.lane_launches}; capacity {p.info.lane_fm_ws_capacity} CTAs of 32 voices, "
          f"{p.info.lane_fm_capacity} of 64)")
    del p
