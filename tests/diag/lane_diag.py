"""Diagnostics for the lane-per-voice kernel: where do partitioned renders / the oracle differ?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["TUUN_B200_LANE_MIN_VOICES"] = "1"
from oracle.binding import OracleProgram
from tuun_b200.generator import Program
from tuun_b200.workloads import fm_filter_params, fm_filter_voice

V, N = 200, 6000
ids = (np.arange(V) * 40503 + 49230) % 65536
w, params = fm_filter_voice(), fm_filter_params(ids)
one = np.zeros((V, N), np.float32)
Program(w, 44100).render(one, params=params)
p = Program(w, 44100)
parts = np.zeros((V, N), np.float32)
a = 0
for n in (1000, 273, 16, 2500, 2211):
    blk = np.zeros((V, n), np.float32)
    p.render(blk, params=params)
    parts[:, a:a + n] = blk
    a += n
ref, _, _, _ = OracleProgram(w, 44100).render_batch(params, V, N)
os.environ["TUUN_B200_LANES"] = "0"
warp = np.zeros((V, N), np.float32)
Program(w, 44100).render(warp, params=params)
for name, x, y in (("parts-one", parts, one), ("one-oracle", one, ref), ("parts-oracle", parts, ref), ("warp-oracle", warp, ref)):
    d = np.abs(x - y)
    v, t = np.unravel_index(np.argmax(d), d.shape)
    print(name, "max", d.max(), "voice", v, "id", ids[v], "t", t, "params", params[v])
    pv = d[v]
    print("   first t with d>1e-6:", int(np.argmax(pv > 1e-6)), " d at boundaries:", [float(pv[k]) for k in (255, 256, 271, 272, 999, 1000, 1255, 1256, 1273, 1289, 3789, 5999)])
