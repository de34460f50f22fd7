"""Re-run one seed of tests/test_gpu_lanes_fuzz.py and show where the kernels differ from the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_gpu_lanes_fuzz import Gen, SR, TAU
from oracle.binding import OracleProgram
from tuun_b200.generator import Program
from tuun_b200.waveform import Const, Fin, Time, f32, sub
seed = int(sys.argv[1])
g = Gen(1000 + seed)
w = g.tree(3)
if seed % 4 == 3:
    w = Fin(sub(Time(), Const(f32(g.r.uniform(0.005, 0.02)))), w)
print(str(w))
V, N = 70, 256 + 16 * 21 + 5 + 16 * 9 + 3
rng = np.random.default_rng(seed)
params = np.stack([TAU * rng.uniform(30, 2500, V), TAU * rng.uniform(0.5, 40, V), rng.uniform(-1, 1, V), rng.uniform(0.1, 2, V)], axis=1).astype(np.float32)
ref = np.zeros((V, N), np.float32)
o = OracleProgram(w, SR)
for v in range(V):
    o.initialize_state(); o.seed_noise(77, v); o.set_params(params[v]); r = o.render(N); ref[v, :len(r)] = r
res = {}
for lanes in ("1", "0"):
    os.environ["TUUN_B200_LANES"] = lanes
    os.environ["TUUN_B200_LANE_MIN_VOICES"] = "1"
    os.environ["TUUN_B200_DEBUG"] = "1" if lanes == "1" else "0"
    p = Program(w, SR); p.seed_noise(77, 0)
    got = np.zeros((V, N), np.float32)
    p.render(got, params=params)
    res[lanes] = got
    d = np.abs(got - ref) / np.maximum(1, np.max(np.abs(ref), axis=1, keepdims=True))
    v, t = np.unravel_index(np.argmax(d), d.shape)
    print("lanes", lanes, "max", d.max(), "voice", v, "t", t, "bad", np.count_nonzero(d > 2e-4), "bad voices", np.unique(np.nonzero(d > 2e-4)[0])[:10])
    print("   got", got[v, t - 2:t + 3], "ref", ref[v, t - 2:t + 3], "params", params[v])
    bt = np.nonzero(d[v] > 2e-4)[0]
    print("   bad t of that voice:", bt[:20])
