"""Re-run one seed of test_random_tree_batch (tests/test_gpu_fuzz_general.py) and show where a voice differs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_gpu_fuzz_general import GenP, SR, TAU
from oracle.binding import OracleProgram
from tuun_b200.generator import Program
seed, depth = int(sys.argv[1]), int(sys.argv[2])
g = GenP(17000 + seed)
w = g.tree(depth)
V = 11
n1, n2 = int(g.r.integers(300, 3000)), int(g.r.integers(1, 2500))
n = n1 + n2
params = np.stack([TAU * g.r.uniform(30, 2500, V), TAU * g.r.uniform(2, 60, V),
                   -g.r.choice([0.0, 0.0007, 0.003, 0.011, 0.03, 0.09], V) * g.r.uniform(0.5, 1.0, V),
                   g.r.uniform(-1, 1, V)], axis=1).astype(np.float32)
if len(sys.argv) > 3:  # "trigger": render only the Alt's trigger; "short": the same with the FIR cut to 9 taps
    from tuun_b200.waveform import Filter
    w = w.trigger
    if sys.argv[3] == "short":
        w = Filter(w.waveform, w.feed_forward[:9], w.feedback)
    if sys.argv[3] == "input":
        w = w.waveform
print(str(w)[:600]); print("n1", n1, "n2", n2, "params v5", params[5])
p = Program(w, SR); p.seed_noise(0x7475756E2545F491, 0)
a = np.zeros((V, n1), np.float32); b = np.zeros((V, n2), np.float32)
l1 = np.asarray(p.render(a, params=params)).astype(np.int64); l2 = np.asarray(p.render(b, params=params)).astype(np.int64)
for v in range(V):
    o = OracleProgram(w, SR); o.seed_noise(0x7475756E2545F491, v); o.set_params(params[v])
    r = o.render(n, block=1024)
    got = np.concatenate([a[v, :l1[v]], b[v, :l2[v]] if l1[v] == n1 else np.zeros(0, np.float32)])
    m = min(len(got), len(r))
    d = np.abs(got[:m] - r[:m]); bad = np.nonzero(d > 1e-4 * max(1.0, np.abs(r).max(initial=0)))[0]
    print("voice", v, "len", len(got), len(r), "bad", len(bad), bad[:8], bad[-3:] if len(bad) else "")
    if len(bad) > 6:
        t = bad[0]
        print("  got", got[t - 2:t + 6]); print("  ref", r[t - 2:t + 6])
        # runs of bad samples
        runs = np.split(bad, np.nonzero(np.diff(bad) > 1)[0] + 1)
        print("  runs", [(int(x[0]), len(x)) for x in runs][:12])
if len(sys.argv) > 4:  # show the values of voice V around sample T
    v, t = int(sys.argv[4]), int(sys.argv[5])
    o = OracleProgram(w, SR); o.seed_noise(0x7475756E2545F491, v); o.set_params(params[v])
    r = o.render(n, block=1024)
    got = np.concatenate([a[v], b[v]])
    np.set_printoptions(precision=4, linewidth=200)
    print("got", got[t - 6:t + 8]); print("ref", r[t - 6:t + 8])
