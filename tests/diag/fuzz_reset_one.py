"""Re-run one seed of test_random_tree_under_reset and show the mismatch (TUUN_B200_SPLIT honoured)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_gpu_fuzz_general import GenR, SR, f32, mul
from oracle.binding import OracleProgram
from tuun_b200.generator import Generator, lower_check
from tuun_b200.waveform import Const, Reset, Sine
seed = int(sys.argv[1])
g = GenR(31000 + seed)
r = g.r
trig = Sine(g.hz(8, 400), Const(f32(r.uniform(0, 6))))
w = Reset(trig, g.tree(3))
if r.random() < 0.3:
    w = mul(w, Const(f32(0.5)))
n = int(g.r.integers(600, 6000))
print(str(w)); print("n", n, "split_passes", lower_check(w).split_passes)
o = OracleProgram(w, SR); o.seed_noise(0x7475756E2545F491, 0)
ref = o.render(n, block=1024)
gen = Generator(SR); p = gen.initialize_state(w)
out = np.full(n, np.inf, np.float32)
done = gen.generate(p, out)
i = p.info
print("len gpu", done, "oracle", len(ref), "split rounds", i.split_rounds, "segments", i.split_segments, i.split_seg_samples)
d = np.abs(out[:done] - ref[:done])
bad = np.nonzero(d > 1e-4 * max(1.0, np.abs(ref).max()))[0]
print("bad", len(bad), "first", bad[:10], "last", bad[-5:])
for t in bad[:3]:
    print(t, "gpu", out[max(0, t - 2):t + 3], "ref", ref[max(0, t - 2):t + 3])
