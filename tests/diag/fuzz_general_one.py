"""Re-run one seed of tests/test_gpu_fuzz_general.py (optionally another depth) and show the mismatch."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_gpu_fuzz_general import Gen, SR
from oracle.binding import OracleProgram
from tuun_b200.generator import Generator
seed, depth = int(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = Gen(9000 + seed)
w = g.tree(depth)
n = int(g.r.integers(600, 6000))
print(str(w)[:1500]); print("n", n)
o = OracleProgram(w, SR); o.seed_noise(0x7475756E2545F491, 0)
ref = o.render(n, block=1024)
gen = Generator(SR); p = gen.initialize_state(w)
out = np.full(n, np.inf, np.float32)
blk = int(sys.argv[3]) if len(sys.argv) > 3 else n
done = 0
while done < n:
    want = min(blk, n - done)
    got = gen.generate(p, out[done:done + want])
    done += got
    if got < want:
        break
print("len gpu", done, "oracle", len(ref))
m = min(done, len(ref))
d = np.abs(out[:m] - ref[:m])
bad = np.nonzero(d > 1e-4 * max(1.0, np.abs(ref).max()))[0]
print("bad", len(bad), "first", bad[:10], "last", bad[-5:])
for t in bad[:3]:
    print(t, "gpu", out[max(0, t - 2):t + 3], "ref", ref[max(0, t - 2):t + 3])
for blk in (1024, 256, 64):
    o2 = OracleProgram(w, SR); o2.seed_noise(0x7475756E2545F491, 0)
    r2 = o2.render(n, block=blk)
    print("oracle block", blk, "len", len(r2), "max diff vs block 1024", float(np.max(np.abs(r2[:min(len(r2), len(ref))] - ref[:min(len(r2), len(ref))]), initial=0)))
