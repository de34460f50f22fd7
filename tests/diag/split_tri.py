import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle.binding import OracleProgram
from tuun_b200.builder import Std, to_waveform
from tuun_b200.generator import Program
from tuun_b200.optimizer import optimize
s = Std()
w = optimize(to_waveform(s.triangle(55)))
n = 256 + 64 * 512 + 77
ref = OracleProgram(w, 44100).render(n)
for env in ({"TUUN_B200_SPLIT": "0"}, {"TUUN_B200_SPLIT": "4"}, {"TUUN_B200_SPLIT": "0", "TUUN_B200_LANE_MIN_VOICES": "1"}):
    os.environ.pop("TUUN_B200_LANE_MIN_VOICES", None)
    os.environ.update(env)
    out = np.zeros((1, n), dtype=np.float32)
    p = Program(w, 44100)
    p.render(out)
    d = np.abs(out[0] - ref)
    bad = np.nonzero(d > 1e-4)[0]
    print(env, "bad", len(bad), bad[:5], bad[-5:] if len(bad) else "", "launches", p.info.kernel_launches, p.info.lane_launches)
    print("  got", out[0, :6], out[0, 398:404])
print("  ref", ref[:6], ref[398:404])
