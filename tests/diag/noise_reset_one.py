"""Noise gated by a Fin inside a Reset: device vs oracle (which samples of the node's stream each run draws)."""
import math, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.binding import OracleProgram
from tuun_b200.generator import Generator
from tuun_b200.waveform import Append, Const, Fin, Noise, Reset, Sine, Time, add, f32, mul
SR = 44100
trees = {
    "fin-noise": Reset(Sine(Const(f32(863.0)), Const(f32(2.488))), Append(Fin(add(Time(), Const(f32(-0.0029))), mul(Noise(), Const(0.3))), Const(f32(-0.346)))),
    "fin-only": Reset(Sine(Const(f32(863.0)), Const(f32(2.488))), Fin(add(Time(), Const(f32(-0.0029))), mul(Noise(), Const(0.3)))),
    "plain": Reset(Sine(Const(f32(863.0)), Const(f32(2.488))), mul(Noise(), Const(0.3))),
}
n = 2554
for name, w in trees.items():
    o = OracleProgram(w, SR); o.seed_noise(0x7475756E2545F491, 0)
    ref = o.render(n, block=1024)
    gen = Generator(SR)
    try:
        p = gen.initialize_state(w)
    except Exception as e:
        print(name, "not taken:", e)
        continue
    out = np.full(n, np.inf, np.float32)
    done = gen.generate(p, out)
    m = min(done, len(ref))
    bad = np.nonzero(np.abs(out[:m] - ref[:m]) > 1e-6)[0]
    print(name, "len", done, len(ref), "bad", len(bad), bad[:5])
