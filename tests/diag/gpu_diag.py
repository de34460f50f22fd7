"""GPU diagnostics: error of individual building blocks against the CPU oracle."""
import math, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.binding import OracleProgram
from tuun_b200.generator import Program
from tuun_b200.waveform import *
from tuun_b200.waveform import add, mul
from tuun_b200.workloads import fm_filter_params, fm_filter_voice, lpf

SR = 44100
def cmp(name, w, n, params=None):
    p = Program(w, SR)
    out = np.zeros((1, n), np.float32)
    p.render(out, params=None if params is None else params[None, :])
    o = OracleProgram(w, SR)
    if params is not None: o.set_params(params)
    ref = o.render(n)
    d = np.abs(out[0, :len(ref)] - ref)
    nz = np.count_nonzero(out[0, :len(ref)] != ref)
    print(f"{name:40s} n={n} max={d.max():.3e} mean={d.mean():.3e} mismatches={nz} ({nz/len(ref):.2%}) argmax={d.argmax()}", flush=True)
    return out[0], ref

P = fm_filter_params(np.arange(65536))
for v in (300, 3900, 40000, 65535, 4095):
    pr = P[v]
    print("voice", v, pr)
    mod = Sine(Const(float(pr[0])), Const(f32(np.float32(3.14159265) / np.float32(2))))
    cmp(" modulator", mod, 44100)
    fr = add(mul(mod, Const(float(pr[1]))), Const(float(pr[2])))
    cmp(" freq expr", fr, 44100)
    car = Sine(fr, Const(0.0))
    cmp(" carrier(no filter)", car, 44100)
    cmp(" full voice", fm_filter_voice(), 44100, pr)
    cmp(" filter over const-sine", Filter(Sine(Const(float(pr[2])), Const(0.0)), [Const(float(pr[3])), Const(float(pr[4])), Const(float(pr[5]))], [Const(float(pr[6])), Const(float(pr[7]))]), 44100)
