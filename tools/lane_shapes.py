"""Throughput of the two kernels on several batch shapes (65,536 voices): the lane-per-voice kernel
(interpreter and fused paths) against the warp-per-voice kernel.  usage: python tools/lane_shapes.py [seconds]"""
import math, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tuun_b200.generator import Program
from tuun_b200.waveform import Alt, Const, Filter, Fin, Noise, Reset, Sine, Time, add, f32, mul, sub
from tuun_b200.workloads import fm_filter_params, fm_filter_voice, lpf

SR, V = 44100, 65536
N = int(float(sys.argv[1]) * SR) if len(sys.argv) > 1 else 2 * SR
N -= N % 4
TAU = f32(2 * math.pi)
rng = np.random.default_rng(0)
fr = (TAU * rng.uniform(50, 2000, (V, 3))).astype(np.float32)
p5 = fm_filter_params(np.arange(V))
fm = Sine(add(mul(Sine(Const(1.0, param=0), Const(f32(math.pi / 2))), Const(1.0, param=1)), Const(1.0, param=2)), Const(0.0))
fo = rng.uniform(30.0, 1800.0, V).astype(np.float32)
osc = np.stack([TAU * fo, -fo, f32(4) * fo, f32(-4) * fo], axis=1).astype(np.float32)
_trig = lambda: Sine(Const(1.0, param=0), Const(0.0))
saw = mul(add(Reset(_trig(), mul(Time(), Const(1.0, param=1))), Const(0.5)), Const(2.0))      # std.tuun:19
tri = Alt(_trig(), Reset(_trig(), add(mul(Time(), Const(1.0, param=2)), Const(-1.0))),
          Reset(_trig(), add(mul(Time(), Const(1.0, param=3)), Const(3.0))))                  # std.tuun:23-28
def _adsr(dur):
    """attack / decay / sustain ramp / release, each a Fin of literal length; longer than the note."""
    from tuun_b200.waveform import Append
    ramp = lambda d, m, a: Fin(sub(Time(), Const(f32(d))), add(mul(Time(), Const(f32(m))), Const(f32(a))))
    return Append(ramp(0.02, 50.0, 0.0), Append(ramp(0.1, -3.0, 1.0), Append(ramp(dur, -0.1 / dur, 0.7), ramp(0.5, -1.2, 0.6))))


shapes = [
    ("sine", Sine(Const(1.0, param=0), Const(0.0)), fr),
    ("pm", Sine(Const(1.0, param=0), mul(Sine(Const(1.0, param=1), Const(0.0)), Const(6.0))), fr),
    ("square|lpf", lpf(Alt(Sine(Const(1.0, param=0), Const(0.0)), Const(1.0), Const(-1.0)), 0.707, 2000), fr),
    ("noise*0.1|lpf", lpf(mul(Noise(), Const(0.1)), 0.7, 2000), None),
    ("fm", fm, p5),
    ("fm|lpf (cfg5)", fm_filter_voice(), p5),
    ("fm|lpf|lpf|lpf", lpf(lpf(fm_filter_voice(), 2.0, 1600), 1.0, 3200), p5),
    ("$f * note(N)", Fin(sub(Time(), Const(f32(N / SR))), Sine(Const(1.0, param=0), Const(0.0))), fr),   # cfg1 shape, swept f
    ("(fm|lpf) * note(N)", Fin(sub(Time(), Const(f32(N / SR))), fm_filter_voice()), p5),
    ("sawtooth(f)", saw, osc),
    ("pulse(0.3,f)|lpf", lpf(Alt(sub(saw, Const(0.3)), Const(1.0), Const(-1.0)), 0.707, 2000), osc),
    ("triangle(f)", tri, osc),
    # hard sync (a Reset nested in a Reset) and a note under an ADSR of literal-length pieces (a timeline): lanes.cuh
    # ST_RESET_CLK with an enclosing clock, ST_SEG_CLK / ST_SEG_SEL
    ("sync(pulse f, saw 4f)", Reset(Alt(sub(saw, Const(0.9)), Const(1.0), Const(-1.0)),
                                    mul(add(Reset(Sine(Const(1.0, param=2), Const(0.0)), mul(Time(), Const(1.0, param=3))), Const(0.125)), Const(2.0))), osc),
    ("(fm|lpf) * adsr note(N)", Fin(sub(Time(), Const(f32(N / SR))), mul(fm_filter_voice(), _adsr(N / SR))), p5),
]
out = torch.empty((V, N), dtype=torch.float32, device="cuda")


def run(w, params, env):
    for k in ("TUUN_B200_LANES", "TUUN_B200_LANE_FUSE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    p = Program(w, SR)
    pd = None if params is None else torch.from_numpy(params).cuda()
    ts = []
    for i in range(4):
        p.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        p.render(out, params=pd)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return V * N / min(ts[1:]), p.info.lane_launches > 0


print(f"{V} voices x {N} samples; voice-samples/s")
for name, w, params in shapes:
    warp, _ = run(w, params, {"TUUN_B200_LANES": "0"})
    lanes, used = run(w, params, {})
    unf, _ = run(w, params, {"TUUN_B200_LANE_FUSE": "0"})
    print(f"{name:16s} warp-per-voice {warp:9.3e}   lane-per-voice {lanes:9.3e} ({'lanes' if used else 'not used'})   unfused {unf:9.3e}   ratio {lanes / warp:4.2f}")
