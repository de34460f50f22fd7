"""Random steady-state trees through both kernels against the oracle: every combination the lane
program can be made of (point operators with waveform and constant operands, all sine forms, Alt,
constant filters, noise, Reset over clock-closed-form trees, a root Fin), with per-voice parameters,
streamed in two calls.  Seeds are fixed; trees whose triggers sit within rounding of zero may move an
edge by one sample (hard part 1), so a handful of differing samples per tree is allowed."""
import math

import numpy as np
import pytest

from oracle.binding import OracleProgram
from tuun_b200.waveform import (Alt, BinaryPointOp, Const, Filter, Fin, Noise, Operator, Reset, Sine, Time, add, f32, mul, sub)

pytestmark = pytest.mark.gpu
SR = 44100
TAU = f32(2 * math.pi)
N_PARAMS = 4


class Gen:
    def __init__(self, seed):
        self.r = np.random.default_rng(seed)

    def const(self, lo=-1.0, hi=1.0):
        if self.r.random() < 0.3:
            return Const(1.0, param=int(self.r.integers(0, N_PARAMS)))
        return Const(f32(self.r.uniform(lo, hi)))

    def rate(self):  # rad/s
        if self.r.random() < 0.4:
            return Const(1.0, param=int(self.r.integers(0, 2)))  # params 0, 1 are rates
        return Const(f32(TAU * self.r.uniform(20, 3000)))

    def clocked(self, depth):  # closed-form in a Reset's clock
        k = self.r.integers(0, 5 if depth > 0 else 3)
        if k == 0:
            return add(mul(Time(), Const(f32(self.r.uniform(-300, 300)))), Const(f32(self.r.uniform(-1, 1))))
        if k == 1:
            return Sine(self.rate(), Const(f32(self.r.uniform(0, 6))))
        if k == 2:
            return Const(f32(self.r.uniform(-1, 1)))
        if k == 3:
            return BinaryPointOp(Operator(int(self.r.integers(0, 3))), self.clocked(depth - 1), self.clocked(depth - 1))
        return Alt(self.clocked(depth - 1), self.clocked(depth - 1), Const(f32(self.r.uniform(-1, 1))))

    def tree(self, depth):
        k = self.r.integers(0, 10 if depth > 0 else 3)
        if k == 0:
            return Sine(self.rate(), Const(f32(self.r.uniform(0, 6))))
        if k == 1:
            return mul(Noise(), Const(f32(self.r.uniform(0.05, 0.5))))
        if k == 2:
            return add(mul(Time(), Const(f32(self.r.uniform(-3, 3)))), Const(f32(self.r.uniform(-1, 1))))
        if k == 3:  # point operator, both waveforms
            op = Operator(int(self.r.choice([0, 1, 2])))
            return BinaryPointOp(op, self.tree(depth - 1), self.tree(depth - 1))
        if k == 4:  # constant right-hand sides, incl. divide
            x = self.tree(depth - 1)
            return BinaryPointOp(Operator.Divide, mul(x, self.const()), Const(f32(self.r.uniform(0.5, 4))))
        if k == 5:  # FM / PM
            m = add(mul(self.tree(depth - 1), Const(f32(self.r.uniform(50, 4000)))), Const(f32(TAU * self.r.uniform(100, 2000))))
            if self.r.random() < 0.5:
                return Sine(m, Const(f32(self.r.uniform(0, 3))))
            return Sine(self.rate(), mul(self.tree(depth - 1), Const(f32(self.r.uniform(0.5, 5)))))
        if k == 6:
            return Alt(self.tree(depth - 1), self.tree(depth - 1) if self.r.random() < 0.5 else Const(1.0),
                       self.tree(depth - 1) if self.r.random() < 0.5 else Const(-1.0))
        if k == 7:  # stable filters: biquad, one-pole, FIR
            x = self.tree(depth - 1)
            kind = self.r.integers(0, 3)
            if kind == 0:
                rad, th = self.r.uniform(0.3, 0.97), self.r.uniform(0.05, 2.5)
                return Filter(x, [Const(f32(0.2)), Const(f32(0.3)), Const(f32(0.2))],
                              [Const(f32(-2 * rad * math.cos(th))), Const(f32(rad * rad))])
            if kind == 1:
                return Filter(x, [Const(f32(0.5)), Const(f32(0.25))], [Const(f32(self.r.uniform(-0.9, 0.9)))])
            return Filter(x, [Const(f32(c)) for c in self.r.uniform(-0.4, 0.4, int(self.r.integers(1, 6)))], [])
        if k == 8:
            return Reset(Sine(self.rate(), Const(0.0)), self.clocked(2))
        return add(mul(self.tree(depth - 1), Const(f32(self.r.uniform(-2, 2)))), self.const())


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("TUUN_FUZZ_SEEDS", "36"))))
def test_random_steady_tree(monkeypatch, seed):
    from tuun_b200.generator import Program, lower_check
    g = Gen(1000 + seed)
    w = g.tree(3)
    if seed % 4 == 3:
        w = Fin(sub(Time(), Const(f32(g.r.uniform(0.005, 0.02)))), w)
    if lower_check(w).lane_smem_bytes == 0:
        pytest.skip("not a lane program")
    V, N1, N2 = 70, 256 + 16 * 21 + 5, 16 * 9 + 3
    rng = np.random.default_rng(seed)
    params = np.stack([TAU * rng.uniform(30, 2500, V), TAU * rng.uniform(0.5, 40, V), rng.uniform(-1, 1, V),
                       rng.uniform(0.1, 2, V)], axis=1).astype(np.float32)
    ref = np.zeros((V, N1 + N2), dtype=np.float32)
    rlen = np.zeros(V, dtype=np.int64)
    o = OracleProgram(w, SR)
    for v in range(V):
        o.initialize_state()
        o.seed_noise(77, v)
        o.set_params(params[v])
        r = o.render(N1 + N2)
        ref[v, :len(r)] = r
        rlen[v] = len(r)
    for lanes in (True, False):
        monkeypatch.setenv("TUUN_B200_LANES", "1" if lanes else "0")
        monkeypatch.setenv("TUUN_B200_LANE_MIN_VOICES", "1")
        p = Program(w, SR)
        p.seed_noise(77, 0)
        got = np.zeros((V, N1 + N2), dtype=np.float32)
        a = np.zeros((V, N1), dtype=np.float32)
        l1 = p.render(a, params=params)
        b = np.zeros((V, N2), dtype=np.float32)
        l2 = p.render(b, params=params)
        got[:, :N1] = a
        got[:, N1:] = b
        assert ((l1 + l2).astype(np.int64) == rlen).all(), (seed, lanes)
        if lanes:
            assert p.info.lane_launches >= 1
        mask = np.arange(N1 + N2)[None, :] < rlen[:, None]
        d = np.abs(got - ref) * mask
        scale = np.maximum(1.0, np.max(np.abs(ref), axis=1, keepdims=True))  # tolerance relative to the voice's peak
        bad = int(np.count_nonzero(d / scale > 2e-4))
        assert bad <= 6, (seed, lanes, bad, float(d.max()), str(w)[:200])


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("TUUN_FUZZ_SEEDS", "36"))))
def test_random_steady_tree_split_in_time(monkeypatch, seed):
    """The same random steady trees with every voice cut into segments (TUUN_B200_SPLIT forced: split.cu): phase
    prefix sums, affine scans of filter histories, Reset clocks — against the oracle, through both kernel families
    (a handful of voices: warp kernel; TUUN_B200_LANE_MIN_VOICES=1: lane kernels, 2^k segments), in two calls."""
    from tuun_b200.generator import Program, lower_check
    g = Gen(1000 + seed)
    w = g.tree(3)
    info = lower_check(w)
    if info.split_passes == 0:
        pytest.skip("not a steady program")
    V, N1, N2 = 5, 256 + 512 * 9 + 37, 512 * 5 + 100
    rng = np.random.default_rng(seed)
    params = np.stack([TAU * rng.uniform(30, 2500, V), TAU * rng.uniform(0.5, 40, V), rng.uniform(-1, 1, V),
                       rng.uniform(0.1, 2, V)], axis=1).astype(np.float32)
    ref = np.zeros((V, N1 + N2), dtype=np.float32)
    o = OracleProgram(w, SR)
    for v in range(V):
        o.initialize_state()
        o.seed_noise(77, v)
        o.set_params(params[v])
        ref[v] = o.render(N1 + N2)
    scale = np.maximum(1.0, np.max(np.abs(ref), axis=1, keepdims=True))
    for lanes, segs in ((False, 7), (True, 4)):
        if lanes and info.lane_smem_bytes == 0:
            continue
        monkeypatch.setenv("TUUN_B200_LANES", "1" if lanes else "0")
        monkeypatch.setenv("TUUN_B200_LANE_MIN_VOICES", "1")
        monkeypatch.setenv("TUUN_B200_SPLIT", str(segs))
        p = Program(w, SR)
        p.seed_noise(77, 0)
        a = np.zeros((V, N1), dtype=np.float32)
        b = np.zeros((V, N2), dtype=np.float32)
        l1 = p.render(a, params=params)
        l2 = p.render(b, params=params)
        assert (l1 == N1).all() and (l2 == N2).all()
        if not lanes and info.tile == 256:
            assert p.info.split_rounds == 0      # clocked words (a Reset): split only on the lane kernels
            continue
        assert p.info.split_rounds >= 2, (seed, lanes)
        d = np.abs(np.concatenate([a, b], axis=1) - ref)
        bad = int(np.count_nonzero(d / scale > 2e-4))
        assert bad <= 6, (seed, lanes, bad, float(d.max()), str(w)[:200])
