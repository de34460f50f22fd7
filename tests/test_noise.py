"""Noise (generator.rs:113-118).  The reference draws `fastrand::f32() * 2 - 1` from an unseeded
thread-local generator, so no reference value can be pinned ("parity unpinned"); what IS pinned:
the oracle's restatement of fastrand 2.3.0's published generator (wyrand) as per-node, per-voice
streams, its range and moments, and — on the GPU — bit-exact agreement with those streams."""
import numpy as np
import pytest

from oracle.binding import OracleProgram
from tuun_b200.waveform import Alt, BinaryPointOp, Const, Filter, Fin, Noise, Operator, Reset, Sine, Time, add, mul, sub

SR = 44100


def wyrand_f32(state: int) -> float:
    """fastrand 2.3.0: Rng::gen_u64 (wyrand) + Rng::f32, written out independently in Python ints."""
    m = (1 << 64) - 1
    t = (state & m) * ((state ^ 0x8bb84b93962eacc9) & m)
    r = (t & m) ^ (t >> 64)
    bits = 0x3F800000 | ((r & 0xffffffff) >> 9)
    return float(np.array([bits], dtype=np.uint32).view(np.float32)[0] - np.float32(1.0))


def test_oracle_streams_are_wyrand():
    seed, voice, node = 0x1234, 3, 0
    o = OracleProgram(Noise(), SR)
    o.seed_noise(seed, voice)
    got = o.render(64)
    m = (1 << 64) - 1
    base = (seed + 0x9e3779b97f4a7c15 * (node + 1) + 0xd6e8feb86659fd93 * voice) & m
    want = [np.float32(np.float32(wyrand_f32((base + 0x2d358dccaa6c78a5 * (k + 1)) & m)) * np.float32(2) - np.float32(1))
            for k in range(64)]
    np.testing.assert_array_equal(got, np.asarray(want, dtype=np.float32))


def test_oracle_noise_range_and_moments():
    o = OracleProgram(Noise(), SR)
    x = o.render(200000, block=1024)
    assert x.min() >= -1.0 and x.max() < 1.0          # wasm.rs:458-466 pins only the range
    assert abs(float(x.mean())) < 0.01 and abs(float(x.var()) - 1.0 / 3.0) < 0.01
    assert abs(float(np.corrcoef(x[:-1], x[1:])[0, 1])) < 0.01
    # block size does not matter; a fresh tree replays; another voice / node draws another stream
    np.testing.assert_array_equal(OracleProgram(Noise(), SR).render(5000, block=7), x[:5000])
    o.initialize_state()
    np.testing.assert_array_equal(o.render(100), x[:100])
    o2 = OracleProgram(Noise(), SR)
    o2.seed_noise(0x7475756E2545F491, 1)
    assert not np.array_equal(o2.render(100), x[:100])
    shifted = OracleProgram(BinaryPointOp(Operator.Add, Const(0.0), Noise()), SR).render(100)
    assert not np.array_equal(shifted, x[:100])  # the Noise is node 1 of this op list, not node 0


@pytest.mark.gpu
def test_gpu_noise_bit_exact_with_oracle_streams():
    from tuun_b200.generator import Program
    for w in (Noise(), mul(Noise(), Const(0.1)),
              Fin(sub(Time(), Const(0.25)), Noise()),
              Alt(Sine(Const(2764.6016), Const(0.0)), Noise(), mul(Noise(), Const(-0.5))),
              Reset(Sine(Const(2764.6016), Const(0.0)), add(Noise(), Time()))):
        p = Program(w, SR)
        p.seed_noise(99, 5)
        out = np.zeros((3, 20000), dtype=np.float32)
        lens = p.render(out)
        for v in range(3):
            o = OracleProgram(w, SR)
            o.seed_noise(99, 5 + v)
            ref = o.render(20000, block=1024)
            assert lens[v] == len(ref)
            np.testing.assert_array_equal(out[v, :len(ref)], ref, err_msg=str(w))


@pytest.mark.gpu
def test_gpu_filtered_noise_and_streaming():
    from tuun_b200.generator import Generator
    from tuun_b200.workloads import lpf
    w = lpf(mul(Noise(), Const(0.1)), 0.7, 2000)  # `noise*0.1 | lpf(0.7, 2000)` (config 4)
    ref = OracleProgram(w, SR).render(3 * SR, block=1024)
    g = Generator(SR)
    for block in (None, 1024, 777):
        p = g.initialize_state(w)
        out = np.zeros(3 * SR, dtype=np.float32)
        if block is None:
            assert g.generate(p, out) == len(out)
        else:
            for a in range(0, len(out), block):
                b = min(len(out), a + block)
                assert g.generate(p, out[a:b]) == b - a
        assert np.max(np.abs(out - ref)) <= 1e-4
