"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/tuun_b200.h
declares, validates op lists before touching CUDA, and fails loudly (TB_ERR_CUDA) instead of
falling back when no device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

from tuun_b200 import _abi
from tuun_b200.waveform import (BinaryPointOp, Const, Filter, Fixed, Noise, Operator, Reset, Sine,
                                TbNode, Time, flatten)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    header = open(os.path.join(ROOT, "include", "tuun_b200.h")).read()
    declared = set(re.findall(r"\b(tb_[a-z_]+)\s*\(", header))
    assert declared == set(_abi.EXPORTS)
    L = _abi.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.tb_abi_version() == 3


def test_node_layout():
    assert ctypes.sizeof(TbNode) == 64
    assert TbNode.fixed_off.offset == 48 and TbNode.value.offset == 20


def _create(ops, sample_rate=44100):
    h = ctypes.c_void_p()
    L = _abi.lib()
    rc = L.tb_program_create(ops.nodes, ops.n_nodes,
                             ops.lists.ctypes.data_as(ctypes.c_void_p) if ops.lists.size else None,
                             len(ops.lists),
                             ops.fixed_pool.ctypes.data_as(ctypes.c_void_p) if ops.fixed_pool.size else None,
                             len(ops.fixed_pool), sample_rate, -1, ctypes.byref(h))
    return rc, h, L.tb_last_error().decode()


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_invalid_op_list_rejected_before_cuda():
    ops = flatten(BinaryPointOp(Operator.Add, Time(), Const(1.0)))
    ops.nodes[2].a = 7  # child index out of range
    rc, h, msg = _create(ops)
    assert rc == _abi.TB_ERR_INVALID and "malformed" in msg
    ops = flatten(Time())
    rc, h, msg = _create(ops, sample_rate=0)
    assert rc == _abi.TB_ERR_INVALID


def test_unsupported_is_reported_not_emulated():
    from tuun_b200.generator import lower_check
    rc, h, msg = _create(flatten(Filter(Time(), [Const(1.0)] * 34, [])))
    assert rc == _abi.TB_ERR_UNSUPPORTED and "feed-forward taps" in msg
    rc, h, msg = _create(flatten(Filter(Time(), [Const(1.0)], [Const(0.1)] * 9)))
    assert rc == _abi.TB_ERR_UNSUPPORTED and "feedback taps" in msg
    assert lower_check(Filter(Time(), [Const(1.0)], [Const(0.1)] * 8)).lane_smem_bytes == 0
    rc, h, msg = _create(flatten(Filter(Time(), [Const(1.0)] * 11 + [Time()], [])))
    assert rc == _abi.TB_ERR_UNSUPPORTED and "waveforms" in msg
    long_fir = lower_check(Filter(Time(), [Const(1.0)] * 33, [Const(0.5)]))  # moving_average(32) and a pole
    assert long_fir.n_code_words > 0 and long_fir.lane_smem_bytes == 0       # general interpreter only


def test_reset_over_any_tree_lowers():
    """What the all-runs-at-once form of a Reset does not take (a Filter, an Append whose first part has no
    analytic length, a Fin with a rendered length, a Noise a run draws only part of) lowers to the run-by-run
    form (generator.rs:288-316) instead of being refused."""
    from tuun_b200.generator import lower_check
    from tuun_b200.waveform import Append, Fin, Fixed, add, mul
    trig = Sine(Const(1.0), Const(0.0))
    burst = Fin(add(Time(), Const(-0.003)), mul(Noise(), Const(0.3)))
    seg = lower_check(Reset(trig, mul(Noise(), Fin(add(Time(), Const(-0.003)), Const(1.0))))).n_code_words
    for inner in (Filter(Noise(), [Const(0.5), Const(0.5)], []),
                  Filter(Time(), [Const(1.0)], []),
                  Append(Fixed([1.0, 2.0]), Const(0.0)),
                  Append(Fin(Sine(Const(3.0), Const(0.0)), Const(1.0)), Const(0.0)),
                  Fin(Sine(Const(3.0), Const(0.0)), Time()),
                  burst,
                  mul(Fin(add(Time(), Const(-0.003)), Const(1.0)), Noise()),
                  Reset(Sine(Const(9.0), Const(0.0)), burst)):
        info = lower_check(Reset(trig, inner))
        assert info.n_code_words > 0 and info.state_words >= 1
    assert seg > 0


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_no_device_fails_loudly():
    rc, h, msg = _create(flatten(Sine(Const(2764.6), Const(0.0))))
    assert rc == _abi.TB_ERR_CUDA and "CUDA" in msg
    assert not h.value


def test_lane_plan_of_the_headline_voice():
    """Host half only: the FM + low-pass voice of config 5 qualifies for the lane-per-voice kernel
    (csrc/lanes.cuh) and its lane program fuses into one LN_FM word."""
    from tuun_b200.generator import lower_check
    from tuun_b200.workloads import cfg1_sine, fm_filter_voice
    info = lower_check(fm_filter_voice())
    assert info.tile == 512 and info.lane_smem_bytes > 0
    # code (LN_FM two words + END) + 9 rotation units + the 8-chunk accumulator buffer + W words, 64 voices per CTA
    assert info.lane_smem_bytes == 3 * 16 + 9 * 64 * 16 + 8 * 65 * 16 + 37 * 64 * 4
    # a note of fixed duration (root Fin with an analytic length over a steady tree) qualifies too, though
    # not for the warp-per-voice steady interpreter (tile 256); a rendered length (Fin over a Sine) does not
    fin = lower_check(cfg1_sine())
    assert fin.lane_smem_bytes > 0 and fin.tile == 256
    from tuun_b200.waveform import Const, Fin, Sine
    assert lower_check(Fin(Sine(Const(1.0), Const(0.0)), Sine(Const(440.0), Const(0.0)))).lane_smem_bytes == 0


def test_lane_plan_of_instruments():
    """Host half only: a Reset nested in a Reset (hard sync) and a timeline of literal-length pieces under a root Fin
    (an ADSR over a note) lower into the lane program (lower.cpp emit_steady / emit_timeline); so does a note of config
    2's harmonica, with one rotation table for each of its two frequencies.  A timeline that ends inside its note, a
    piece of per-voice length and a Filter under a piece's clock keep the general interpreter."""
    from tuun_b200.generator import lower_check
    from tuun_b200.waveform import Alt, Append, Const, Filter, Fin, Reset, Sine, Time, add, mul
    from tuun_b200.workloads import cfg2_harmonica
    saw = lambda f: mul(add(Reset(Sine(Const(6.2831855 * f), Const(0.0)), mul(Time(), Const(-f))), Const(0.5)), Const(2.0))
    pulse = lambda f, w: Alt(add(saw(f), Const(w)), Const(1.0), Const(-1.0))
    sync = Reset(pulse(440.0, -0.93), pulse(701.0, 0.3))
    assert lower_check(sync).lane_smem_bytes > 0 and lower_check(sync).split_passes == 0  # no time-axis split
    assert lower_check(pulse(440.0, -0.93)).split_passes > 0                              # (one clock level: split)
    def adsr(c0, c1, c2=None, sustain=None):
        last = Const(sustain) if sustain is not None else Fin(add(Time(), Const(-0.3)), add(mul(Time(), Const(-1.0)), Const(0.8)))
        return Append(Fin(add(Time(), c0), mul(Time(), Const(20.0))),
                      Append(Fin(add(Time(), c1), add(mul(Time(), Const(-1.0)), Const(1.0))), last))
    tone = lambda: Sine(Const(1.0, param=0), Const(0.0))
    note = lambda dur, env: Fin(add(Time(), Const(-dur)), mul(tone(), env))
    assert lower_check(note(0.4, adsr(Const(-0.05), Const(-0.2)))).lane_smem_bytes > 0
    assert lower_check(note(5.0, adsr(Const(-0.05), Const(-0.2), sustain=0.7))).lane_smem_bytes > 0   # held level: never ends
    assert lower_check(note(1.5, adsr(Const(-0.05), Const(-0.2)))).lane_smem_bytes == 0               # ends at 0.55 s
    assert lower_check(note(0.4, adsr(Const(-0.05, param=1), Const(-0.2)))).lane_smem_bytes == 0      # per-voice length
    lp = Filter(mul(Time(), Const(20.0)), [Const(0.5), Const(0.5)], [])
    filtered = Append(Fin(add(Time(), Const(-0.05)), lp), Const(1.0))
    assert lower_check(note(0.4, filtered)).lane_smem_bytes == 0
    from tuun_b200.waveform import Noise
    breath = Append(Fin(add(Time(), Const(-0.05)), mul(Noise(), Const(0.1))), Const(1.0))  # a draw count is no clock
    assert lower_check(note(0.4, breath)).lane_smem_bytes == 0
    h = lower_check(cfg2_harmonica(2).a)
    # 18 Q units: two rotation tables of nine for four constant-rate sines (440 Hz twice, the 1.6 Hz vibrato twice)
    assert h.lane_smem_bytes > 0 and h.lane_smem_bytes < 76 * 1024 and h.tile == 256


def test_nesting_is_checked_against_the_kernel_control_stack():
    """A sequence is a right-nested Append chain (optimizer.rs:212-229): two control-stack words per note when it has to
    be ONE program — per-voice note lengths here (a root sequence of voice-independent lengths lowers part by part and
    has no such limit: lower.h sequence_parts)."""
    from tuun_b200.generator import lower_check
    from tuun_b200.waveform import Append, Fin, add
    def chain(notes, per_voice):
        w = Const(0.0)
        for k in range(notes):
            length = Const(-0.01, param=0) if per_voice else Const(-0.01)
            w = Append(Fin(add(Time(), length), Sine(Const(2000.0 + k), Const(0.0))), w)
        return w
    assert lower_check(chain(100, True)).n_code_words > 0
    with pytest.raises(_abi.TuunB200Error) as e:
        lower_check(chain(400, True))
    assert e.value.status == _abi.TB_ERR_UNSUPPORTED and "nested too deeply" in e.value.message
    info = lower_check(chain(400, False))
    assert info.sequence_parts == 401 and info.n_code_words > 0


def test_op_list_must_be_a_tree_and_fixed_ranges_must_not_wrap():
    """tb_lower_check (no device): a node with two parents, an unreachable node, a Fixed range that wraps u64."""
    import ctypes
    import numpy as np
    from tuun_b200 import _abi
    from tuun_b200.waveform import BinaryPointOp, Const, Fixed, Operator, Sine, flatten
    L = _abi.lib()
    info = _abi.TbProgramInfo()

    def check(ops, fixed_len=None):
        lists = np.asarray(ops.lists, dtype=np.int32)
        lp = lists.ctypes.data_as(ctypes.c_void_p) if lists.size else None
        return L.tb_lower_check(ops.nodes, ops.n_nodes, lp, len(ops.lists),
                                len(ops.fixed_pool) if fixed_len is None else fixed_len, ctypes.byref(info))

    ops = flatten(BinaryPointOp(Operator.Add, Sine(Const(440.0), Const(0.0)), Const(1.0)))
    assert check(ops) == _abi.TB_OK
    shared = flatten(BinaryPointOp(Operator.Add, Sine(Const(440.0), Const(0.0)), Const(1.0)))
    shared.nodes[shared.n_nodes - 1].b = shared.nodes[shared.n_nodes - 1].a      # the sine twice, the constant orphaned
    assert check(shared) == _abi.TB_ERR_INVALID
    assert b"parent" in L.tb_last_error() or b"reachable" in L.tb_last_error()
    fx = flatten(Fixed([1.0, 2.0, 3.0]))
    assert check(fx) == _abi.TB_OK
    fx.nodes[0].fixed_off = 2 ** 64 - 2                                            # off + len wraps to 1 <= 3
    assert check(fx) == _abi.TB_ERR_INVALID


def test_lowering_plans_without_a_device():
    """tb_lower_check reports what a render would do with a tree: time-axis split passes (0: not a steady program) and
    the parts of a root sequence."""
    from tuun_b200 import workloads as W
    from tuun_b200.generator import lower_check
    from tuun_b200.builder import Std, to_waveform
    from tuun_b200.optimizer import optimize
    assert lower_check(W.fm_filter_voice()).split_passes == 3          # modulator analytic, carrier sums, filter, samples
    assert lower_check(W.fm_pair_voice()).split_passes == 2
    assert lower_check(W.cfg1_from_source()).split_passes == 0         # a note that ends: not steady
    progs = dict(W.cfg4_filters())
    assert lower_check(progs["square220-lpf"]).split_passes == 2
    assert lower_check(progs["square-cascade"]).split_passes == 4      # one more pass per filter in the chain
    assert lower_check(progs["pulse-filter_4_3"]).split_passes == 3    # Reset clocks, then the filter
    assert lower_check(progs["noise-lpf"]).split_passes == 0           # a Fixed buffer can end
    s = Std()
    assert lower_check(optimize(to_waveform(s.sawtooth(220)))).split_passes == 2
    seq = lower_check(W.cfg2_harmonica(4))
    assert seq.sequence_parts == 4 and seq.split_passes == 0
    marks = [w for name, w, _ in W.tracker_benches() if name == "marks_4_40"][0]
    assert lower_check(marks).sequence_parts == 160                    # left-nested Appends of Marked sequences
    assert lower_check(W.fm_filter_voice()).sequence_parts == 0
