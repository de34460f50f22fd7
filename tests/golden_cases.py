"""Known-answer vectors of the reference's own generator tests, restated over the Python
Waveform mirror.  Every case cites the reference test it comes from
(/root/reference/src/lib/generator.rs:1353-1925).  All run at sample_rate = 1 through
`run_tests` (generator.rs:1284-1351): chunk sizes 1, 2, 4, 8 must reproduce `expected` exactly.
"""
import math

import numpy as np

from tuun_b200.waveform import (Alt, Append, BinaryPointOp, Const, Filter, Fin, Fixed, Marked,
                                Operator, Reset, Sine, Time, f32)

F32_TAU = f32(2 * math.pi)  # f32::consts::TAU
F32_PI = f32(math.pi)


def sin_waveform(frequency, phase):  # generator.rs:1467-1477
    return Sine(BinaryPointOp(Operator.Multiply, Const(F32_TAU), Const(frequency)), Const(phase))


def time_minus(c):
    return BinaryPointOp(Operator.Subtract, Time(), Const(c))


def cases():
    out = []

    def case(name, w, expected):
        out.append((name, w, np.asarray(expected, dtype=np.float32)))

    # test_time :1354
    case("time", Time(), [0, 1, 2, 3, 4, 5, 6, 7])
    # test_fixed :1360
    case("fixed", Fixed([1, 2, 3, 4, 5]), [1, 2, 3, 4, 5])
    # test_fin :1375-1396 (the Marked makes the length dynamic -> rendered "Maybe" path)
    case("fin_marked",
         BinaryPointOp(Operator.Multiply, Const(2.0),
                       Append(Fin(BinaryPointOp(Operator.Subtract, Time(), Marked(1, Const(4.0))),
                                  Const(1.0)),
                              Fixed([1.0, 0.75, 0.5, 0.25]))),
         [2, 2, 2, 2, 2, 1.5, 1, 0.5])
    # test_reset :1543-1599
    case("reset_time", Reset(sin_waveform(0.25, 0.0), Time()), [0, 1, 2, 3, 0, 1, 2, 3])
    case("reset_fin_trigger", Reset(Fin(time_minus(6.0), sin_waveform(0.25, 0.0)), Time()),
         [0, 1, 2, 3, 0, 1])
    case("reset_fin_inner", Reset(sin_waveform(0.25, 0.0), Fin(time_minus(3.0), Time())),
         [0, 1, 2, 0, 0, 1, 2, 0])
    case("reset_phase_pi", Reset(sin_waveform(0.25, F32_PI), Time()), [0, 1, 0, 1, 2, 3, 0, 1])
    case("reset_16", Reset(sin_waveform(0.25, 0.0), Time()), [0, 1, 2, 3] * 4)
    # test_append :1602-1613
    case("append", Append(Fixed([1.0] * 3), Fixed([2.0] * 3)), [1, 1, 1, 2, 2, 2])
    # test_sum :1624-1675
    case("add_consts", BinaryPointOp(Operator.Add, Const(1.0), Const(2.0)), [3.0] * 8)
    case("add_fixed_const", BinaryPointOp(Operator.Add, Fixed([1, 2, 3]), Const(10.0)), [11, 12, 13])
    case("add_short_long", BinaryPointOp(Operator.Add, Fixed([1, 2]), Fixed([10, 20, 30])), [11, 22])
    case("add_long_short", BinaryPointOp(Operator.Add, Fixed([1, 2, 3]), Fixed([10, 20])), [11, 22])
    case("add_fin", Fin(time_minus(4.0), BinaryPointOp(Operator.Add, Const(1.0), Const(2.0))), [3.0] * 4)
    case("add_empty", BinaryPointOp(Operator.Add, Fixed([]), Const(5.0)), [])
    # test_dot_product :1678-1736
    case("mul_fin", Fin(time_minus(8.0), BinaryPointOp(Operator.Multiply, Const(3.0), Const(2.0))), [6.0] * 8)
    case("mul_fixed_const", BinaryPointOp(Operator.Multiply, Fixed([3, 4, 5]), Const(2.0)), [6, 8, 10])
    case("mul_short_long", BinaryPointOp(Operator.Multiply, Fixed([3, 4]), Fixed([2, 5, 1])), [6, 20])
    case("mul_empty", BinaryPointOp(Operator.Multiply, Fixed([]), Const(5.0)), [])
    # test_merge :1739-1777
    case("merge_consts", BinaryPointOp(Operator.Merge, Const(1.0), Const(2.0)), [3.0] * 8)
    case("merge_short_long", BinaryPointOp(Operator.Merge, Fixed([1, 2]), Fixed([10, 20, 30])), [11, 22, 30])
    case("merge_fixed_const", BinaryPointOp(Operator.Merge, Fixed([1, 2]), Const(10.0)),
         [11, 12, 10, 10, 10, 10, 10, 10])
    case("merge_same", BinaryPointOp(Operator.Merge, Fixed([1, 2]), Fixed([10, 20])), [11, 22])
    case("merge_empty", BinaryPointOp(Operator.Merge, Fixed([]), Fixed([10, 20])), [10, 20])
    # test_filter :1780-1903
    two = lambda n: [Const(2.0) for _ in range(n)]
    case("fir3_time", Filter(Time(), two(3), []), [6, 12, 18, 24, 30, 36, 42, 48])
    case("fir3_fin5", Filter(Fin(time_minus(5.0), Time()), two(3), []), [6, 12, 18, 14, 8])
    case("fir5_fin8", Filter(Fin(time_minus(8.0), Time()), two(5), []), [20, 30, 40, 50, 44, 36, 26, 14])
    case("fir2_over_reset",
         Filter(Reset(sin_waveform(1.0 / 3.0, 3.0 * F32_PI / 2.0), Time()), two(2), []),
         [0, 2, 6, 4, 2, 6, 4, 2])
    case("moving_average", Filter(Const(1.0), [Const(0.2) for _ in range(5)], []), [1.0] * 8)
    case("iir_1_1", Filter(Time(), [Const(0.5)], [Const(-0.5)]),
         [0.0, 0.5, 1.25, 2.125, 3.0625, 4.03125, 5.015625, 6.0078125])
    case("iir_cascade",
         Filter(Filter(Time(), [Const(0.5)], [Const(-0.5)]), [Const(0.4)], [Const(-0.6)]),
         [0.0, 0.2, 0.62, 1.222, 1.9582, 2.7874203, 3.6787024, 4.610347])
    case("fir_time_coeff", Filter(Const(1.0), [Const(1.0), Time()], []), [1, 2, 3, 4, 5, 6, 7, 8])
    case("fir_finite_coeffs",
         Filter(Fixed([1.0] * 3), [Const(1.0), Fixed([2.0]), Fixed([3.0] * 2)], []), [6, 3, 0])
    return out


def sine_cases(sample_rate=44100, n=100):
    """test_sine :1498-1540 — tolerance 1e-5 against the f64 analytic value."""
    tau = 2 * math.pi
    out = []
    t = np.arange(n, dtype=np.float64) / sample_rate
    out.append(("sine_1hz", sin_waveform(1.0, 0.0), np.sin(tau * np.arange(n) / sample_rate).astype(np.float32)))
    chirp = Sine(BinaryPointOp(Operator.Multiply,
                               BinaryPointOp(Operator.Add, Time(), Const(10.0)), Const(F32_TAU)),
                 Const(0.0))
    out.append(("sine_chirp", chirp, np.sin(tau * (0.5 * t * t + 10.0 * t)).astype(np.float32)))
    out.append(("sine_phase_pi", sin_waveform(0.25, F32_PI),
                np.sin(tau * 0.25 * np.arange(n) / sample_rate + math.pi).astype(np.float32)))
    return out


def length_cases():
    """check_length calls outside run_tests: (waveform, position, expected, max)."""
    app = Append(Fixed([1.0] * 3), Fixed([2.0] * 3))
    return [
        ("append_0", app, 0, 6, 1000),  # :1610
        ("append_2", app, 2, 4, 1000),  # :1611
        ("append_4", app, 4, 2, 1000),  # :1612
        ("fir5_fixed3", Filter(Fixed([1.0, 2.0, 3.0]), [Const(2.0)] * 5, []), 0, 3, 5),  # :1813
        ("fir5_fin8", Filter(Fin(time_minus(8.0), Time()), [Const(2.0)] * 5, []), 0, 8, 1000),  # :1829
    ]
