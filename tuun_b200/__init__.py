"""tuun_b200 — B200-native renderer for Tuun's waveform-generation hot path.

Layout: csrc/ (CUDA kernels, lowering, C ABI -> libtuun_b200.so), waveform.py (the Waveform IR
mirror and its flattening to the ABI's op list), generator.py (host mirror of the reference's
Generator interface), optimizer.py / std.py (config builders: the reference optimizer and the
lib/v0/std.tuun definitions the named workloads are written in).
"""
from .waveform import (Alt, Append, BinaryPointOp, Captured, Const, Filter, Fin, Fixed, Marked,  # noqa: F401
                       Noise, Operator, Reset, Sine, Time, Waveform, flatten)
