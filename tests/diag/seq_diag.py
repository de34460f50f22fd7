import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from tuun_b200 import workloads as W
from tuun_b200.generator import Program
from tuun_b200 import _abi
w = W.cfg2_harmonica(4)
p = Program(w, 44100)
print("parts", p.info.sequence_parts, "err", _abi.lib().tb_last_error())
out = np.zeros((1, 100000), dtype=np.float32)
print(p.render(out), "seq renders", p.info.sequence_renders, "launches", p.info.kernel_launches)
