import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["TUUN_B200_LANE_MIN_VOICES"]="1"; os.environ["TUUN_B200_SPLIT"]="0"
import numpy as np
from tuun_b200 import workloads as W
from tuun_b200.generator import Program
seq = W.cfg2_harmonica(4)
for V,N in ((130,88200),(130,100000)):
    try:
        s = Program(seq, 44100)
        out = np.zeros((V, N), dtype=np.float32)
        print(V, N, np.asarray(s.render(out))[:3], flush=True)
    except Exception as e:
        print(V, N, "ERR", e, flush=True)
