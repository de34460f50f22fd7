"""One harmonica note (config 2) over 16,384 voices on the lane interpreter kernel: the launch ncu looks at."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tuun_b200 import workloads as W
from tuun_b200.generator import Program
note = W.cfg2_harmonica(2).a
V, N = 16384, 22050
p = Program(note, 44100)
out = torch.empty((V, N), dtype=torch.float32, device="cuda")
for rep in range(2):
    p.reset()
    p.render(out)
    torch.cuda.synchronize()
print("lane launches", p.info.lane_launches)
