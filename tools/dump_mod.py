"""Dump the GPU render of a constant-rate sine (cfg5 modulator of voice 300) for offline comparison
against exact arithmetic.  usage: python tools/dump_mod.py out.npy"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tuun_b200.generator import Program
from tuun_b200.waveform import Sine, Const
from tuun_b200.workloads import fm_filter_params
pr = fm_filter_params(np.arange(65536))[300]
w = Sine(Const(float(pr[0])), Const(float(np.float32(np.float32(3.14159265) / np.float32(2)))))
p = Program(w, 44100)
out = np.zeros((1, 441000), np.float32)
p.render(out)
np.save(sys.argv[1], out[0])
