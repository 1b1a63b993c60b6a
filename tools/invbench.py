#!/usr/bin/env python
"""Kernel timing of the LUT inversion (device-resident): M LUT entries x D features against Q observations.
usage: python tools/invbench.py [M] [Q] [D]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pyrism_b200 as pb
from pyrism_b200 import inversion
M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
Q = int(float(sys.argv[2])) if len(sys.argv) > 2 else 16384
D = int(sys.argv[3]) if len(sys.argv) > 3 else 15
eng = pb.Engine(0)
rng = np.random.default_rng(0)
lut = eng.empty((M, D), np.float32).upload(rng.random((M, D), dtype=np.float32))
obs = eng.empty((Q, D), np.float32).upload(rng.random((Q, D), dtype=np.float32))
bi, bc = eng.empty((Q,), np.int64), eng.empty((Q,), np.float32)
step = lambda: inversion.invert_device(eng, M, D, lut, D, Q, obs, D, bi, bc)
for _ in range(2):
    step()
eng.sync()
eng.timer_begin()
reps = 3
for _ in range(reps):
    step()
ms = eng.timer_end() / reps
DP = 8 if D <= 8 else 16 if D <= 16 else 32
print("path:", inversion.last_stats(eng))
print("invert M=%d Q=%d D=%d: %.3f ms  %.3e pairs/s  %.2f TFLOP/s fp32 (2 x %d padded features per pair; sub counted with the fma)"
      % (M, Q, D, ms, M * Q / ms * 1e3, M * Q * 3.0 * DP / ms * 1e3 / 1e12, DP))
