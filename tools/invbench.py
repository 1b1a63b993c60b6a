#!/usr/bin/env python
"""Kernel timing of the LUT inversion (device-resident): M LUT entries x D features against Q observations.
usage: python tools/invbench.py [M] [Q] [D] [random|real]
  real: the LUT is the band-mean table of M Latin-hypercube parameter sets (L8 + ASTER BRF means, the bench's configs[4] product) and the
        observations are the band means of Q held-out sets -- the dense, strongly correlated table the tensor-core error bound must cope with."""
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
KIND = sys.argv[4] if len(sys.argv) > 4 else "random"
if KIND == "real":
    import bench
    from pyrism_b200 import _capi
    g = np.radians(np.array(bench.GEOM_DEG))

    def means(n, seed):
        P = bench.lhs(n, seed)
        return eng.prosail(P["N"], P["Cab"], P["Cxc"], P["Cbr"], P["Cw"], P["Cm"], P["lai"], P["hotspot"], g[0:1], g[1:2], g[2:3], P["soil_refl"],
                           P["soil_moist"], version='5', lidf_type='campbell', a=P["lidf_a"], planes=(), mean_planes=(_capi.O_RSOT,),
                           windows_only=True)["means"][:, 0, 0, :D].astype(np.float32)
    lut_h, obs_h = means(M, 1234), means(Q, 99)
else:
    lut_h, obs_h = rng.random((M, D), dtype=np.float32), rng.random((Q, D), dtype=np.float32)
lut = eng.empty((M, D), np.float32).upload(lut_h)
obs = eng.empty((Q, D), np.float32).upload(obs_h)
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
print("path:", inversion.last_stats(eng), "kappa env:", os.environ.get("PYRISM_B200_TC_KAPPA"), "median cost %.3e" % float(np.median(bc.download())))
print("invert M=%d Q=%d D=%d: %.3f ms  %.3e pairs/s  %.2f TFLOP/s fp32 (2 x %d padded features per pair; sub counted with the fma)"
      % (M, Q, D, ms, M * Q / ms * 1e3, M * Q * 3.0 * DP / ms * 1e3 / 1e12, DP))
