#!/usr/bin/env python
"""The all-zero-pigment leaf (Cab = Cxc = Cbr = Cw = Cm = 0: k = 0, tau = 1, r + t = 1 up to rounding) -- the one case whose parity
test is held to 1e-7 instead of 1e-9 (tests/test_gpu_parity.py::test_prospect_extreme_absorption).  This probe says why: the
reference decides between its Stokes formulas and its zero-absorption branch (models.py:1057-1059) on the last ulp of r + t, and
where it takes the Stokes formulas with 1 - r - t ~ 1e-16 they carry 1e-16 / D ~ 1e-8 of rounding noise (a^2 - 1, b^(2(N-1)) - 1).
Printed per N: how many bands the oracle (= the reference, bit for bit) sends through each branch, and the deviations of the
oracle and of the GPU from the exact zero-absorption limit evaluated with mpmath (50 digits) on the same float64 constants.
usage (GPU box): python tools/zero_leaf_probe.py > gpurun_out/zero_leaf_probe.txt"""
import os, sys
import numpy as np
import mpmath as mp
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pyrism_b200 as pb
from oracle import prosail_oracle as orc

mp.mp.dps = 50
for N in (1.0, 1.5, 2.3, 3.0):
    o = orc.prospect(N, 0.0, 0.0, 0.0, 0.0, 0.0)
    p = pb.PROSPECT(N, 0.0, 0.0, 0.0, 0.0, 0.0)
    T = orc.load_tables()
    KN = T['p5_KN']
    talf = orc.calctav(40, KN); t12 = orc.calctav(90, KN); t21 = t12 / (KN * KN)
    ks_true, kt_true = np.empty(2101), np.empty(2101)
    for i in range(2101):
        ta, tt, t2 = mp.mpf(float(talf[i])), mp.mpf(float(t12[i])), mp.mpf(float(t21[i]))
        r21 = mp.mpf(float(1 - t21[i])); ralf = mp.mpf(float(1.0 - talf[i])); r12 = mp.mpf(float(1. - t12[i]))
        den = 1 - r21 * r21
        Ta = ta * t2 / den; Ra = ralf + r21 * Ta
        t = tt * t2 / den; r = r12 + r21 * t
        Tsub = t / (t + (1 - t) * (mp.mpf(N) - 1)); Rsub = 1 - Tsub      # the limit of the Stokes formulas for 1 - r - t -> 0
        d2 = 1 - Rsub * r
        kt_true[i] = float(Ta * Tsub / d2); ks_true[i] = float(Ra + Ta * Rsub * t / d2)
    jz = (o['r'] + o['t']) >= 1.0
    dev = lambda a, b: float(np.max(np.abs(a - b)))
    print("N = %.1f: oracle takes the zero-absorption branch in %d bands, the Stokes formulas in %d (1 - r - t there: %.1e .. %.1e)"
          % (N, int(jz.sum()), int((~jz).sum()), float(np.min((1 - o['r'] - o['t'])[~jz])) if (~jz).any() else 0.0,
             float(np.max((1 - o['r'] - o['t'])[~jz])) if (~jz).any() else 0.0))
    print("    |oracle - limit|  ks %.2e  kt %.2e   (Stokes bands only: ks %.2e, zero-absorption bands only: ks %.2e)"
          % (dev(o['ks'], ks_true), dev(o['kt'], kt_true), dev(o['ks'][~jz], ks_true[~jz]) if (~jz).any() else 0.0,
             dev(o['ks'][jz], ks_true[jz]) if jz.any() else 0.0))
    print("    |GPU    - limit|  ks %.2e  kt %.2e" % (dev(p.ks, ks_true), dev(p.kt, kt_true)))
    print("    |GPU    - oracle| ks %.2e  kt %.2e" % (dev(p.ks, o['ks']), dev(p.kt, o['kt'])))
