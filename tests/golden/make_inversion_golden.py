#!/usr/bin/env python
"""Generates tests/golden/inversion_lut.npz: a band-mean look-up table and observations made from spectra of the LIVE,
UNMODIFIED reference (ibaris/pyrism from /root/reference, or its copy oracle/_ref), plus the arg-min of an independent
float64 brute-force search.  This pins the LUT inversion (which has no counterpart in the reference) to reference DATA and to
an arithmetic that shares nothing with the kernel's float32 contract:

    lut      float64 [M, 15]  L8 B2..B7 + ASTER B1..B9 band means of BRF for M Latin-hypercube parameter sets (reference output)
    params   float64 [M, 11]  the parameter sets (N Cab Cxc Cbr Cw Cm lai hotspot campbell_a soil_refl soil_moist)
    obs      float64 [Q, 15]  Q/2 held-out reference spectra + Q/2 LUT rows with 0.1 % multiplicative noise
    best     int64   [Q]      argmin_m sum_d (obs[q, d] - lut[m, d])^2 in float64, lowest index on ties
    cost     float64 [Q]      that minimum
    margin   float64 [Q]      (second smallest cost - smallest cost): float32 evaluation may only pick another row where this
                              is below float32 rounding of the cost

    python tests/golden/make_inversion_golden.py          (needs the reference: about 10 s)
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ref_runner as rr          # noqa: E402
from scipy.stats import qmc                  # noqa: E402

LO = np.array([1.0, 5.0, 1.0, 0.0, 0.001, 0.001, 0.1, 0.01, 20.0, 0.3, 0.0])
HI = np.array([3.0, 90.0, 25.0, 1.0, 0.05, 0.02, 8.0, 0.5, 80.0, 1.2, 1.0])
GEOM = (35.0, 30.0, 50.0)


def band_means(p):
    _, _, c = rr.prosail(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], GEOM[0], GEOM[1], GEOM[2], p[9], p[10], lidf_type='campbell', a=p[8])
    return np.array([getattr(c.BRF.L8, "B%d" % i) for i in range(2, 8)] + [getattr(c.BRF.ASTER, "B%d" % i) for i in range(1, 10)], dtype=np.float64)


def main():
    M, Q = 600, 100
    P = qmc.scale(qmc.LatinHypercube(d=11, seed=20261019 + 5).random(M), LO, HI)
    lut = np.stack([band_means(p) for p in P])
    held = qmc.scale(qmc.LatinHypercube(d=11, seed=20261019 + 6).random(Q // 2), LO, HI)
    rng = np.random.default_rng(20261019 + 7)
    rows = rng.integers(0, M, Q // 2)
    obs = np.concatenate([np.stack([band_means(p) for p in held]), lut[rows] * (1.0 + 1e-3 * rng.standard_normal((Q // 2, 15)))])
    c = ((obs[:, None, :] - lut[None, :, :]) ** 2).sum(axis=2)          # float64, [Q, M]
    best = np.argmin(c, axis=1).astype(np.int64)
    srt = np.sort(c, axis=1)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "inversion_lut.npz")
    np.savez_compressed(out, lut=lut, params=P, obs=obs, best=best, cost=srt[:, 0], margin=srt[:, 1] - srt[:, 0], noisy_rows=rows)
    print("wrote %s: LUT %s, %d observations; %d of the noisy rows are recovered; smallest relative margin %.2e"
          % (out, lut.shape, Q, int((best[Q // 2:] == rows).sum()), float(np.min((srt[:, 1] - srt[:, 0]) / srt[:, 1]))))


if __name__ == "__main__":
    main()
