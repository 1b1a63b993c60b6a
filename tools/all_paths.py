#!/usr/bin/env python
"""A tiny pass through every kernel and entry point (all modes, both precisions, one and several geometries, means,
windows-only, both prologue mappings, inversion): a first-contact check on a GPU box (compute-sanitizer is closed on this pool;
the out-of-bounds check of the test-suite is the canary test in tests/test_gpu_parity.py)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pyrism_b200 as pb
from pyrism_b200 import _capi as capi, inversion

rng = np.random.default_rng(3)
eng = pb.get_engine(0)
n = 37                                             # ragged: not a multiple of the 8-item chunk
P = dict(N=rng.uniform(1, 3, n), Cab=rng.uniform(5, 90, n), Cxc=rng.uniform(1, 25, n), Cbr=rng.uniform(0, 1, n),
         Cw=rng.uniform(0.001, 0.05, n), Cm=rng.uniform(0.001, 0.02, n))
lai, hot = rng.uniform(0.0, 8, n), rng.uniform(0.0, 0.5, n)
lai[3] = 0.0
sr, sm = rng.uniform(0.3, 1.2, n), rng.uniform(0, 1, n)
acc = 0.0
for dt in (np.float64, np.float32):
    for thr in (2048, 0):
        eng.set_prologue_threshold(thr)
        leaf = pb.PROSPECT(dtype=dt, **P)
        acc += float(leaf.ks.sum() + leaf.ka.sum() + leaf.L8.B5.ks.sum())
        leafd = pb.PROSPECT(Can=rng.uniform(0.1, 8, n), version='D', dtype=dt, **P)
        soil = pb.LSM(sr, sm, dtype=dt)
        acc += float(soil.ref.sum() + soil.L8.B2.sum() + leafd.kt.sum())
        for lidf, a, b in (('campbell', rng.uniform(20, 80, n), 0.0), ('verhoef', rng.uniform(-0.6, 0.6, n), rng.uniform(-0.3, 0.3, n))):
            m = pb.PROSAIL(lai=lai, hotspot=hot, iza=[35.0, 10.0, 0.0], vza=[30.0, 45.0, 0.0], raa=[50.0, 120.0, 0.0], soil_reflectance=sr,
                           soil_moisture=sm, lidf_type=lidf, a=a, b=b, keep_leaf=True, dtype=dt, **P)
            acc += float(m.BRF.ref.sum() + m.BHR.L8.B4.sum() + m.canopy.DHT.sum())
            c = pb.SAIL(iza=35, vza=30, raa=50, ks=leaf.ks, kt=leaf.kt, lai=lai, hotspot=hot, rho_surface=soil.ref, lidf_type=lidf, a=a, b=b, dtype=dt)
            acc += float(c.BRF.ref.sum())
            g = (np.radians([35.0]), np.radians([30.0]), np.radians([50.0]))
            args = (P['N'], P['Cab'], P['Cxc'], P['Cbr'], P['Cw'], P['Cm'], lai, hot) + g + (sr, sm)
            for kw in (dict(planes=(capi.O_RSOT,)), dict(planes=(capi.O_RSOT, capi.O_RDDT, capi.O_RSDT, capi.O_RDOT)),
                       dict(mean_planes=(capi.O_RSOT,), windows_only=True)):
                r = eng.prosail(*args, lidf_type=lidf, a=a, b=b, dtype=dt, **kw)
                acc += float(sum(v.sum() for v in r.values()))
        vs = pb.VolScatt([35.0, 10.0], [30.0, 45.0], [50.0, 120.0])
        vs.coef(lidf_type='campbell', a=57)
        acc += float(vs.ks.sum() + pb.LIDF.verhoef(-0.35, -0.15).sum())
        vs.coef(lidf_type='verhoef', n_elements=36, a=-0.35, b=-0.15)      # the general leaf-class kernel
        acc += float(vs.ks.sum() + pb.LIDF.campbell(57.0, n_elements=7).sum())
eng.set_prologue_threshold(2048)
for m_, q_, d_ in ((1, 1, 1), (700, 1030, 15), (513, 5, 32), (300, 9, 6)):
    lutf = rng.random((m_, d_)).astype(np.float32)
    obs = rng.random((q_, d_)).astype(np.float32)
    idx, cost = inversion.invert(eng, lutf, obs)
    acc += float(cost.sum()) + float(idx.sum())
print("all paths done, checksum %.6f" % acc)
