#!/usr/bin/env python
"""Kernel-level timing of the fused PROSAIL kernel (device-resident inputs/outputs) for tuning experiments.
usage: [PYRISM_B200_LIB=path.so] python tools/kbench.py [samples] [reps] [planes: brf|four|leaf|means|multi] [f64|f32]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
import pyrism_b200 as pb
from pyrism_b200 import _capi as capi
S = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
what = sys.argv[3] if len(sys.argv) > 3 else "brf"
dt = np.float32 if (len(sys.argv) > 4 and sys.argv[4] == "f32") else np.float64
PITCH = int(sys.argv[5]) if len(sys.argv) > 5 else 2101      # row pitch of the output planes in elements (2112 = 128-byte aligned rows)
eng = pb.Engine(0)
P = bench.lhs(S, 123)
geom = np.deg2rad(np.array(bench.GEOM_DEG))
dev = {k: eng.to_device(v) for k, v in P.items()}
leaf = {k: dev[k] for k in ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm")}
soil = {"reflectance": dev["soil_refl"], "moisture": dev["soil_moist"]}
canopy = {"lidf_a": dev["lidf_a"], "lai": dev["lai"], "hotspot": dev["hotspot"], "iza": eng.to_device(geom[0:1]),
          "vza": eng.to_device(geom[1:2]), "raa": eng.to_device(geom[2:3])}
if what == "leaf":
    ks, kt = eng.empty((S, PITCH), dt), eng.empty((S, PITCH), dt)
    step = lambda: eng.prospect_device(S, leaf, ks, kt, pitch=PITCH, dtype=dt)
elif what == "multi":                      # BASELINE config 4 shape: S sets x 64 geometries, Verhoef LIDF, BRF plane
    G = 64
    iza, vza, raa = np.meshgrid([15.0, 30.0, 45.0, 60.0], [0.0, 15.0, 30.0, 45.0], [0.0, 60.0, 120.0, 180.0], indexing="ij")
    can4 = dict(canopy, iza=eng.to_device(np.deg2rad(iza.ravel())), vza=eng.to_device(np.deg2rad(vza.ravel())), raa=eng.to_device(np.deg2rad(raa.ravel())),
                lidf_a=eng.to_device(np.random.default_rng(1).uniform(-0.6, 0.6, S)), lidf_b=eng.to_device(np.random.default_rng(2).uniform(-0.35, 0.35, S)))
    out4 = eng.empty((S * G, PITCH), dt)
    step = lambda: eng.prosail_device(S, leaf, soil, can4, {capi.O_RSOT: out4}, pitch=PITCH, lidf_type='verhoef', n_geom=G, dtype=dt)
elif what == "means":
    means = eng.empty((S, 1, len(eng.windows)), dt)
    step = lambda: eng.prosail_device(S, leaf, soil, canopy, None, mean_planes=(capi.O_RSOT,), means=means, windows_only=True, dtype=dt)
else:
    planes = {capi.O_RSOT: eng.empty((S, PITCH), dt)}
    if what == "four":
        for p in (capi.O_RDDT, capi.O_RSDT, capi.O_RDOT):
            planes[p] = eng.empty((S, PITCH), dt)
    step = lambda: eng.prosail_device(S, leaf, soil, canopy, planes, pitch=PITCH, dtype=dt)
for _ in range(3):
    step()
eng.sync()
eng.profile(True); eng.profile_read()
for _ in range(reps):
    step()
pr = eng.profile_read()
ms = pr['bands'][0] / pr['bands'][1]
if what == "multi":
    S = S * 64
print("%s %s pitch %d lib=%s S=%d kernel %.3f ms  -> %.3e spectra/s   prologue %.3f ms" % (what, np.dtype(dt).name, PITCH, os.path.basename(capi.LIB_PATH), S, ms, S / ms * 1e3, pr['prologue'][0] / max(pr['prologue'][1], 1)))
