#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Throughput of every BASELINE.json configuration on one B200 (device-resident inputs/outputs, CUDA-event timing on
the launching stream; warm-up first).  One JSON line per measurement; the bench.py line stays the contract, this is
the table behind DESIGN.md section 6.

    python tools/bench_configs.py [--quick] > gpurun_out/configs.jsonl

Synthetic Latin-hypercube parameter sets with the SURVEY.md 8(d) bounds, seed 20261019 + config index.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import pyrism_b200 as pb                      # noqa: E402
from pyrism_b200 import _capi as capi         # noqa: E402

NB = 2101
SEED = 20261019


def lhs(n, names, lo, hi, seed):
    from scipy.stats import qmc
    x = qmc.scale(qmc.LatinHypercube(d=len(names), seed=seed).random(n), lo, hi)
    return {k: np.ascontiguousarray(x[:, i]) for i, k in enumerate(names)}


LEAF = (("N", 1.0, 3.0), ("Cab", 5.0, 90.0), ("Cxc", 1.0, 25.0), ("Cbr", 0.0, 1.0), ("Cw", 0.001, 0.05), ("Cm", 0.001, 0.02))
CANOPY = (("lai", 0.1, 8.0), ("hotspot", 0.01, 0.5), ("soil_refl", 0.3, 1.2), ("soil_moist", 0.0, 1.0))


def draw(n, extra, seed):
    spec = LEAF + CANOPY + tuple(extra)
    return lhs(n, [s[0] for s in spec], [s[1] for s in spec], [s[2] for s in spec], seed)


def timed(eng, fn, reps, warm=3):
    for _ in range(warm):
        fn()
    eng.sync()
    eng.timer_begin()
    for _ in range(reps):
        fn()
    return eng.timer_end() / reps


def emit(**kw):
    print(json.dumps(kw))
    sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="1/10 of the sizes (smoke run)")
    args = ap.parse_args()
    q = 10 if args.quick else 1
    eng = pb.Engine(0)
    info = eng.device_info()

    # ---- config 1: the README example through the drop-in classes (latency of one scalar call chain) ----------------
    for ver, kw in (("5", {}), ("D", {"Can": 1.0, "version": "D"})):
        def readme():
            leaf = pb.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055, **kw)
            soil = pb.LSM(reflectance=3.14 / 4, moisture=0.15)
            can = pb.SAIL(iza=35, vza=30, raa=50, ks=leaf.ks, kt=leaf.kt, rho_surface=soil.ref, lidf_type='campbell', lai=3, hotspot=0.25)
            return can.BRF.ref[0]
        readme()
        t0 = time.perf_counter()
        for _ in range(20):
            readme()
        ms = (time.perf_counter() - t0) / 20 * 1e3
        emit(config=1, what="README example PROSPECT-%s + LSM + SAIL (scalar drop-in calls, host in / host out)" % ver,
             ms_per_call_chain=ms, spectra_per_s=1e3 / ms, dtype="f64", device=info["name"])

    # ---- config 2: PROSPECT-5 leaf-only, 1e7 sets x 2101 bands (ks + kt), streamed through a 1e6-set output slab -----
    n2, slab = 10000000 // q, 1000000 // q
    P = draw(n2, (), SEED + 1)
    for dt, name in ((np.float64, "f64"), (np.float32, "f32")):
        ks, kt = eng.empty((slab, NB), dt), eng.empty((slab, NB), dt)
        dev = {k: eng.to_device(P[k]) for k in ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm")}

        def leaf_pass():
            for s0 in range(0, n2, slab):
                leaf = {k: dev[k].ptr + 8 * s0 for k in dev}
                eng.prospect_device(min(slab, n2 - s0), leaf, ks, kt, dtype=dt)
        ms = timed(eng, leaf_pass, 2, warm=1)
        emit(config=2, what="PROSPECT-5 leaf-only, %d parameter sets x 2101 bands, ks + kt planes" % n2, dtype=name, ms=ms,
             spectra_per_s=n2 / ms * 1e3, output_bytes_per_s=n2 * 2 * NB * np.dtype(dt).itemsize / ms * 1e3)
        ks.free(); kt.free()

    # ---- config 3: PROSAIL LUT, 1e6 LHS sets x 2101 bands, BRF only and the four reflectance factors, fp64 and fp32 ----
    n3 = 1000000 // q
    P = draw(n3, (("lidf_a", 20.0, 80.0),), SEED + 2)
    dev = {k: eng.to_device(v) for k, v in P.items()}
    geom = np.deg2rad([35.0, 30.0, 50.0])
    leaf = {k: dev[k] for k in ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm")}
    soil = {"reflectance": dev["soil_refl"], "moisture": dev["soil_moist"]}
    canopy = {"lidf_a": dev["lidf_a"], "lai": dev["lai"], "hotspot": dev["hotspot"], "iza": eng.to_device(geom[0:1]),
              "vza": eng.to_device(geom[1:2]), "raa": eng.to_device(geom[2:3])}
    for dt, name in ((np.float64, "f64"), (np.float32, "f32")):
        for planes in ((capi.O_RSOT,), (capi.O_RSOT, capi.O_RDDT, capi.O_RSDT, capi.O_RDOT)):
            bufs = {p: eng.empty((n3, NB), dt) for p in planes}
            ms = timed(eng, lambda: eng.prosail_device(n3, leaf, soil, canopy, bufs, dtype=dt), 5)
            emit(config=3, what="PROSAIL LUT %d LHS sets x 2101 bands, campbell, %d output plane(s)" % (n3, len(planes)), dtype=name, ms=ms,
                 spectra_per_s=n3 / ms * 1e3, output_bytes_per_s=n3 * len(planes) * NB * np.dtype(dt).itemsize / ms * 1e3)
            for b in bufs.values():
                b.free()

    # ---- config 5 (one GPU's share): band-mean LUT, fused window means of the BRF, nothing else written ----------------
    n5 = 12500000 // q
    slab5 = 1250000 // q
    P5 = draw(n5, (("lidf_a", 20.0, 80.0),), SEED + 4)
    dev5 = {k: eng.to_device(v) for k, v in P5.items()}
    for dt, name in ((np.float64, "f64"), (np.float32, "f32")):
        means = eng.empty((n5, 1, len(eng.windows)), dt)

        def means_pass():
            for s0 in range(0, n5, slab5):
                sl = lambda k: dev5[k].ptr + 8 * s0
                lf = {k: sl(k) for k in ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm")}
                so = {"reflectance": sl("soil_refl"), "moisture": sl("soil_moist")}
                ca = {"lidf_a": sl("lidf_a"), "lai": sl("lai"), "hotspot": sl("hotspot"), "iza": canopy["iza"], "vza": canopy["vza"], "raa": canopy["raa"]}
                eng.prosail_device(min(slab5, n5 - s0), lf, so, ca, None, mean_planes=(capi.O_RSOT,),
                                   means=means.ptr + s0 * len(eng.windows) * np.dtype(dt).itemsize, windows_only=True, dtype=dt)
        ms = timed(eng, means_pass, 2, warm=1)
        emit(config=5, what="L8/ASTER band-mean PROSAIL LUT, %d sets (one GPU's share of 1e8 over 8), 15 window means of the BRF, "
             "only the 782 wavelengths inside a window are evaluated" % n5, dtype=name, ms=ms, spectra_per_s=n5 / ms * 1e3)
        means.free()
    for d in list(dev5.values()):
        d.free()

    # ---- config 4: multi-angle 4SAIL, 1e5 sets x 64 geometries, Verhoef LIDF, BRF plane (streamed in slabs of 1e4 sets) ---
    n4, G, slab4 = 100000 // q, 64, 10000 // q
    P4 = draw(n4, (("lidf_a", -0.6, 0.6), ("lidf_b", -0.35, 0.35)), SEED + 3)
    dev4 = {k: eng.to_device(v) for k, v in P4.items()}
    iza, vza, raa = np.meshgrid([15.0, 30.0, 45.0, 60.0], [0.0, 15.0, 30.0, 45.0], [0.0, 60.0, 120.0, 180.0], indexing="ij")
    gd = {"iza": eng.to_device(np.deg2rad(iza.ravel())), "vza": eng.to_device(np.deg2rad(vza.ravel())), "raa": eng.to_device(np.deg2rad(raa.ravel()))}
    for dt, name in ((np.float64, "f64"), (np.float32, "f32")):
        out = eng.empty((slab4 * G, NB), dt)

        def multi_pass():
            for s0 in range(0, n4, slab4):
                sl = lambda k: dev4[k].ptr + 8 * s0
                lf = {k: sl(k) for k in ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm")}
                so = {"reflectance": sl("soil_refl"), "moisture": sl("soil_moist")}
                ca = {"lidf_a": sl("lidf_a"), "lidf_b": sl("lidf_b"), "lai": sl("lai"), "hotspot": sl("hotspot"), **gd}
                eng.prosail_device(min(slab4, n4 - s0), lf, so, ca, {capi.O_RSOT: out}, lidf_type='verhoef', n_geom=G, dtype=dt)
        eng.profile(True); eng.profile_read()
        ms = timed(eng, multi_pass, 2, warm=1)
        pr = eng.profile_read(); eng.profile(False)
        emit(config=4, what="multi-angle PROSAIL, %d sets x %d geometries, Verhoef LIDF, BRF plane" % (n4, G), dtype=name, ms=ms,
             spectra_per_s=n4 * G / ms * 1e3, prologue_share=pr['prologue'][0] / max(pr['prologue'][0] + pr['bands'][0], 1e-9),
             output_bytes_per_s=n4 * G * NB * np.dtype(dt).itemsize / ms * 1e3)
        out.free()


if __name__ == "__main__":
    main()
