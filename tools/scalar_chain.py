"""Latency of the README call chain PROSPECT -> LSM -> SAIL through the scalar drop-in API (BASELINE config 1), with a cProfile
breakdown of the host side.  usage (GPU box): python tools/scalar_chain.py"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import pyrism_b200 as pb
def chain():
    p = pb.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055)
    s = pb.LSM(reflectance=3.14 / 4, moisture=0.15)
    c = pb.SAIL(iza=35, vza=30, raa=50, ks=p.ks, kt=p.kt, rho_surface=s.ref, lidf_type='campbell', lai=3, hotspot=0.25)
    return c.BRF.ref[100]
for _ in range(20): chain()
t0 = time.perf_counter()
for _ in range(200): chain()
print("chain %.3f ms" % ((time.perf_counter() - t0) / 200 * 1e3))
for name, f in (("PROSPECT", lambda: pb.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055)), ("LSM", lambda: pb.LSM(reflectance=3.14 / 4, moisture=0.15))):
    t0 = time.perf_counter()
    for _ in range(200): f()
    print(name, "%.3f ms" % ((time.perf_counter() - t0) / 200 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): chain()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
