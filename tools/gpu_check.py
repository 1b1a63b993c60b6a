#!/usr/bin/env python
"""First-contact diagnostic on a GPU box: prints max |GPU - oracle| for every piece of the path."""
import os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import mpmath as mp
import pyrism_b200 as pb
from pyrism_b200 import _capi as capi
from oracle import prosail_oracle as orc

eng = pb.get_engine(0)
print(eng.device_info())
rng = np.random.default_rng(0)

# ---- math primitives
def rel(a, b): return float(np.max(np.abs(a - b) / np.abs(b)))
x = np.concatenate([rng.uniform(1e-3, 1e3, 20000), 10.0 ** rng.uniform(-30, 30, 20000)])
print("rcp   rel", rel(eng.test_math(0, x), 1.0 / x))
print("sqrt  rel", rel(eng.test_math(1, x), np.sqrt(x)))
xe = -np.concatenate([rng.uniform(0, 50, 20000), 10.0 ** rng.uniform(-12, 2.8, 20000)])
print("exp   rel", rel(eng.test_math(2, xe), np.exp(xe)))
print("log2  abs", float(np.max(np.abs(eng.test_math(3, x) - np.log2(x)))), "rel", rel(eng.test_math(3, x[x != 1]), np.log2(x[x != 1])))
y = rng.uniform(-300, 500, 40000)
print("exp2  rel", rel(eng.test_math(4, y), np.exp2(y)))
k = np.concatenate([10.0 ** rng.uniform(-12, 1.8, 4000), rng.uniform(0, 70, 2000), [0.0, -1.0, 64.0, 65.0, 100.0, 2.0 ** -30, 2.0, 1.0]])
tau_ref = np.array([float((1 - mp.mpf(v)) * mp.exp(-mp.mpf(v)) + mp.mpf(v) ** 2 * mp.e1(mp.mpf(v))) if v > 0 else 1.0 for v in k])
tg = eng.test_math(5, k)
print("tau   abs", float(np.max(np.abs(tg - tau_ref))), "rel(k<64)", rel(tg[(k > 0) & (k < 64)], tau_ref[(k > 0) & (k < 64)]))

# ---- PROSPECT
n = 64
P = dict(N=rng.uniform(1, 3, n), Cab=rng.uniform(5, 90, n), Cxc=rng.uniform(1, 25, n), Cbr=rng.uniform(0, 1, n),
         Cw=rng.uniform(0.001, 0.05, n), Cm=rng.uniform(0.001, 0.02, n))
Can = rng.uniform(0.1, 8, n)
for ver in ('5', 'D'):
    t0 = time.time()
    p = pb.PROSPECT(version=ver, Can=Can if ver == 'D' else 0, **P)
    dt = time.time() - t0
    w = [0, 0, 0]
    for i in range(n):
        o = orc.prospect(P['N'][i], P['Cab'][i], P['Cxc'][i], P['Cbr'][i], P['Cw'][i], P['Cm'][i], Can=Can[i] if ver == 'D' else 0, version=ver)
        w[0] = max(w[0], np.max(np.abs(p.ks[i] - o['ks']))); w[1] = max(w[1], np.max(np.abs(p.kt[i] - o['kt'])))
        l8 = np.array([o['L8'][b] for b, _, _ in orc.L8_WINDOWS]); gl8 = np.array([[getattr(getattr(p.L8, b), f)[i] for f in ('ks','kt','ka','ke','omega')] for b, _, _ in orc.L8_WINDOWS])
        w[2] = max(w[2], np.max(np.abs(l8 - gl8)))
    print("PROSPECT-%s  ks %.3e  kt %.3e  L8 means(f32) %.3e   (%.1f ms)" % (ver, w[0], w[1], w[2], dt * 1e3))

# ---- LSM
s = pb.LSM(reflectance=rng.uniform(0.3, 1.2, 5), moisture=rng.uniform(0, 1, 5))
o = orc.lsm(s.sRef[2], s.moisture[2])
print("LSM ref %.3e" % np.max(np.abs(s.ref[2] - o['ref'])), "L8.B2 %.3e" % abs(s.L8.B2[2] - o['L8']['B2']))

# ---- LIDF / VolScatt KATs
print("campbell(0)", np.max(np.abs(pb.LIDF.campbell(0.0) - orc.lidf_campbell(0.0))), "campbell(57)", np.max(np.abs(pb.LIDF.campbell(57.0) - orc.lidf_campbell(57.0))))
for ab in ((0, 0), (-.35, -.15), (1, 0), (0.4, -0.3), (1.5, 0)):
    print("verhoef", ab, np.max(np.abs(pb.LIDF.verhoef(*ab) - orc.lidf_verhoef(*ab))))
v = pb.VolScatt(50, 30, 50); v.coef(a=0, b=0, lidf_type='verhoef')
print("VolScatt verhoef(0,0)@50/30/50", v.ks, v.ko, v.bf, v.Fs, v.Ft, "KAT 0.823188863879162 0.6867716425258737 0.5 0.6217351737543753 0.011166182392885724")

# ---- SAIL and fused PROSAIL
lai, hot = rng.uniform(0.1, 8, n), rng.uniform(0.01, 0.5, n)
sr, sm = rng.uniform(0.3, 1.2, n), rng.uniform(0, 1, n)
iza, vza, raa = np.array([35.0, 10.0, 60.0]), np.array([30.0, 45.0, 0.0]), np.array([50.0, 120.0, 180.0])
for lt, a, b in (('campbell', rng.uniform(20, 80, n), np.zeros(n)), ('verhoef', rng.uniform(-.6, .6, n), rng.uniform(-.35, .35, n))):
    t0 = time.time()
    m = pb.PROSAIL(lai=lai, hotspot=hot, iza=iza, vza=vza, raa=raa, soil_reflectance=sr, soil_moisture=sm, lidf_type=lt, a=a, b=b, keep_leaf=True, **P)
    dt = time.time() - t0
    w = {}
    for i in range(n):
        for g in range(3):
            lf, so, c = orc.prosail(P['N'][i], P['Cab'][i], P['Cxc'][i], P['Cbr'][i], P['Cw'][i], P['Cm'][i], lai[i], hot[i], iza[g], vza[g], raa[g], sr[i], sm[i], lidf_type=lt, a=a[i], b=b[i])
            for nm, got, ref in (('BRF', m.BRF.ref[i, g], c['rsot']), ('BHR', m.BHR.ref[i, g], c['rddt']), ('DHR', m.DHR.ref[i, g], c['rsdt']), ('HDR', m.HDR.ref[i, g], c['rdot']),
                                 ('rso', m.canopy.BRF[i, g], c['rso']), ('rdd', m.canopy.BHR[i, g], c['rdd']), ('tdd', m.canopy.BHT[i, g], c['tdd']), ('rsd', m.canopy.DHR[i, g], c['rsd']),
                                 ('tsd', m.canopy.DHT[i, g], c['tsd']), ('rdo', m.canopy.HDR[i, g], c['rdo']), ('tdo', m.canopy.HDT[i, g], c['tdo']), ('ke', m.ke[i, g], c['ke']),
                                 ('tsstoo', m.kt[i, g], c['tsstoo']), ('tss', m.kt_iza[i, g], c['tss']), ('volks', m.VollScat.ks[i, g], c['vol_ks']), ('Fs', m.VollScat.Fs[i, g], c['Fs']),
                                 ('L8.B7', m.BRF.L8.B7[i, g], c['L8']['BRF']['B7']), ('AST.B9 DHR', m.DHR.ASTER.B9[i, g], c['ASTER']['DHR']['B9'])):
                w[nm] = max(w.get(nm, 0), float(np.max(np.abs(got - ref))))
    print("PROSAIL", lt, "(%.1f ms)" % (dt * 1e3), " ".join("%s %.1e" % kv for kv in w.items()))

# drop-in scalar path (README example)
p = pb.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055)
l = pb.LSM(reflectance=3.14 / 4, moisture=0.15)
c = pb.SAIL(iza=35, vza=30, raa=50, ks=p.ks, kt=p.kt, rho_surface=l.ref, lidf_type='campbell', lai=3, hotspot=0.25)
lf, so, oc = orc.prosail(1.5, 35, 5, 0.15, 0.003, 0.0055, 3, 0.25, 35, 30, 50, 3.14 / 4, 0.15)
print("README example: BRF %.3e  shapes" % np.max(np.abs(c.BRF.ref - oc['rsot'])), c.BRF.ref.shape, c.kt, c.kt_iza, c.VollScat.ks, c.VollScat.bf, c.BRF.L8.B2, oc['L8']['BRF']['B2'])
print("fp64 peak TFLOP/s:", eng.measure_fp64_peak(200))
print("launches", eng.launch_count())
