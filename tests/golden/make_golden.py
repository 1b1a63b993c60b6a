#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the LIVE reference.

Runs only in the build container (needs /root/reference).  The reference is
imported unmodified; the only accommodation is the 2-line shim for the removed
``scipy.misc.factorial`` name, applied before import.  Every
scalar is passed as ``np.float64``.

Outputs (all small, committed):
  ref_cases.npz      (spectra at every 2nd band, band means over all bands) inputs + reference outputs for N_LEAF leaf-only cases (P5 and D) and
                     N_CANOPY PROSPECT->LSM->SAIL cases (campbell and verhoef, random geometry),
                     drawn from a seeded Latin hypercube over the SURVEY.md 8(d) bounds, plus
                     the README example (config 1, as P5 and as D with Can=1) and edge cases.
  ref_fixtures.npz   the reference's own golden data for this path: prospect5_spectrum.txt,
                     prospect_d_test.mat (LRT), REFL_CAN.txt, tav_alpha40.mat.
Also prints the max |oracle - reference| per quantity, i.e. the pin of oracle/prosail_oracle.py.
"""
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
import scipy.special, scipy.misc                                    # noqa: E401
scipy.misc.factorial = scipy.special.factorial                       # reference models.py:10 imports the removed name
REF = os.environ.get("PYRISM_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
os.chdir("/tmp")
import pyrism                                                        # noqa: E402
from scipy.io import loadmat                                         # noqa: E402
from scipy.stats import qmc                                          # noqa: E402

from oracle import prosail_oracle as orc                             # noqa: E402

f = np.float64
N_LEAF, N_CANOPY = 12, 20
BANDS = np.arange(0, 2101, 2)          # spectra are stored at every 2nd band (size); band means use all bands
SEED = 20261019

L8N = [w[0] for w in orc.L8_WINDOWS]
ASN = [w[0] for w in orc.ASTER_WINDOWS]

# ---- bounds of SURVEY.md 8(d): N Cab Cxc Cbr Cw Cm Can lai hot lidf_a(campbell) va vb srefl smoist iza vza raa
LO = np.array([1.0, 5, 1, 0, 0.001, 0.001, 0.1, 0.1, 0.01, 20, -0.6, -0.35, 0.3, 0, 0, 0, 0])
HI = np.array([3.0, 90, 25, 1, 0.05, 0.02, 8, 8, 0.5, 80, 0.6, 0.35, 1.2, 1, 70, 60, 180])
X = qmc.scale(qmc.LatinHypercube(d=len(LO), seed=SEED).random(N_CANOPY), LO, HI)

worst = {}


def track(name, a, b):
    d = float(np.max(np.abs(np.asarray(a, dtype=f) - np.asarray(b, dtype=f))))
    worst[name] = max(worst.get(name, 0.0), d)


def ref_leaf(x, version):
    kw = dict(N=f(x[0]), Cab=f(x[1]), Cxc=f(x[2]), Cbr=f(x[3]), Cw=f(x[4]), Cm=f(x[5]), version=version)
    if version == 'D':
        kw['Can'] = f(x[6])
    return pyrism.PROSPECT(**kw), kw


def leaf_means(p):
    l8 = np.array([[getattr(getattr(p.L8, b), q) for q in ('ks', 'kt', 'ka', 'ke', 'omega')] for b in L8N], dtype=np.float32)
    a = np.array([[getattr(getattr(p.ASTER, b), q) for q in ('ks', 'kt', 'ka', 'ke', 'omega')] for b in ASN], dtype=np.float32)
    return l8, a


out = {'band_index': BANDS}

# ------------------------------------------------------------------ leaf-only cases
leaf_in, leaf_ver, leaf_out, leaf_l8, leaf_as = [], [], [], [], []
for i in range(N_LEAF):
    ver = '5' if i % 2 == 0 else 'D'
    p, kw = ref_leaf(X[i], ver)
    o = orc.prospect(**kw)
    for q in ('ks', 'kt', 'ka', 'ke', 'om', 'r', 't', 'Ra', 'Ta'):
        track('leaf.' + q, getattr(p, q), o[q])
    l8, a = leaf_means(p)
    track('leaf.L8(f32)', l8, np.array([o['L8'][b] for b in L8N]))
    track('leaf.ASTER(f32)', a, np.array([o['ASTER'][b] for b in ASN]))
    leaf_in.append([X[i][0], X[i][1], X[i][2], X[i][3], X[i][4], X[i][5], X[i][6] if ver == 'D' else 0.0])
    leaf_ver.append(ver)
    leaf_out.append(np.stack([p.ks, p.kt, p.ka, p.ke, p.om])[:, BANDS])
    leaf_l8.append(l8)
    leaf_as.append(a)
out.update(leaf_in=np.array(leaf_in), leaf_version=np.array(leaf_ver), leaf_out=np.array(leaf_out),
           leaf_L8=np.array(leaf_l8), leaf_ASTER=np.array(leaf_as))

# per-band constants of the reference (name-mangled privates), both table sets, alpha = 40 and 90
pr5, _ = ref_leaf(X[0], '5')
prd, _ = ref_leaf(X[0], 'D')
for tag, pr, KN in (('p5', pr5, pyrism.models.models.lib.p5.KN), ('pd', prd, pyrism.models.models.lib.pd.KN)):
    out['tav40_' + tag] = pr._PROSPECT__calctav(40, KN)
    out['tav90_' + tag] = pr._PROSPECT__calctav(90, KN)
    out['tav57_' + tag] = pr._PROSPECT__calctav(57, KN)
    track('tav', out['tav40_' + tag], orc.calctav(40, KN))
    track('tav', out['tav90_' + tag], orc.calctav(90, KN))
    track('tav', out['tav57_' + tag], orc.calctav(57, KN))


# ------------------------------------------------------------------ canopy cases
def ref_canopy(x, ver, lidf_type, leafkw=None, geom=None, lai=None, hot=None, soil=None, ab=None):
    if leafkw is None:
        p, leafkw = ref_leaf(x, ver)
    else:
        p = pyrism.PROSPECT(**leafkw)
    sr, sm = (f(x[12]), f(x[13])) if soil is None else soil
    s = pyrism.LSM(reflectance=sr, moisture=sm)
    iza, vza, raa = (f(x[14]), f(x[15]), f(x[16])) if geom is None else geom
    lai = f(x[7]) if lai is None else lai
    hot = f(x[8]) if hot is None else hot
    if ab is None:
        ab = (f(x[9]), f(0.0)) if lidf_type == 'campbell' else (f(x[10]), f(x[11]))
    kw = dict(iza=iza, vza=vza, raa=raa, lai=lai, hotspot=hot, lidf_type=lidf_type, a=ab[0], b=ab[1])
    c = pyrism.SAIL(ks=p.ks, kt=p.kt, rho_surface=s.ref, **kw)
    ol, os_, oc = orc.prosail(leafkw['N'], leafkw['Cab'], leafkw['Cxc'], leafkw['Cbr'], leafkw['Cw'], leafkw['Cm'],
                              lai, hot, iza, vza, raa, sr, sm, Can=leafkw.get('Can', 0), version=leafkw['version'],
                              lidf_type=lidf_type, a=ab[0], b=ab[1])
    track('soil.ref', s.ref, os_['ref'])
    for q, r in (('rsot', c.BRF.ref), ('rddt', c.BHR.ref), ('rsdt', c.DHR.ref), ('rdot', c.HDR.ref),
                 ('rdd', c.canopy.BHR), ('tdd', c.canopy.BHT), ('rsd', c.canopy.DHR), ('tsd', c.canopy.DHT),
                 ('rdo', c.canopy.HDR), ('tdo', c.canopy.HDT), ('rso', c.canopy.BRF), ('ke', c.ke),
                 ('tsstoo', c.kt), ('tss', c.kt_iza), ('too', c.kt_vza), ('vol_ks', c.VollScat.ks),
                 ('vol_ko', c.VollScat.ko), ('bf', c.VollScat.bf), ('Fs', c.VollScat.Fs), ('Ft', c.VollScat.Ft)):
        track('canopy.' + q, np.broadcast_to(r, np.shape(oc[q])) if np.ndim(oc[q]) else np.ravel(r)[0], oc[q])
    prods = (c.BRF, c.BRDF, c.BHR, c.DHR, c.HDR)
    l8 = np.array([[getattr(pp.L8, b) for b in L8N] for pp in prods])
    a = np.array([[getattr(pp.ASTER, b) for b in ASN] for pp in prods])
    track('canopy.L8', l8, np.array([[oc['L8'][k][b] for b in L8N] for k in ('BRF', 'BRDF', 'BHR', 'DHR', 'HDR')]))
    track('canopy.ASTER', a, np.array([[oc['ASTER'][k][b] for b in ASN] for k in ('BRF', 'BRDF', 'BHR', 'DHR', 'HDR')]))
    inp = [leafkw['N'], leafkw['Cab'], leafkw['Cxc'], leafkw['Cbr'], leafkw['Cw'], leafkw['Cm'], leafkw.get('Can', 0.0),
           lai, hot, ab[0], ab[1], sr, sm, iza, vza, raa]
    spectra = np.stack([np.broadcast_to(np.asarray(v, dtype=f), (2101,)) for v in
                        (c.BRF.ref, c.BHR.ref, c.DHR.ref, c.HDR.ref, c.canopy.BHR, c.canopy.BHT, c.canopy.DHR,
                         c.canopy.DHT, c.canopy.HDR, c.canopy.HDT, c.canopy.BRF, c.ke, s.ref, p.ks, p.kt)])[:, BANDS]
    scal = np.array([np.ravel(v)[0] for v in (c.kt, c.kt_iza, c.kt_vza, c.VollScat.ks, c.VollScat.ko, c.VollScat.bf,
                                               c.VollScat.Fs, c.VollScat.Ft)], dtype=f)
    return np.array(inp, dtype=f), spectra, scal, l8, a


can_in, can_ver, can_lidf, can_spec, can_scal, can_l8, can_as = [], [], [], [], [], [], []


def add(res, ver, lidf_type):
    can_in.append(res[0]); can_spec.append(res[1]); can_scal.append(res[2]); can_l8.append(res[3]); can_as.append(res[4])
    can_ver.append(ver); can_lidf.append(lidf_type)


for i in range(N_CANOPY):
    ver = '5' if (i // 2) % 2 == 0 else 'D'
    lt = 'campbell' if i % 2 == 0 else 'verhoef'
    add(ref_canopy(X[i], ver, lt), ver, lt)

# README example = BASELINE config 1 (examples/prospect_sail.py), as P5 and as D with Can=1
readme = dict(N=f(1.5), Cab=f(35), Cxc=f(5), Cbr=f(0.15), Cw=f(0.003), Cm=f(0.0055))
geom = (f(35), f(30), f(50))
add(ref_canopy(None, '5', 'campbell', leafkw=dict(readme, version='5'), geom=geom, lai=f(3), hot=f(0.25),
               soil=(f(3.14 / 4), f(0.15)), ab=(f(57), f(0))), '5', 'campbell')
add(ref_canopy(None, 'D', 'campbell', leafkw=dict(readme, version='D', Can=f(1.0)), geom=geom, lai=f(3), hot=f(0.25),
               soil=(f(3.14 / 4), f(0.15)), ab=(f(57), f(0))), 'D', 'campbell')
# the reference's own PROSAIL test case (tests/test_prospect_prosail.py:81-86)
tcase = dict(N=f(1.5), Cab=f(40), Cxc=f(8), Cbr=f(0), Cw=f(0.01), Cm=f(0.009), version='5')
add(ref_canopy(None, '5', 'verhoef', leafkw=tcase, geom=(f(30), f(10), f(0)), lai=f(3), hot=f(0.01),
               soil=(f(1), f(1)), ab=(f(-0.35), f(-0.15))), '5', 'verhoef')
# edge cases: lai<=0 (bare soil), hotspot<=0 (alf=1e36), exact hotspot geometry (alf==0), negative vza flip,
# verhoef a>1 shortcut, nadir/nadir, large zenith
edge = [
    dict(geom=(f(30), f(20), f(10)), lai=f(0.0), hot=f(0.1), lt='campbell', ab=(f(57), f(0))),
    dict(geom=(f(30), f(20), f(10)), lai=f(2.0), hot=f(0.0), lt='campbell', ab=(f(57), f(0))),
    dict(geom=(f(30), f(30), f(0)), lai=f(2.5), hot=f(0.2), lt='campbell', ab=(f(45), f(0))),
    dict(geom=(f(40), f(-25), f(30)), lai=f(4.0), hot=f(0.05), lt='verhoef', ab=(f(-0.35), f(-0.15))),
    dict(geom=(f(20), f(35), f(120)), lai=f(1.0), hot=f(0.3), lt='verhoef', ab=(f(1.5), f(0.0))),
    dict(geom=(f(0), f(0), f(0)), lai=f(5.0), hot=f(0.1), lt='campbell', ab=(f(30), f(0))),
    dict(geom=(f(75), f(70), f(180)), lai=f(7.5), hot=f(0.02), lt='verhoef', ab=(f(0.0), f(0.0))),
    dict(geom=(f(10), f(50), f(90)), lai=f(0.05), hot=f(0.5), lt='campbell', ab=(f(80), f(0))),
]
for e in edge:
    add(ref_canopy(None, '5', e['lt'], leafkw=tcase, geom=e['geom'], lai=e['lai'], hot=e['hot'],
                   soil=(f(0.8), f(0.4)), ab=e['ab']), '5', e['lt'])

out.update(can_in=np.array(can_in), can_version=np.array(can_ver), can_lidf=np.array(can_lidf),
           can_spectra=np.array(can_spec), can_scalars=np.array(can_scal), can_L8=np.array(can_l8),
           can_ASTER=np.array(can_as))
out['can_in_cols'] = np.array("N Cab Cxc Cbr Cw Cm Can lai hotspot lidf_a lidf_b soil_refl soil_moist iza vza raa".split())
out['can_spectra_rows'] = np.array("rsot rddt rsdt rdot rdd tdd rsd tsd rdo tdo rso ke rho ks kt".split())
out['can_scalars_cols'] = np.array("tsstoo tss too vol_ks vol_ko bf Fs Ft".split())

# LIDF / VolScatt at the reference's own KAT points and a few more
lidf_cases = []
for a_ in (0.0, 20.0, 57.0, 80.0, 33.3):
    r = pyrism.LIDF.campbell(f(a_), 18)
    track('lidf.campbell', r, orc.lidf_campbell(f(a_)))
    lidf_cases.append(r)
out['lidf_campbell_a'] = np.array([0.0, 20.0, 57.0, 80.0, 33.3])
out['lidf_campbell'] = np.array(lidf_cases)
vab = [(0.0, 0.0), (-0.35, -0.15), (1.0, 0.0), (-1.0, 0.0), (0.0, -1.0), (0.0, 1.0), (0.4, -0.3), (1.5, 0.0)]
lidf_cases = []
for a_, b_ in vab:
    r = pyrism.LIDF.verhoef(f(a_), f(b_), 18)
    track('lidf.verhoef', r, orc.lidf_verhoef(f(a_), f(b_)))
    lidf_cases.append(r)
out['lidf_verhoef_ab'] = np.array(vab)
out['lidf_verhoef'] = np.array(lidf_cases)

np.savez_compressed(os.path.join(HERE, "ref_cases.npz"), **out)

# ------------------------------------------------------------------ the reference's own golden files
tdir = os.path.join(REF, "tests", "data")
fx = dict(prospect5_spectrum=np.loadtxt(os.path.join(tdir, "prospect5_spectrum.txt")),
          prospect_d_LRT=loadmat(os.path.join(tdir, "prospect_d_test.mat"))['LRT'],
          REFL_CAN=np.loadtxt(os.path.join(tdir, "REFL_CAN.txt")),
          tav_alpha40=loadmat(os.path.join(tdir, "tav_alpha40.mat"))['tav'])
np.savez_compressed(os.path.join(HERE, "ref_fixtures.npz"), **fx)

print("max |oracle - live reference| per quantity (the oracle pin):")
for k in sorted(worst):
    print("  %-18s %.3e" % (k, worst[k]))
print("numpy", np.__version__, "scipy", scipy.__version__)
for n in ("ref_cases.npz", "ref_fixtures.npz"):
    print(n, os.path.getsize(os.path.join(HERE, n)), "bytes")
