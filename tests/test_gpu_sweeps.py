"""Large parity sweeps on the GPU box (the `-m gpu` form of tools/parity_sweep.py and tools/edge_sweep.py).

test_parity_sweep_60000_sets: 60 000 random parameter sets with random geometry (PROSPECT-5 / D x campbell / verhoef) through the
CUDA path, every plane of the path against the CPU oracle evaluated on all host cores (about 70 s on the 16-core box).
Bars: fp64 <= 1e-10 (the round-2 error budget of the kernel, ten times inside the 1e-9 parity bar; measured 6.4e-12),
fp32 mode <= 1e-5 (north star; measured 2.1e-6).

test_edge_of_domain_cases: 30 edge-of-domain inputs (lai 1e-6 .. 5000, grazing angles, hotspot 1e-9 .. 10, LIDF extremes,
pigment-free / opaque leaves, black and very bright soil) through the drop-in PROSAIL class in both precisions.
"""
import multiprocessing as mp
import os

import numpy as np
import pytest

import pyrism_b200 as pb
from oracle import prosail_oracle as orc

pytestmark = pytest.mark.gpu
QUANT = ('rsot', 'rddt', 'rsdt', 'rdot', 'rso', 'rdd', 'tdd', 'rsd', 'tsd', 'rdo', 'tdo', 'ke')
PL = {'rsot': 0, 'rddt': 1, 'rsdt': 2, 'rdot': 3, 'rso': 4, 'rdd': 5, 'tdd': 6, 'rsd': 7, 'tsd': 8, 'rdo': 9, 'tdo': 10, 'ke': 11, 'leaf_r': 12, 'leaf_t': 13,
      'rho': 14}
FP64_BAR, FP32_BAR = 1e-10, 1e-5


def _oracle(args):
    import warnings
    warnings.filterwarnings("ignore")
    rows, version, lidf = args
    out = []
    for r in rows:
        (N, Cab, Cxc, Cbr, Cw, Cm, Can, lai, hot, a, b, sr, sm, iza, vza, raa) = r
        leaf, soil, can = orc.prosail(N, Cab, Cxc, Cbr, Cw, Cm, lai, hot, iza, vza, raa, sr, sm, Can=Can if version == 'D' else 0, version=version,
                                      lidf_type=lidf, a=a, b=b)
        out.append(np.stack([can[q] * np.ones(2101) for q in QUANT] + [leaf['ks'], leaf['kt'], soil['ref']]))
    return np.stack(out)


def test_parity_sweep_60000_sets(engine):
    n = int(os.environ.get("PYRISM_SWEEP_N", "60000"))
    rng = np.random.default_rng(20261019 + 99)
    cores = os.cpu_count() or 1
    pool = mp.get_context("fork").Pool(cores)
    worst64 = worst32 = 0.0
    try:
        for version in ('5', 'D'):
            for lidf in ('campbell', 'verhoef'):
                m = n // 4
                P = np.stack([rng.uniform(1, 3, m), rng.uniform(5, 90, m), rng.uniform(1, 25, m), rng.uniform(0, 1, m), rng.uniform(0.001, 0.05, m),
                              rng.uniform(0.001, 0.02, m), rng.uniform(0.1, 8, m), rng.uniform(0.1, 8, m), rng.uniform(0.01, 0.5, m),
                              rng.uniform(20, 80, m) if lidf == 'campbell' else rng.uniform(-0.6, 0.6, m),
                              np.zeros(m) if lidf == 'campbell' else rng.uniform(-0.35, 0.35, m), rng.uniform(0.3, 1.2, m), rng.uniform(0, 1, m),
                              rng.uniform(0, 70, m), rng.uniform(0, 60, m), rng.uniform(0, 180, m)], axis=1)
                for c0 in range(0, m, 5000):                  # 5000 sets x 15 planes x 2101 doubles = 1.3 GB of oracle output at a time
                    Pc = P[c0:c0 + 5000]
                    ref = np.concatenate(pool.map(_oracle, [(Pc[i::cores], version, lidf) for i in range(cores)]))
                    order = np.concatenate([np.arange(len(Pc))[i::cores] for i in range(cores)])
                    Pc = Pc[order]
                    for dt in (np.float64, np.float32):
                        r = engine.prosail(Pc[:, 0], Pc[:, 1], Pc[:, 2], Pc[:, 3], Pc[:, 4], Pc[:, 5], Pc[:, 7], Pc[:, 8], np.radians(Pc[:, 13]),
                                           np.radians(Pc[:, 14]), np.radians(Pc[:, 15]), Pc[:, 11], Pc[:, 12], Can=Pc[:, 6], version=version, lidf_type=lidf,
                                           a=Pc[:, 9], b=Pc[:, 10], planes=tuple(range(15)), geom_per_sample=True, dtype=dt)
                        w = max(float(np.max(np.abs(r[q][:, 0].astype(np.float64) - ref[:, PL[q]]))) for q in PL)
                        if dt == np.float64:
                            worst64 = max(worst64, w)
                        else:
                            worst32 = max(worst32, w)
                    del ref, r
    finally:
        pool.close()
    line = "parity sweep, %d sets: worst fp64 %.3e (bar %.0e), worst fp32 %.3e (bar %.0e)" % (n, worst64, FP64_BAR, worst32, FP32_BAR)
    print(line)
    if os.path.isdir("gpurun_out"):                   # the measured maxima travel back from the GPU box (copied to profiles/ by hand)
        with open(os.path.join("gpurun_out", "parity_sweep.txt"), "w") as fh:
            fh.write(line + "\n")
    assert worst64 <= FP64_BAR, worst64
    assert worst32 <= FP32_BAR, worst32


BASE = dict(N=1.5, Cab=40.0, Cxc=8.0, Cbr=0.1, Cw=0.01, Cm=0.009, lai=3.0, hotspot=0.1, iza=35.0, vza=30.0, raa=50.0, sr=1.0, sm=0.5,
            lidf_type='campbell', a=57.0, b=0.0)
CASES = [("base", {}), ("lai 50", dict(lai=50.0)), ("lai 1e-6", dict(lai=1e-6)), ("lai 300", dict(lai=300.0)), ("hotspot 1e-9", dict(hotspot=1e-9)),
         ("hotspot 10", dict(hotspot=10.0)), ("iza 89.9", dict(iza=89.9)), ("vza 89", dict(vza=89.0)), ("iza 0 vza 0", dict(iza=0.0, vza=0.0, raa=0.0)),
         ("raa 360", dict(raa=360.0)), ("raa 180 hot", dict(iza=30.0, vza=30.0, raa=180.0)), ("vza -20", dict(vza=-20.0)), ("campbell a 1", dict(a=1.0)),
         ("campbell a 89.9", dict(a=89.9)), ("verhoef -1 0", dict(lidf_type='verhoef', a=-1.0, b=0.0)), ("verhoef .99 0", dict(lidf_type='verhoef', a=0.99, b=0.0)),
         ("verhoef 0 1", dict(lidf_type='verhoef', a=0.0, b=-1.0)), ("N 1", dict(N=1.0)), ("N 6", dict(N=6.0)), ("no pigments", dict(Cab=0.0, Cxc=0.0, Cbr=0.0)),
         ("dry thin leaf", dict(Cw=1e-5, Cm=1e-4)), ("thick leaf", dict(Cw=0.2, Cm=0.1, Cab=200.0)), ("black soil", dict(sr=0.0)),
         ("bright wet soil", dict(sr=2.5, sm=1.0)), ("dry soil", dict(sm=0.0)), ("lai 0", dict(lai=0.0)),
         # round 2n: beyond the underflow floor of the scaled exp (exp(-m lai) < 1e-300), exponents of b^(N-1) of both signs
         ("lai 5000", dict(lai=5000.0)), ("N 0.8", dict(N=0.8)), ("N 1.0001", dict(N=1.0001)), ("weak absorption", dict(Cab=0.4, Cxc=0.08, Cbr=0.001, Cw=1e-4, Cm=9e-5))]


@pytest.mark.parametrize("name,over", CASES, ids=[c[0] for c in CASES])
def test_edge_of_domain_cases(name, over):
    import warnings
    warnings.filterwarnings("ignore")
    p = dict(BASE, **over)
    _, _, o = orc.prosail(p['N'], p['Cab'], p['Cxc'], p['Cbr'], p['Cw'], p['Cm'], p['lai'], p['hotspot'], p['iza'], p['vza'], p['raa'], p['sr'], p['sm'],
                          lidf_type=p['lidf_type'], a=p['a'], b=p['b'])
    for dt, bar in ((np.float64, FP64_BAR), (np.float32, FP32_BAR)):
        m = pb.PROSAIL(N=p['N'], Cab=p['Cab'], Cxc=p['Cxc'], Cbr=p['Cbr'], Cw=p['Cw'], Cm=p['Cm'], lai=p['lai'], hotspot=p['hotspot'], iza=p['iza'],
                       vza=p['vza'], raa=p['raa'], soil_reflectance=p['sr'], soil_moisture=p['sm'], lidf_type=p['lidf_type'], a=p['a'], b=p['b'], dtype=dt)
        for got, key in ((m.BRF.ref, 'rsot'), (m.BHR.ref, 'rddt'), (m.DHR.ref, 'rsdt'), (m.HDR.ref, 'rdot')):
            ref = np.asarray(o[key], dtype=np.float64) * np.ones(2101)
            assert np.array_equal(np.isnan(got), np.isnan(ref)), (name, key)
            d = np.abs(got.astype(np.float64) - ref)
            assert np.nanmax(d) <= bar, (name, key, np.dtype(dt).name, float(np.nanmax(d)))
