"""fp32 mode of the CUDA path (``dtype=np.float32`` / the ``prosail_*_f32`` entry points of the C ABI) against the float64
CPU oracle.  Tolerance: 1e-5 absolute in reflectance / transmittance (BASELINE.json north star, fp32 mode)."""
import numpy as np
import pytest

import pyrism_b200 as pb
from pyrism_b200 import _capi as capi
from oracle import prosail_oracle as orc

pytestmark = pytest.mark.gpu
TOL32 = 1e-5
f32 = np.float32
rng = np.random.default_rng(20261019 + 32)


def lhs(n):
    return dict(N=rng.uniform(1, 3, n), Cab=rng.uniform(5, 90, n), Cxc=rng.uniform(1, 25, n), Cbr=rng.uniform(0, 1, n),
                Cw=rng.uniform(0.001, 0.05, n), Cm=rng.uniform(0.001, 0.02, n))


@pytest.mark.parametrize("version", ['5', 'D'])
def test_prospect_f32_vs_oracle(version):
    n = 200
    P = lhs(n)
    Can = rng.uniform(0.1, 8, n)
    p = pb.PROSPECT(version=version, Can=Can if version == 'D' else 0, dtype=f32, **P)
    assert p.ks.shape == (n, 2101) and p.ks.dtype == f32 and p.kt.dtype == f32
    worst = 0.0
    for i in range(n):
        o = orc.prospect(P['N'][i], P['Cab'][i], P['Cxc'][i], P['Cbr'][i], P['Cw'][i], P['Cm'][i], Can=Can[i] if version == 'D' else 0,
                         version=version)
        worst = max(worst, np.max(np.abs(p.ks[i] - o['ks'])), np.max(np.abs(p.kt[i] - o['kt'])))
        got = np.array([getattr(p.L8, b).ks[i] for b in ('B2', 'B3', 'B4', 'B5', 'B6', 'B7')])
        assert np.max(np.abs(got - np.array([o['L8'][b][0] for b in ('B2', 'B3', 'B4', 'B5', 'B6', 'B7')]))) <= TOL32
    assert worst <= TOL32, worst


def test_prospect_f32_scalar_call_and_extremes():
    p = pb.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055, dtype=f32)
    o = orc.prospect(np.float64(1.5), np.float64(35), np.float64(5), np.float64(0.15), np.float64(0.003), np.float64(0.0055))
    assert p.ks.shape == (2101,) and p.ks.dtype == f32
    assert np.max(np.abs(p.ks - o['ks'])) <= TOL32 and np.max(np.abs(p.kt - o['kt'])) <= TOL32
    # extreme absorption (tau table limits / rare path) and almost no absorption
    for kw in (dict(N=3.0, Cab=300., Cxc=80., Cbr=5., Cw=0.4, Cm=0.2), dict(N=1.0, Cab=0.01, Cxc=0.0, Cbr=0.0, Cw=1e-6, Cm=1e-6),
               dict(N=2.5, Cab=0., Cxc=0., Cbr=0., Cw=0., Cm=0.)):
        p = pb.PROSPECT(dtype=f32, **kw)
        o = orc.prospect(**{k: np.float64(v) for k, v in kw.items()})
        assert np.all(np.isfinite(p.ks)) and np.all(np.isfinite(p.kt))
        assert np.max(np.abs(p.ks - o['ks'])) <= TOL32 and np.max(np.abs(p.kt - o['kt'])) <= TOL32, kw


@pytest.mark.parametrize("lidf_type", ['campbell', 'verhoef'])
def test_prosail_f32_multi_geometry_vs_oracle(lidf_type):
    n = 64
    P = lhs(n)
    lai, hot = rng.uniform(0.1, 8, n), rng.uniform(0.01, 0.5, n)
    sr, sm = rng.uniform(0.3, 1.2, n), rng.uniform(0, 1, n)
    a = rng.uniform(20, 80, n) if lidf_type == 'campbell' else rng.uniform(-0.6, 0.6, n)
    b = np.zeros(n) if lidf_type == 'campbell' else rng.uniform(-0.35, 0.35, n)
    iza, vza, raa = np.array([35.0, 10.0, 60.0, 0.0]), np.array([30.0, 45.0, 0.0, 0.0]), np.array([50.0, 120.0, 180.0, 0.0])
    m = pb.PROSAIL(lai=lai, hotspot=hot, iza=iza, vza=vza, raa=raa, soil_reflectance=sr, soil_moisture=sm, lidf_type=lidf_type, a=a, b=b,
                   keep_leaf=True, dtype=f32, **P)
    assert m.BRF.ref.shape == (n, 4, 2101) and m.BRF.ref.dtype == f32
    worst = {}
    for i in range(n):
        for g in range(4):
            _, _, o = orc.prosail(P['N'][i], P['Cab'][i], P['Cxc'][i], P['Cbr'][i], P['Cw'][i], P['Cm'][i], lai[i], hot[i], iza[g], vza[g], raa[g],
                                  sr[i], sm[i], lidf_type=lidf_type, a=a[i], b=b[i])
            pairs = dict(rsot=m.BRF.ref, rddt=m.BHR.ref, rsdt=m.DHR.ref, rdot=m.HDR.ref, rdd=m.canopy.BHR, tdd=m.canopy.BHT, rsd=m.canopy.DHR,
                         tsd=m.canopy.DHT, rdo=m.canopy.HDR, tdo=m.canopy.HDT, rso=m.canopy.BRF, ke=m.ke)
            for k, arr in pairs.items():
                worst[k] = max(worst.get(k, 0.0), float(np.max(np.abs(arr[i, g] - o[k]))))
            got = np.array([getattr(m.BRF.L8, bnd)[i, g] for bnd in ('B2', 'B3', 'B4', 'B5', 'B6', 'B7')])
            assert np.max(np.abs(got - np.array([o['L8']['BRF'][bnd] for bnd in ('B2', 'B3', 'B4', 'B5', 'B6', 'B7')]))) <= TOL32
    assert max(worst.values()) <= TOL32, worst


def test_sail_f32_spectra_from_memory_and_edge_cases():
    leaf = pb.PROSPECT(N=1.5, Cab=40, Cxc=8., Cbr=0.0, Cw=0.01, Cm=0.009)
    soil = pb.LSM(1.0, 0.5)
    cases = [dict(iza=30, vza=10, raa=30, lai=3.0, hotspot=0.01, lidf_type='verhoef', a=-0.35, b=-0.15),     # the reference's own PROSAIL test
             dict(iza=30, vza=10, raa=30, lai=0.0, hotspot=0.1),                                             # bare soil, models.py:609-634
             dict(iza=30, vza=10, raa=30, lai=2.0, hotspot=0.0),                                             # no hotspot, models.py:687-694
             dict(iza=30, vza=30, raa=0, lai=2.0, hotspot=0.2),                                              # exact hotspot, models.py:696-699
             dict(iza=48, vza=10, raa=30, lai=0.05, hotspot=0.05), dict(iza=30, vza=10, raa=30, lai=0.02, hotspot=0.05),   # J1 around k = m
             dict(iza=75, vza=70, raa=10, lai=8.0, hotspot=0.5), dict(iza=0, vza=0, raa=0, lai=1.0, hotspot=0.3)]
    for kw in cases:
        c = pb.SAIL(ks=leaf.ks, kt=leaf.kt, rho_surface=soil.ref, dtype=f32, **kw)
        o = orc.sail(kw['iza'], kw['vza'], kw['raa'], leaf.ks, leaf.kt, kw['lai'], kw['hotspot'], soil.ref, lidf_type=kw.get('lidf_type', 'campbell'),
                     a=kw.get('a', 57), b=kw.get('b', 0))
        assert c.BRF.ref.dtype == f32
        for got, key in ((c.BRF.ref, 'rsot'), (c.BHR.ref, 'rddt'), (c.DHR.ref, 'rsdt'), (c.HDR.ref, 'rdot'), (c.canopy.DHT, 'tsd')):
            assert np.max(np.abs(got - o[key])) <= TOL32, (kw, key, np.max(np.abs(got - o[key])))


def test_f32_lut_matches_the_f64_path_at_lut_size(engine):
    """Size-independent check at LUT size: the fp64 kernels are pinned to the oracle at 1e-9, so at 50 000 parameter sets the
    fp32 mode is compared with them (1e-5), for the BRF-only, four-plane and windows-only instances."""
    n = 50000
    P = lhs(n)
    args = (P['N'], P['Cab'], P['Cxc'], P['Cbr'], P['Cw'], P['Cm'], rng.uniform(0.1, 8, n), rng.uniform(0.01, 0.5, n), np.radians([35.0]),
            np.radians([30.0]), np.radians([50.0]), rng.uniform(0.3, 1.2, n), rng.uniform(0, 1, n))
    kw = dict(lidf_type='campbell', a=rng.uniform(20, 80, n))
    planes = (capi.O_RSOT, capi.O_RDDT, capi.O_RSDT, capi.O_RDOT)
    d64 = engine.prosail(*args, planes=planes, mean_planes=(capi.O_RSOT,), **kw)
    d32 = engine.prosail(*args, planes=planes, dtype=f32, **kw)
    brf = engine.prosail(*args, planes=(capi.O_RSOT,), dtype=f32, **kw)
    wo = engine.prosail(*args, mean_planes=(capi.O_RSOT,), windows_only=True, dtype=f32, **kw)
    for k in ('rsot', 'rddt', 'rsdt', 'rdot'):
        assert d32[k].dtype == f32 and np.all(np.isfinite(d32[k]))
        assert np.max(np.abs(d32[k] - d64[k])) <= TOL32, (k, np.max(np.abs(d32[k] - d64[k])))
    assert np.array_equal(brf['rsot'], d32['rsot'])                    # specialised instances: same arithmetic
    assert wo['means'].dtype == f32 and np.max(np.abs(wo['means'] - d64['means'])) <= TOL32


def test_f32_entry_points_reject_bad_arguments(engine):
    lib = engine.lib
    assert lib.prosail_prosail_f32(engine.h, 0, 1, None, None, None, None, 1, None) == -1      # PROSAIL_E_ARG: null structs
    with pytest.raises(ValueError):
        engine.prosail(1.5, 40, 8, 0, 0.01, 0.009, 3.0, 0.1, [0.5], [0.2], [0.3], 1.0, 0.5, dtype=np.float16)


def test_prospect_f32_derived_planes_come_from_the_kernel():
    """ka, ke, om (models.py:1066-1068) are output planes of the PROSPECT kernel in both precisions."""
    n = 50
    P = lhs(n)
    for dt, tol in ((np.float64, 1e-9), (f32, TOL32)):
        p = pb.PROSPECT(dtype=dt, **P)
        assert p.ka.dtype == dt and p.ka.shape == (n, 2101)
        for i in (0, 17, n - 1):
            o = orc.prospect(P['N'][i], P['Cab'][i], P['Cxc'][i], P['Cbr'][i], P['Cw'][i], P['Cm'][i])
            for q in ('ka', 'ke', 'om'):
                assert np.max(np.abs(getattr(p, q)[i] - o[q])) <= tol, (q, i, dt)
        assert p.int.shape == (n, 2101, 6) and p.int.dtype == f32


@pytest.mark.parametrize("over", [dict(lai=300.0), dict(iza=89.9), dict(lai=50.0, hotspot=1e-9), dict(vza=89.0, lai=8.0), dict(lai=1e-6)])
def test_f32_underflowing_transmittances_stay_finite(over):
    """tss = exp(-k lai) and e1 underflow in float for lai of hundreds or a grazing sun: no 0 * inf may appear (the J1 term is
    taken as a direct difference there), and the result stays within the fp32 tolerance of the float64 oracle."""
    p = dict(N=1.5, Cab=40.0, Cxc=8.0, Cbr=0.1, Cw=0.01, Cm=0.009, lai=3.0, hotspot=0.1, iza=35.0, vza=30.0, raa=50.0)
    p.update(over)
    _, _, o = orc.prosail(p['N'], p['Cab'], p['Cxc'], p['Cbr'], p['Cw'], p['Cm'], p['lai'], p['hotspot'], p['iza'], p['vza'], p['raa'], 1.0, 0.5)
    for dt, tol in ((np.float64, 1e-9), (f32, TOL32)):
        m = pb.PROSAIL(N=p['N'], Cab=p['Cab'], Cxc=p['Cxc'], Cbr=p['Cbr'], Cw=p['Cw'], Cm=p['Cm'], lai=p['lai'], hotspot=p['hotspot'], iza=p['iza'],
                       vza=p['vza'], raa=p['raa'], soil_reflectance=1.0, soil_moisture=0.5, dtype=dt)
        for got, key in ((m.BRF.ref, 'rsot'), (m.BHR.ref, 'rddt'), (m.DHR.ref, 'rsdt'), (m.HDR.ref, 'rdot')):
            assert np.all(np.isfinite(got)) and np.max(np.abs(got - o[key])) <= tol, (over, key, dt)
