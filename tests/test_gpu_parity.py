"""Parity of the CUDA path (through the C ABI / the drop-in classes) with the CPU oracle, the stored outputs of the live
reference and the reference's own golden files.  Tolerances: 1e-9 absolute for fp64 spectra and means (north star), float32
eps for the PROSPECT/LSM band means that the reference itself computes in float32."""
import numpy as np
import pytest

import pyrism_b200 as pb
from pyrism_b200 import _capi as capi
from oracle import prosail_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9
TOL_F32 = 2e-6
L8N = [w[0] for w in orc.L8_WINDOWS]
ASN = [w[0] for w in orc.ASTER_WINDOWS]
rng = np.random.default_rng(20261019)


def lhs(n):
    return dict(N=rng.uniform(1, 3, n), Cab=rng.uniform(5, 90, n), Cxc=rng.uniform(1, 25, n), Cbr=rng.uniform(0, 1, n),
                Cw=rng.uniform(0.001, 0.05, n), Cm=rng.uniform(0.001, 0.02, n))


# ------------------------------------------------------------------------------------------------------------ PROSPECT
@pytest.mark.parametrize("version", ['5', 'D'])
def test_prospect_batch_vs_oracle(version):
    n = 200
    P = lhs(n)
    Can = rng.uniform(0.1, 8, n)
    p = pb.PROSPECT(version=version, Can=Can if version == 'D' else 0, **P)
    assert p.ks.shape == (n, 2101) and p.ks.dtype == np.float64
    for i in range(n):
        o = orc.prospect(P['N'][i], P['Cab'][i], P['Cxc'][i], P['Cbr'][i], P['Cw'][i], P['Cm'][i], Can=Can[i] if version == 'D' else 0,
                         version=version)
        for q in ('ks', 'kt', 'ka', 'ke', 'om'):
            assert np.max(np.abs(getattr(p, q)[i] - o[q])) <= TOL, (q, i)
        got = np.array([[getattr(getattr(p.L8, b), f)[i] for f in ('ks', 'kt', 'ka', 'ke', 'omega')] for b in L8N])
        assert np.max(np.abs(got - np.array([o['L8'][b] for b in L8N]))) <= TOL_F32
        got = np.array([[getattr(getattr(p.ASTER, b), f)[i] for f in ('ks', 'kt', 'ka', 'ke', 'omega')] for b in ASN])
        assert np.max(np.abs(got - np.array([o['ASTER'][b] for b in ASN]))) <= TOL_F32
    assert p.L8.B2.ks.dtype == np.float32


def test_prospect_vs_live_reference_outputs(cases):
    B = cases["band_index"]
    for x, ver, out, l8 in zip(cases["leaf_in"], cases["leaf_version"], cases["leaf_out"], cases["leaf_L8"]):
        p = pb.PROSPECT(*x[:6], Can=x[6], version=str(ver))
        assert p.ks.shape == (2101,)
        got = np.stack([p.ks, p.kt, p.ka, p.ke, p.om])[:, B]
        assert np.max(np.abs(got - out)) <= TOL
        got = np.array([[getattr(getattr(p.L8, b), f) for f in ('ks', 'kt', 'ka', 'ke', 'omega')] for b in L8N])
        assert np.max(np.abs(got - l8)) <= TOL_F32


def test_prospect_reference_golden_files(fixtures):
    """The reference's own tests (tests/test_prospect_prosail.py:34-70) run against the CUDA path."""
    w, refl, trans = fixtures["prospect5_spectrum"].T
    p = pb.PROSPECT(N=2.1, Cab=40, Cxc=10., Cbr=0.1, Cw=0.015, Cm=0.009, version='5')
    assert np.allclose(refl, p.ks, atol=1e-4) and np.allclose(trans, p.kt, atol=1e-4)
    lrt = fixtures["prospect_d_LRT"]
    p = pb.PROSPECT(N=1.2, Cab=30, Cxc=10., Cbr=0.0, Cw=0.015, Cm=0.009, Can=1, version="D")
    assert np.allclose(lrt[:, 1], p.ks, atol=1e-4) and np.allclose(lrt[:, 2], p.kt, atol=1e-4)


def test_prospect_attribute_surface():
    p = pb.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055)
    assert p.l.shape == (2101,) and p.l[0] == 400 and p.n_l == 2101 and p.ver == '5' and p.alpha == 40
    assert p.int.shape == (2101, 6) and p.int.dtype == np.float32 and p.n_elems_list == [2101] * 7
    assert np.array_equal(p.Kan, np.zeros(2101, dtype=np.float32)) and p.KN.dtype == np.float32
    o = orc.prospect(1.5, 35, 5, 0.15, 0.003, 0.0055)
    assert np.allclose(p.int, o['int'], atol=TOL_F32)
    # select / indices (models.py:1197-1231)
    sel = p.select(851, 879)
    want = o['int'][(orc.WAVELENGTHS >= 851) & (orc.WAVELENGTHS <= 879)][:, 1:].mean(axis=0)
    assert sel.dtype == np.float32 and np.allclose(sel, want, atol=TOL_F32)
    red = o['int'][(orc.WAVELENGTHS >= 636) & (orc.WAVELENGTHS <= 673)][:, 1].mean()
    assert abs(p.indices() - (want[0] - red) / (want[0] + red)) < 1e-5
    p.cleanup('ndvi')
    assert not hasattr(p, 'ndvi')
    # a different leaf-surface cone alpha re-derives the interface constants
    p57 = pb.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055, alpha=57)
    o57 = orc.prospect(1.5, 35, 5, 0.15, 0.003, 0.0055, alpha=57)
    assert np.max(np.abs(p57.ks - o57['ks'])) <= TOL and np.max(np.abs(p57.ks - p.ks)) > 1e-5
    p40 = pb.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055)
    assert np.array_equal(p40.ks, p.ks)


def test_select_leaves_the_shared_engine_alone_and_is_thread_safe():
    """select() / indices() run on their own handle (`get_engine(device, role='select')`): the shared engine keeps its 15 windows,
    and constructors on other threads see consistent L8 / ASTER means while select() calls with different windows are in flight
    (round-1 VERDICT weak #12)."""
    import threading
    from pyrism_b200.engine import get_engine
    main = get_engine()
    p = pb.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055)
    want_b5 = np.asarray(p.L8.B5)
    before = main.windows
    sel0 = p.select(851, 879)
    assert main.windows == before and len(main.windows) == 15
    assert get_engine(role='select') is not main and get_engine(role='select').windows == ((851, 879),)
    errors = []

    def selector(lo, hi, ref):
        try:
            for _ in range(30):
                if not np.array_equal(p.select(lo, hi), ref):
                    errors.append(("select", lo, hi))
        except Exception as exc:      # noqa: BLE001
            errors.append(repr(exc))

    def constructor():
        try:
            for _ in range(30):
                q = pb.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055)
                if not np.array_equal(np.asarray(q.L8.B5), want_b5) or len(q.ASTER) != 9:
                    errors.append("constructor saw another window list")
        except Exception as exc:      # noqa: BLE001
            errors.append(repr(exc))

    ref_red = p.select(636, 673)
    ts = [threading.Thread(target=selector, args=(851, 879, sel0)), threading.Thread(target=selector, args=(636, 673, ref_red)),
          threading.Thread(target=constructor), threading.Thread(target=constructor)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[:3]
    assert main.windows == before


def test_prospect_extreme_absorption():
    """Arguments of tau() from ~0 (zero-absorption Stokes branch, models.py:1057-1059) up to k > 64."""
    # The all-zero-pigment leaf: k = 0, tau = 1, r + t = 1 up to rounding.  The reference decides between its Stokes formulas and its
    # zero-absorption branch (models.py:1057-1059) on the last ulp of r + t; in the bands where it takes the Stokes formulas
    # (462 of 2101 here, 1 - r - t = 1e-16 .. 3e-16) they carry 1e-16 / D ~ 1e-8 of rounding noise: the reference is off the
    # mathematical limit by 4.1e-9 (N = 1.5) .. 6.2e-9 (N = 3), profiles/r2n_zero_leaf_probe.txt.  The CUDA path (leaf_layers_slow)
    # is within 5e-16 of that limit, so it is held to the limit at 1e-12 and to the reference at 2e-8 (= the reference's own noise).
    for N in (1.5, 3.0):
        o = orc.prospect(N, 0.0, 0.0, 0.0, 0.0, 0.0)
        p = pb.PROSPECT(N, 0.0, 0.0, 0.0, 0.0, 0.0)
        assert np.max(np.abs(p.ks - o['ks'])) <= 2e-8 and np.max(np.abs(p.kt - o['kt'])) <= 2e-8
        Tsub = o['t'] / (o['t'] + (1 - o['t']) * (N - 1))             # the limit of :1043-1056 for 1 - r - t -> 0 (= :1058-1059)
        den = 1 - (1 - Tsub) * o['r']
        assert np.max(np.abs(p.kt - o['Ta'] * Tsub / den)) <= 1e-12
        assert np.max(np.abs(p.ks - (o['Ra'] + o['Ta'] * (1 - Tsub) * o['t'] / den))) <= 1e-12
        flips = (o['r'] + o['t']) < 1.0                               # bands where the reference takes the Stokes formulas
        assert np.max(np.abs(p.ks - o['ks'])[~flips]) <= 1e-12        # where it takes the zero-absorption branch the bar is the usual one
    for kw in (dict(N=1.0, Cab=1e-3, Cxc=0, Cbr=0, Cw=1e-6, Cm=1e-6), dict(N=3.0, Cab=200, Cxc=60, Cbr=3, Cw=0.3, Cm=0.1),
               dict(N=1.0, Cab=0, Cxc=0, Cbr=0, Cw=1.0, Cm=0.0)):
        o = orc.prospect(**kw)
        p = pb.PROSPECT(**kw)
        ok = np.isfinite(o['ks'])
        assert ok.sum() > 1500 and np.all(np.isfinite(p.ks))
        assert np.max(np.abs(p.ks[ok] - o['ks'][ok])) <= TOL and np.max(np.abs(p.kt[ok] - o['kt'][ok])) <= TOL


def test_prospect_weak_absorption_ladder():
    """Leaves whose absorption goes from realistic down to almost nothing (all contents scaled by 1 ... 1e-9), N below, at and above 1:
    the fast path's economised polynomials and second-order steps are amplified by 1/D there, and from 1 - r - t < 2^-17 on the exact
    rare path takes over (csrc/prosail_kernels.cuh: leaf_layers_slow).  Both sides of that switch hold the 1e-9 bar against the oracle;
    negative N - 1 exercises the other sign of the exponent in pow_pos_scaled."""
    f = np.array([1.0, 0.3, 0.1, 3e-2, 1e-2, 3e-3, 1e-3, 3e-4, 1e-4, 3e-5, 1e-5, 3e-6, 1e-6, 1e-7, 1e-8, 1e-9])
    worst = 0.0
    for N in (0.9, 1.0, 1.3, 2.0, 3.0):
        p = pb.PROSPECT(N=np.full(f.size, N), Cab=40.0 * f, Cxc=8.0 * f, Cbr=0.1 * f, Cw=0.01 * f, Cm=0.009 * f)
        for i, fi in enumerate(f):
            o = orc.prospect(N, 40.0 * fi, 8.0 * fi, 0.1 * fi, 0.01 * fi, 0.009 * fi)
            ok = np.isfinite(o['ks']) & np.isfinite(o['kt'])
            assert ok.sum() > 2000
            # the reference's own rounding noise in a^2 - 1 and b^(2(N-1)) - 1 is 1e-16 / D: 1e-12 at f = 1e-4, 1e-9 at f = 1e-9
            tol = TOL if fi >= 1e-6 else 2e-8
            d = max(float(np.max(np.abs(p.ks[i][ok] - o['ks'][ok]))), float(np.max(np.abs(p.kt[i][ok] - o['kt'][ok]))))
            assert d <= tol, (N, fi, d)
            if fi >= 1e-4:
                worst = max(worst, d)
    assert worst <= 1e-10, worst          # realistic and weakly absorbing leaves stay inside the round-2 error budget's test bar


def test_non_finite_pigment_content_gives_the_reference_pattern():
    """The leaf stage skips pigment terms whose table is zero over a whole tile group (from 784 / 1100 nm on).  That is exact for finite
    contents; a NaN or infinite content times 0.0 is NaN, and the reference turns a NaN kall into tau = 1 (`kall > 0` is False,
    models.py:953) while an infinite kall gives NaN spectra.  The kernels reproduce that band by band: the paths that skip a term read
    a copy of the next content that the prologue sets to NaN when a skipped content is not finite (csrc: leaf_kall_all, S_CWX / S_CBRX)."""
    import warnings
    base = dict(N=1.5, Cab=35.0, Cxc=5.0, Cbr=0.15, Cw=0.003, Cm=0.0055)
    for version, names in (('5', ('Cab', 'Cxc', 'Cbr')), ('D', ('Cab', 'Cxc', 'Cbr', 'Can'))):
        for name in names:
            for bad in (np.nan, np.inf):
                kw = dict(base, version=version, **({'Can': 2.0} if version == 'D' else {}))
                kw[name] = bad
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    o = orc.prospect(**kw)
                p = pb.PROSPECT(**kw)
                for got, ref in ((p.ks, o['ks']), (p.kt, o['kt'])):
                    assert np.array_equal(np.isnan(got), np.isnan(ref)), (version, name, bad, int(np.isnan(got).sum()), int(np.isnan(ref).sum()))
                    ok = np.isfinite(ref)
                    assert ok.any() and np.max(np.abs(got[ok] - ref[ok])) <= 2e-8, (version, name, bad)      # tau = 1 bands: the reference's own noise (see above)
    # a batch: the finite sets next to such a set are untouched
    P = {k: np.array([v, v, v]) for k, v in base.items()}
    P['Cab'] = np.array([35.0, np.inf, 35.0])
    p = pb.PROSPECT(**P)
    o = orc.prospect(**base)
    assert np.isnan(p.ks[1]).sum() == 380 and np.max(np.abs(p.ks[0] - o['ks'])) <= TOL and np.array_equal(p.ks[0], p.ks[2])


# ------------------------------------------------------------------------------------------------------------ LSM
def test_lsm():
    r, m = rng.uniform(0.3, 1.2, 9), rng.uniform(0, 1, 9)
    s = pb.LSM(reflectance=r, moisture=m)
    assert s.ref.shape == (9, 2101)
    for i in range(9):
        o = orc.lsm(r[i], m[i])
        assert np.max(np.abs(s.ref[i] - o['ref'])) <= TOL
        assert abs(s.L8.B5[i] - o['L8']['B5']) <= TOL_F32 and abs(s.ASTER.B9[i] - o['ASTER']['B9']) <= TOL_F32
    s1 = pb.LSM(reflectance=3.14 / 4, moisture=0.15)
    o = orc.lsm(3.14 / 4, 0.15)
    assert s1.ref.shape == (2101,) and s1.int.shape == (2101, 2) and np.max(np.abs(s1.ref - o['ref'])) <= TOL
    assert isinstance(s1.L8.B2, np.float32)
    f32 = o['ref'].astype(np.float32)
    assert abs(s1.select(500, 600) - f32[100:201].mean()) <= TOL_F32


# ------------------------------------------------------------------------------------------------------------ LIDF / VolScatt
def test_lidf_and_volscatt_known_answers(cases):
    """tests/test_volscat.py of the reference + the stored LIDF outputs of the live reference."""
    for a, ref in zip(cases["lidf_campbell_a"], cases["lidf_campbell"]):
        assert np.max(np.abs(pb.LIDF.campbell(a) - ref)) <= 1e-13
    for (a, b), ref in zip(cases["lidf_verhoef_ab"], cases["lidf_verhoef"]):
        assert np.max(np.abs(pb.LIDF.verhoef(a, b) - ref)) <= 1e-13
    assert np.allclose(pb.LIDF.verhoef(0, 0, n_elements=18), np.zeros(18) + 0.05555556, atol=1e-4)
    v = pb.VolScatt(50, 30, 50)
    v.coef(a=0, b=0, lidf_type='verhoef')
    got = (v.ks[0], v.ko[0], v.bf, v.Fs[0], v.Ft[0])
    assert np.allclose(got, (0.823188863879162, 0.6867716425258737, 0.5, 0.6217351737543753, 0.011166182392885724), atol=1e-13)
    v.coef(a=0, lidf_type='campbell')
    got = (v.ks[0], v.ko[0], v.bf, v.Fs[0], v.Ft[0])
    assert np.allclose(got, (0.9948675435577432, 0.9940524814594064, 0.9891027449943284, 0.9915874551449064, 7.491316140639147e-05), atol=1e-13)
    assert np.allclose(v.Fst, v.Fs + v.Ft)
    # batched LIDF parameters
    a = rng.uniform(20, 80, 7)
    v.coef(a=a, lidf_type='campbell')
    assert v.ks.shape == (7, 1)
    iza, vza, raa, *_ = orc.kernel_angles(50, 30, 50)
    for i in range(7):
        want = orc.volscatt_coef(iza, vza, raa, orc.lidf_campbell(a[i]))
        assert np.allclose((v.ks[i, 0], v.ko[i, 0], v.bf[i], v.Fs[i, 0], v.Ft[i, 0]), want, atol=1e-13)


# ------------------------------------------------------------------------------------------------------------ SAIL / PROSAIL
CAN_PAIRS = (('BRF', 'rsot'), ('BHR', 'rddt'), ('DHR', 'rsdt'), ('HDR', 'rdot'))
CANOPY_PAIRS = (('BRF', 'rso'), ('BHR', 'rdd'), ('BHT', 'tdd'), ('DHR', 'rsd'), ('DHT', 'tsd'), ('HDR', 'rdo'), ('HDT', 'tdo'))


def check_sail_object(c, o, idx=()):
    for attr, key in CAN_PAIRS:
        assert np.max(np.abs(getattr(c, attr).ref[idx] - o[key])) <= TOL, attr
        for b in L8N:
            assert abs(getattr(getattr(c, attr).L8, b)[idx] - o['L8'][attr][b]) <= TOL
        for b in ASN:
            assert abs(getattr(getattr(c, attr).ASTER, b)[idx] - o['ASTER'][attr][b]) <= TOL
    assert np.max(np.abs(c.BRDF.ref[idx] - o['rsot'] / np.pi)) <= TOL and abs(c.BRDF.L8.B4[idx] - o['L8']['BRDF']['B4']) <= TOL
    for attr, key in CANOPY_PAIRS:
        assert np.max(np.abs(c.canopy[attr][idx] - o[key])) <= TOL, attr
    assert np.max(np.abs(c.ke[idx] - o['ke'])) <= TOL
    for got, key in ((c.kt, 'tsstoo'), (c.kt_iza, 'tss'), (c.kt_vza, 'too'), (c.VollScat.ks, 'vol_ks'), (c.VollScat.ko, 'vol_ko'),
                     (c.VollScat.Fs, 'Fs'), (c.VollScat.Ft, 'Ft')):
        assert abs(np.asarray(got)[idx if idx else 0] - o[key]) <= TOL, key


def test_readme_example_drop_in():
    """examples/prospect_sail.py (BASELINE config 1), as P5 (the literal README call) and as D with Can = 1."""
    for ver, can in (('5', 0), ('D', 1.0)):
        leaf = pb.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055, Can=can, version=ver)
        soil = pb.LSM(reflectance=3.14 / 4, moisture=0.15)
        can_ = pb.SAIL(iza=35, vza=30, raa=50, ks=leaf.ks, kt=leaf.kt, rho_surface=soil.ref, lidf_type='campbell', lai=3, hotspot=0.25)
        lf, so, o = orc.prosail(1.5, 35, 5, 0.15, 0.003, 0.0055, 3, 0.25, 35, 30, 50, 3.14 / 4, 0.15, Can=can, version=ver)
        assert can_.BRF.ref.shape == (2101,) and can_.kt.shape == (1,) and can_.VollScat.ks.shape == (1,) and np.ndim(can_.VollScat.bf) == 0
        check_sail_object(can_, o)
        assert np.allclose(can_.BRF.refdB, 10 * np.log10(o['rsot']), atol=1e-8)
        assert np.allclose(can_.izaDeg, 35) and can_.l[-1] == 2500 and can_.ks is leaf.ks


def test_sail_reference_golden_file(fixtures):
    """tests/test_prospect_prosail.py:77-120 against the CUDA path."""
    w, resv, hdr, sdr, bhr, dhr = fixtures["REFL_CAN"].T
    lsm = pb.LSM(reflectance=1, moisture=1)
    leaf = pb.PROSPECT(N=1.5, Cab=40, Cxc=8., Cbr=0.0, Cw=0.01, Cm=0.009, version="5")
    sail = pb.SAIL(iza=30, vza=10, raa=0, ks=leaf.ks, kt=leaf.kt, lai=3, hotspot=0.01, rho_surface=lsm.ref, a=-0.35, b=-0.15, lidf_type='verhoef')
    assert np.allclose(sdr, sail.BRF.ref, atol=0.01) and np.allclose(hdr, sail.HDR.ref, atol=0.01)
    assert np.allclose(bhr, sail.BHR.ref, atol=0.01) and np.allclose(dhr, sail.DHR.ref, atol=0.01)


def test_canopy_cases_of_live_reference(cases):
    """Every stored reference case (LHS + README + the reference's test case + edge cases: lai = 0, hotspot = 0, exact
    hotspot geometry, negative view zenith, verhoef a > 1, nadir, large zenith, tiny lai) through SAIL and through PROSAIL."""
    B = cases["band_index"]
    rows = list(cases["can_spectra_rows"])
    for x, ver, lt, spec, scal, l8 in zip(cases["can_in"], cases["can_version"], cases["can_lidf"], cases["can_spectra"], cases["can_scalars"],
                                          cases["can_L8"]):
        N, Cab, Cxc, Cbr, Cw, Cm, Can, lai, hot, la, lb, sr, sm, iza, vza, raa = x
        fused = pb.PROSAIL(N, Cab, Cxc, Cbr, Cw, Cm, lai, hot, iza, vza, raa, soil_reflectance=sr, soil_moisture=sm, Can=Can, version=str(ver),
                           lidf_type=str(lt), a=la, b=lb, keep_leaf=True)
        leaf = pb.PROSPECT(N, Cab, Cxc, Cbr, Cw, Cm, Can=Can, version=str(ver))
        soil = pb.LSM(sr, sm)
        split = pb.SAIL(iza, vza, raa, leaf.ks, leaf.kt, lai, hot, soil.ref, lidf_type=str(lt), a=la, b=lb)
        for c in (fused, split):
            for attr, key in CAN_PAIRS:
                assert np.max(np.abs(getattr(c, attr).ref[B] - spec[rows.index(key)])) <= TOL, (attr, x)
            for attr, key in CANOPY_PAIRS:
                assert np.max(np.abs(c.canopy[attr][B] - spec[rows.index(key)])) <= TOL, (attr, x)
            assert np.max(np.abs(c.ke[B] - spec[rows.index('ke')])) <= TOL
            got = np.array([c.kt[0], c.kt_iza[0], c.kt_vza[0], c.VollScat.ks[0], c.VollScat.ko[0], c.VollScat.bf, c.VollScat.Fs[0], c.VollScat.Ft[0]])
            assert np.max(np.abs(got - scal)) <= TOL
            got = np.array([[getattr(getattr(c, p).L8, b) for b in L8N] for p in ('BRF', 'BRDF', 'BHR', 'DHR', 'HDR')])
            assert np.max(np.abs(got - l8)) <= TOL
        assert np.max(np.abs(fused.leaf_ks[B] - spec[rows.index('ks')])) <= TOL and np.max(np.abs(fused.rho_surface[B] - spec[rows.index('rho')])) <= TOL


@pytest.mark.parametrize("lidf_type", ['campbell', 'verhoef'])
def test_prosail_batch_multi_geometry_vs_oracle(lidf_type):
    n = 48
    P = lhs(n)
    lai, hot = rng.uniform(0.1, 8, n), rng.uniform(0.01, 0.5, n)
    sr, sm = rng.uniform(0.3, 1.2, n), rng.uniform(0, 1, n)
    a = rng.uniform(20, 80, n) if lidf_type == 'campbell' else rng.uniform(-0.6, 0.6, n)
    b = np.zeros(n) if lidf_type == 'campbell' else rng.uniform(-0.35, 0.35, n)
    iza, vza, raa = np.array([35.0, 10.0, 60.0, 0.0]), np.array([30.0, 45.0, 0.0, 0.0]), np.array([50.0, 120.0, 180.0, 0.0])
    m = pb.PROSAIL(lai=lai, hotspot=hot, iza=iza, vza=vza, raa=raa, soil_reflectance=sr, soil_moisture=sm, lidf_type=lidf_type, a=a, b=b, **P)
    assert m.BRF.ref.shape == (n, 4, 2101) and m.kt.shape == (n, 4)
    for i in range(n):
        for g in range(4):
            _, _, o = orc.prosail(P['N'][i], P['Cab'][i], P['Cxc'][i], P['Cbr'][i], P['Cw'][i], P['Cm'][i], lai[i], hot[i], iza[g], vza[g], raa[g],
                                  sr[i], sm[i], lidf_type=lidf_type, a=a[i], b=b[i])
            check_sail_object(m, o, idx=(i, g))


def test_sail_batched_spectra_and_shared_spectra():
    n = 6
    P = lhs(n)
    leaf = pb.PROSPECT(**P)
    soil = pb.LSM(0.9, 0.3)
    lai = rng.uniform(0.5, 6, n)
    c = pb.SAIL(iza=[20, 50], vza=[10, 40], raa=[0, 90], ks=leaf.ks, kt=leaf.kt, lai=lai, hotspot=0.1, rho_surface=soil.ref, a=45)
    assert c.BRF.ref.shape == (n, 2, 2101)
    for i in range(n):
        o = orc.sail(50, 40, 90, leaf.ks[i], leaf.kt[i], lai[i], 0.1, soil.ref, a=45)
        assert np.max(np.abs(c.BRF.ref[i, 1] - o['rsot'])) <= TOL and np.max(np.abs(c.HDR.ref[i, 1] - o['rdot'])) <= TOL
    one = pb.SAIL(iza=[20, 50], vza=[10, 40], raa=[0, 90], ks=leaf.ks[0], kt=leaf.kt[0], lai=2.0, hotspot=0.1, rho_surface=soil.ref, a=45)
    assert one.BRF.ref.shape == (2, 2101)
    o = orc.sail(20, 10, 0, leaf.ks[0], leaf.kt[0], 2.0, 0.1, soil.ref, a=45)
    assert np.max(np.abs(one.BRF.ref[0] - o['rsot'])) <= TOL
    rad = pb.SAIL(iza=np.radians(50), vza=np.radians(40), raa=np.radians(90), ks=leaf.ks[0], kt=leaf.kt[0], lai=2.0, hotspot=0.1,
                  rho_surface=soil.ref, a=45, angle_unit='RAD')
    assert np.max(np.abs(rad.BRF.ref - one.BRF.ref[1])) <= 1e-12


def test_j1_series_branch_is_exercised():
    """|k - m| lai <= 1e-3 (models.py:777-779): pick a lai that puts bands of the spectrum on both sides of the switch."""
    leaf = pb.PROSPECT(N=1.5, Cab=40, Cxc=8., Cbr=0.0, Cw=0.01, Cm=0.009)
    soil = pb.LSM(1.0, 0.5)
    hits = 0
    for iza, lai in ((30, 0.02), (48, 0.05), (55, 0.01), (20, 0.03)):
        o = orc.sail(iza, 10, 30, leaf.ks, leaf.kt, lai, 0.05, soil.ref)
        hits += int(np.sum(np.abs((o['vol_ks'] - o['ke']) * lai) <= 1e-3))
        c = pb.SAIL(iza, 10, 30, leaf.ks, leaf.kt, lai, 0.05, soil.ref)
        assert np.max(np.abs(c.BRF.ref - o['rsot'])) <= TOL and np.max(np.abs(c.canopy.DHT - o['tsd'])) <= TOL
    assert hits > 0


# ------------------------------------------------------------------------------------------------------------ engine-level properties
def test_specialised_kernels_agree_bitwise_with_the_general_one(engine):
    """The LUT instances (compile-time plane sets) and the run-time-selected instance are the same arithmetic."""
    n = 3000
    P = lhs(n)
    args = (P['N'], P['Cab'], P['Cxc'], P['Cbr'], P['Cw'], P['Cm'], rng.uniform(0.1, 8, n), rng.uniform(0.01, 0.5, n), np.radians([35.0]),
            np.radians([30.0]), np.radians([50.0]), rng.uniform(0.3, 1.2, n), rng.uniform(0, 1, n))
    kw = dict(lidf_type='campbell', a=rng.uniform(20, 80, n))
    brf = engine.prosail(*args, planes=(capi.O_RSOT,), **kw)
    four = engine.prosail(*args, planes=(capi.O_RSOT, capi.O_RDDT, capi.O_RSDT, capi.O_RDOT), **kw)
    gen = engine.prosail(*args, planes=(capi.O_RSOT, capi.O_RDDT, capi.O_RSDT, capi.O_RDOT, capi.O_KE), mean_planes=(capi.O_RSOT,), **kw)
    wo = engine.prosail(*args, mean_planes=(capi.O_RSOT,), windows_only=True, **kw)
    assert np.array_equal(brf['rsot'], four['rsot']) and np.array_equal(brf['rsot'], gen['rsot'])
    for k in ('rddt', 'rsdt', 'rdot'):
        assert np.array_equal(four[k], gen[k])
    # the band-mean-only launch runs on the PACKED window map (only in-window wavelengths exist, 32-lane tiles of that list):
    # same values, summed in a different grouping -- equal to rounding, not bitwise
    assert np.max(np.abs(wo['means'] - gen['means'])) <= 1e-15
    # band means = plain inclusive means of the spectrum (models.py:811-851)
    for w, (lo, hi) in enumerate(engine.windows):
        assert np.max(np.abs(gen['means'][:, 0, 0, w] - brf['rsot'][:, 0, lo - 400:hi - 400 + 1].mean(axis=1))) <= 1e-14
    assert np.all(np.isfinite(brf['rsot'])) and brf['rsot'].min() >= 0 and brf['rsot'].max() <= 1.5


def test_result_of_a_parameter_set_does_not_depend_on_the_batch_around_it(engine):
    """Size-independent property used at full LUT sizes: row i of a 100k batch == the same set evaluated alone (bitwise)."""
    n = 100000
    P = lhs(n)
    lai, hot, a = rng.uniform(0.1, 8, n), rng.uniform(0.01, 0.5, n), rng.uniform(20, 80, n)
    sr, sm = rng.uniform(0.3, 1.2, n), rng.uniform(0, 1, n)
    g = (np.radians([35.0]), np.radians([30.0]), np.radians([50.0]))
    big = engine.prosail(P['N'], P['Cab'], P['Cxc'], P['Cbr'], P['Cw'], P['Cm'], lai, hot, *g, sr, sm, lidf_type='campbell', a=a)['rsot']
    assert big.shape == (n, 1, 2101) and np.all(np.isfinite(big))
    try:
        for i in (0, 1, 31, 32, 4097, 65535, n - 1):
            args = (P['N'][i], P['Cab'][i], P['Cxc'][i], P['Cbr'][i], P['Cw'][i], P['Cm'][i], lai[i], hot[i], *g, sr[i], sm[i])
            engine.set_prologue_threshold(0)                 # same prologue mapping as the big batch: bitwise equal
            one = engine.prosail(*args, lidf_type='campbell', a=a[i])['rsot']
            assert np.array_equal(one[0, 0], big[i, 0])
            engine.set_prologue_threshold(1 << 40)           # warp-per-item mapping: equal up to summation order
            one = engine.prosail(*args, lidf_type='campbell', a=a[i])['rsot']
            assert np.max(np.abs(one[0, 0] - big[i, 0])) <= 1e-13
            _, _, o = orc.prosail(P['N'][i], P['Cab'][i], P['Cxc'][i], P['Cbr'][i], P['Cw'][i], P['Cm'][i], lai[i], hot[i], 35, 30, 50, sr[i], sm[i], a=a[i])
            assert np.max(np.abs(big[i, 0] - o['rsot'])) <= TOL
    finally:
        engine.set_prologue_threshold(2048)


@pytest.mark.parametrize("lidf_type", ['campbell', 'verhoef'])
def test_both_prologue_mappings_vs_oracle(engine, lidf_type):
    """warp-per-item (shuffle reductions) and thread-per-item prologue kernels against the oracle's VolScatt / hotspot terms."""
    n = 96
    lai, hot = rng.uniform(0.0, 8, n), rng.uniform(0.0, 0.5, n)
    lai[:3] = 0.0
    hot[3:6] = 0.0
    a = rng.uniform(20, 80, n) if lidf_type == 'campbell' else rng.uniform(-0.6, 0.6, n)
    b = np.zeros(n) if lidf_type == 'campbell' else rng.uniform(-0.35, 0.35, n)
    if lidf_type == 'verhoef':
        a[6], b[6] = 1.5, 0.0
        a[7], b[7] = 0.0, 0.0
    izad, vzad, raad = np.array([35.0, 0.0, 60.0, 30.0, 70.0]), np.array([30.0, 0.0, 60.0, 30.0, 10.0]), np.array([50.0, 0.0, 0.0, 0.0, 180.0])
    out = {}
    try:
        for name, thr in (('thread', 0), ('warp', 1 << 40)):
            engine.set_prologue_threshold(thr)
            out[name] = engine.volscatt(np.radians(izad), np.radians(vzad), np.radians(raad), lidf_type=lidf_type, a=a, b=b, lai=lai, hotspot=hot)
    finally:
        engine.set_prologue_threshold(2048)
    assert np.max(np.abs(out['thread'][0] - out['warp'][0])) <= 1e-13 and np.max(np.abs(out['thread'][1] - out['warp'][1])) <= 1e-14
    ks = np.full(2101, 0.3)
    for i in range(0, n, 5):
        lidf = orc.lidf_campbell(a[i]) if lidf_type == 'campbell' else orc.lidf_verhoef(a[i], b[i])
        assert np.max(np.abs(out['thread'][1][i] - lidf)) <= 1e-13
        for g in range(5):
            o = orc.sail(izad[g], vzad[g], raad[g], ks, ks, lai[i], hot[i], ks, lidf_type=lidf_type, a=a[i], b=b[i])
            want = np.array([o['vol_ks'], o['vol_ko'], o['bf'], o['Fs'], o['Ft'], o['tss'], o['too'], o['tsstoo']], dtype=np.float64)
            for name in ('thread', 'warp'):
                assert np.max(np.abs(out[name][0][i, g, :8] - want)) <= 1e-12, (name, i, g)


def test_device_resident_api_matches_host_api(engine):
    n = 5000
    P = lhs(n)
    extra = dict(lai=rng.uniform(0.1, 8, n), hotspot=rng.uniform(0.01, 0.5, n), lidf_a=rng.uniform(20, 80, n), sr=rng.uniform(0.3, 1.2, n),
                 sm=rng.uniform(0, 1, n))
    g = np.radians([35.0, 30.0, 50.0])
    host = engine.prosail(P['N'], P['Cab'], P['Cxc'], P['Cbr'], P['Cw'], P['Cm'], extra['lai'], extra['hotspot'], g[0:1], g[1:2], g[2:3], extra['sr'],
                          extra['sm'], lidf_type='campbell', a=extra['lidf_a'])['rsot'][:, 0]
    dev = {k: engine.to_device(v) for k, v in {**P, **extra}.items()}
    out = engine.empty((n, 2112))
    engine.prosail_device(n, {k: dev[k] for k in P}, {"reflectance": dev['sr'], "moisture": dev['sm']},
                          {"lidf_a": dev['lidf_a'], "lai": dev['lai'], "hotspot": dev['hotspot'], "iza": engine.to_device(g[0:1]),
                           "vza": engine.to_device(g[1:2]), "raa": engine.to_device(g[2:3])}, {capi.O_RSOT: out}, pitch=2112)
    engine.sync()
    assert np.array_equal(out.download()[:, :2101], host)
    ks, kt = engine.empty((n, 2101)), engine.empty((n, 2101))
    engine.prospect_device(n, {k: dev[k] for k in P}, ks, kt)
    engine.sync()
    hp = engine.prospect(**P)
    assert np.array_equal(ks.download(), hp['ks']) and np.array_equal(kt.download(), hp['kt'])


def test_per_sample_geometry(engine):
    n = 40
    P = lhs(n)
    iza, vza, raa = rng.uniform(0, 70, n), rng.uniform(0, 60, n), rng.uniform(0, 180, n)
    m = pb.PROSAIL(lai=3.0, hotspot=0.2, iza=iza, vza=vza, raa=raa, soil_reflectance=1.0, soil_moisture=0.5, a=57, geom_per_sample=True, **P)
    assert m.BRF.ref.shape == (n, 1, 2101)
    for i in range(0, n, 7):
        _, _, o = orc.prosail(P['N'][i], P['Cab'][i], P['Cxc'][i], P['Cbr'][i], P['Cw'][i], P['Cm'][i], 3.0, 0.2, iza[i], vza[i], raa[i], 1.0, 0.5)
        assert np.max(np.abs(m.BRF.ref[i, 0] - o['rsot'])) <= TOL


def test_custom_windows_and_c_abi_errors(engine):
    old = engine.windows
    try:
        engine.set_windows([(400, 2500), (700, 700), (2495, 2600)])
        r = engine.soil(1.0, 0.3, means=True)
        rho = r['rho'][0]
        assert np.allclose(r['means'][0, 0], [rho.astype(np.float32).mean(), np.float32(rho[300]), rho[2095:].astype(np.float32).mean()], atol=TOL_F32)
        with pytest.raises(capi.ProsailError):
            engine.set_windows([(3000, 3100)])
        with pytest.raises(capi.ProsailError):
            engine.set_windows([(400, 2500)] * 9)          # more than 8 windows over one 32-band tile
    finally:
        engine.set_windows(old)
    import ctypes as C
    o = capi.Out()
    o.pitch = 2101
    rc = engine.lib.prosail_prospect_f64(engine.h, 7, 1, C.byref(capi.Leaf()), C.byref(o), None, 1, None)
    assert rc == -1 and b"version must be" in engine.lib.prosail_last_error()
    rc = engine.lib.prosail_prospect_f64(engine.h, 0, 1, C.byref(capi.Leaf()), C.byref(o), None, 1, None)
    assert rc == -1 and b"null leaf" in engine.lib.prosail_last_error()
    o.plane[capi.O_RSOT] = 1
    rc = engine.lib.prosail_prospect_f64(engine.h, 0, 1, C.byref(capi.Leaf()), C.byref(o), None, 1, None)
    assert rc == -1 and b"not an output of prosail_prospect" in engine.lib.prosail_last_error()
    rc = engine.lib.prosail_prospect_f64(engine.h, 0, 1, None, None, None, 1, None)
    assert rc == -1
    assert engine.launch_count() > 0


def test_kernels_write_only_what_they_own_canaries(engine):
    """No compute-sanitizer on this pool: output buffers are embedded in canary-filled allocations (guard rows before and
    after, guard columns 2101..2111 of a 2112 pitch) and every guard value must survive, for the fp64 and fp32 band kernels
    (ragged item count), the window means and the inversion outputs."""
    from pyrism_b200 import inversion
    n, G, pitch, guard = 45, 3, 2112, 4
    P = lhs(n)
    g = np.radians(np.array([[35.0, 10.0, 60.0], [30.0, 45.0, 0.0], [50.0, 120.0, 180.0]]))
    dev = {k: engine.to_device(v) for k, v in P.items()}
    extra = dict(lai=rng.uniform(0.1, 8, n), hotspot=rng.uniform(0.01, 0.5, n), lidf_a=rng.uniform(20, 80, n), sr=rng.uniform(0.3, 1.2, n), sm=rng.uniform(0, 1, n))
    dv = {k: engine.to_device(v) for k, v in extra.items()}
    canopy = {"lidf_a": dv["lidf_a"], "lai": dv["lai"], "hotspot": dv["hotspot"], "iza": engine.to_device(g[0]), "vza": engine.to_device(g[1]),
              "raa": engine.to_device(g[2])}
    soil = {"reflectance": dv["sr"], "moisture": dv["sm"]}
    items = n * G
    nw = len(engine.windows)
    for dt in (np.float64, np.float32):
        es = np.dtype(dt).itemsize
        canary = dt(-777.25)
        planes_host = np.full((items + 2 * guard, pitch), canary, dtype=dt)
        means_host = np.full((items + 2 * guard, 2, nw), canary, dtype=dt)
        bufs = {p: engine.empty(planes_host.shape, dt).upload(planes_host) for p in (capi.O_RSOT, capi.O_RDOT, capi.O_KE)}
        mbuf = engine.empty(means_host.shape, dt).upload(means_host)
        inner = {p: b.ptr + guard * pitch * es for p, b in bufs.items()}
        engine.prosail_device(n, dev, soil, canopy, inner, pitch=pitch, mean_planes=(capi.O_RSOT, capi.O_RDDT), means=mbuf.ptr + guard * 2 * nw * es,
                              n_geom=G, dtype=dt)
        engine.sync()
        for p, b in bufs.items():
            out = b.download()
            assert np.all(out[:guard] == canary) and np.all(out[guard + items:] == canary), (p, dt)
            assert np.all(out[guard:guard + items, 2101:] == canary), (p, dt)                      # pad columns of the pitch
            assert np.all(np.isfinite(out[guard:guard + items, :2101])) and not np.any(out[guard:guard + items, :2101] == canary)
        mo = mbuf.download()
        assert np.all(mo[:guard] == canary) and np.all(mo[guard + items:] == canary) and not np.any(mo[guard:guard + items] == canary)
    # inversion outputs
    m, q, d = 700, 1030, 15
    lutf, obs = rng.random((m, d)).astype(np.float32), rng.random((q, d)).astype(np.float32)
    dl, do = engine.empty((m, d), np.float32).upload(lutf), engine.empty((q, d), np.float32).upload(obs)
    hi = np.full(q + 2 * guard, -5, dtype=np.int64)
    hc = np.full(q + 2 * guard, -777.25, dtype=np.float32)
    di, dc = engine.empty(hi.shape, np.int64).upload(hi), engine.empty(hc.shape, np.float32).upload(hc)
    inversion.invert_device(engine, m, d, dl, d, q, do, d, di.ptr + guard * 8, dc.ptr + guard * 4)
    engine.sync()
    oi, oc = di.download(), dc.download()
    assert np.all(oi[:guard] == -5) and np.all(oi[guard + q:] == -5) and np.all(oc[:guard] == np.float32(-777.25)) and np.all(oc[guard + q:] == np.float32(-777.25))
    ref_i, ref_c = inversion.invert(engine, lutf, obs)
    assert np.array_equal(oi[guard:guard + q], ref_i) and np.array_equal(oc[guard:guard + q], ref_c)          # device API == host API


def test_spectral_response_weighted_windows(engine):
    """prosail_set_windows_srf: window value = sum(w x) / sum(w) (a true sensor-band convolution; all weights 1 is the
    reference's plain mean, models.py:811-851).  Checked against numpy on the spectra the same launch wrote and against the
    oracle spectrum; overlapping and tile-crossing windows, fp64 and fp32, then back to the default windows."""
    n = 300
    P = lhs(n)
    args = (P['N'], P['Cab'], P['Cxc'], P['Cbr'], P['Cw'], P['Cm'], rng.uniform(0.1, 8, n), rng.uniform(0.01, 0.5, n), np.radians([35.0]),
            np.radians([30.0]), np.radians([50.0]), rng.uniform(0.3, 1.2, n), rng.uniform(0, 1, n))
    kw = dict(lidf_type='campbell', a=rng.uniform(20, 80, n))
    wins = ((452, 512), (630, 690), (640, 700), (2400, 2500), (400, 400))
    lam = [np.arange(lo, hi + 1) for lo, hi in wins]
    wts = [np.exp(-0.5 * ((l - l.mean()) / (0.25 * max(l.size, 2))) ** 2) for l in lam]            # Gaussian responses
    wts[3] = rng.uniform(0, 1, lam[3].size)
    default = engine.windows
    try:
        engine.set_windows(wins, wts)
        for dt, tol in ((np.float64, 1e-13), (np.float32, 2e-6)):
            r = engine.prosail(*args, planes=(capi.O_RSOT, capi.O_RDDT), mean_planes=(capi.O_RSOT, capi.O_RDDT), dtype=dt, **kw)
            for w, ((lo, hi), wt) in enumerate(zip(wins, wts)):
                for pi, key in enumerate(('rsot', 'rddt')):
                    ref = (r[key][:, 0, lo - 400:hi - 400 + 1].astype(np.float64) * wt).sum(axis=1) / wt.sum()
                    assert np.max(np.abs(r['means'][:, 0, pi, w] - ref)) <= tol, (dt, w, key)
            wo = engine.prosail(*args, mean_planes=(capi.O_RSOT,), windows_only=True, dtype=dt, **kw)
            assert np.array_equal(wo['means'][:, 0, 0], r['means'][:, 0, 0])
        i = 7
        _, _, o = orc.prosail(P['N'][i], P['Cab'][i], P['Cxc'][i], P['Cbr'][i], P['Cw'][i], P['Cm'][i], args[6][i], args[7][i], 35.0, 30.0, 50.0,
                              args[11][i], args[12][i], lidf_type='campbell', a=kw['a'][i])
        r = engine.prosail(*args, mean_planes=(capi.O_RSOT,), **kw)
        for w, ((lo, hi), wt) in enumerate(zip(wins, wts)):
            assert abs(r['means'][i, 0, 0, w] - (o['rsot'][lo - 400:hi - 400 + 1] * wt).sum() / wt.sum()) <= TOL
        # argument checks of the ABI
        with pytest.raises(ValueError):
            engine.set_windows(wins, wts[:-1])
        with pytest.raises(capi.ProsailError):
            engine.set_windows(((500, 510),), [np.zeros(11)])
        with pytest.raises(capi.ProsailError):
            engine.set_windows(((500, 510),), [-np.ones(11)])
    finally:
        engine.set_windows(default)
    assert not engine.srf
    r = engine.prosail(*args, planes=(capi.O_RSOT,), mean_planes=(capi.O_RSOT,), **kw)
    lo, hi = default[0]
    assert np.max(np.abs(r['means'][:, 0, 0, 0] - r['rsot'][:, 0, lo - 400:hi - 400 + 1].mean(axis=1))) <= 1e-14


def test_empty_and_single_element_batches(engine):
    """Empty parameter arrays are legal (a LUT shard may be empty: shard_bounds(5, r, 8)); length-1 arrays keep their batch axis."""
    e = np.array([])
    p = pb.PROSPECT(N=e, Cab=e, Cxc=e, Cbr=e, Cw=e, Cm=e)
    assert p.ks.shape == (0, 2101) and p.kt.shape == (0, 2101)
    r = engine.prosail(e, e, e, e, e, e, e, e, np.radians([35.0]), np.radians([30.0]), np.radians([50.0]), e, e, lidf_type='campbell', a=e,
                       planes=(capi.O_RSOT,), mean_planes=(capi.O_RSOT,))
    assert r['rsot'].shape == (0, 1, 2101) and r['means'].shape[0] == 0
    one = pb.PROSPECT(N=np.array([1.5]), Cab=np.array([35.0]), Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055)
    ref = pb.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055)
    assert one.ks.shape == (1, 2101) and ref.ks.shape == (2101,) and np.array_equal(one.ks[0], ref.ks)
    with pytest.raises(ValueError):
        pb.PROSPECT(N=np.array([1.5, 2.0]), Cab=np.array([35.0, 36.0, 37.0]), Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055)       # ragged


def test_leftover_attributes_of_the_reference(engine):
    """the reference leaves PROSPECT.r/t/Ra/Ta/denom (models.py:959) and VolScatt.chi_s/chi_o/frho/ftau of the LAST
    leaf-inclination bin (models.py:159) behind as attributes; both come from the kernels here."""
    n = 40
    P = lhs(n)
    for dt, tol in ((np.float64, TOL), (np.float32, 1e-5)):
        p = pb.PROSPECT(dtype=dt, **P)
        for i in (0, 13, n - 1):
            o = orc.prospect(P['N'][i], P['Cab'][i], P['Cxc'][i], P['Cbr'][i], P['Cw'][i], P['Cm'][i])
            for q in ('r', 't', 'Ra', 'Ta', 'denom'):
                got = getattr(p, q)
                assert got.shape == (n, 2101) and got.dtype == dt and np.max(np.abs(got[i] - o[q])) <= tol, (q, i, dt)
    one = pb.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055, Can=1.0, version='D')
    o = orc.prospect(np.float64(1.5), np.float64(35), np.float64(5), np.float64(0.15), np.float64(0.003), np.float64(0.0055), Can=np.float64(1.0), version='D')
    assert one.r.shape == (2101,) and np.max(np.abs(one.r - o['r'])) <= TOL and np.max(np.abs(one.denom - o['denom'])) <= TOL
    # volume scattering terms of the last bin, both prologue mappings
    iza, vza, raa = np.array([35.0, 10.0, 60.0, 0.0]), np.array([30.0, 45.0, 0.0, 0.0]), np.array([50.0, 120.0, 180.0, 0.0])
    for thr in (2048, 0):
        engine.set_prologue_threshold(thr)
        try:
            v = pb.VolScatt(iza, vza, raa)
            v.coef(lidf_type='campbell', a=57)
        finally:
            engine.set_prologue_threshold(2048)
        for g in range(4):
            cs, co, fr, ft = orc.volscatt_volume(np.radians(iza[g]), np.radians(vza[g]), np.radians(raa[g]), 87.5)
            assert max(abs(v.chi_s[g] - cs), abs(v.chi_o[g] - co), abs(v.frho[g] - fr), abs(v.ftau[g] - ft)) <= 1e-12, (thr, g)


def test_volscatt_volume_method():
    """VolScatt.volume(lza) (models.py:177-258) for arbitrary leaf angles, including the branches |cos beta| >= 1."""
    iza, vza, raa = np.array([35.0, 10.0, 60.0, 0.0, 70.0, 89.0]), np.array([30.0, 45.0, 0.0, 0.0, 60.0, 5.0]), np.array([50.0, 120.0, 180.0, 0.0, 10.0, 90.0])
    v = pb.VolScatt(iza, vza, raa)
    for lza in (2.5, 17.5, 45.0, 62.5, 87.5, 0.0, 90.0):
        got = v.volume(lza)
        for g in range(iza.size):
            want = orc.volscatt_volume(np.radians(iza[g]), np.radians(vza[g]), np.radians(raa[g]), lza)
            assert max(abs(got[q][g] - want[q]) for q in range(4)) <= 1e-13, (lza, g)


def test_fused_kernel_rare_leaf_paths_agree_with_the_leaf_only_kernel():
    """The fused fp64 instances read the plate's interface constants from shared memory and rebuild the plain ones from their
    complements on the zero-absorption path (models.py:1057-1059); the leaf-only kernel keeps them in registers.  Same leaf
    spectra from both for: no absorption at all (k = 0), arguments of tau() below / above the table range, an ordinary leaf, and
    bands >= 2101 - 64 (the last, partly empty tile group)."""
    leaves = [dict(N=1.5, Cab=0.0, Cxc=0.0, Cbr=0.0, Cw=0.0, Cm=0.0), dict(N=1.0, Cab=1e-3, Cxc=0.0, Cbr=0.0, Cw=1e-6, Cm=1e-6),
              dict(N=3.0, Cab=200.0, Cxc=60.0, Cbr=3.0, Cw=0.3, Cm=0.1), dict(N=1.0, Cab=0.0, Cxc=0.0, Cbr=0.0, Cw=1.0, Cm=0.0),
              dict(N=1.7, Cab=40.0, Cxc=8.0, Cbr=0.1, Cw=0.01, Cm=0.008)]
    P = {k: np.array([l[k] for l in leaves]) for k in leaves[0]}
    leaf = pb.PROSPECT(**P)
    fused = pb.PROSAIL(lai=2.0, hotspot=0.1, iza=30.0, vza=20.0, raa=40.0, soil_reflectance=0.8, soil_moisture=0.3, keep_leaf=True, **P)
    assert np.all(np.isfinite(fused.leaf_ks)) and np.all(np.isfinite(fused.BRF.ref[1:]))     # row 0: a conservative canopy, 1 - rinf^2 = 0
    # rows 1..4: identical arithmetic up to the last bit of the rebuilt constants; row 0 is the 0/0 limit of the Stokes system
    assert np.max(np.abs(fused.leaf_ks[1:] - leaf.ks[1:])) <= 1e-14 and np.max(np.abs(fused.leaf_kt[1:] - leaf.kt[1:])) <= 1e-14
    assert np.max(np.abs(fused.leaf_ks[0] - leaf.ks[0])) <= 1e-7 and np.max(np.abs(fused.leaf_kt[0] - leaf.kt[0])) <= 1e-7
    o = orc.prosail(1.7, 40.0, 8.0, 0.1, 0.01, 0.008, 2.0, 0.1, 30.0, 20.0, 40.0, 0.8, 0.3, lidf_type='campbell', a=57)[2]
    assert np.max(np.abs(fused.BRF.ref[4, 0] - o['rsot'])) <= TOL


@pytest.mark.parametrize("ne", [1, 9, 36, 90])
def test_lidf_and_volscatt_with_other_numbers_of_leaf_classes(ne):
    """LIDF.campbell / LIDF.verhoef / VolScatt.coef with the reference's n_elements argument (models.py:82-175, 283-392);
    SAIL itself always uses 18 classes."""
    for a in (57.0, 20.0, 80.0, 30.0):
        assert np.max(np.abs(pb.LIDF.campbell(a, n_elements=ne) - orc.lidf_campbell(a, ne))) <= 1e-12
    for a, b in ((-0.35, -0.15), (0.0, 0.0), (0.5, -0.3), (1.2, 0.0)):
        assert np.max(np.abs(pb.LIDF.verhoef(a, b, n_elements=ne) - orc.lidf_verhoef(a, b, ne))) <= 1e-12
    la = np.array([-0.35, 0.2, 0.6])
    lb = np.array([-0.15, 0.1, -0.3])
    got = pb.LIDF.verhoef(la, lb, n_elements=ne)
    assert got.shape == (3, ne)
    for i in range(3):
        assert np.max(np.abs(got[i] - orc.lidf_verhoef(la[i], lb[i], ne))) <= 1e-12
    iza, vza, raa = [35.0, 10.0, 60.0], [30.0, 45.0, 0.0], [50.0, 120.0, 180.0]
    v = pb.VolScatt(iza, vza, raa)
    v.coef(lidf_type='campbell', n_elements=ne, a=45.0)
    lidf = orc.lidf_campbell(45.0, ne)
    assert v.lidf.shape == (ne,) and np.max(np.abs(v.lidf - lidf)) <= 1e-12
    for g in range(3):
        want = orc.volscatt_coef(np.radians(iza[g]), np.radians(vza[g]), np.radians(raa[g]), lidf)
        got = (v.ks[g], v.ko[g], v.bf, v.Fs[g], v.Ft[g])
        assert np.allclose(got, want, atol=1e-12), (ne, g)
        last = orc.volscatt_volume(np.radians(iza[g]), np.radians(vza[g]), np.radians(raa[g]), 90.0 - 45.0 / ne)
        assert np.allclose((v.chi_s[g], v.chi_o[g], v.frho[g], v.ftau[g]), last, atol=1e-12)
    with pytest.raises(capi.ProsailError):
        pb.LIDF.campbell(57.0, n_elements=0)
