"""The CPU oracle (oracle/prosail_oracle.py) against (i) outputs of the live reference stored by
tests/golden/make_golden.py, (ii) the reference's own golden files, (iii) the known-answer constants of the reference's
tests.  No GPU."""
import numpy as np
import pytest

from oracle import prosail_oracle as orc

L8N = [w[0] for w in orc.L8_WINDOWS]
ASN = [w[0] for w in orc.ASTER_WINDOWS]


def test_tables_are_float32_and_complete():
    T = orc.load_tables()
    for k in ("p5_KN", "p5_Kab", "p5_Kxc", "p5_Kbr", "p5_Kw", "p5_Km", "pd_KN", "pd_Kab", "pd_Kxc", "pd_Kbr", "pd_Kw", "pd_Km", "pd_Kan",
              "rsoil1", "rsoil2"):
        assert T[k].shape == (2101,) and T[k].dtype == np.float32
    # library.py: rsoil1 is the bright (dry) spectrum, 0.2377 at 400 nm
    assert abs(float(T["rsoil1"][0]) - 0.2377) < 1e-6


def test_tav_matches_reference_bitwise(cases):
    T = orc.load_tables()
    for tag, KN in (("p5", T["p5_KN"]), ("pd", T["pd_KN"])):
        for alpha in (40, 90, 57):
            assert np.array_equal(orc.calctav(alpha, KN), cases["tav%d_%s" % (alpha, tag)])


def test_tav_fixture_of_reference(fixtures):
    # tests/data/tav_alpha40.mat is the clean-fp64 tav for the D table; the reference's float32 intermediates
    # move it by ~4e-6
    d = np.max(np.abs(orc.calctav(40, orc.load_tables()["pd_KN"]) - fixtures["tav_alpha40"][:, 0]))
    assert 1e-7 < d < 1e-5


def test_leaf_cases_bitwise(cases):
    B = cases["band_index"]
    for x, ver, out, l8, aster in zip(cases["leaf_in"], cases["leaf_version"], cases["leaf_out"], cases["leaf_L8"], cases["leaf_ASTER"]):
        o = orc.prospect(*x[:6], Can=x[6], version=str(ver))
        got = np.stack([o[k] for k in ("ks", "kt", "ka", "ke", "om")])[:, B]
        assert np.array_equal(got, out)
        assert np.array_equal(np.array([o["L8"][b] for b in L8N]), l8)
        assert np.array_equal(np.array([o["ASTER"][b] for b in ASN]), aster)
        assert o["L8"]["B2"].dtype == np.float32


def test_canopy_cases_bitwise(cases):
    B = cases["band_index"]
    rows = list(cases["can_spectra_rows"])
    for x, ver, lt, spec, scal, l8, aster in zip(cases["can_in"], cases["can_version"], cases["can_lidf"], cases["can_spectra"],
                                                 cases["can_scalars"], cases["can_L8"], cases["can_ASTER"]):
        N, Cab, Cxc, Cbr, Cw, Cm, Can, lai, hot, la, lb, sr, sm, iza, vza, raa = x
        leaf, soil, c = orc.prosail(N, Cab, Cxc, Cbr, Cw, Cm, lai, hot, iza, vza, raa, sr, sm, Can=Can, version=str(ver),
                                    lidf_type=str(lt), a=la, b=lb)
        for name in ("rsot", "rddt", "rsdt", "rdot", "rdd", "tdd", "rsd", "tsd", "rdo", "tdo", "rso", "ke"):
            got = np.broadcast_to(np.asarray(c[name], dtype=np.float64), (2101,))[B]
            assert np.array_equal(got, spec[rows.index(name)]), name
        assert np.array_equal(soil["ref"][B], spec[rows.index("rho")])
        assert np.array_equal(leaf["ks"][B], spec[rows.index("ks")]) and np.array_equal(leaf["kt"][B], spec[rows.index("kt")])
        got = np.array([c[k] for k in ("tsstoo", "tss", "too", "vol_ks", "vol_ko", "bf", "Fs", "Ft")], dtype=np.float64)
        assert np.array_equal(got, scal)
        prods = ("BRF", "BRDF", "BHR", "DHR", "HDR")
        assert np.array_equal(np.array([[c["L8"][p][b] for b in L8N] for p in prods]), l8)
        assert np.array_equal(np.array([[c["ASTER"][p][b] for b in ASN] for p in prods]), aster)


def test_lidf_cases_bitwise(cases):
    for a, ref in zip(cases["lidf_campbell_a"], cases["lidf_campbell"]):
        assert np.array_equal(orc.lidf_campbell(np.float64(a)), ref)
    for (a, b), ref in zip(cases["lidf_verhoef_ab"], cases["lidf_verhoef"]):
        assert np.array_equal(orc.lidf_verhoef(np.float64(a), np.float64(b)), ref)


# ---- the reference's own tests, restated against the oracle ---------------------------------------------------------------
def test_reference_prospect5_golden(fixtures):
    """tests/test_prospect_prosail.py:34-52 (opticleaf web run, atol 1e-4)."""
    w, refl, trans = fixtures["prospect5_spectrum"].T
    o = orc.prospect(2.1, 40, 10., 0.1, 0.015, 0.009, version='5')
    assert np.allclose(refl, o["ks"], atol=1e-4) and np.allclose(trans, o["kt"], atol=1e-4)


def test_reference_prospectd_golden(fixtures):
    """tests/test_prospect_prosail.py:54-70 (MATLAB LRT, atol 1e-4)."""
    lrt = fixtures["prospect_d_LRT"]
    o = orc.prospect(1.2, 30, 10., 0.0, 0.015, 0.009, Can=1, version='D')
    assert np.allclose(lrt[:, 1], o["ks"], atol=1e-4) and np.allclose(lrt[:, 2], o["kt"], atol=1e-4)


def test_reference_prosail_golden(fixtures):
    """tests/test_prospect_prosail.py:77-120 (REFL_CAN.txt, atol 1e-2)."""
    w, resv, hdr, sdr, bhr, dhr = fixtures["REFL_CAN"].T
    leaf, soil, c = orc.prosail(1.5, 40, 8., 0.0, 0.01, 0.009, 3, 0.01, 30, 10, 0, 1, 1, version='5', lidf_type='verhoef', a=-0.35, b=-0.15)
    assert np.allclose(sdr, c["rsot"], atol=0.01) and np.allclose(hdr, c["rdot"], atol=0.01)
    assert np.allclose(bhr, c["rddt"], atol=0.01) and np.allclose(dhr, c["rsdt"], atol=0.01)


def test_reference_errors():
    """tests/test_prospect_prosail.py:72-74, 123-155 and models.py:925-926."""
    with pytest.raises(ValueError):
        orc.prospect(2.1, 40, 10., 0.1, 0.015, 0.009, Can=1, version="d")
    with pytest.raises(AssertionError):
        orc.prospect(2.1, 40, 10., 0.1, 0.015, 0.009, Can=0, version="D")
    leaf = orc.prospect(1.5, 40, 8., 0.0, 0.01, 0.009)
    soil = orc.lsm(1, 1)
    short = np.atleast_1d(0.0)
    for ks, kt, rho in ((short, leaf["kt"], soil["ref"]), (leaf["ks"], short, soil["ref"]), (leaf["ks"], leaf["kt"], short)):
        with pytest.raises(AssertionError):
            orc.sail(30, 10, 0, ks, kt, 3, 0.01, rho, lidf_type='verhoef', a=-0.35, b=-0.15)
    with pytest.raises(AssertionError):
        orc.sail(30, 10, 0, leaf["ks"], leaf["kt"], 3, 0.01, soil["ref"], lidf_type='nilson')


CAMPBELL0 = np.array([8.31370856e-01, 1.18374448e-01, 2.60866256e-02, 9.68442862e-03, 4.68681271e-03, 2.66412346e-03, 1.68943850e-03,
                      1.16093072e-03, 8.49048566e-04, 6.53091633e-04, 5.24062841e-04, 4.36148636e-04, 3.74878009e-04, 3.31739471e-04,
                      3.01550650e-04, 2.81097639e-04, 2.68401688e-04, 2.62316948e-04])


def test_reference_lidf_kat():
    """tests/test_volscat.py:6-25."""
    assert np.allclose(orc.lidf_verhoef(0, 0), np.zeros(18) + 0.05555556, atol=1e-4)
    assert np.allclose(orc.lidf_campbell(0), CAMPBELL0, atol=1e-4)
    assert np.max(np.abs(orc.lidf_campbell(0) - CAMPBELL0)) < 1e-9


@pytest.mark.parametrize("lidf, kat", [
    (("verhoef", 0, 0), (0.823188863879162, 0.6867716425258737, 0.5, 0.6217351737543753, 0.011166182392885724)),   # test_volscat.py:28-38
    (("campbell", 0, 0), (0.9948675435577432, 0.9940524814594064, 0.9891027449943284, 0.9915874551449064, 7.491316140639147e-05)),  # :41-52
])
def test_reference_volscatt_kat(lidf, kat):
    lt, a, b = lidf
    l = orc.lidf_verhoef(a, b) if lt == "verhoef" else orc.lidf_campbell(a)
    iza, vza, raa, *_ = orc.kernel_angles(50, 30, 50)
    got = orc.volscatt_coef(iza, vza, raa, l)
    assert np.allclose(got, kat, atol=1e-4)
    assert np.max(np.abs(np.array(got) - np.array(kat))) < 1e-14      # the constants are 16-digit values


def test_band_windows_are_inclusive_and_unweighted():
    x = np.arange(2101, dtype=np.float64)
    m = orc.band_window_means(x, orc.L8_WINDOWS + orc.ASTER_WINDOWS)
    assert m["B2"] == np.mean(np.arange(630 - 400, 690 - 400 + 1))     # later ASTER B2 overwrites L8 B2 in this flat dict
    l8 = orc.band_window_means(x, orc.L8_WINDOWS)
    assert l8["B2"] == (452 + 512) / 2 - 400 and l8["B7"] == (2107 + 2294) / 2 - 400


def test_edge_branches():
    leaf = orc.prospect(1.5, 40, 8., 0.0, 0.01, 0.009)
    soil = orc.lsm(0.8, 0.4)
    bare = orc.sail(30, 20, 10, leaf["ks"], leaf["kt"], 0.0, 0.1, soil["ref"])
    assert np.array_equal(bare["rsot"], soil["ref"]) and bare["tss"] == 1
    nohot = orc.sail(30, 20, 10, leaf["ks"], leaf["kt"], 2.0, 0.0, soil["ref"])
    assert nohot["alf"] == 1e36 and np.all(np.isfinite(nohot["rsot"]))
    hot = orc.sail(30, 30, 0, leaf["ks"], leaf["kt"], 2.5, 0.2, soil["ref"])
    assert hot["alf"] == 0.0 and hot["tsstoo"] == hot["tss"]
    flipped = orc.sail(40, -25, 30, leaf["ks"], leaf["kt"], 4.0, 0.05, soil["ref"], lidf_type='verhoef', a=-0.35, b=-0.15)
    same = orc.sail(40, 25, 210, leaf["ks"], leaf["kt"], 4.0, 0.05, soil["ref"], lidf_type='verhoef', a=-0.35, b=-0.15)
    assert np.allclose(flipped["rsot"], same["rsot"], atol=1e-12)
    # zero absorption (Cw = Cm = 0 above the pigment range): the Stokes branch r + t >= 1 keeps ks + kt = 1
    z = orc.prospect(1.5, 0.0, 0.0, 0.0, 0.0, 0.0)
    assert np.allclose(z["ks"] + z["kt"], 1.0, atol=1e-12)
