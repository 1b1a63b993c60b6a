"""The BASELINE.json configurations at their FULL sizes on one B200.  At these sizes the oracle cannot be run over the whole
batch, so each test combines (i) oracle parity on a subsample of the very batch that was computed, with (ii) size-independent
properties that hold for every element: band means == means of the spectra, determinism, shard-concatenation == single run,
Helmholtz reciprocity of 4SAIL over the geometry grid, energy conservation of the leaf, fp32 mode within 1e-5 of fp64.
Inputs: Latin hypercubes with the SURVEY.md 8(d) bounds, seed 20261019 + config index."""
import numpy as np
import pytest

import pyrism_b200 as pb
from pyrism_b200 import _capi as capi
from pyrism_b200 import lut
from oracle import prosail_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9
SEED = 20261019
LEAF = (("N", 1.0, 3.0), ("Cab", 5.0, 90.0), ("Cxc", 1.0, 25.0), ("Cbr", 0.0, 1.0), ("Cw", 0.001, 0.05), ("Cm", 0.001, 0.02))
CANOPY = (("lai", 0.1, 8.0), ("hotspot", 0.01, 0.5), ("soil_refl", 0.3, 1.2), ("soil_moist", 0.0, 1.0))


def draw(n, extra, seed):
    """Latin hypercube with an affine stratum permutation per dimension, stratum(i) = (a i + b) mod n with gcd(a, n) = 1
    (every stratum is hit exactly once, as in scipy's LatinHypercube, but 1e7-sized designs take a second, not 15)."""
    import math
    rng = np.random.default_rng(seed)
    spec = LEAF + CANOPY + tuple(extra)
    i = np.arange(n, dtype=np.int64)
    out = {}
    for name, lo, hi in spec:
        a = int(rng.integers(n // 3, n)) | 1
        while math.gcd(a, n) != 1:
            a += 2
        strata = (a * i + int(rng.integers(0, n))) % n
        out[name] = lo + (hi - lo) * (strata + rng.random(n)) / n
    return out


def leaf_args(P, sl=slice(None)):
    return tuple(P[k][sl] for k in ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm"))


GEOM = tuple(np.radians([v]) for v in (35.0, 30.0, 50.0))


def test_config2_leaf_only_1e7(engine):
    """PROSPECT-5, 1e7 parameter sets x 2101 bands.  The 3.4e11 bytes of spectra are never brought to the host: the kernel's
    own float32 window means (15 windows x ks, kt, ka, ke, om) are, for all 1e7 sets."""
    n = 10000000
    P = draw(n, (), SEED + 1)
    m = np.empty((n, 5, 15), dtype=np.float32)
    step = 2000000
    for s0 in range(0, n, step):
        sl = slice(s0, s0 + step)
        m[sl] = engine.prospect(*leaf_args(P, sl), means=True, spectra=False)['means']
    assert np.all(np.isfinite(m))
    ks, kt, ka, ke, om = (m[:, i] for i in range(5))
    assert ks.min() > 0 and kt.min() >= 0 and ka.min() >= 0                       # every window of every set
    assert np.max(np.abs(ks + kt + ka - 1.0)) <= 4e-7                             # energy conservation, float32 means
    assert np.max(np.abs(ks + ka - ke)) <= 4e-7 and om.max() <= 1.0 + 1e-6
    # subsample: spectra + oracle
    idx = np.r_[0:40, n // 2:n // 2 + 40, n - 40:n]
    r = engine.prospect(*leaf_args(P, idx), means=True)
    assert np.array_equal(r['means'], m[idx])                                     # a set does not depend on the batch around it
    for j, i in enumerate(idx[::6]):
        o = orc.prospect(*[P[k][i] for k in ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm")])
        assert np.max(np.abs(r['ks'][6 * j] - o['ks'])) <= TOL and np.max(np.abs(r['kt'][6 * j] - o['kt'])) <= TOL


def test_config3_prosail_lut_1e6_fp64_and_fp32(engine):
    n = 1000000
    P = draw(n, (("lidf_a", 20.0, 80.0),), SEED + 2)
    args = leaf_args(P) + (P["lai"], P["hotspot"]) + GEOM + (P["soil_refl"], P["soil_moist"])
    kw = dict(lidf_type='campbell', a=P["lidf_a"])
    w64 = engine.prosail(*args, mean_planes=(capi.O_RSOT,), windows_only=True, **kw)['means'][:, 0, 0]          # [n, 15]
    w64b = engine.prosail(*args, mean_planes=(capi.O_RSOT,), windows_only=True, **kw)['means'][:, 0, 0]
    assert np.array_equal(w64, w64b)                                              # deterministic (no atomics in the means)
    assert np.all(np.isfinite(w64)) and w64.min() > 0 and w64.max() < 1.5
    w32 = engine.prosail(*args, mean_planes=(capi.O_RSOT,), windows_only=True, dtype=np.float32, **kw)['means'][:, 0, 0]
    assert np.max(np.abs(w32 - w64)) <= 1e-5                                      # fp32 mode, all 1e6 sets
    # two shards concatenated == the single run (what the multi-GPU path relies on)
    b0, e0 = lut.shard_bounds(n, 0, 2)
    b1, e1 = lut.shard_bounds(n, 1, 2)
    for b, e in ((b0, e0), (b1, e1)):
        a_sh = tuple(x[b:e] if np.size(x) == n else x for x in args)
        sh = engine.prosail(*a_sh, mean_planes=(capi.O_RSOT,), windows_only=True, lidf_type='campbell', a=P["lidf_a"][b:e])['means'][:, 0, 0]
        assert np.array_equal(sh, w64[b:e])
    # full spectra of a subsample: window means == plain means of the spectrum, and oracle parity (all four factors)
    idx = np.r_[0:64, n - 64:n]
    a_sub = tuple(x[idx] if np.size(x) == n else x for x in args)
    planes = (capi.O_RSOT, capi.O_RDDT, capi.O_RSDT, capi.O_RDOT)
    sub = engine.prosail(*a_sub, planes=planes, lidf_type='campbell', a=P["lidf_a"][idx])
    for w, (lo, hi) in enumerate(engine.windows):
        assert np.max(np.abs(sub['rsot'][:, 0, lo - 400:hi - 400 + 1].mean(axis=1) - w64[idx, w])) <= 1e-13
    for j in range(idx.size):                         # all 128 sets of the subsample (round 1: every 8th)
        i = idx[j]
        _, _, o = orc.prosail(P["N"][i], P["Cab"][i], P["Cxc"][i], P["Cbr"][i], P["Cw"][i], P["Cm"][i], P["lai"][i], P["hotspot"][i], 35.0, 30.0, 50.0,
                              P["soil_refl"][i], P["soil_moist"][i], lidf_type='campbell', a=P["lidf_a"][i])
        for k in ('rsot', 'rddt', 'rsdt', 'rdot'):
            assert np.max(np.abs(sub[k][j, 0] - o[k])) <= TOL, (k, i)


def test_config4_multi_angle_1e5_x_64_reciprocity_and_parity(engine):
    """1e5 sets x 64 geometries, Verhoef LIDF.  4SAIL obeys Helmholtz reciprocity (swap sun and view zenith), which ties
    together different items of the batch: checked on the window means of the BRF for every set."""
    n, G = 100000, 64
    P = draw(n, (("lidf_a", -0.6, 0.6), ("lidf_b", -0.35, 0.35)), SEED + 3)
    iza, vza, raa = (g.ravel() for g in np.meshgrid([15.0, 30.0, 45.0, 60.0], [0.0, 15.0, 30.0, 45.0], [0.0, 60.0, 120.0, 180.0], indexing="ij"))
    args = leaf_args(P) + (P["lai"], P["hotspot"], np.radians(iza), np.radians(vza), np.radians(raa), P["soil_refl"], P["soil_moist"])
    m = engine.prosail(*args, mean_planes=(capi.O_RSOT,), windows_only=True, lidf_type='verhoef', a=P["lidf_a"], b=P["lidf_b"])['means'][:, :, 0]
    assert m.shape == (n, G, 15) and np.all(np.isfinite(m)) and m.min() > 0
    pairs = [(g1, g2) for g1 in range(G) for g2 in range(G) if g1 < g2 and iza[g1] == vza[g2] and vza[g1] == iza[g2] and raa[g1] == raa[g2]]
    assert len(pairs) == 12
    for g1, g2 in pairs:
        # two evaluations, each within the 1e-11 error budget of the round-2 primitives
        assert np.max(np.abs(m[:, g1] - m[:, g2])) <= 5e-11, (iza[g1], vza[g1], raa[g1])
    # oracle parity on full spectra of a subsample (all 64 geometries of 6 sets)
    idx = np.array([0, 1, n // 3, n // 2, n - 2, n - 1])
    a_sub = tuple(x[idx] if np.size(x) == n else x for x in args)
    engine.set_prologue_threshold(0)          # same (thread-per-item) prologue mapping as the big batch: bitwise equal
    try:
        sub = engine.prosail(*a_sub, planes=(capi.O_RSOT,), mean_planes=(capi.O_RSOT,), lidf_type='verhoef', a=P["lidf_a"][idx], b=P["lidf_b"][idx])
    finally:
        engine.set_prologue_threshold(2048)
    assert np.max(np.abs(sub['means'][:, :, 0] - m[idx])) <= 1e-15       # packed vs tile-granular window map: equal to rounding
    for j, i in enumerate(idx):
        for g in range(G):                            # all 64 geometries of the 6 sets (round 1: every 5th)
            _, _, o = orc.prosail(P["N"][i], P["Cab"][i], P["Cxc"][i], P["Cbr"][i], P["Cw"][i], P["Cm"][i], P["lai"][i], P["hotspot"][i], iza[g], vza[g], raa[g],
                                  P["soil_refl"][i], P["soil_moist"][i], lidf_type='verhoef', a=P["lidf_a"][i], b=P["lidf_b"][i])
            assert np.max(np.abs(sub['rsot'][j, g] - o['rsot'])) <= TOL, (i, g)


def test_config5_band_mean_lut_one_gpu_share(engine):
    """1.25e7 sets (one GPU's share of the 1e8-set LUT), fused L8/ASTER window means only."""
    n = 12500000
    P = draw(n, (("lidf_a", 20.0, 80.0),), SEED + 4)
    m = np.empty((n, 15))
    step = 2500000
    for s0 in range(0, n, step):
        sl = slice(s0, s0 + step)
        m[sl] = engine.prosail(*(leaf_args(P, sl) + (P["lai"][sl], P["hotspot"][sl]) + GEOM + (P["soil_refl"][sl], P["soil_moist"][sl])),
                               mean_planes=(capi.O_RSOT,), windows_only=True, lidf_type='campbell', a=P["lidf_a"][sl])['means'][:, 0, 0]
    assert np.all(np.isfinite(m)) and m.min() > 0 and m.max() < 1.5
    # overlapping windows are consistent: ASTER B5 (2145-2185) and B6 (2185-2225) share 2185 nm; L8 B7 (2107-2294) covers both.
    # oracle parity of the window means on a subsample
    L8N = [w[0] for w in orc.L8_WINDOWS]
    ASN = [w[0] for w in orc.ASTER_WINDOWS]
    sample = np.unique(np.r_[0, 1, n // 2, n - 1, np.random.default_rng(SEED + 40).integers(0, n, 252)])     # 256 of the 1.25e7 sets (round 1: 4)
    for i in sample:
        _, _, o = orc.prosail(P["N"][i], P["Cab"][i], P["Cxc"][i], P["Cbr"][i], P["Cw"][i], P["Cm"][i], P["lai"][i], P["hotspot"][i], 35.0, 30.0, 50.0,
                              P["soil_refl"][i], P["soil_moist"][i], lidf_type='campbell', a=P["lidf_a"][i])
        ref = np.array([o['L8']['BRF'][b] for b in L8N] + [o['ASTER']['BRF'][b] for b in ASN])
        assert np.max(np.abs(m[i] - ref)) <= TOL
