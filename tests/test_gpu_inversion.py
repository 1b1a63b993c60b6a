"""prosail_invert_f32 on the GPU against the contract oracle (oracle/invert_oracle.py): bit-exact indices and costs, ties,
weights, ragged sizes, and an end-to-end inversion of synthetic observations against a PROSAIL band-mean LUT."""
import numpy as np
import pytest

import pyrism_b200 as pb
from pyrism_b200 import inversion, lut
from oracle import invert_oracle as inv

pytestmark = pytest.mark.gpu
rng = np.random.default_rng(20261019 + 41)


@pytest.mark.parametrize("m, q, d", [(1, 1, 1), (257, 3, 15), (5000, 1025, 15), (20000, 300, 6), (3001, 700, 32), (4096, 1024, 8), (70000, 64, 15)])
def test_invert_matches_the_contract_oracle_bitwise(engine, m, q, d):
    lutf = rng.random((m, d)).astype(np.float32)
    obs = rng.random((q, d)).astype(np.float32)
    if m > 300:
        lutf[m - 1] = lutf[5]                               # duplicates: lowest index wins
        obs[0] = lutf[5]
    w = rng.uniform(0.2, 3.0, d).astype(np.float32) if d % 2 else None
    idx, cost = inversion.invert(engine, lutf, obs, w, index_offset=7)
    ref_idx, ref_cost = inv.invert(lutf, obs, w, index_offset=7)
    same = idx == ref_idx
    if not np.all(same):                                    # only possible on an fmaf double-rounding midpoint: must still be optimal
        assert np.all(inv.is_optimal(lutf, obs[~same], idx[~same], w, index_offset=7)) and (~same).sum() <= 2
    assert np.array_equal(cost[same], ref_cost[same])
    if m > 300:
        assert idx[0] == 5 + 7 and cost[0] == 0.0


def test_invert_edge_cases(engine):
    lutf = rng.random((100, 15)).astype(np.float32)
    obs = rng.random((4, 15)).astype(np.float32)
    obs[1] = np.nan
    idx, cost = inversion.invert(engine, lutf, obs)
    assert idx[1] == -1 and np.isinf(cost[1]) and idx[0] >= 0
    i0, _ = inversion.invert(engine, lutf, obs[:0])
    assert i0.shape == (0,)
    with pytest.raises(pb._capi.ProsailError):
        inversion.invert(engine, rng.random((10, 33)).astype(np.float32), rng.random((2, 33)).astype(np.float32))
    with pytest.raises(pb._capi.ProsailError):
        inversion.invert(engine, lutf, obs, weights=-np.ones(15, dtype=np.float32))
    with pytest.raises(ValueError):
        inversion.invert(engine, lutf, obs[:, :3])


def test_sharded_fold_is_independent_of_the_sharding(engine):
    m, q = 50001, 500
    lutf = rng.random((m, 15)).astype(np.float32)
    obs = rng.random((q, 15)).astype(np.float32)
    ref = inversion.invert(engine, lutf, obs)
    for world in (2, 8):
        parts = [inversion.invert(engine, lutf[b:e], obs, index_offset=b) for b, e in (lut.shard_bounds(m, r, world) for r in range(world))]
        idx, cost = inversion.reduce_best(np.stack([p[1] for p in parts]), np.stack([p[0] for p in parts]))
        assert np.array_equal(idx, ref[0]) and np.array_equal(cost, ref[1])


def test_end_to_end_band_lut_inversion_recovers_the_generating_entries(engine):
    """A 40 000-entry PROSAIL band-mean LUT (fp32 mode); observations = LUT rows (must be found exactly, cost 0) and LUT rows
    with 0.5% noise (the retrieved entry must be at least as close as the generating one)."""
    n = 40000
    P = lut.latin_hypercube(n, ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm", "lai", "hotspot", "lidf_a", "soil_reflectance", "soil_moisture"), 20261019 + 5)
    bi = inversion.BandLutInverter(engine, P, (35.0, 30.0, 50.0))
    assert bi.features.shape == (n, 15) and bi.features.dtype == np.float32
    pick = rng.integers(0, n, 200)
    idx, cost, best = bi.invert(bi.features[pick])
    assert np.all(cost == 0.0) and np.allclose(best["lai"], P["lai"][idx])
    assert np.all(inv.costs(bi.features[idx], bi.features[pick]).diagonal() == 0.0)
    noisy = (bi.features[pick] * (1 + 0.005 * rng.standard_normal((200, 15)))).astype(np.float32)
    idx2, cost2, _ = bi.invert(noisy)
    gen_cost = ((noisy.astype(np.float64) - bi.features[pick]) ** 2).sum(axis=1)
    assert np.all(cost2 <= gen_cost * (1 + 1e-5) + 1e-12)
    ref_idx, ref_cost = inv.invert(bi.features, noisy)
    assert np.mean(idx2 == ref_idx) >= 0.99 and np.all(inv.is_optimal(bi.features, noisy, idx2))


def test_band_lut_inverter_shards_fold_to_the_single_gpu_result(engine):
    """Two BandLutInverter shards (rank 0 / rank 1 of a world of 2, both on this GPU) folded with reduce_best == one inverter."""
    n = 30001
    P = lut.latin_hypercube(n, ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm", "lai", "hotspot", "lidf_a", "soil_reflectance", "soil_moisture"), 20261019 + 6)
    full = inversion.BandLutInverter(engine, P, (35.0, 30.0, 50.0))
    obs = (full.features[rng.integers(0, n, 300)] * (1 + 0.01 * rng.standard_normal((300, 15)))).astype(np.float32)
    ref_idx, ref_cost = full.invert_local(obs)
    parts = [inversion.BandLutInverter(engine, P, (35.0, 30.0, 50.0), rank=r, world=2) for r in range(2)]
    assert parts[0].end == parts[1].begin and np.array_equal(np.concatenate([p.features for p in parts]), full.features)
    res = [p.invert_local(obs) for p in parts]
    idx, cost = inversion.reduce_best(np.stack([r[1] for r in res]), np.stack([r[0] for r in res]))
    assert np.array_equal(idx, ref_idx) and np.array_equal(cost, ref_cost)
    assert np.array_equal(inv.invert(full.features, obs)[1], ref_cost)
    w = np.linspace(0.5, 2.0, 15).astype(np.float32)
    assert np.array_equal(full.invert_local(obs, w)[1], inv.invert(full.features, obs, w)[1])


@pytest.mark.parametrize("m, q, d, k", [(5, 3, 15, 8), (300, 257, 15, 1), (5000, 300, 15, 16), (20000, 100, 6, 5), (3001, 70, 32, 10), (70000, 33, 15, 16)])
def test_topk_matches_the_contract_oracle(engine, m, q, d, k):
    lutf = rng.random((m, d)).astype(np.float32)
    obs = rng.random((q, d)).astype(np.float32)
    if m > 100:
        lutf[m - 1] = lutf[7]                               # duplicate rows: equal costs, ascending index order
        lutf[m // 2] = lutf[7]
        obs[0] = lutf[7]
    w = rng.uniform(0.2, 3.0, d).astype(np.float32) if d == 6 else None
    idx, cost = inversion.invert_topk(engine, lutf, obs, k, w, index_offset=3)
    ref_idx, ref_cost = inv.invert_topk(lutf, obs, k, w, index_offset=3)
    assert idx.shape == (q, k)
    same = np.all(idx == ref_idx, axis=1)
    assert same.mean() >= 0.98 and np.array_equal(cost[same], ref_cost[same])           # the rest: fmaf double-rounding near-ties
    assert np.all(cost[:, :-1] <= cost[:, 1:])
    if m > 100 and k >= 3:
        assert idx[0, :3].tolist() == [7 + 3, m // 2 + 3, m - 1 + 3] and np.all(cost[0, :3] == 0)
    if m < k:
        assert np.all(idx[:, m:] == -1) and np.all(np.isinf(cost[:, m:]))
    if k == 1:
        i1, c1 = inversion.invert(engine, lutf, obs, w, index_offset=3)
        assert np.array_equal(i1, idx[:, 0]) and np.array_equal(c1, cost[:, 0])


def test_topk_shard_fold_and_band_lut_inverter(engine):
    m, q, k = 40001, 200, 8
    lutf = rng.random((m, 15)).astype(np.float32)
    obs = rng.random((q, 15)).astype(np.float32)
    ref = inversion.invert_topk(engine, lutf, obs, k)
    parts = [inversion.invert_topk(engine, lutf[b:e], obs, k, index_offset=b) for b, e in (lut.shard_bounds(m, r, 4) for r in range(4))]
    idx, cost = inversion.reduce_topk(np.stack([p[1] for p in parts]), np.stack([p[0] for p in parts]), k)
    assert np.array_equal(idx, ref[0]) and np.array_equal(cost, ref[1])
    n = 20000
    P = lut.latin_hypercube(n, ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm", "lai", "hotspot", "lidf_a", "soil_reflectance", "soil_moisture"), 20261019 + 8)
    bi = inversion.BandLutInverter(engine, P, (35.0, 30.0, 50.0))
    pick = rng.integers(0, n, 50)
    i2, c2, best = bi.invert_topk(bi.features[pick], 4)
    assert np.array_equal(i2[:, 0], pick) and np.all(c2[:, 0] == 0) and best["lai"].shape == (50, 4) and np.allclose(best["lai"][:, 0], P["lai"][pick])
    with pytest.raises(pb._capi.ProsailError):
        inversion.invert_topk(engine, lutf, obs, 17)


def test_inversion_pinned_to_float64_brute_force_on_reference_spectra(engine):
    """The GPU inversion against an INDEPENDENT float64 arg-min on a LUT made of live-reference spectra
    (tests/golden/inversion_lut.npz, generated by tests/golden/make_inversion_golden.py): same indices, costs within float32
    rounding.  Then the same with the LUT regenerated on the GPU from the stored parameter sets: the GPU's band means agree with the
    reference's to <= 1e-9 and the inversion picks the same entries."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inversion_lut.npz"))
    lut32, obs32 = g["lut"].astype(np.float32), g["obs"].astype(np.float32)
    idx, cost = inversion.invert(engine, lut32, obs32)
    assert np.array_equal(idx, g["best"])
    assert np.allclose(cost, g["cost"], rtol=2e-4, atol=1e-12)
    idx5, cost5 = inversion.invert_topk(engine, lut32, obs32, 5)
    srt = np.argsort(((g["obs"][:, None, :] - g["lut"][None, :, :]) ** 2).sum(axis=2), axis=1, kind="stable")[:, :5]
    assert np.array_equal(idx5[:, 0], g["best"]) and np.mean(idx5 == srt) > 0.99     # deeper ranks may tie within float32 rounding
    P = g["params"]
    m = engine.prosail(P[:, 0], P[:, 1], P[:, 2], P[:, 3], P[:, 4], P[:, 5], P[:, 6], P[:, 7], np.radians([35.0]), np.radians([30.0]), np.radians([50.0]),
                       P[:, 9], P[:, 10], lidf_type='campbell', a=P[:, 8], planes=(), mean_planes=(pb._capi.O_RSOT,), windows_only=True)['means'][:, 0, 0]
    assert np.max(np.abs(m - g["lut"])) <= 1e-9
    idx2, _ = inversion.invert(engine, m.astype(np.float32), obs32)
    assert np.array_equal(idx2, g["best"])


def test_tensor_core_candidate_pass_keeps_the_bit_exact_contract(engine):
    """prosail_invert_f32 with <= 15 features runs its candidate pass on the tcgen05 tensor cores (csrc/invert_tc.cuh: split-tf32 MMAs
    into TMEM, then an exact float32 re-rank).  The result must be the contract's, bit for bit -- on random tables, with weights, on
    a tightly clustered table (thousands of near-ties inside the error bound: the exact fix-up scan has to take over) and on a table
    of identical rows (lowest index wins)."""
    for m, q, d, use_w in ((5000, 1025, 15, False), (70001, 300, 15, True), (20000, 257, 6, True), (131072, 128, 11, False)):
        lutf = rng.random((m, d)).astype(np.float32)
        obs = rng.random((q, d)).astype(np.float32)
        obs[:5] = lutf[[7, 7, m - 1, 0, m // 2]]
        w = rng.uniform(0.2, 3.0, d).astype(np.float32) if use_w else None
        idx, cost = inversion.invert(engine, lutf, obs, w, index_offset=3)
        st = inversion.last_stats(engine)
        assert st["tensor_cores"] and not st["timed_out"], st
        assert st["bound_violations"] == 0 and st["error_ratio"] <= 2e-6, st     # the monitored error of the tensor-core brackets: KAPPA / 4
        ref_idx, ref_cost = inv.invert(lutf, obs, w, index_offset=3)
        same = idx == ref_idx
        if not np.all(same):                                # only possible on an fmaf double-rounding midpoint of the numpy oracle
            assert np.all(inv.is_optimal(lutf, obs[~same], idx[~same], w, index_offset=3)) and (~same).sum() <= 2
        assert np.array_equal(cost[same], ref_cost[same])
        assert idx[0] == 7 + 3 and cost[0] == 0.0 and idx[3] == 0 + 3
    # clustered table: every row within 1e-6 of one point -> far more near-ties than the candidate lists hold
    base = rng.random(15).astype(np.float32)
    lutf = (base[None, :] + 1e-6 * rng.standard_normal((6000, 15))).astype(np.float32)
    obs = (base[None, :] + 1e-6 * rng.standard_normal((130, 15))).astype(np.float32)
    idx, cost = inversion.invert(engine, lutf, obs)
    st = inversion.last_stats(engine)
    assert st["tensor_cores"] and not st["timed_out"], st
    ref_idx, ref_cost = inv.invert(lutf, obs)
    same = idx == ref_idx
    assert same.mean() > 0.97 and np.all(inv.is_optimal(lutf, obs[~same], idx[~same])) if not np.all(same) else True
    assert np.array_equal(cost[same], ref_cost[same])
    # identical rows: the lowest index
    lutf = np.tile(rng.random((1, 15)).astype(np.float32), (3000, 1))
    idx, cost = inversion.invert(engine, lutf, rng.random((200, 15)).astype(np.float32), index_offset=11)
    assert np.all(idx == 11) and inversion.last_stats(engine)["redone_exactly"] == 200
    # a table with no finite row at all, and one whose only finite row is the last
    lutf = np.full((1000, 15), np.nan, dtype=np.float32)
    idx, cost = inversion.invert(engine, lutf, rng.random((130, 15)).astype(np.float32))
    assert np.all(idx == -1) and np.all(np.isinf(cost))
    lutf[999] = 0.5
    idx, cost = inversion.invert(engine, lutf, rng.random((130, 15)).astype(np.float32), index_offset=5)
    assert np.all(idx == 999 + 5) and np.all(np.isfinite(cost))
    # NaN observation and NaN row
    lutf = rng.random((4000, 15)).astype(np.float32)
    lutf[17, 3] = np.nan
    obs = rng.random((129, 15)).astype(np.float32)
    obs[1] = np.nan
    idx, cost = inversion.invert(engine, lutf, obs)
    ref_idx, ref_cost = inv.invert(lutf, obs)
    assert idx[1] == -1 and np.isinf(cost[1]) and np.array_equal(idx, ref_idx) and np.array_equal(cost, ref_cost)


def test_tensor_core_path_shapes_pitches_and_zero_weights(engine):
    """Odd sizes around the tile boundaries (128 observations / 128 rows per tile), row pitches larger than the feature count
    (device-resident call), zero weights (a feature that must not count) and a large index offset."""
    from pyrism_b200.inversion import invert_device
    for m, q, d in ((512, 1, 15), (513, 127, 15), (640, 129, 1), (1025, 385, 14), (33000, 1000, 15)):
        pl, po = d + 3, d + 1                                   # padded rows: the extra columns hold garbage that must be ignored
        lutp = rng.random((m, pl)).astype(np.float32) * 10
        obsp = rng.random((q, po)).astype(np.float32) * 10
        lutp[:, :d] = rng.random((m, d)).astype(np.float32)
        obsp[:, :d] = rng.random((q, d)).astype(np.float32)
        w = rng.uniform(0.5, 2.0, d).astype(np.float32)
        if d > 2:
            w[1] = 0.0
        d_lut = engine.empty((m, pl), np.float32).upload(lutp)
        d_obs = engine.empty((q, po), np.float32).upload(obsp)
        d_idx, d_cost = engine.empty((q,), np.int64), engine.empty((q,), np.float32)
        invert_device(engine, m, d, d_lut, pl, q, d_obs, po, d_idx, d_cost, w, 10 ** 12)
        idx, cost = d_idx.download(), d_cost.download()
        st = inversion.last_stats(engine)
        assert st["tensor_cores"] and not st["timed_out"], (m, q, d, st)
        ref_idx, ref_cost = inv.invert(lutp[:, :d], obsp[:, :d], w, index_offset=10 ** 12)
        same = idx == ref_idx
        if not np.all(same):
            assert np.all(inv.is_optimal(lutp[:, :d], obsp[~same][:, :d], idx[~same], w, index_offset=10 ** 12)) and (~same).sum() <= 2
        assert np.array_equal(cost[same], ref_cost[same]), (m, q, d)
        for b in (d_lut, d_obs, d_idx, d_cost):
            b.free()


def test_tensor_core_path_on_a_real_band_mean_table(engine):
    """The table the path exists for: L8 + ASTER BRF band means of 300 000 Latin-hypercube parameter sets (dense and strongly
    correlated: thousands of rows within a few 1e-5 of any observation) against the band means of 4096 held-out sets.  The
    per-observation error bound must keep the exact fix-up scan rare (< 2 % of the observations), the monitored tensor-core error
    must stay under KAPPA / 4, and the result must equal the independent exact float32 kernel (prosail_invert_topk_f32, k = 1)."""
    import bench
    g = np.radians(np.array(bench.GEOM_DEG))

    def means(n, seed):
        P = bench.lhs(n, seed)
        return engine.prosail(P["N"], P["Cab"], P["Cxc"], P["Cbr"], P["Cw"], P["Cm"], P["lai"], P["hotspot"], g[0:1], g[1:2], g[2:3], P["soil_refl"],
                              P["soil_moist"], version='5', lidf_type='campbell', a=P["lidf_a"], planes=(), mean_planes=(pb._capi.O_RSOT,),
                              windows_only=True)["means"][:, 0, 0, :].astype(np.float32)
    lutf, obs = means(300000, 4321), means(4096, 77)
    idx, cost = inversion.invert(engine, lutf, obs)
    st = inversion.last_stats(engine)
    assert st["tensor_cores"] and not st["timed_out"] and st["bound_violations"] == 0 and st["error_ratio"] <= 2e-6, st
    assert st["redone_exactly"] <= 0.02 * len(obs), st
    idx1, cost1 = inversion.invert_topk(engine, lutf, obs, 1)
    assert np.array_equal(idx, idx1[:, 0]) and np.array_equal(cost, cost1[:, 0])
    sub = slice(0, 64)
    ref_idx, ref_cost = inv.invert(lutf, obs[sub])
    assert np.array_equal(idx[sub], ref_idx) and np.array_equal(cost[sub], ref_cost)
