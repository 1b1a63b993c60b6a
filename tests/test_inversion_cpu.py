"""Host logic of the LUT inversion: the contract oracle itself, the shard-fold rule, and the world-size-2 gloo path with the
per-rank kernel call replaced by the oracle (no GPU here; the GPU parity of prosail_invert_f32 is in tests/test_gpu_inversion.py)."""
import os

import numpy as np
import pytest

from oracle import invert_oracle as inv
from pyrism_b200 import inversion, lut

rng = np.random.default_rng(20261019 + 40)


def test_oracle_picks_the_first_minimum_and_respects_weights():
    lutf = rng.random((500, 15)).astype(np.float32)
    lutf[123] = lutf[77]                                   # an exact duplicate: the lower index must win
    obs = lutf[[77, 400, 123]] + np.float32(1e-4)
    idx, cost = inv.invert(lutf, obs)
    assert idx.tolist() == [77, 400, 77] and np.all(cost < 1e-6)
    # zero weight on a feature makes it irrelevant
    lut2 = lutf.copy()
    lut2[:, 3] = rng.random(500).astype(np.float32)
    w = np.ones(15, dtype=np.float32)
    w[3] = 0
    assert np.array_equal(inv.invert(lut2, obs, w)[0], inv.invert(lutf, obs, w)[0])
    # offset and the no-finite-cost marker
    bad = np.full((2, 15), np.nan, dtype=np.float32)
    assert inv.invert(lutf, bad, index_offset=1000)[0].tolist() == [-1, -1]
    assert inv.invert(lutf, obs, index_offset=1000)[0].tolist() == [1077, 1400, 1077]
    assert np.all(inv.is_optimal(lutf, obs, idx))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_fold_of_shards_equals_the_unsharded_result(world):
    m, q = 4001, 257
    lutf = rng.random((m, 6)).astype(np.float32)
    lutf[3000] = lutf[10]                                  # tie across shards
    obs = np.concatenate([lutf[[10, 3000]], rng.random((q - 2, 6)).astype(np.float32)])
    ref_idx, ref_cost = inv.invert(lutf, obs)
    costs, idxs = [], []
    for r in range(world):
        b, e = lut.shard_bounds(m, r, world)
        i, c = inv.invert(lutf[b:e], obs, index_offset=b)
        idxs.append(i); costs.append(c)
    idx, cost = inversion.reduce_best(np.stack(costs), np.stack(idxs))
    assert np.array_equal(idx, ref_idx) and np.array_equal(cost, ref_cost)
    assert idx[0] == 10 and idx[1] == 10
    # a shard without any finite cost does not win
    costs[0][:] = np.nan
    idxs[0][:] = -1
    idx2, _ = inversion.reduce_best(np.stack(costs), np.stack(idxs))
    assert np.all((idx2 >= lut.shard_bounds(m, 0, world)[1]) | (world == 1) | (idx2 == -1))


def test_topk_oracle_and_fold():
    m, q, k = 2000, 40, 6
    lutf = rng.random((m, 9)).astype(np.float32)
    lutf[1500] = lutf[3]
    obs = np.concatenate([lutf[[3]], rng.random((q - 1, 9)).astype(np.float32)])
    ref_i, ref_c = inv.invert_topk(lutf, obs, k)
    assert ref_i[0, :2].tolist() == [3, 1500] and np.all(np.diff(ref_c, axis=1) >= 0)
    assert np.array_equal(ref_i[:, 0], inv.invert(lutf, obs)[0])
    parts = [inv.invert_topk(lutf[b:e], obs, k, index_offset=b) for b, e in (lut.shard_bounds(m, r, 3) for r in range(3))]
    idx, cost = inversion.reduce_topk(np.stack([p[1] for p in parts]), np.stack([p[0] for p in parts]), k)
    assert np.array_equal(idx, ref_i) and np.array_equal(cost, ref_c)
    small_i, small_c = inv.invert_topk(lutf[:4], obs, k)
    assert np.all(small_i[:, 4:] == -1) and np.all(np.isinf(small_c[:, 4:]))


def _gloo_worker(rank, world, port, lutf, obs, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeEngine(object):          # the kernel call of this rank, played by the contract oracle
        pass
    orig = inversion.invert
    inversion.invert = lambda eng, l, o, w=None, off=0: inv.invert(l, o, w, off)
    try:
        b, e = lut.shard_bounds(lutf.shape[0], rank, world)
        idx, cost = inversion.invert_sharded(FakeEngine(), lutf[b:e], b, obs)
    finally:
        inversion.invert = orig
    if rank == 0:
        np.savez(out, idx=idx, cost=cost)
    dist.barrier()
    dist.destroy_process_group()


def test_invert_sharded_over_gloo_world_size_2(tmp_path):
    import torch.multiprocessing as mp
    lutf = rng.random((3000, 15)).astype(np.float32)
    obs = rng.random((100, 15)).astype(np.float32)
    out = str(tmp_path / "inv.npz")
    mp.spawn(_gloo_worker, args=(2, 29641, lutf, obs, out), nprocs=2, join=True)
    z = np.load(out)
    ref_idx, ref_cost = inv.invert(lutf, obs)
    assert np.array_equal(z["idx"], ref_idx) and np.array_equal(z["cost"], ref_cost)


def test_contract_oracle_agrees_with_float64_brute_force_on_reference_spectra():
    """Pin of the inversion to reference DATA: tests/golden/inversion_lut.npz holds a 600-entry band-mean LUT and 100 observations
    made from spectra of the live reference, with the arg-min of an independent float64 brute-force search
    (tests/golden/make_inversion_golden.py).  Every margin between best and second-best cost is > 2e-3 relative, far above float32
    rounding, so the float32 contract must pick exactly the float64 minimiser."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inversion_lut.npz"))
    assert np.min(g["margin"] / (g["cost"] + g["margin"])) > 1e-3
    idx, cost = inv.invert(g["lut"].astype(np.float32), g["obs"].astype(np.float32))
    assert np.array_equal(idx, g["best"])
    assert np.allclose(cost, g["cost"], rtol=2e-4, atol=1e-12)
    assert np.array_equal(g["best"][50:], g["noisy_rows"])                  # the noisy LUT rows are recovered
    assert np.all(inv.is_optimal(g["lut"], g["obs"], idx))
