# -*- coding: utf-8 -*-
"""LUT inversion on top of the engine: nearest LUT entry (weighted squared distance over band features) for a batch of
observations, on one GPU or over LUT shards spread across ranks.

This is the consumer of the look-up tables the band kernel generates (BASELINE.json north star).  The
reference has no counterpart -- it stops at the forward model -- so the contract is the one stated in
``include/prosail_b200.h`` (``prosail_invert_f32``): float32 cost in a fixed evaluation order, ties to the lowest index.

Multi-GPU: every rank holds a contiguous shard of the LUT (``lut.shard_bounds``) and inverts the SAME observations against
it; the per-rank (cost, global index) pairs are all-gathered (8 + 4 bytes per observation and rank -- NCCL over NVLink on the
GPU box, gloo in the CPU tests) and folded with :func:`reduce_best`.  This arg-min is the only cross-GPU reduction of the
whole path.
"""
import ctypes as C

import numpy as np

from . import _capi as capi


def _f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def invert(engine, lut_features, observations, weights=None, index_offset=0):
    """Best LUT row per observation on ``engine``'s GPU (host arrays in, host arrays out).

    lut_features float32 [M, D], observations float32 [Q, D], weights [D] or None (D <= 32).
    Returns (index int64 [Q] (+ index_offset; -1 where no finite cost exists), cost float32 [Q]).
    """
    lut_features, observations = _f32c(lut_features), _f32c(observations)
    if lut_features.ndim != 2 or observations.ndim != 2 or lut_features.shape[1] != observations.shape[1]:
        raise ValueError("lut_features [M, D] and observations [Q, D] must share D")
    m, d = lut_features.shape
    q = observations.shape[0]
    w = None if weights is None else _f32c(weights).reshape(-1)
    if w is not None and w.size != d:
        raise ValueError("weights must have D = %d values" % d)
    idx = np.empty(q, dtype=np.int64)
    cost = np.empty(q, dtype=np.float32)
    with engine._lock:
        capi.check(engine.lib.prosail_invert_f32(engine.h, m, d, lut_features.ctypes.data, d, q, observations.ctypes.data, d,
                                                 None if w is None else w.ctypes.data, int(index_offset), idx.ctypes.data,
                                                 cost.ctypes.data, 1, None))
    return idx, cost


def invert_device(engine, n_lut, n_feat, lut, pitch_lut, n_obs, obs, pitch_obs, best_index, best_cost, weights=None,
                  index_offset=0, stream=None):
    """Device-resident variant: ``lut``, ``obs``, ``best_index`` (int64) and ``best_cost`` (float32) are DeviceArrays or raw
    device pointers; asynchronous on ``stream``.  ``weights`` stays a host array."""
    w = None if weights is None else _f32c(weights).reshape(-1)
    capi.check(engine.lib.prosail_invert_f32(engine.h, int(n_lut), int(n_feat), C.c_void_p(engine._addr(lut)), int(pitch_lut), int(n_obs),
                                             C.c_void_p(engine._addr(obs)), int(pitch_obs), None if w is None else w.ctypes.data,
                                             int(index_offset), C.c_void_p(engine._addr(best_index)), C.c_void_p(engine._addr(best_cost)),
                                             0, stream))


def reduce_best(costs, indices):
    """Fold per-shard results: costs [W, Q], indices [W, Q] (global, -1 = none) -> (index [Q], cost [Q]).
    Smallest cost wins, equal costs go to the smallest index -- the same rule the kernel applies inside a shard, so the result
    does not depend on how the LUT was sharded."""
    costs = np.asarray(costs, dtype=np.float32)
    indices = np.asarray(indices, dtype=np.int64)
    c = np.where(indices >= 0, costs, np.float32(np.inf))
    best = c.min(axis=0)
    big = np.iinfo(np.int64).max
    cand = np.where((c == best[None, :]) & (indices >= 0), indices, big)
    idx = cand.min(axis=0)
    idx = np.where(idx == big, -1, idx)
    return idx, best


def invert_sharded(engine, lut_local, index_offset, observations, weights=None, group=None):
    """This rank's shard ``lut_local`` starts at global row ``index_offset``; every rank passes the same observations and
    receives the global result.  Without an initialised process group it is the single-GPU :func:`invert`."""
    idx, cost = invert(engine, lut_local, observations, weights, index_offset)
    try:
        import torch
        import torch.distributed as dist
    except ImportError:
        return idx, cost
    if not dist.is_available() or not dist.is_initialized():
        return idx, cost
    world = dist.get_world_size(group)
    use_cuda = dist.get_backend(group) == "nccl"
    ti, tc = torch.from_numpy(idx), torch.from_numpy(cost)
    if use_cuda:
        ti, tc = ti.cuda(), tc.cuda()
    gi = [torch.empty_like(ti) for _ in range(world)]
    gc = [torch.empty_like(tc) for _ in range(world)]
    dist.all_gather(gi, ti, group=group)
    dist.all_gather(gc, tc, group=group)
    return reduce_best(np.stack([t.cpu().numpy() for t in gc]), np.stack([t.cpu().numpy() for t in gi]))


class BandLutInverter(object):
    """Generate a band-mean PROSAIL LUT for this rank's shard of ``params`` (fp32 mode, only the wavelengths inside a
    window are evaluated) and invert observations against it.

    params: dict with N, Cab, Cxc, Cbr, Cw, Cm[, Can], lai, hotspot, lidf_a[, lidf_b], soil_reflectance, soil_moisture -- arrays of
    one common length (the GLOBAL LUT; each rank keeps only its shard).  geometry_deg = (iza, vza, raa) of the observations.
    """

    def __init__(self, engine, params, geometry_deg, version='5', lidf_type='campbell', rank=0, world=1):
        from . import lut as lutmod
        self.engine = engine
        self.params = {k: np.asarray(v) for k, v in params.items()}
        means, (self.begin, self.end) = lutmod.generate_band_lut(engine, params, geometry_deg, version=version, lidf_type=lidf_type,
                                                                 products=("rsot",), windows_only=True, rank=rank, world=world,
                                                                 dtype=np.float32)
        self.features = np.ascontiguousarray(means.reshape(means.shape[0], -1), dtype=np.float32)      # [n_local, n_windows]

    def invert(self, observations, weights=None, group=None):
        """-> (global index [Q], cost [Q], {parameter name: value of the best entry [Q]})"""
        idx, cost = invert_sharded(self.engine, self.features, self.begin, observations, weights, group)
        n = max(np.size(v) for v in self.params.values())
        best = {k: (v[np.clip(idx, 0, n - 1)] if np.size(v) == n else v) for k, v in self.params.items()}
        return idx, cost, best
