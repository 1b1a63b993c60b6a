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


def invert_topk(engine, lut_features, observations, k, weights=None, index_offset=0):
    """The k best LUT rows per observation (1 <= k <= 16): (index int64 [Q, k], cost float32 [Q, k]), ascending cost, equal costs
    by ascending index; -1 / inf pad when fewer than k rows have a finite cost."""
    lut_features, observations = _f32c(lut_features), _f32c(observations)
    if lut_features.ndim != 2 or observations.ndim != 2 or lut_features.shape[1] != observations.shape[1]:
        raise ValueError("lut_features [M, D] and observations [Q, D] must share D")
    m, d = lut_features.shape
    q = observations.shape[0]
    w = None if weights is None else _f32c(weights).reshape(-1)
    if w is not None and w.size != d:
        raise ValueError("weights must have D = %d values" % d)
    idx = np.empty((q, int(k)), dtype=np.int64)
    cost = np.empty((q, int(k)), dtype=np.float32)
    with engine._lock:
        capi.check(engine.lib.prosail_invert_topk_f32(engine.h, int(k), m, d, lut_features.ctypes.data, d, q, observations.ctypes.data, d,
                                                      None if w is None else w.ctypes.data, int(index_offset), idx.ctypes.data,
                                                      cost.ctypes.data, 1, None))
    return idx, cost


def reduce_topk(costs, indices, k):
    """Fold per-shard k-best lists: costs [W, Q, k], indices [W, Q, k] (global, -1 = none) -> (index [Q, k], cost [Q, k]) with
    the kernel's order (ascending cost, equal costs by ascending index)."""
    costs = np.asarray(costs, dtype=np.float32)
    indices = np.asarray(indices, dtype=np.int64)
    w, q, kk = costs.shape
    c = np.moveaxis(costs, 0, 1).reshape(q, w * kk)
    i = np.moveaxis(indices, 0, 1).reshape(q, w * kk)
    c = np.where(i >= 0, c, np.float32(np.inf))
    key_i = np.where(i >= 0, i, np.iinfo(np.int64).max)
    order = np.lexsort((key_i, c), axis=1)[:, :k]
    rows = np.arange(q)[:, None]
    out_i, out_c = i[rows, order], c[rows, order]
    return np.where(np.isfinite(out_c), out_i, -1), out_c


def invert_device(engine, n_lut, n_feat, lut, pitch_lut, n_obs, obs, pitch_obs, best_index, best_cost, weights=None,
                  index_offset=0, stream=None):
    """Device-resident variant: ``lut``, ``obs``, ``best_index`` (int64) and ``best_cost`` (float32) are DeviceArrays or raw
    device pointers; asynchronous on ``stream``.  ``weights`` stays a host array."""
    w = None if weights is None else _f32c(weights).reshape(-1)
    capi.check(engine.lib.prosail_invert_f32(engine.h, int(n_lut), int(n_feat), C.c_void_p(engine._addr(lut)), int(pitch_lut), int(n_obs),
                                             C.c_void_p(engine._addr(obs)), int(pitch_obs), None if w is None else w.ctypes.data,
                                             int(index_offset), C.c_void_p(engine._addr(best_index)), C.c_void_p(engine._addr(best_cost)),
                                             0, stream))


def last_stats(engine):
    """How the last ``invert`` / ``invert_device`` on this engine ran (waits for it): dict(tensor_cores=bool -- the candidate pass ran
    on tcgen05 (``csrc/invert_tc.cuh``; n_feat <= 15 and >= 512 LUT rows), redone_exactly=int -- observations rescanned exactly because
    their candidate list could not be proven complete, timed_out=bool, error_ratio=float -- the largest measured error of the
    tensor-core brackets relative to sum |A_k||B_k| over the kept candidates (the model assumes 8e-6), bound_violations=int --
    candidates outside the model's bound (expected 0; their observations were rescanned exactly)).  The results obey the same
    bit-exact contract either way."""
    used, redone, to, ratio, viol = C.c_int(0), C.c_longlong(0), C.c_int(0), C.c_double(0.0), C.c_longlong(0)
    capi.check(engine.lib.prosail_invert_last_stats(engine.h, C.byref(used), C.byref(redone), C.byref(to), C.byref(ratio), C.byref(viol)))
    return dict(tensor_cores=bool(used.value), redone_exactly=int(redone.value), timed_out=bool(to.value), error_ratio=float(ratio.value),
                bound_violations=int(viol.value))


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


def _fold_over_ranks(idx, cost, group=None):
    """all_gather the per-rank (index, cost) pairs and fold them; identity without an initialised process group."""
    import sys
    torch = sys.modules.get("torch")          # a process group can only exist if the caller imported torch: never import it here (~2 s)
    if torch is None:
        return idx, cost
    import torch.distributed as dist
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


def invert_sharded(engine, lut_local, index_offset, observations, weights=None, group=None):
    """This rank's shard ``lut_local`` starts at global row ``index_offset``; every rank passes the same observations and
    receives the global result.  Without an initialised process group it is the single-GPU :func:`invert`."""
    idx, cost = invert(engine, lut_local, observations, weights, index_offset)
    return _fold_over_ranks(idx, cost, group)


class BandLutInverter(object):
    """Generate a band-mean PROSAIL LUT for this rank's shard of ``params`` (fp32 mode, only the wavelengths inside a
    window are evaluated) and invert observations against it.  The LUT features are produced on the GPU and STAY there
    (``features_dev``, float32 [n_local, n_windows]); an inversion call moves only the observations and the results.

    params: dict with N, Cab, Cxc, Cbr, Cw, Cm[, Can], lai, hotspot, lidf_a[, lidf_b], soil_reflectance, soil_moisture -- arrays of
    one common length (the GLOBAL LUT; each rank keeps only its shard).  geometry_deg = (iza, vza, raa) of the observations.
    """

    def __init__(self, engine, params, geometry_deg, version='5', lidf_type='campbell', rank=0, world=1):
        from . import lut as lutmod
        from .core import Kernel
        self.engine = engine
        self.params = {k: np.asarray(v) for k, v in params.items()}
        local, (self.begin, self.end) = lutmod.shard(params, rank, world)
        n = self.end - self.begin
        self.n_local, self.n_feat = n, len(engine.windows)
        self.features_dev = engine.empty((max(n, 1), self.n_feat), np.float32)
        self._features = None
        if n > 0:
            k = Kernel(*geometry_deg)
            if k.iza.size != 1:
                raise ValueError("BandLutInverter takes one sun/view geometry")
            dev = lambda name, default=None: engine.to_device(np.broadcast_to(np.asarray(local.get(name, default), dtype=np.float64), (n,)))
            leaf = {q: dev(q) for q in ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm")}
            if version == 'D':
                leaf["Can"] = dev("Can")
            soil = {"reflectance": dev("soil_reflectance"), "moisture": dev("soil_moisture")}
            canopy = {"lidf_a": dev("lidf_a"), "lidf_b": dev("lidf_b", 0.0), "lai": dev("lai"), "hotspot": dev("hotspot"),
                      "iza": engine.to_device(k.iza), "vza": engine.to_device(k.vza), "raa": engine.to_device(k.raa)}
            with engine._lock:
                engine.set_alpha(version, 40)
                engine.prosail_device(n, leaf, soil, canopy, None, mean_planes=(capi.O_RSOT,), means=self.features_dev, version=version,
                                      lidf_type=lidf_type, windows_only=True, dtype=np.float32)
                engine.sync()

    @property
    def features(self):
        """Host copy of the LUT features (float32 [n_local, n_windows]), downloaded on first use."""
        if self._features is None:
            self._features = self.features_dev.download()[:self.n_local]
        return self._features

    def invert_local(self, observations, weights=None):
        """Best entry of THIS rank's shard: (global index [Q], cost [Q]); only the observations travel to the GPU."""
        obs = _f32c(observations)
        if obs.ndim != 2 or obs.shape[1] != self.n_feat:
            raise ValueError("observations must be [Q, %d]" % self.n_feat)
        q = obs.shape[0]
        idx, cost = np.full(q, -1, dtype=np.int64), np.full(q, np.inf, dtype=np.float32)
        if q == 0 or self.n_local == 0:
            return idx, cost
        eng = self.engine
        with eng._lock:
            d_obs = eng.empty(obs.shape, np.float32).upload(obs)
            d_idx, d_cost = eng.empty((q,), np.int64), eng.empty((q,), np.float32)
            invert_device(eng, self.n_local, self.n_feat, self.features_dev, self.n_feat, q, d_obs, self.n_feat, d_idx, d_cost, weights, self.begin)
            idx, cost = d_idx.download(), d_cost.download()
            for b in (d_obs, d_idx, d_cost):
                b.free()
        return idx, cost

    def invert_topk(self, observations, k, weights=None):
        """k best entries of THIS rank's shard: (global index [Q, k], cost [Q, k], {parameter name: [Q, k] values});
        fold several ranks with :func:`reduce_topk`."""
        obs = _f32c(observations)
        if obs.ndim != 2 or obs.shape[1] != self.n_feat:
            raise ValueError("observations must be [Q, %d]" % self.n_feat)
        eng = self.engine
        q = obs.shape[0]
        if q == 0 or self.n_local == 0:          # empty shard (world > LUT rows) or nothing to invert: padded result, like invert_local
            idx = np.full((q, int(k)), -1, dtype=np.int64)
            cost = np.full((q, int(k)), np.inf, dtype=np.float32)
            return idx, cost, {kk: np.full((q, int(k)), np.nan) for kk in self.params}
        with eng._lock:
            d_obs = eng.empty(obs.shape, np.float32).upload(obs)
            d_idx, d_cost = eng.empty((q, int(k)), np.int64), eng.empty((q, int(k)), np.float32)
            w = None if weights is None else _f32c(weights).reshape(-1)
            capi.check(eng.lib.prosail_invert_topk_f32(eng.h, int(k), self.n_local, self.n_feat, C.c_void_p(self.features_dev.ptr), self.n_feat, q,
                                                       C.c_void_p(d_obs.ptr), self.n_feat, None if w is None else w.ctypes.data, self.begin,
                                                       C.c_void_p(d_idx.ptr), C.c_void_p(d_cost.ptr), 0, None))
            idx, cost = d_idx.download(), d_cost.download()
            for b in (d_obs, d_idx, d_cost):
                b.free()
        n = max(np.size(v) for v in self.params.values())
        best = {kk: (v[np.clip(idx, 0, n - 1)] if np.size(v) == n else v) for kk, v in self.params.items()}
        return idx, cost, best

    def invert(self, observations, weights=None, group=None):
        """-> (global index [Q], cost [Q], {parameter name: value of the best entry [Q]}); folds over the ranks of ``group``."""
        idx, cost = self.invert_local(observations, weights)
        idx, cost = _fold_over_ranks(idx, cost, group)
        n = max(np.size(v) for v in self.params.values())
        best = {k: (v[np.clip(idx, 0, n - 1)] if np.size(v) == n else v) for k, v in self.params.items()}
        return idx, cost, best
