# -*- coding: utf-8 -*-
"""Look-up-table generation on top of the engine: Latin-hypercube parameter sets, contiguous sharding of the
parameter-set index over ranks (one process per GPU), and the optional final gather of band-mean tables.

The path shards trivially: every (parameter set, geometry) is independent, so there is NO
collective on the data path.  The only communication is the optional ``gather_means`` at the end, which moves
15 numbers per spectrum, not spectra.
"""
import numpy as np

# SURVEY.md 8(d) sampling bounds
BOUNDS = {
    "N": (1.0, 3.0), "Cab": (5.0, 90.0), "Cxc": (1.0, 25.0), "Cbr": (0.0, 1.0), "Cw": (0.001, 0.05), "Cm": (0.001, 0.02),
    "Can": (0.1, 8.0), "lai": (0.1, 8.0), "hotspot": (0.01, 0.5), "lidf_a": (20.0, 80.0),
    "verhoef_a": (-0.6, 0.6), "verhoef_b": (-0.35, 0.35), "soil_reflectance": (0.3, 1.2), "soil_moisture": (0.0, 1.0),
    "iza": (0.0, 70.0), "vza": (0.0, 60.0), "raa": (0.0, 180.0),
}
BASE_SEED = 20261019


def latin_hypercube(n, names, seed, bounds=BOUNDS):
    """dict name -> float64[n], scaled Latin hypercube (scipy.stats.qmc) over ``bounds``."""
    from scipy.stats import qmc
    lo = np.array([bounds[k][0] for k in names], dtype=np.float64)
    hi = np.array([bounds[k][1] for k in names], dtype=np.float64)
    x = qmc.scale(qmc.LatinHypercube(d=len(names), seed=seed).random(int(n)), lo, hi)
    return {k: np.ascontiguousarray(x[:, i]) for i, k in enumerate(names)}


def shard_bounds(n, rank, world):
    """Contiguous, balanced [begin, end) of parameter-set indices owned by ``rank``; the shards tile [0, n)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_seed(config_index, rank):
    """One independent LHS per shard: seed = 20261019 + config_index + 1000 * rank."""
    return BASE_SEED + int(config_index) + 1000 * int(rank)


def shard(params, rank, world):
    """Slice every per-sample array of ``params`` to this rank's contiguous block."""
    n = max(np.size(v) for v in params.values())
    b, e = shard_bounds(n, rank, world)
    return {k: (np.ascontiguousarray(np.asarray(v)[b:e]) if np.size(v) == n else v) for k, v in params.items()}, (b, e)


def gather_means(local_means, n_total, group=None):
    """All-gather the per-rank band-mean tables ``[n_local, ...]`` into ``[n_total, ...]`` on every rank.

    Works with any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).  Shards
    may differ by one row, so rows are padded to the largest shard for the collective and trimmed afterwards.
    """
    import sys
    torch = sys.modules.get("torch")          # no process group without torch already imported by the caller
    if torch is None:
        return np.asarray(local_means)
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return np.asarray(local_means)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    local = np.ascontiguousarray(local_means)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    if local.shape[0] != sizes[rank][1] - sizes[rank][0]:
        raise ValueError("rank %d holds %d rows, expected %d" % (rank, local.shape[0], sizes[rank][1] - sizes[rank][0]))
    rows = max(e - b for b, e in sizes)
    use_cuda = dist.get_backend(group) == "nccl"
    t = torch.zeros((rows,) + local.shape[1:], dtype=torch.float64)
    t[:local.shape[0]] = torch.from_numpy(local.astype(np.float64, copy=False))
    if use_cuda:
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    parts = [o.cpu().numpy()[:e - b] for o, (b, e) in zip(out, sizes)]
    return np.concatenate(parts, axis=0)


def generate_band_lut(engine, params, geometry_deg, version='5', lidf_type='campbell', products=("rsot",), windows_only=True,
                      rank=0, world=1, dtype=np.float64):
    """Band-mean PROSAIL LUT for this rank's shard of ``params`` (BASELINE config 5 shape).

    ``params``: dict with N, Cab, Cxc, Cbr, Cw, Cm[, Can], lai, hotspot, lidf_a[, lidf_b], soil_reflectance,
    soil_moisture (arrays of one common length).  Returns (means [n_local, G, n_products, n_windows], (begin, end)).
    """
    from . import _capi as capi
    from .core import Kernel
    local, span = shard(params, rank, world)
    k = Kernel(*geometry_deg)
    planes = tuple(capi.PLANE_NAMES.index(p) for p in products)
    if len(set(planes)) != len(planes):
        raise ValueError("products must not repeat: %r" % (products,))
    res = engine.prosail(local["N"], local["Cab"], local["Cxc"], local["Cbr"], local["Cw"], local["Cm"], local["lai"],
                         local["hotspot"], k.iza, k.vza, k.raa, local["soil_reflectance"], local["soil_moisture"],
                         Can=local.get("Can"), version=version, lidf_type=lidf_type, a=local["lidf_a"], b=local.get("lidf_b", 0.0),
                         planes=() if windows_only else planes, mean_planes=planes, windows_only=windows_only, dtype=dtype)
    # the kernel emits the mean planes in plane-id order; hand them back in the order of `products`
    order = np.argsort(np.argsort(planes))
    return np.ascontiguousarray(res["means"][..., order, :]), span
