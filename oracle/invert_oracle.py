# -*- coding: utf-8 -*-
"""CPU statement of the LUT-inversion contract of ``prosail_invert_f32`` (include/prosail_b200.h).

TEST INFRASTRUCTURE ONLY (same rule as prosail_oracle.py: nothing under pyrism_b200/ imports it).

The reference (ibaris/pyrism) has no inversion.  Pin: tests/golden/inversion_lut.npz (tests/golden/make_inversion_golden.py) holds a
band-mean LUT and observations made from spectra of the LIVE reference together with the arg-min of an independent float64 brute-force
search; this module and the GPU kernels must both reproduce it (tests/test_inversion_cpu.py, tests/test_gpu_inversion.py).  Beyond
that pin, parity is against the contract itself -- float32 features pre-scaled by
sqrtf(w_d), cost accumulated in ascending d with one fused multiply-add per feature, ties to the lowest index.  fmaf is
reproduced in numpy as  float32(float64(diff) * float64(diff) + float64(acc)):  the product of two float32 values is exact in
float64 and the sum is rounded once to 53 bits and once to 24 -- identical to a true fmaf except when the 53-bit sum lands
exactly on a float32 rounding midpoint (probability ~2^-29 per operation); callers that need a hard guarantee use
:func:`is_optimal`, which accepts any index whose float64 cost ties with the minimum within float32 rounding.
"""
import numpy as np


def costs(lut, obs, weights=None):
    """float32 cost matrix [Q, M] in the kernel's evaluation order."""
    lut = np.asarray(lut, dtype=np.float32)
    obs = np.asarray(obs, dtype=np.float32)
    d = lut.shape[1]
    s = np.ones(d, dtype=np.float32) if weights is None else np.sqrt(np.asarray(weights, dtype=np.float32))
    ls = (lut * s[None, :]).astype(np.float32)
    os_ = (obs * s[None, :]).astype(np.float32)
    acc = np.zeros((obs.shape[0], lut.shape[0]), dtype=np.float32)
    for k in range(d):
        diff = (os_[:, k:k + 1] - ls[None, :, k]).astype(np.float32)
        acc = (diff.astype(np.float64) * diff.astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
    return acc


def invert(lut, obs, weights=None, index_offset=0):
    """(index int64 [Q], cost float32 [Q]); lowest index among equal costs; -1 when no cost is finite."""
    c = costs(lut, obs, weights)
    c = np.where(np.isfinite(c), c, np.float32(np.inf))
    idx = np.argmin(c, axis=1).astype(np.int64)          # numpy returns the first minimum
    best = c[np.arange(c.shape[0]), idx]
    idx = np.where(np.isfinite(best), idx + index_offset, -1)
    return idx, best


def invert_topk(lut, obs, k, weights=None, index_offset=0):
    """(index int64 [Q, k], cost float32 [Q, k]): ascending cost, equal costs by ascending index, -1 / inf padding."""
    c = costs(lut, obs, weights)
    c = np.where(np.isfinite(c), c, np.float32(np.inf))
    order = np.argsort(c, axis=1, kind='stable')[:, :k]          # stable: equal costs keep index order
    if order.shape[1] < k:
        order = np.pad(order, ((0, 0), (0, k - order.shape[1])), constant_values=0)
    rows = np.arange(c.shape[0])[:, None]
    best = c[rows, order]
    if lut.shape[0] < k:
        best[:, lut.shape[0]:] = np.inf
    idx = np.where(np.isfinite(best), order + index_offset, -1)
    return idx, best


def is_optimal(lut, obs, idx, weights=None, index_offset=0, rel=4e-7):
    """True per observation if ``idx`` is a minimiser of the float64 cost up to float32 rounding of the accumulation."""
    lut64, obs64 = np.asarray(lut, dtype=np.float64), np.asarray(obs, dtype=np.float64)
    w = np.ones(lut64.shape[1]) if weights is None else np.asarray(weights, dtype=np.float64)
    out = np.zeros(obs64.shape[0], dtype=bool)
    for q in range(obs64.shape[0]):
        c = ((obs64[q][None, :] - lut64) ** 2 * w[None, :]).sum(axis=1)
        out[q] = c[idx[q] - index_offset] <= c.min() * (1.0 + 16 * rel) + 1e-30
    return out
