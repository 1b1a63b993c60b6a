"""The band kernel's work distribution, restated in Python and checked for what must hold whatever the shape: every (tile group,
chunk) pair is visited exactly once, the walk of a warp is a contiguous run of the group-major list (contiguous mode) or one group
with a fixed stride (strided mode), and the shares differ by at most one pair.

Mirrors pyrism_b200/csrc/prosail_api.cu::launch_bands_g (host: nwarps, nwarps_main, nchunks_main, grid) and
pyrism_b200/csrc/prosail_kernels.cuh::k_bands (device: the walk set-up and the advance at the end of the chunk loop).  The GPU tests
prove the same through results (shard concatenation == single run, bitwise agreement of the instances, ragged sizes); this file
pins the arithmetic on the CPU."""
import itertools

import pytest


def host_launch(ngroups, nchunks, resident_warps, warps_per_cta, contiguous):
    units = ngroups * nchunks
    if contiguous or units <= resident_warps or resident_warps < ngroups:
        warps = max(1, min(units, resident_warps))
        nwarps_main, nchunks_main = 0, 0
    else:
        warps = (resident_warps // ngroups) * ngroups
        nwarps_main, nchunks_main = warps, nchunks
    grid = (warps + warps_per_cta - 1) // warps_per_cta
    return dict(nwarps=warps, nwarps_main=nwarps_main, nchunks_main=nchunks_main, grid=grid)


def device_walk(gw, warp, block, L, ngroups, nchunks, warps_per_cta):
    """The (slot, chunk) pairs warp `gw` visits, in order."""
    remaining, chunk, step, wrap_lo, wrap_hi, tslot = 0, 0, 1, 0, nchunks, 0
    if gw < L['nwarps_main']:
        R = L['nwarps_main'] // ngroups
        rng = gw // ngroups
        tslot = gw - rng * ngroups
        chunk, step, wrap_hi = rng, R, L['nchunks_main']
        remaining = (wrap_hi - chunk + R - 1) // R if chunk < wrap_hi else 0
    elif gw < L['nwarps']:
        full = L['nwarps'] == L['grid'] * warps_per_cta and L['nwarps_main'] == 0
        sh = warp * L['grid'] + block if full else gw
        e, E, nx = sh - L['nwarps_main'], L['nwarps'] - L['nwarps_main'], nchunks - L['nchunks_main']
        nunits = ngroups * nx
        t0, t1 = nunits * e // E, nunits * (e + 1) // E
        tslot = t0 // nx
        wrap_lo = L['nchunks_main']
        chunk = wrap_lo + (t0 - tslot * nx)
        remaining = t1 - t0
    out = []
    while remaining > 0:
        out.append((tslot, chunk))
        nxt, new_group = chunk + step, False
        if nxt >= wrap_hi:
            nxt, new_group = wrap_lo, True
        chunk = nxt
        tslot += 1 if new_group else 0
        remaining -= 1
    return out


SHAPES = [(22, 125000 // 50, 1776, 12), (33, 777, 2368, 16), (13, 1000, 2368, 16), (22, 1, 1776, 12), (33, 3, 3552, 24), (22, 81, 1776, 12),
          (5, 10000, 3, 12), (22, 80, 1776, 12), (1, 17, 64, 16)]


@pytest.mark.parametrize("shape,contiguous", list(itertools.product(SHAPES, (False, True))))
def test_every_pair_once_and_balanced(shape, contiguous):
    ngroups, nchunks, resident, wpc = shape
    L = host_launch(ngroups, nchunks, resident, wpc, contiguous)
    assert 1 <= L['nwarps'] <= max(resident, 1) and L['grid'] * wpc >= L['nwarps']
    seen = {}
    loads = []
    for block in range(L['grid']):
        for warp in range(wpc):
            gw = block * wpc + warp
            walk = device_walk(gw, warp, block, L, ngroups, nchunks, wpc)
            loads.append(len(walk))
            for pair in walk:
                assert 0 <= pair[0] < ngroups and 0 <= pair[1] < nchunks, pair
                assert pair not in seen, ("visited twice", pair, gw, seen[pair])
                seen[pair] = gw
            if L['nwarps_main'] == 0 and walk:          # contiguous mode: one run of the group-major list
                flat = [s * nchunks + c for s, c in walk]
                assert flat == list(range(flat[0], flat[0] + len(flat)))
            if gw < L['nwarps_main'] and walk:          # strided mode: one group, fixed stride
                assert len({s for s, _ in walk}) == 1
    assert len(seen) == ngroups * nchunks
    working = [n for n in loads if n]
    if L['nwarps_main'] == 0:
        assert max(working) - min(working) <= 1 and len(working) == min(L['nwarps'], ngroups * nchunks)
    else:
        assert max(working) - min(working) <= 1 and len(working) == L['nwarps_main']
