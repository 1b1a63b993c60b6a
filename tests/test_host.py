"""Host-side logic of pyrism_b200 that needs no GPU: constants, angle handling, result containers, sharding helpers,
and that the C-ABI library loads and exports every symbol declared in include/prosail_b200.h."""
import ctypes
import os
import re

import numpy as np
import pytest

import pyrism_b200 as pb
from pyrism_b200 import _capi as capi
from pyrism_b200 import constants as K
from pyrism_b200 import lut
from pyrism_b200.engine import batch_size, _vec
from oracle import prosail_oracle as orc

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def test_tables_match_reference_loader_layout():
    assert K.lib.p5._fields == ('KN', 'Kab', 'Kxc', 'Kbr', 'Kw', 'Km')
    assert K.lib.pd._fields == ('KN', 'Kab', 'Kxc', 'Kbr', 'Kw', 'Km', 'Kan')
    T = orc.load_tables()
    for f in K.lib.p5._fields:
        assert np.array_equal(getattr(K.lib.p5, f), T['p5_' + f]) and getattr(K.lib.p5, f).dtype == np.float32
    for f in K.lib.pd._fields:
        assert np.array_equal(getattr(K.lib.pd, f), T['pd_' + f])
    assert np.array_equal(K.lib.soil.rsoil1, T['rsoil1']) and np.array_equal(K.lib.soil.rsoil2, T['rsoil2'])


@pytest.mark.parametrize("version", ['5', 'D'])
@pytest.mark.parametrize("alpha", [40, 57, 90, 10.5])
def test_interface_constants_bitwise(version, alpha, cases):
    """The product's host constants == the oracle's calctav == the live reference."""
    KN = (K.lib.p5 if version == '5' else K.lib.pd).KN
    talf, t12, t21 = K.interface_constants(version, alpha)
    assert np.array_equal(talf, orc.calctav(alpha, KN))
    assert np.array_equal(t12, orc.calctav(90, KN))
    assert np.array_equal(t21, orc.calctav(90, KN) / (KN * KN))
    tag = 'p5' if version == '5' else 'pd'
    if alpha in (40, 57, 90):
        assert np.array_equal(talf, cases['tav%d_%s' % (alpha, tag)])
    assert talf.dtype == np.float64 and talf.flags.c_contiguous


def test_default_windows_are_the_reference_bands():
    assert K.DEFAULT_WINDOWS[:6] == tuple((lo, hi) for _, lo, hi in orc.L8_WINDOWS)
    assert K.DEFAULT_WINDOWS[6:] == tuple((lo, hi) for _, lo, hi in orc.ASTER_WINDOWS)


# ---- Kernel: the reference's tests/test_core.py ---------------------------------------------------------------------------
def test_kernel_deg_rad():
    k = pb.Kernel(35, 30, 50)
    assert np.allclose(k.iza, np.radians(35), atol=1e-12) and np.allclose(k.vzaDeg, 30) and k.iza.shape == (1,)
    k = pb.Kernel(np.radians(35), np.radians(30), np.radians(50), angle_unit='RAD')
    assert np.allclose(k.izaDeg, 35, atol=1e-4) and np.allclose(k.raaDeg, 50, atol=1e-4)
    assert np.allclose(k.B, 1 / np.cos(np.radians(35)) + 1 / np.cos(np.radians(30)))


def test_kernel_align_and_errors():
    k = pb.Kernel([10, 20, 30], 5, [0, 90])
    assert k.iza.shape == k.vza.shape == k.raa.shape == (3,)
    assert np.allclose(k.vzaDeg, [5, 5, 5]) and np.allclose(k.raaDeg, [0, 90, 90])       # padded with the last value
    with pytest.raises(AssertionError):
        pb.Kernel([10, 20], 5, 0, align=False)
    with pytest.raises(AssertionError):
        pb.Kernel(10, 5, 0, angle_unit='GRAD')


def test_kernel_negative_zenith_and_normalize():
    k = pb.Kernel(40, -25, 30)
    ref = orc.kernel_angles(40, -25, 30)
    assert np.array_equal(k.iza, [ref[0]]) and np.array_equal(k.vza, [ref[1]]) and np.array_equal(k.raa, [ref[2]])
    assert k.vzaDeg[0] == -25                       # DEG mode keeps the user's value in *Deg (core/_core.py:123-146)
    assert np.allclose(k.phi, (np.radians(30) + np.pi) % (2 * np.pi))
    k = pb.Kernel(40, 25, 30, normalize=True, nbar=10.0)
    assert k.iza.shape == (2,) and np.allclose(k.izaDeg, [40, 10]) and np.allclose(k.vzaDeg, [25, 0])
    assert np.allclose(k.normalization(kernel=np.array([3.0, 1.0])), [2.0, 0.0])
    with pytest.raises(ValueError):
        k.normalization()


def test_sail_result_container():
    r = pb.SailResult(ref=np.array([1.0, 0.1]))
    assert np.allclose(r.refdB, [0.0, -10.0]) and 'refdB' in r
    r.extra = 3
    assert r['extra'] == 3 and 'extra' in dir(r)
    with pytest.raises(AttributeError):
        r.nope
    assert np.allclose(pb.dB(pb.linear(-3.0)), -3.0) and np.allclose(pb.deg(pb.rad(33.0)), 33.0)


def test_batch_helpers():
    assert batch_size(1.0, 2, None) == 1 and batch_size(np.ones(5), 3.0, np.zeros(5)) == 5
    with pytest.raises(ValueError):
        batch_size(np.ones(5), np.ones(4))
    assert np.array_equal(_vec(2.0, 3, 'x'), [2.0, 2.0, 2.0]) and _vec([1, 2], 2, 'x').dtype == np.float64
    with pytest.raises(ValueError):
        _vec([1, 2, 3], 2, 'x')


def test_model_argument_errors_need_no_gpu():
    """Exception types of the reference constructors are raised before any device work."""
    with pytest.raises(ValueError):
        pb.PROSPECT(N=2.1, Cab=40, Cxc=10., Cbr=0.1, Cw=0.015, Cm=0.009, Can=1, version="d")      # models.py:915-917
    with pytest.raises(AssertionError):
        pb.PROSPECT(N=2.1, Cab=40, Cxc=10., Cbr=0.1, Cw=0.015, Cm=0.009, Can=0, version="D")      # models.py:925-926
    ok = np.zeros(2101)
    for ks, kt, rho in ((np.zeros(1), ok, ok), (ok, np.zeros(1), ok), (ok, ok, np.zeros(1))):      # models.py:534-547
        with pytest.raises(AssertionError):
            pb.SAIL(iza=30, vza=10, raa=0, ks=ks, kt=kt, lai=3, hotspot=0.01, rho_surface=rho)
    with pytest.raises(AssertionError):
        pb.SAIL(iza=30, vza=10, raa=0, ks=ok, kt=ok, lai=3, hotspot=0.01, rho_surface=ok, lidf_type='nilson')   # models.py:565
    v = pb.VolScatt(50, 30, 50)
    with pytest.raises(ValueError):
        v.coef(lidf_type='verhoef', a=0)                                                          # models.py:130-131
    with pytest.raises(ValueError):
        v.coef(lidf_type='campbell')                                                              # models.py:136-137
    with pytest.raises(TypeError):
        v.coef(lidf_type='campbell', a=1, c=2)                                                    # models.py:126-127
    with pytest.raises(AttributeError):
        v.coef(lidf_type='nilson', a=1, b=2)                                                      # models.py:142


# ---- C ABI ------------------------------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    header = open(os.path.join(ROOT, "include", "prosail_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(prosail_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(capi.EXPORTED_SYMBOLS)       # the ctypes layer binds exactly the header's functions
    assert b"sm_100a" in capi.load().prosail_version()


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(capi.Leaf) == 7 * 8 and ctypes.sizeof(capi.Soil) == 2 * 8
    assert capi.Canopy.lidf_a.offset == 8 and capi.Canopy.n_geom.offset == 40 and capi.Canopy.iza.offset == 48
    assert ctypes.sizeof(capi.Canopy) == 72
    assert capi.Out.pitch.offset == 23 * 8 and capi.Out.mean_mask.offset == 24 * 8 and capi.Out.means.offset == 25 * 8
    assert ctypes.sizeof(capi.Out) == 28 * 8
    assert capi.NOUT == 23 and len(capi.PLANE_NAMES) == 23 and capi.NGEOM == 12


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    """Without a GPU the product path fails loudly; it never routes through the oracle."""
    with pytest.raises(capi.ProsailError) as e:
        pb.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    import sys
    assert "oracle.prosail_oracle" not in {m for m in sys.modules if m.startswith("pyrism_b200")}
    src = "".join(open(os.path.join(ROOT, "pyrism_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "pyrism_b200")) if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src


# ---- sharding ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n, world", [(10, 1), (10, 3), (1000003, 8), (5, 8), (0, 4)])
def test_shard_bounds_tile_the_index_range(n, world):
    spans = [lut.shard_bounds(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    sizes = [e - b for b, e in spans]
    assert max(sizes) - min(sizes) <= 1


def test_lhs_is_seeded_and_bounded():
    names = ("N", "Cab", "lai")
    a = lut.latin_hypercube(64, names, lut.shard_seed(2, 0))
    b = lut.latin_hypercube(64, names, lut.shard_seed(2, 0))
    c = lut.latin_hypercube(64, names, lut.shard_seed(2, 1))
    assert all(np.array_equal(a[k], b[k]) for k in names) and not np.array_equal(a["N"], c["N"])
    for k in names:
        lo, hi = lut.BOUNDS[k]
        assert a[k].min() >= lo and a[k].max() <= hi
        # one sample per stratum: the defining property of a Latin hypercube
        assert len(set(np.floor((a[k] - lo) / (hi - lo) * 64).astype(int))) == 64
