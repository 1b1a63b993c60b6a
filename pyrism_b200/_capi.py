# -*- coding: utf-8 -*-
"""ctypes binding of ``libprosail_b200.so`` (C ABI: ``include/prosail_b200.h``).

The library is the only compute path of this package.  If it cannot be loaded,
or no B200-class GPU is visible, every entry point raises -- there is no CPU
fallback (the numpy restatement under ``oracle/`` is test infrastructure and is
never imported from here).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYRISM_B200_LIB") or os.path.join(_HERE, "lib", "libprosail_b200.so")

NBANDS = 2101
NLIDF = 18
MAX_WINDOWS = 64

VERSION_5, VERSION_D = 0, 1
LIDF_CAMPBELL, LIDF_VERHOEF = 0, 1

# output planes (prosail_b200.h: PROSAIL_O_*)
(O_RSOT, O_RDDT, O_RSDT, O_RDOT, O_RSO, O_RDD, O_TDD, O_RSD, O_TSD, O_RDO, O_TDO, O_KE, O_LEAF_R, O_LEAF_T,
 O_RHO, O_LEAF_KA, O_LEAF_KE, O_LEAF_OM, O_LAYER_R, O_LAYER_T, O_LAYER_RA, O_LAYER_TA, O_LAYER_DENOM, NOUT) = range(24)
PLANE_NAMES = ("rsot", "rddt", "rsdt", "rdot", "rso", "rdd", "tdd", "rsd", "tsd", "rdo", "tdo", "ke", "leaf_r",
               "leaf_t", "rho", "leaf_ka", "leaf_ke", "leaf_om", "layer_r", "layer_t", "layer_ra", "layer_ta", "layer_denom")
NGEOM = 12      # prosail_b200.h: PROSAIL_NGEOM

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)


class Leaf(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm", "Can")]


class Soil(C.Structure):
    _fields_ = [("reflectance", C.c_void_p), ("moisture", C.c_void_p)]


class Canopy(C.Structure):
    _fields_ = [("lidf_type", C.c_int), ("lidf_a", C.c_void_p), ("lidf_b", C.c_void_p), ("lai", C.c_void_p),
                ("hotspot", C.c_void_p), ("n_geom", C.c_int), ("geom_per_sample", C.c_int), ("iza", C.c_void_p),
                ("vza", C.c_void_p), ("raa", C.c_void_p)]


class Out(C.Structure):
    _fields_ = [("plane", C.c_void_p * NOUT), ("pitch", C.c_longlong), ("mean_mask", C.c_uint), ("means", C.c_void_p),
                ("geom", C.c_void_p), ("windows_only", C.c_int)]


class ProsailError(RuntimeError):
    """Raised when the CUDA library reports an error (code, message)."""

    def __init__(self, code, message):
        super(ProsailError, self).__init__("libprosail_b200: [%d] %s" % (code, message))
        self.code = code


_lib = None

# name -> (restype, argtypes)
_SIGNATURES = {
    "prosail_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "prosail_destroy": (None, [C.c_void_p]),
    "prosail_last_error": (C.c_char_p, []),
    "prosail_version": (C.c_char_p, []),
    "prosail_set_leaf_tables": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 9),
    "prosail_set_soil_tables": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "prosail_set_windows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "prosail_set_windows_srf": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "prosail_prospect_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.POINTER(Leaf), C.POINTER(Out), C.c_void_p, C.c_int,
                                       C.c_void_p]),
    "prosail_soil_f64": (C.c_int, [C.c_void_p, C.c_longlong, C.POINTER(Soil), C.c_void_p, C.c_longlong, C.c_void_p,
                                   C.c_int, C.c_void_p]),
    "prosail_volscatt_f64": (C.c_int, [C.c_void_p, C.c_longlong, C.POINTER(Canopy), C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p]),
    "prosail_sail_f64": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong,
                                   C.c_void_p, C.c_longlong, C.POINTER(Canopy), C.POINTER(Out), C.c_int, C.c_void_p]),
    "prosail_prosail_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.POINTER(Leaf), C.POINTER(Soil),
                                      C.POINTER(Canopy), C.POINTER(Out), C.c_int, C.c_void_p]),
    "prosail_dev_alloc": (C.c_int, [C.c_void_p, C.c_longlong, C.POINTER(C.c_void_p)]),
    "prosail_dev_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "prosail_host_alloc": (C.c_int, [C.c_longlong, C.POINTER(C.c_void_p)]),
    "prosail_host_free": (C.c_int, [C.c_void_p]),
    "prosail_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "prosail_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "prosail_sync": (C.c_int, [C.c_void_p, C.c_void_p]),
    "prosail_timer_begin": (C.c_int, [C.c_void_p, C.c_void_p]),
    "prosail_timer_end": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    "prosail_launch_count": (C.c_longlong, [C.c_void_p]),
    "prosail_set_prologue_threshold": (C.c_int, [C.c_void_p, C.c_longlong]),
    "prosail_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "prosail_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "prosail_device_info": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "prosail_volscatt_bins_f64": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "prosail_volume_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p]),
    "prosail_invert_f32": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_longlong,
                                     C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "prosail_invert_last_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                  C.POINTER(C.c_longlong)]),
    "prosail_invert_topk_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p,
                                          C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "prosail_device_pci_bus_id": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "prosail_measure_fp64_peak": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    "prosail_test_math": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int]),
}
# fp32 mode: same argument lists (prosail_out_f32_t has the layout of prosail_out_t with float planes / means)
for _n in ("prosail_prospect", "prosail_soil", "prosail_sail", "prosail_prosail"):
    _SIGNATURES[_n + "_f32"] = _SIGNATURES[_n + "_f64"]
EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))


def load():
    """Load the shared library (once) and attach the prototypes.  Raises ImportError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `make` (or __graft_entry__.build()). "
                          "pyrism_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise ProsailError(rc, (load().prosail_last_error() or b"").decode("utf-8", "replace"))
