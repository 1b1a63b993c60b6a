# -*- coding: utf-8 -*-
"""pyrism_b200 -- B200-native (sm_100a) batched PROSPECT-5/D -> 4SAIL engine behind pyrism's own API.

    import pyrism_b200 as pyrism
    leaf = pyrism.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055)
    soil = pyrism.LSM(reflectance=3.14 / 4, moisture=0.15)
    can  = pyrism.SAIL(iza=35, vza=30, raa=50, ks=leaf.ks, kt=leaf.kt, rho_surface=soil.ref, lai=3, hotspot=0.25)

Importing the package needs no GPU; the first model call loads ``lib/libprosail_b200.so`` and creates a CUDA
handle.  There is no CPU fallback: without the library or a B200 the call raises.
"""
from .core import Kernel, SailResult, rad, deg, sec, dB, linear, align_all, asarrays
from .models import PROSPECT, LSM, SAIL, PROSAIL, VolScatt, LIDF
from .engine import Engine, get_engine
from . import constants
from . import inversion, lut

__version__ = "0.1.0"
__all__ = ["PROSPECT", "LSM", "SAIL", "PROSAIL", "VolScatt", "LIDF", "Kernel", "SailResult", "Engine", "get_engine",
           "rad", "deg", "sec", "dB", "linear", "align_all", "asarrays", "constants", "inversion", "lut"]
