#!/usr/bin/env python
"""Pack the four 2101-row spectral coefficient tables into one float32 .npz.

The tables are scientific input data (PROSPECT-5/D specific absorption
coefficients, two soil spectra, two light spectra), not code.  They are read
with ``np.loadtxt(..., dtype=np.float32)`` exactly as the reference's loader
does (reference: pyrism/models/library.py:43,50,57,64), so the float32 bit
patterns are the ones the reference computes with.

Usage (in the build container, where /root/reference is mounted):
    python tools/make_tables.py [/root/reference/pyrism/models/data]
"""
import os
import sys

import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/pyrism/models/data"
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pyrism_b200", "data", "tables.npz")

f32 = np.float32
# PROSPECT-D file: leading wavelength column, then KN Kab Kxc Kan Kbr Kw Km (library.py:43)
_, d_KN, d_Kab, d_Kxc, d_Kan, d_Kbr, d_Kw, d_Km = np.loadtxt(os.path.join(src, "prospect_d_spectra.txt"), unpack=True, dtype=f32)
# PROSPECT-5 file: KN Kab Kxc Kbr Kw Km (library.py:50)
p_KN, p_Kab, p_Kxc, p_Kbr, p_Kw, p_Km = np.loadtxt(os.path.join(src, "prospect5_spectra.txt"), unpack=True, dtype=f32)
rsoil1, rsoil2 = np.loadtxt(os.path.join(src, "soil_reflectance.txt"), unpack=True, dtype=f32)
es, ed = np.loadtxt(os.path.join(src, "light_spectra.txt"), unpack=True, dtype=f32)

arrays = dict(
    p5_KN=p_KN, p5_Kab=p_Kab, p5_Kxc=p_Kxc, p5_Kbr=p_Kbr, p5_Kw=p_Kw, p5_Km=p_Km,
    pd_KN=d_KN, pd_Kab=d_Kab, pd_Kxc=d_Kxc, pd_Kbr=d_Kbr, pd_Kw=d_Kw, pd_Km=d_Km, pd_Kan=d_Kan,
    rsoil1=rsoil1, rsoil2=rsoil2, es=es, ed=ed,
)
for k, v in arrays.items():
    assert v.shape == (2101,) and v.dtype == f32, (k, v.shape, v.dtype)
np.savez_compressed(dst, **arrays)
print("wrote", os.path.normpath(dst), os.path.getsize(dst), "bytes")
