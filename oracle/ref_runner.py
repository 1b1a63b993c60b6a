# -*- coding: utf-8 -*-
"""Runs the UNMODIFIED reference (ibaris/pyrism 0.0.5) from ``oracle/_ref/pyrism``.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Nothing in ``pyrism_b200/`` imports this module; only ``tests/`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do.

``oracle/_ref/`` is a build product, not source: ``__graft_entry__.build()`` copies the reference's own package
directory ``/root/reference/pyrism`` there when the reference is mounted (this container).  It is git-ignored (never in
history) but travels to the GPU box with the ``gpurun`` snapshot, so the reference itself -- not the port -- can be timed
on the box's host cores next to the GPU arm (``cpu_baseline.kind == "reference"``) and used to re-check the oracle there.

The only accommodation is the two-line shim of SURVEY.md F1: ``pyrism/models/models.py:10`` imports
``scipy.misc.factorial``, a name that scipy >= 1.3 no longer has; it is set to ``scipy.special.factorial`` before the import.
No reference file is edited.  Parameters are passed as ``numpy.float64`` scalars (with Python floats the
reference computes parts of PROSPECT in float32 under numpy >= 2).
"""
from __future__ import division

import os
import sys
import warnings

import numpy as np

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mod = None


def available():
    """True when oracle/_ref/pyrism exists (built here by __graft_entry__.build(); shipped to the GPU box by gpurun)."""
    return os.path.isfile(os.path.join(REF_DIR, "pyrism", "__init__.py"))


def load():
    """Import the reference package (once) and return the module."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        raise ImportError("oracle/_ref/pyrism is absent: run __graft_entry__.build() where /root/reference is mounted")
    warnings.filterwarnings("ignore")
    import scipy.misc
    import scipy.special
    if not hasattr(scipy.misc, "factorial"):
        scipy.misc.factorial = scipy.special.factorial          # SURVEY.md F1; the reference file itself is untouched
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import pyrism
    if not os.path.abspath(pyrism.__file__).startswith(os.path.abspath(REF_DIR)):
        raise ImportError("a different pyrism (%s) shadows oracle/_ref" % pyrism.__file__)
    _mod = pyrism
    return pyrism


def prosail(N, Cab, Cxc, Cbr, Cw, Cm, lai, hotspot, iza, vza, raa, soil_reflectance, soil_moisture, Can=0, version='5',
            lidf_type='campbell', a=57, b=0):
    """PROSPECT -> LSM -> SAIL through the reference's own constructors, exactly as examples/prospect_sail.py:9-28 chains
    them.  Returns the three reference objects (leaf, soil, canopy)."""
    pyrism = load()
    f = np.float64
    kw = dict(N=f(N), Cab=f(Cab), Cxc=f(Cxc), Cbr=f(Cbr), Cw=f(Cw), Cm=f(Cm))
    if version == 'D':
        kw.update(Can=f(Can), version='D')
    leaf = pyrism.PROSPECT(**kw)
    soil = pyrism.LSM(reflectance=f(soil_reflectance), moisture=f(soil_moisture))
    ckw = dict(a=f(a), b=f(b)) if lidf_type == 'verhoef' else dict(a=f(a))
    can = pyrism.SAIL(iza=f(iza), vza=f(vza), raa=f(raa), ks=leaf.ks, kt=leaf.kt, rho_surface=soil.ref, lidf_type=lidf_type,
                      lai=f(lai), hotspot=f(hotspot), **ckw)
    return leaf, soil, can
