# -*- coding: utf-8 -*-
"""Spectral coefficient tables and the per-band interface constants.

``lib`` mirrors the reference's import-time global of the same name
(pyrism/models/models.py:16-19, pyrism/models/library.py:13-17): a ``Spectra``
namedtuple ``(p5, pd, soil, light)`` of float32 arrays with 2101 entries.

``interface_constants`` produces ``talf, t12, t21`` for one (table set, alpha).
They depend on the wavelength only, so they are evaluated once on the host and
uploaded; to stay bit-compatible with the reference they must carry its float32
intermediates: the quantities derived from the float32 ``KN``
table alone are float32 under numpy >= 2 promotion, and only mix into float64
where ``sin(alpha)`` enters.  (reference: PROSPECT.__calctav, models.py:961-998,
and __refl_trans_one_layer, models.py:1011-1015.)
"""
import os
from collections import namedtuple

import numpy as np

Spectra = namedtuple('Spectra', 'p5 pd soil light')
P5S = namedtuple('P5S', 'KN Kab Kxc Kbr Kw Km')
PDS = namedtuple('PDS', 'KN Kab Kxc Kbr Kw Km Kan')
SoilS = namedtuple("SoilS", "rsoil1 rsoil2")
LightS = namedtuple("LightS", "es ed")

NBANDS = 2101
WAVELENGTHS = np.arange(400, 2501)

# Inclusive windows in nm (reference models.py:836-841 and 801-809).
L8_BANDS = (("B2", 452, 512), ("B3", 533, 590), ("B4", 636, 673), ("B5", 851, 879), ("B6", 1566, 1651),
            ("B7", 2107, 2294))
ASTER_BANDS = (("B1", 520, 600), ("B2", 630, 690), ("B3", 760, 860), ("B4", 1600, 1700), ("B5", 2145, 2185),
               ("B6", 2185, 2225), ("B7", 2235, 2285), ("B8", 2295, 2365), ("B9", 2360, 2430))
DEFAULT_WINDOWS = tuple((lo, hi) for _, lo, hi in L8_BANDS + ASTER_BANDS)


def _load():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "tables.npz")
    with np.load(path) as z:
        t = {k: np.ascontiguousarray(z[k], dtype=np.float32) for k in z.files}
    return Spectra(P5S(*(t['p5_' + n] for n in P5S._fields)), PDS(*(t['pd_' + n] for n in PDS._fields)),
                   SoilS(t['rsoil1'], t['rsoil2']), LightS(t['es'], t['ed']))


lib = _load()


def _tav(alpha, n):
    """Stern/Allen mean transmissivity of a dielectric interface for incidence cone ``alpha`` [deg].

    ``n`` is the float32 refractive-index table; ``q``-prefixed names are float32 on purpose.
    """
    if n.dtype != np.float32:
        raise TypeError("the refractive index table must stay float32")
    qn2 = n * n
    qp = qn2 + 1
    qm = qn2 - 1
    qa = (n + 1) * (n + 1) / 2.
    qk = -(qn2 - 1) * (qn2 - 1) / 4.
    s = np.sin(np.deg2rad(alpha))
    s2 = s * s
    half = s2 - qp / 2
    root = np.sqrt(half * half + qk) if alpha != 90 else 0.
    b = root - half
    qa3 = qa ** 3
    ts = (qk ** 2 / (6 * b ** 3) + qk / b - b / 2) - (qk ** 2. / (6 * qa3) + qk / qa - qa / 2)
    num_b = 2 * qp * b - qm ** 2
    num_a = 2 * qp * qa - qm ** 2
    parts = (-2 * qn2 * (b - qa) / (qp ** 2),
             -2 * qn2 * qp * np.log(b / qa) / (qm ** 2),
             qn2 * (1 / b - 1 / qa) / 2,
             16 * qn2 ** 2 * (qn2 ** 2 + 1) * np.log(num_b / num_a) / (qp ** 3 * qm ** 2),
             16 * qn2 ** 3 * (1. / num_b - 1 / num_a) / (qp ** 3))
    tp = parts[0] + parts[1] + parts[2] + parts[3] + parts[4]
    return (ts + tp) / (2 * s ** 2)


_CACHE = {}


def interface_constants(version, alpha=40):
    """(talf, t12, t21) float64[2101] for the table set ``version`` ('5' | 'D') and leaf-surface cone ``alpha``."""
    key = (version, float(alpha))
    if key not in _CACHE:
        KN = (lib.p5 if version == '5' else lib.pd).KN
        talf = np.asarray(_tav(alpha, KN), dtype=np.float64)
        t12 = np.asarray(_tav(90, KN), dtype=np.float64)
        t21 = np.asarray(t12 / (KN * KN), dtype=np.float64)
        _CACHE[key] = (np.ascontiguousarray(talf), np.ascontiguousarray(t12), np.ascontiguousarray(t21))
    return _CACHE[key]
