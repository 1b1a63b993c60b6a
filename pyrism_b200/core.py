# -*- coding: utf-8 -*-
"""Angle handling, result containers and unit helpers used by the optical hot path.

Mirrors the parts of ``pyrism/core`` the path touches: ``Kernel`` (core/_core.py:17-179),
``SailResult`` (core/auxiliary.py:81-131) and ``rad, deg, sec, dB, linear, align_all, asarrays``
(core/auxiliary.py:182-348).  Everything here is per-call host bookkeeping (O(number of geometries)),
not per-wavelength work.
"""
import numpy as np


def rad(angle):
    """Degrees -> radians (auxiliary.py:182-192)."""
    return angle * np.pi / 180.0


def deg(angle):
    """Radians -> degrees (auxiliary.py:195-205)."""
    return angle * 180. / np.pi


def sec(angle):
    return 1 / np.cos(angle)


def dB(x):
    """Linear -> dB (auxiliary.py:222-227)."""
    with np.errstate(invalid='ignore', divide='ignore'):
        return 10 * np.log10(x)


def linear(x):
    return 10 ** (x / 10)


def asarrays(data):
    return [np.asarray(item).flatten() for item in data]


def max_length(data):
    return np.max([len(item) for item in data])


def align_all(data, constant_values='default'):
    """Pad every 1-D array to the longest one, repeating its last value (auxiliary.py:332-340)."""
    n = max(len(item) for item in data)
    out = np.empty((len(data), n), dtype=np.result_type(*data))
    for k, item in enumerate(data):
        out[k, :len(item)] = item
        out[k, len(item):] = item[-1] if isinstance(constant_values, str) and constant_values == 'default' else constant_values
    return out


class SailResult(dict):
    """dict with attribute access (auxiliary.py:81-131).  ``refdB`` is derived from ``ref`` on first use."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __missing__(self, key):
        if key == 'refdB' and 'ref' in self:
            self[key] = dB(self['ref'])
            return dict.__getitem__(self, key)
        raise KeyError(key)

    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__

    def __repr__(self):
        if self.keys():
            m = max(map(len, list(self.keys()))) + 1
            return '\n'.join([k.rjust(m) + ': ' + repr(v) for k, v in sorted(self.items())])
        return self.__class__.__name__ + "()"

    def __dir__(self):
        return list(self.keys())


class Kernel(object):
    """Angle normalisation base class (core/_core.py:60-173).

    Attributes: ``iza, vza, raa`` [rad], ``izaDeg, vzaDeg, raaDeg``, ``phi`` (raa folded into [0, 2 pi)),
    ``B = sec(mean iza) + sec(mean vza)``, ``normalize, nbar, angle_unit``.  Negative zenith angles are
    replaced by their absolute value with pi added to ``raa``.
    """

    def __init__(self, iza, vza, raa, normalize=False, nbar=0.0, angle_unit='DEG', align=True):
        if angle_unit != 'DEG' and angle_unit != 'RAD':
            raise AssertionError("angle_unit must be 'DEG' or 'RAD', but angle_unit is: {}".format(str(angle_unit)))
        self.normalize = normalize
        self.nbar = nbar
        self.angle_unit = angle_unit

        iza, vza, raa = asarrays((iza, vza, raa))
        if align:
            iza, vza, raa = align_all((iza, vza, raa))
        elif len(vza) != len(iza) or len(vza) != len(raa):
            raise AssertionError("Input dimensions must agree. The actual dimensions are iza: {0}, vza: {1} and "
                                 "raa: {2}".format(len(iza), len(vza), len(raa)))
        iza, vza, raa = (np.array(x, dtype=np.float64).flatten() for x in (iza, vza, raa))

        if normalize:           # nadir term appended (core/_core.py:127-132, 156-161)
            iza_m, vza_m = iza, vza
            iza = np.append(iza, nbar)
            vza = np.append(vza, 0.0)
            raa = np.append(raa, 0.0)
        else:
            iza_m, vza_m = iza, vza

        if angle_unit == 'DEG':
            self.izaDeg, self.vzaDeg, self.raaDeg = iza.copy(), vza.copy(), raa.copy()
            self.B = sec(np.mean(rad(iza_m))) + sec(np.mean(rad(vza_m)))
            iza, vza, raa = rad(iza), rad(vza), rad(raa)
        else:
            self.B = sec(np.mean(iza_m)) + sec(np.mean(vza_m))

        w = np.where(vza < 0)[0]
        vza[w] = -vza[w]
        raa[w] = raa[w] + np.pi
        w = np.where(iza < 0)[0]
        iza[w] = -iza[w]
        raa[w] = raa[w] + np.pi
        self.iza, self.vza, self.raa = iza, vza, raa
        if angle_unit == 'RAD':
            self.izaDeg, self.vzaDeg, self.raaDeg = deg(iza), deg(vza), deg(raa)
        self.phi = np.abs((self.raa % (2. * np.pi)))

    def normalization(self, kernel=None, args=None):
        """core/_core.py:80-96."""
        if args is None and kernel is None:
            raise ValueError("kernel or/ and args must be defined.")
        if args is None:
            return kernel - kernel[-1]
        if kernel is None:
            return [item[0:-1] for item in args]
        kernel = kernel - kernel[-1]
        return [item[0:-1] for item in tuple(list(args) + [kernel])]
