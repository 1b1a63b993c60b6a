# -*- coding: utf-8 -*-
"""Host-side engine: owns one ``prosail_handle_t`` per device and turns numpy (or raw device) arrays into
calls of the C ABI.  It contains no arithmetic of the hot path -- only argument marshalling, broadcasting of
scalar parameters, and buffer management.
"""
import ctypes as C
import threading
import weakref

import numpy as np

from . import _capi as capi
from . import constants as K

_f64 = np.float64
_f32 = np.float32


def _real(dtype):
    """Result dtype of a call: float64 (the parity path) or float32 (the fp32 mode, <= 1e-5 absolute)."""
    dt = np.dtype(dtype)
    if dt not in (np.dtype(_f64), np.dtype(_f32)):
        raise ValueError("dtype must be float64 or float32, not %s" % dt)
    return dt


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _vec(x, n, name):
    """Broadcast a scalar / length-1 / length-n parameter to a contiguous float64[n]."""
    a = np.asarray(x, dtype=_f64).reshape(-1)
    if a.size == 1 and n != 1:
        a = np.full(n, a[0], dtype=_f64)
    if a.size != n:
        raise ValueError("parameter %s has %d values, expected 1 or %d" % (name, a.size, n))
    return np.ascontiguousarray(a)


def batch_size(*params):
    """Common length of array-valued parameters (1 if all are scalars)."""
    n = 1
    for p in params:
        if p is None:
            continue
        s = np.size(p)
        if s != 1:
            if n != 1 and s != n:
                raise ValueError("array-valued parameters must share one length (%d vs %d)" % (n, s))
            n = s
    return n


class PinnedPool(object):
    """Page-locked host arrays (cudaHostAlloc) exposed as numpy arrays; freed when the last view dies."""

    SMALL = 1 << 20      # below this a pageable numpy array is cheaper than cudaHostAlloc
    MAX_PINNED = 8 << 30  # one request above this is not page-locked (a 1e5-set SAIL call with all planes asks for ~20 GB)

    @staticmethod
    def empty(shape, dtype=_f64, force=False):
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        if nbytes < PinnedPool.SMALL and not force:
            return np.empty(shape, dtype=dtype)
        lib = capi.load()
        p = C.c_void_p()
        if nbytes > PinnedPool.MAX_PINNED or lib.prosail_host_alloc(nbytes, C.byref(p)) != 0 or not p.value:
            if force:
                raise capi.ProsailError(-1, "cudaHostAlloc of %d bytes failed: %s" % (nbytes, (lib.prosail_last_error() or b"").decode("utf-8", "replace")))
            return np.empty(shape, dtype=dtype)         # pageable memory: slower copies, same results
        buf = (C.c_char * nbytes).from_address(p.value)
        weakref.finalize(buf, lib.prosail_host_free, C.c_void_p(p.value))
        return np.frombuffer(buf, dtype=dtype).reshape(shape)


class DeviceArray(object):
    """A float64 device buffer owned by an Engine (used by the device-resident API and the bench)."""

    def __init__(self, engine, shape, dtype=_f64):
        self.engine = engine
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = C.c_void_p()
        capi.check(engine.lib.prosail_dev_alloc(engine.h, self.nbytes, C.byref(p)))
        self.ptr = p.value
        self._fin = weakref.finalize(self, DeviceArray._free, engine.lib, engine.h, self.ptr)

    @staticmethod
    def _free(lib, h, ptr):
        try:
            lib.prosail_dev_free(h, C.c_void_p(ptr))
        except Exception:
            pass

    def upload(self, host, stream=None):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        assert host.nbytes == self.nbytes, (host.nbytes, self.nbytes)
        capi.check(self.engine.lib.prosail_memcpy_h2d(self.engine.h, C.c_void_p(self.ptr), _ptr(host), self.nbytes, stream))
        self.engine.sync(stream)
        return self

    def download(self, out=None, stream=None):
        if out is None:
            out = np.empty(self.shape, dtype=self.dtype)
        capi.check(self.engine.lib.prosail_memcpy_d2h(self.engine.h, _ptr(out), C.c_void_p(self.ptr), self.nbytes, stream))
        self.engine.sync(stream)
        return out

    def download_rows(self, row0, nrows, stream=None):
        """Rows [row0, row0 + nrows) of a 2-D (or higher) buffer as a host array (one D2H copy of just those rows)."""
        row0, nrows = int(row0), int(nrows)
        if not (0 <= row0 and row0 + nrows <= self.shape[0]):
            raise IndexError("rows %d..%d outside a buffer of %d rows" % (row0, row0 + nrows, self.shape[0]))
        out = np.empty((nrows,) + self.shape[1:], dtype=self.dtype)
        rowbytes = self.nbytes // max(self.shape[0], 1)
        if nrows:
            capi.check(self.engine.lib.prosail_memcpy_d2h(self.engine.h, _ptr(out), C.c_void_p(self.ptr + row0 * rowbytes), nrows * rowbytes, stream))
            self.engine.sync(stream)
        return out

    @property
    def __cuda_array_interface__(self):
        """Zero-copy hand-over to torch / cupy (torch.as_tensor(buf, device='cuda')): the optional torch plumbing of lut / bench."""
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (int(self.ptr), False), "version": 2, "strides": None}

    def free(self):
        self._fin()


class Engine(object):
    """One CUDA handle + uploaded constants on one device."""

    def __init__(self, device=0, alpha=40, windows=K.DEFAULT_WINDOWS):
        self.lib = capi.load()
        h = C.c_void_p()
        capi.check(self.lib.prosail_create(int(device), C.byref(h)))
        self.h = h
        self.device = int(device)
        self._fin = weakref.finalize(self, self.lib.prosail_destroy, h)
        self._alpha = {}
        self._lock = threading.RLock()
        self.numa_node = self._bind_to_gpu_numa_node()
        capi.check(self.lib.prosail_set_soil_tables(self.h, _ptr(K.lib.soil.rsoil1), _ptr(K.lib.soil.rsoil2)))
        self.set_alpha('5', alpha)
        self.set_alpha('D', alpha)
        self.windows = None
        self.srf = False
        self._weights = None
        self.set_windows(windows)

    def _bind_to_gpu_numa_node(self):
        """Multi-GPU boxes have two sockets: pin this process to the CPUs of the NUMA node the GPU is attached to, so that
        the pinned host buffers it allocates afterwards (first touch) and the D2H copies into them stay on that socket.
        Only done when several ranks share the box (LOCAL_RANK set) or PYRISM_B200_NUMA=1; never fatal."""
        import os
        if os.environ.get("PYRISM_B200_NUMA", "1" if "LOCAL_RANK" in os.environ else "0") != "1":
            return None
        try:
            buf = C.create_string_buffer(32)
            capi.check(self.lib.prosail_device_pci_bus_id(self.h, buf, 32))
            bus = buf.value.decode().lower()
            node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
            if node < 0:
                return None
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
            return node
        except Exception:
            return None

    # ---- constants -------------------------------------------------------------------------------------
    def set_alpha(self, version, alpha):
        """(Re)upload the coefficient set of ``version`` with the interface constants for ``alpha``."""
        with self._lock:
            if self._alpha.get(version) == float(alpha):
                return
            t = K.lib.p5 if version == '5' else K.lib.pd
            talf, t12, t21 = K.interface_constants(version, alpha)
            kan = t.Kan if version == 'D' else None
            capi.check(self.lib.prosail_set_leaf_tables(self.h, 0 if version == '5' else 1, _ptr(t.Kab), _ptr(t.Kxc), _ptr(t.Kbr),
                                                        _ptr(t.Kw), _ptr(t.Km), _ptr(kan), _ptr(talf), _ptr(t12), _ptr(t21)))
            self._alpha[version] = float(alpha)

    def set_windows(self, windows, weights=None):
        """Inclusive (lo_nm, hi_nm) windows for the band means (default: 6 Landsat-8 + 9 ASTER windows).

        weights: None (plain means, the reference's behaviour) or one array per window with hi - lo + 1 spectral-response
        weights (1 nm steps); the window value is then sum(w x) / sum(w)."""
        with self._lock:
            windows = tuple((int(lo), int(hi)) for lo, hi in windows)
            lo = np.ascontiguousarray([w[0] for w in windows], dtype=np.int32)
            hi = np.ascontiguousarray([w[1] for w in windows], dtype=np.int32)
            if weights is None:
                if windows == self.windows and not self.srf:
                    return
                capi.check(self.lib.prosail_set_windows(self.h, len(windows), _ptr(lo), _ptr(hi)))
                self.srf = False
                self._weights = None
            else:
                if len(weights) != len(windows):
                    raise ValueError("one weight array per window")
                ws = [np.asarray(w, dtype=_f64).reshape(-1) for w in weights]
                for (a, b), w in zip(windows, ws):
                    if w.size != b - a + 1:
                        raise ValueError("window (%d, %d) needs %d weights, got %d" % (a, b, b - a + 1, w.size))
                flat = np.ascontiguousarray(np.concatenate(ws))
                capi.check(self.lib.prosail_set_windows_srf(self.h, len(windows), _ptr(lo), _ptr(hi), _ptr(flat)))
                self.srf = True
                self._weights = ws
            self.windows = windows

    # ---- helpers ---------------------------------------------------------------------------------------
    def sync(self, stream=None):
        capi.check(self.lib.prosail_sync(self.h, stream))

    def launch_count(self):
        return int(self.lib.prosail_launch_count(self.h))

    def set_prologue_threshold(self, min_items=2048):
        """Batches with at least ``min_items`` (set x geometry) items use the thread-per-item prologue kernel."""
        capi.check(self.lib.prosail_set_prologue_threshold(self.h, int(min_items)))

    def profile(self, on=True):
        capi.check(self.lib.prosail_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        """-> {'bands': (ms, launches), 'prologue': (...), 'finalize': (...)} since the last read."""
        ms = (C.c_double * 3)()
        cnt = (C.c_longlong * 3)()
        capi.check(self.lib.prosail_profile_read(self.h, ms, cnt))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(('bands', 'prologue', 'finalize'))}

    def device_info(self):
        name = C.create_string_buffer(128)
        sm = C.c_int()
        mem = C.c_longlong()
        capi.check(self.lib.prosail_device_info(self.h, name, 128, C.byref(sm), C.byref(mem)))
        return dict(name=name.value.decode(), sm_count=sm.value, mem_bytes=mem.value)

    def measure_fp64_peak(self, millis=200):
        t = C.c_double()
        capi.check(self.lib.prosail_measure_fp64_peak(self.h, int(millis), C.byref(t)))
        return t.value

    def test_math(self, op, x):
        x = np.ascontiguousarray(x, dtype=_f64)
        y = np.empty_like(x)
        capi.check(self.lib.prosail_test_math(self.h, int(op), x.size, _ptr(x), _ptr(y), 1))
        return y

    def empty(self, shape, dtype=_f64):
        return DeviceArray(self, shape, dtype)

    def to_device(self, host):
        host = np.ascontiguousarray(host, dtype=_f64)
        return DeviceArray(self, host.shape).upload(host)

    def timer_begin(self, stream=None):
        capi.check(self.lib.prosail_timer_begin(self.h, stream))

    def timer_end(self, stream=None):
        ms = C.c_float()
        capi.check(self.lib.prosail_timer_end(self.h, stream, C.byref(ms)))
        return ms.value

    # ---- struct builders (work for host numpy arrays and for DeviceArray) --------------------------------
    @staticmethod
    def _addr(a):
        if a is None:
            return None
        if isinstance(a, DeviceArray):
            return a.ptr
        if isinstance(a, int):
            return a
        return a.ctypes.data

    def _leaf(self, N, Cab, Cxc, Cbr, Cw, Cm, Can):
        return capi.Leaf(*(self._addr(v) for v in (N, Cab, Cxc, Cbr, Cw, Cm, Can)))

    def _soil(self, refl, moist):
        return capi.Soil(self._addr(refl), self._addr(moist))

    def _canopy(self, lidf_type, lidf_a, lidf_b, lai, hotspot, n_geom, geom_per_sample, iza, vza, raa):
        return capi.Canopy(int(lidf_type), self._addr(lidf_a), self._addr(lidf_b), self._addr(lai), self._addr(hotspot),
                           int(n_geom), int(geom_per_sample), self._addr(iza), self._addr(vza), self._addr(raa))

    def _out(self, planes, pitch, mean_mask, means, geom, windows_only):
        o = capi.Out()
        for i in range(capi.NOUT):
            o.plane[i] = self._addr(planes.get(i)) if planes else None
        o.pitch = int(pitch)
        o.mean_mask = int(mean_mask)
        o.means = self._addr(means)
        o.geom = self._addr(geom)
        o.windows_only = int(windows_only)
        return o

    # ---- host-array API ----------------------------------------------------------------------------------
    def _fn(self, name, dt):
        return getattr(self.lib, "%s_%s" % (name, "f32" if dt == np.dtype(_f32) else "f64"))

    _LEAF_PLANES = {'ks': capi.O_LEAF_R, 'kt': capi.O_LEAF_T, 'ka': capi.O_LEAF_KA, 'ke': capi.O_LEAF_KE, 'om': capi.O_LEAF_OM,
                    'r': capi.O_LAYER_R, 't': capi.O_LAYER_T, 'Ra': capi.O_LAYER_RA, 'Ta': capi.O_LAYER_TA, 'denom': capi.O_LAYER_DENOM}

    def prospect(self, N, Cab, Cxc, Cbr, Cw, Cm, Can=None, version='5', alpha=40, means=False, out=None, dtype=_f64,
                 derived=False, spectra=True, layer=False):
        """Batched PROSPECT.  Returns dict of spectra dtype[n, 2101] (+ 'means' float32[n, 5, nw]).

        spectra: ks, kt; derived: ka, ke, om (models.py:1066-1068); layer: the one-plate r, t, Ra, Ta, denom (models.py:959).
        """
        dt = _real(dtype)
        with self._lock:
            self.set_alpha(version, alpha)
            n = batch_size(N, Cab, Cxc, Cbr, Cw, Cm, Can)
            v = [_vec(p, n, nm) for p, nm in ((N, 'N'), (Cab, 'Cab'), (Cxc, 'Cxc'), (Cbr, 'Cbr'), (Cw, 'Cw'), (Cm, 'Cm'))]
            can = _vec(Can, n, 'Can') if (version == 'D') else None
            leaf = self._leaf(v[0], v[1], v[2], v[3], v[4], v[5], can)
            out = out or {}
            res = {}
            want = (('ks', 'kt') if spectra else ()) + (('ka', 'ke', 'om') if derived else ()) + (('r', 't', 'Ra', 'Ta', 'denom') if layer else ())
            for name in want:
                a = out.get(name)
                res[name] = a if a is not None else PinnedPool.empty((n, K.NBANDS), dt)
                assert res[name].dtype == dt
            m = PinnedPool.empty((n, 5, len(self.windows)), np.float32) if means else None
            o = self._out({self._LEAF_PLANES[k]: a for k, a in res.items()}, K.NBANDS, 0, None, None, 0)
            capi.check(self._fn("prosail_prospect", dt)(self.h, 0 if version == '5' else 1, n, C.byref(leaf), C.byref(o), _ptr(m), 1, None))
            if means:
                res['means'] = m
            return res

    def soil(self, reflectance, moisture, means=False, dtype=_f64):
        """Batched LSM.  Returns dict(rho dtype[n, 2101][, means float32[n, 1, nw]])."""
        dt = _real(dtype)
        with self._lock:
            n = batch_size(reflectance, moisture)
            r, mo = _vec(reflectance, n, 'reflectance'), _vec(moisture, n, 'moisture')
            soil = self._soil(r, mo)
            rho = PinnedPool.empty((n, K.NBANDS), dt)
            m = PinnedPool.empty((n, 1, len(self.windows)), np.float32) if means else None
            capi.check(self._fn("prosail_soil", dt)(self.h, n, C.byref(soil), _ptr(rho), K.NBANDS, _ptr(m), 1, None))
            res = dict(rho=rho)
            if means:
                res['means'] = m
            return res

    def _canopy_host(self, n, iza, vza, raa, lai, hotspot, lidf_type, a, b, geom_per_sample):
        iza = np.ascontiguousarray(np.asarray(iza, dtype=_f64).reshape(-1))
        vza = np.ascontiguousarray(np.asarray(vza, dtype=_f64).reshape(-1))
        raa = np.ascontiguousarray(np.asarray(raa, dtype=_f64).reshape(-1))
        if not (iza.size == vza.size == raa.size):
            raise ValueError("iza, vza, raa must have the same length")
        if geom_per_sample:
            if iza.size != n:
                raise ValueError("per-sample geometry needs %d angles" % n)
            G = 1
        else:
            G = iza.size
        lt = {'campbell': capi.LIDF_CAMPBELL, 'verhoef': capi.LIDF_VERHOEF}[lidf_type]
        keep = [iza, vza, raa, _vec(a, n, 'a'), _vec(0.0 if b is None else b, n, 'b'), _vec(lai, n, 'lai'),
                _vec(hotspot, n, 'hotspot')]
        can = self._canopy(lt, keep[3], keep[4], keep[5], keep[6], G, 1 if geom_per_sample else 0, iza, vza, raa)
        return can, G, keep

    def volscatt(self, iza, vza, raa, lidf_type='campbell', a=57, b=0, lai=0.0, hotspot=0.0, geom_per_sample=False):
        """Batched VolScatt.coef: returns (geom float64[n, G, 12], lidf float64[n, 18]); angles in normalised radians.
        geom[..., :] = ks, ko, bf, Fs, Ft, tss, too, tsstoo, chi_s, chi_o, frho, ftau (the last four at the last leaf bin)."""
        with self._lock:
            n = batch_size(a, b, lai, hotspot) if not geom_per_sample else max(batch_size(a, b, lai, hotspot), np.size(iza))
            can, G, keep = self._canopy_host(n, iza, vza, raa, lai, hotspot, lidf_type, a, b, geom_per_sample)
            geom = np.empty((n, G, capi.NGEOM), dtype=_f64)
            lidf = np.empty((n, capi.NLIDF), dtype=_f64)
            capi.check(self.lib.prosail_volscatt_f64(self.h, n, C.byref(can), _ptr(geom), _ptr(lidf), 1, None))
            return geom, lidf

    def volscatt_bins(self, n_elements, lidf_type, a, b=0.0, iza=None, vza=None, raa=None):
        """LIDF / VolScatt.coef with ``n_elements`` leaf-inclination classes instead of SAIL's 18 (the reference's n_elements
        argument).  Returns (geom float64[n, G, 12] or None when no angles are given, lidf float64[n, n_elements])."""
        with self._lock:
            n = batch_size(a, b)
            av, bv = _vec(a, n, 'a'), _vec(0.0 if b is None else b, n, 'b')
            lt = {'campbell': capi.LIDF_CAMPBELL, 'verhoef': capi.LIDF_VERHOEF}[lidf_type]
            lidf = np.empty((n, int(n_elements)), dtype=_f64)
            geom, G, ang = None, 0, [None, None, None]
            if iza is not None:
                ang = [np.ascontiguousarray(np.asarray(x, dtype=_f64).reshape(-1)) for x in (iza, vza, raa)]
                if not (ang[0].size == ang[1].size == ang[2].size):
                    raise ValueError("iza, vza, raa must have the same length")
                G = ang[0].size
                geom = np.empty((n, G, capi.NGEOM), dtype=_f64)
            capi.check(self.lib.prosail_volscatt_bins_f64(self.h, n, lt, _ptr(av), _ptr(bv), G, _ptr(ang[0]), _ptr(ang[1]), _ptr(ang[2]),
                                                          int(n_elements), _ptr(geom), _ptr(lidf), 1, None))
            return geom, lidf

    def volume(self, iza, vza, raa, lza_deg):
        """VolScatt.volume for one leaf zenith angle (degrees): float64 [G, 4] = chi_s, chi_o, frho, ftau; angles in normalised radians."""
        with self._lock:
            a = [np.ascontiguousarray(np.asarray(x, dtype=_f64).reshape(-1)) for x in (iza, vza, raa)]
            if not (a[0].size == a[1].size == a[2].size):
                raise ValueError("iza, vza, raa must have the same length")
            out = np.empty((a[0].size, 4), dtype=_f64)
            capi.check(self.lib.prosail_volume_f64(self.h, a[0].size, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), float(lza_deg), _ptr(out), 1, None))
            return out

    def _alloc_out(self, n_items, planes, mean_planes, want_geom, out, dt=np.dtype(_f64)):
        out = out or {}
        bufs = {}
        for p in planes:
            a = out.get(capi.PLANE_NAMES[p])
            bufs[p] = a if a is not None else PinnedPool.empty((n_items, K.NBANDS), dt)
            assert bufs[p].dtype == dt, "output buffer dtype does not match the requested precision"
        mask = 0
        for p in mean_planes:
            mask |= 1 << p
        nmean = bin(mask).count("1")
        means = (out.get('means') if out.get('means') is not None else
                 PinnedPool.empty((n_items, nmean, len(self.windows)), dt)) if nmean else None
        geom = np.empty((n_items, capi.NGEOM), dtype=_f64) if want_geom else None
        return bufs, mask, means, geom

    def sail(self, ks, kt, rho, iza, vza, raa, lai, hotspot, lidf_type='campbell', a=57, b=0, planes=(capi.O_RSOT,),
             mean_planes=(), geom_per_sample=False, want_geom=True, out=None, dtype=_f64):
        """Batched SAIL with leaf/soil spectra given as arrays ([2101] broadcast or [n, 2101]).

        Returns dict: plane name -> dtype[n, G, 2101]; 'means' dtype[n, G, nmean, nw]; 'geom' float64[n, G, 12].
        """
        dt = _real(dtype)
        with self._lock:
            specs = []
            n_spec = 1
            for name, s_ in (('ks', ks), ('kt', kt), ('rho_surface', rho)):
                s_ = np.ascontiguousarray(s_, dtype=dt)
                if s_.ndim == 1:
                    s_ = s_.reshape(1, -1)
                if s_.shape[-1] != K.NBANDS:
                    raise AssertionError("%s must have 2101 values per spectrum" % name)
                if s_.shape[0] != 1:
                    n_spec = s_.shape[0]
                specs.append(s_)
            n = batch_size(lai, hotspot, a, b)
            if geom_per_sample:
                n = max(n, np.size(iza))
            n = max(n, n_spec)
            for s_ in specs:
                if s_.shape[0] not in (1, n):
                    raise ValueError("spectra must be [2101] or [%d, 2101]" % n)
            can, G, keep = self._canopy_host(n, iza, vza, raa, lai, hotspot, lidf_type, a, b, geom_per_sample)
            bufs, mask, means, geom = self._alloc_out(n * G, planes, mean_planes, want_geom, out, dt)
            o = self._out(bufs, K.NBANDS, mask, means, geom, 0)
            pit = [0 if s_.shape[0] == 1 else K.NBANDS for s_ in specs]
            capi.check(self._fn("prosail_sail", dt)(self.h, n, _ptr(specs[0]), pit[0], _ptr(specs[1]), pit[1], _ptr(specs[2]), pit[2],
                                                 C.byref(can), C.byref(o), 1, None))
            return self._pack(bufs, means, geom, n, G)

    def prosail(self, N, Cab, Cxc, Cbr, Cw, Cm, lai, hotspot, iza, vza, raa, soil_reflectance, soil_moisture, Can=None,
                version='5', alpha=40, lidf_type='campbell', a=57, b=0, planes=(capi.O_RSOT,), mean_planes=(),
                geom_per_sample=False, want_geom=False, windows_only=False, out=None, dtype=_f64):
        """Fused PROSPECT -> LSM -> SAIL for n parameter sets x G geometries (host arrays in, host arrays out)."""
        dt = _real(dtype)
        with self._lock:
            self.set_alpha(version, alpha)
            n = batch_size(N, Cab, Cxc, Cbr, Cw, Cm, Can, lai, hotspot, a, b, soil_reflectance, soil_moisture)
            if geom_per_sample:
                n = max(n, np.size(iza))
            v = [_vec(p, n, nm) for p, nm in ((N, 'N'), (Cab, 'Cab'), (Cxc, 'Cxc'), (Cbr, 'Cbr'), (Cw, 'Cw'), (Cm, 'Cm'))]
            can_ = _vec(Can, n, 'Can') if version == 'D' else None
            leaf = self._leaf(v[0], v[1], v[2], v[3], v[4], v[5], can_)
            sr, smo = _vec(soil_reflectance, n, 'soil_reflectance'), _vec(soil_moisture, n, 'soil_moisture')
            soil = self._soil(sr, smo)
            can, G, keep = self._canopy_host(n, iza, vza, raa, lai, hotspot, lidf_type, a, b, geom_per_sample)
            if windows_only:
                planes = ()
            bufs, mask, means, geom = self._alloc_out(n * G, planes, mean_planes, want_geom, out, dt)
            o = self._out(bufs, K.NBANDS, mask, means, geom, 1 if windows_only else 0)
            capi.check(self._fn("prosail_prosail", dt)(self.h, 0 if version == '5' else 1, n, C.byref(leaf), C.byref(soil), C.byref(can),
                                                    C.byref(o), 1, None))
            return self._pack(bufs, means, geom, n, G)

    # ---- device-resident API (inputs and outputs stay in HBM; asynchronous on `stream`) ---------------------
    def prosail_device(self, n, leaf, soil, canopy, planes, pitch=K.NBANDS, mean_planes=(), means=None, version='5',
                       lidf_type='campbell', n_geom=1, geom_per_sample=False, windows_only=False, stream=None, dtype=_f64):
        """Fused PROSAIL on DeviceArrays.

        leaf: dict N, Cab, Cxc, Cbr, Cw, Cm[, Can]; soil: dict reflectance, moisture; canopy: dict lidf_a[, lidf_b], lai,
        hotspot, iza, vza, raa (normalised radians); planes: {plane index: DeviceArray [n*n_geom, pitch]}.
        """
        lf = self._leaf(leaf['N'], leaf['Cab'], leaf['Cxc'], leaf['Cbr'], leaf['Cw'], leaf['Cm'], leaf.get('Can'))
        so = self._soil(soil['reflectance'], soil['moisture'])
        lt = {'campbell': capi.LIDF_CAMPBELL, 'verhoef': capi.LIDF_VERHOEF}[lidf_type]
        ca = self._canopy(lt, canopy['lidf_a'], canopy.get('lidf_b'), canopy['lai'], canopy['hotspot'], n_geom,
                          1 if geom_per_sample else 0, canopy['iza'], canopy['vza'], canopy['raa'])
        mask = 0
        for p in mean_planes:
            mask |= 1 << p
        o = self._out(planes, pitch, mask, means, None, 1 if windows_only else 0)
        capi.check(self._fn("prosail_prosail", _real(dtype))(self.h, 0 if version == '5' else 1, int(n), C.byref(lf), C.byref(so),
                                                             C.byref(ca), C.byref(o), 0, stream))

    def prospect_device(self, n, leaf, ks, kt, pitch=K.NBANDS, version='5', stream=None, dtype=_f64):
        """Leaf-only batch on DeviceArrays (ks, kt: [n, pitch])."""
        lf = self._leaf(leaf['N'], leaf['Cab'], leaf['Cxc'], leaf['Cbr'], leaf['Cw'], leaf['Cm'], leaf.get('Can'))
        o = self._out({capi.O_LEAF_R: ks, capi.O_LEAF_T: kt}, pitch, 0, None, None, 0)
        capi.check(self._fn("prosail_prospect", _real(dtype))(self.h, 0 if version == '5' else 1, int(n), C.byref(lf), C.byref(o), None, 0, stream))

    @staticmethod
    def _pack(bufs, means, geom, n, G):
        res = {}
        for p, a in bufs.items():
            res[capi.PLANE_NAMES[p]] = a.reshape(n, G, K.NBANDS)
        if means is not None:
            res['means'] = means.reshape(n, G, means.shape[-2], means.shape[-1])
        if geom is not None:
            res['geom'] = geom.reshape(n, G, capi.NGEOM)
        return res


_engines = {}
_engines_lock = threading.Lock()


def get_engine(device=None, role='main'):
    """Process-wide engine for ``device`` (default: env PYRISM_B200_DEVICE or 0).

    ``role='select'`` returns a second handle on the same device that `PROSPECT.select / LSM.select / indices` use for their
    one-window means: its window list changes with every call, the main engine's never does."""
    import os
    if device is None:
        device = int(os.environ.get("PYRISM_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    key = device if role == 'main' else (device, role)
    with _engines_lock:
        if key not in _engines:
            _engines[key] = Engine(device)
        return _engines[key]
