#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Benchmark of the batched PROSAIL hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 0..4] [--samples S]

``--config`` is the index into BASELINE.json ``configs`` (default 2, the configuration the metric "PROSAIL spectra/s
(2101 bands, fp64)" is quoted on; BENCH / SCALE records use the default):

  0  README example: PROSPECT + LSM + 4SAIL through the scalar drop-in classes (latency of one call chain)
  1  PROSPECT-5 leaf-only batch: 1e7 parameter sets x 2101 bands (ks + kt), streamed through a 1e6-set output slab
  2  PROSAIL LUT: 1e6 Latin-hypercube sets per GPU x 2101 bands, BRF spectrum, fp64 (and the fp32 mode as a second figure)
  3  multi-angle 4SAIL: 1e5 sets x 64 geometries, Verhoef LIDF, sharded over the ranks (STRONG scaling: total work fixed)
  4  L8/ASTER band-mean LUT: 1e8 sets over 8 GPUs = 1.25e7 per GPU, fused band means only; plus the timed final
     all-gather of the table (NCCL) and one sharded LUT inversion

One "step" = one pass of the hot path over this rank's synthetic batch (prologue + band kernel).  Multi-GPU: every rank
owns its shard of the parameter sets; there is no data-path collective.

Printed JSON (rank 0): ``value`` = whole-job spectra/s with parameters resident in HBM and the output written to HBM;
``e2e`` = the same through the host-array API (numpy parameters in, numpy results out, H2D + D2H inside the timed
region); ``roofline`` = the band kernel against the measured FP64 (DFMA) peak with the HBM view beside it;
``cpu_baseline`` = the reference itself (oracle/_ref, ``kind: "reference"``) on the host cores, the numpy port as a second
figure; ``parity_max_abs`` = max |GPU - oracle| over rows of the buffer the TIMED loop wrote.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NB = 2101
NB_WINDOWS = 782                # wavelengths inside at least one L8/ASTER window
FLOPS_LEAF, FLOPS_SAIL = 83, 151
FLOPS_PER_BAND = FLOPS_LEAF + FLOPS_SAIL     # div/sqrt/exp/E1/pow counted as ONE flop each
BYTES_PER_SPECTRUM = NB * 8     # one fp64 output plane
# dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of the config-2 kernel, per spectrum
# (round 2n: 4.49 GB written + 1.22 GB read.  The reads are the 256-B item records, staged by each of the 22 tile groups at a different
#  time since every warp walks a contiguous share of the work list -- 4.7 KB per spectrum of L2 misses at 10 % DRAM
#  utilisation; the strided walk of round 1/2k read them once (83.5 MB, ratio 1.010) and was 1 % slower.)
NCU_DRAM_BYTES_PER_SPECTRUM = (4.491783e9 + 1.220478e9) / 262144
NCU_DRAM_SOURCE = "profiles/r2n_k_bands_brf_f64.txt (one ncu --set full capture of 262144 sets, scaled; NOT measured in this run)"
BASE_SEED = 20261019
SEED = BASE_SEED + 2            # config index 2
GEOM_DEG = (35.0, 30.0, 50.0)
METRIC = "PROSAIL spectra/s (2101 bands, fp64)"

# SURVEY.md 8(d) bounds: N Cab Cxc Cbr Cw Cm lai hotspot campbell_a soil_refl soil_moist
LO = np.array([1.0, 5.0, 1.0, 0.0, 0.001, 0.001, 0.1, 0.01, 20.0, 0.3, 0.0])
HI = np.array([3.0, 90.0, 25.0, 1.0, 0.05, 0.02, 8.0, 0.5, 80.0, 1.2, 1.0])
NAMES = ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm", "lai", "hotspot", "lidf_a", "soil_refl", "soil_moist")
LEAF_NAMES = NAMES[:6]
# config 3: Verhoef a, b instead of the Campbell angle
LO3 = np.array([1.0, 5.0, 1.0, 0.0, 0.001, 0.001, 0.1, 0.01, -0.6, -0.35, 0.3, 0.0])
HI3 = np.array([3.0, 90.0, 25.0, 1.0, 0.05, 0.02, 8.0, 0.5, 0.6, 0.35, 1.2, 1.0])
NAMES3 = ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm", "lai", "hotspot", "lidf_a", "lidf_b", "soil_refl", "soil_moist")


def _lhs(n, lo, hi, names, seed):
    from scipy.stats import qmc
    x = qmc.scale(qmc.LatinHypercube(d=len(lo), seed=seed).random(int(n)), lo, hi)
    return {k: np.ascontiguousarray(x[:, i]) for i, k in enumerate(names)}


def lhs(n, seed):
    """Latin-hypercube parameter sets of the PROSAIL LUT configurations (2 and 4)."""
    return _lhs(n, LO, HI, NAMES, seed)


def draw(cfg, n, seed):
    if cfg == 1:
        return _lhs(n, LO[:6], HI[:6], LEAF_NAMES, seed)
    if cfg == 3:
        return _lhs(n, LO3, HI3, NAMES3, seed)
    return lhs(n, seed)


def geometry_grid():
    """config 3: iza in {15,30,45,60} x vza in {0,15,30,45} x raa in {0,60,120,180}."""
    iza, vza, raa = np.meshgrid([15.0, 30.0, 45.0, 60.0], [0.0, 15.0, 30.0, 45.0], [0.0, 60.0, 120.0, 180.0], indexing="ij")
    return iza.ravel(), vza.ravel(), raa.ravel()


WORKLOADS = {
    0: "BASELINE configs[0]: README example PROSPECT-5 (N=1.5,Cab=35,Cxc=5,Cbr=0.15,Cw=0.003,Cm=0.0055) + LSM + 4SAIL campbell, "
       "lai=3, hotspot=0.25, 35/30/50 deg, scalar drop-in calls (host in / host out)",
    1: "BASELINE configs[1]: PROSPECT-5 leaf-only batch, Latin-hypercube leaf parameter sets x 2101 bands, fp64, ks + kt planes",
    2: "BASELINE configs[2]: PROSAIL LUT, Latin-hypercube parameter sets x 2101 bands, fp64, PROSPECT-5 + "
       "LSM soil + 4SAIL campbell, geometry 35/30/50 deg, output BRF spectrum",
    3: "BASELINE configs[3]: multi-angle 4SAIL, 1e5 parameter sets x 64 sun/view geometries, Verhoef LIDF, PROSPECT-5 + LSM soil, "
       "fp64, output BRF spectrum per (set, geometry); sets sharded over the ranks",
    4: "BASELINE configs[4]: Landsat-8/ASTER band-averaged PROSAIL LUT, 1e8 sets over 8 GPUs (1.25e7 per GPU), fused band means of "
       "the BRF only (15 windows, 782 in-window wavelengths evaluated), fp64",
}


# ----------------------------------------------------------------------------------------------------------
# clocks sampled during the timed region
# ----------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []          # (host time, csv line)
        self.t_begin = None
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def begin(self):
        """Mark the start of the timed region: only samples taken from here on are reported."""
        self.t_begin = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.time() + 0.02
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ts, r in self.rows:
            if self.t_begin is not None and not (self.t_begin <= ts <= t_end):
                continue
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------
# CPU arms.  kind "reference": the UNMODIFIED pyrism from oracle/_ref (copied there by __graft_entry__.build(), shipped by
# gpurun), one constructor chain per sample exactly as examples/prospect_sail.py:9-28, np.float64 scalars.
# kind "port": oracle/prosail_oracle.py, the numpy restatement.  Both: one process per host core.
# ----------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    import warnings
    warnings.filterwarnings("ignore")
    kind, cfg, rows, names = args
    if kind == "reference":
        from oracle import ref_runner as rr
        py = rr.load()
    else:
        from oracle import prosail_oracle as orc
    f = np.float64
    acc = 0.0
    gi, gv, gr = geometry_grid()
    for r in rows:
        p = dict(zip(names, r))
        if cfg == 1:                         # leaf only
            if kind == "reference":
                acc += float(py.PROSPECT(N=f(p["N"]), Cab=f(p["Cab"]), Cxc=f(p["Cxc"]), Cbr=f(p["Cbr"]), Cw=f(p["Cw"]), Cm=f(p["Cm"])).ks[0])
            else:
                acc += float(orc.prospect(p["N"], p["Cab"], p["Cxc"], p["Cbr"], p["Cw"], p["Cm"])["ks"][0])
        elif cfg == 3:                       # PROSPECT + LSM once per set, SAIL once per geometry
            if kind == "reference":
                leaf = py.PROSPECT(N=f(p["N"]), Cab=f(p["Cab"]), Cxc=f(p["Cxc"]), Cbr=f(p["Cbr"]), Cw=f(p["Cw"]), Cm=f(p["Cm"]))
                soil = py.LSM(reflectance=f(p["soil_refl"]), moisture=f(p["soil_moist"]))
                for g in range(len(gi)):
                    c = py.SAIL(iza=f(gi[g]), vza=f(gv[g]), raa=f(gr[g]), ks=leaf.ks, kt=leaf.kt, rho_surface=soil.ref, lidf_type='verhoef',
                                a=f(p["lidf_a"]), b=f(p["lidf_b"]), lai=f(p["lai"]), hotspot=f(p["hotspot"]))
                    acc += float(c.BRF.ref[0])
            else:
                leaf = orc.prospect(p["N"], p["Cab"], p["Cxc"], p["Cbr"], p["Cw"], p["Cm"])
                soil = orc.lsm(p["soil_refl"], p["soil_moist"])
                for g in range(len(gi)):
                    c = orc.sail(gi[g], gv[g], gr[g], leaf["ks"], leaf["kt"], p["lai"], p["hotspot"], soil["ref"], lidf_type='verhoef',
                                 a=p["lidf_a"], b=p["lidf_b"])
                    acc += float(c["rsot"][0])
        else:                                # configs 0, 2, 4: the full chain, one geometry (band means are part of SAIL)
            if kind == "reference":
                _, _, c = rr.prosail(p["N"], p["Cab"], p["Cxc"], p["Cbr"], p["Cw"], p["Cm"], p["lai"], p["hotspot"], GEOM_DEG[0], GEOM_DEG[1],
                                     GEOM_DEG[2], p["soil_refl"], p["soil_moist"], lidf_type="campbell", a=p["lidf_a"])
                acc += float(c.BRF.ref[0])
            else:
                _, _, c = orc.prosail(p["N"], p["Cab"], p["Cxc"], p["Cbr"], p["Cw"], p["Cm"], p["lai"], p["hotspot"], GEOM_DEG[0], GEOM_DEG[1],
                                      GEOM_DEG[2], p["soil_refl"], p["soil_moist"], lidf_type="campbell", a=p["lidf_a"])
                acc += float(c["rsot"][0])
    return acc


def cpu_kind():
    from oracle import ref_runner as rr
    return "reference" if rr.available() else "port"


def cpu_units_per_row(cfg):
    return 64 if cfg == 3 else 1


def cpu_rows(cfg, n, seed):
    if cfg == 0:
        row = np.array([1.5, 35.0, 5.0, 0.15, 0.003, 0.0055, 3.0, 0.25, 57.0, 3.14 / 4, 0.15])
        return np.tile(row, (int(n), 1)), NAMES
    P = draw(cfg, n, seed)
    names = tuple(P.keys())
    return np.stack([P[k] for k in names], axis=1), names


def cpu_rate(kind, cfg, n_rows, cores, pool=None):
    """Spectra/s of the CPU arm on `cores` processes: (rate, seconds)."""
    import multiprocessing as mp
    rows, names = cpu_rows(cfg, n_rows, BASE_SEED + cfg + 777)
    chunks = [rows[i::cores] for i in range(cores)]
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(cores)
        pool.map(_cpu_worker, [(kind, cfg, c[:1], names) for c in chunks])      # import + table load outside the timed region
    t0 = time.perf_counter()
    pool.map(_cpu_worker, [(kind, cfg, c, names) for c in chunks])
    dt = time.perf_counter() - t0
    if own:
        pool.close()
    return n_rows * cpu_units_per_row(cfg) / dt, dt


CPU_WHAT = {"reference": "the unmodified reference (oracle/_ref/pyrism: PROSPECT -> LSM -> SAIL constructors incl. L8/ASTER band means, "
                         "np.float64 scalars, scipy.misc.factorial shim only)",
            "port": "oracle/prosail_oracle.py (numpy/scipy restatement of pyrism PROSPECT -> LSM -> SAIL incl. L8/ASTER means)"}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on all host cores (rank 0 only)."""
    if rank != 0:
        return
    import multiprocessing as mp
    cfg = args.config
    kind = cpu_kind()
    cores = os.cpu_count() or 1
    # rows per step: about one second of work per core at ~100 (reference) / ~400 (port) chains per second and core
    per_core = {"reference": 96, "port": 400}[kind]
    if cfg == 3:
        per_core = max(2, per_core // 48)
    if cfg == 1:
        per_core *= 3
    per_step = cores * per_core
    pool = mp.get_context("fork").Pool(cores)
    rows, names = cpu_rows(cfg, cores, 1)
    pool.map(_cpu_worker, [(kind, cfg, rows[i:i + 1], names) for i in range(cores)])      # imports and table load outside the timed steps
    for _ in range(args.warmup):
        cpu_rate(kind, cfg, per_step, cores, pool)
    t = 0.0
    for _ in range(args.steps):
        _, dt = cpu_rate(kind, cfg, per_step, cores, pool)
        t += dt
    pool.close()
    units = per_step * cpu_units_per_row(cfg)
    v = units * args.steps / t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "spectra/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "strong" if cfg == 3 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, per_step, args.gpus),
            "cpu_baseline": {"value": v, "unit": "spectra/s", "cores": cores, "kind": kind,
                             "sample": "%d parameter sets (%d spectra) per step through %s, one process per core" % (per_step, units, CPU_WHAT[kind])},
            "e2e": {"value": v, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(cfg, samples, gpus, extra=None):
    c = {"workload": WORKLOADS[cfg], "index": cfg, "samples_per_gpu": int(samples), "bands": NB,
         "geometries": 64 if cfg == 3 else 1,
         "output_planes": {0: ["rsot"], 1: ["ks", "kt"], 2: ["rsot"], 3: ["rsot"], 4: ["15 window means of rsot"]}[cfg],
         "parallelism": "dp%d (parameter-set shards, no data-path collective)" % gpus,
         "l2": "outputs and records are far larger than the 126 MB L2; no flush needed"}
    if cfg == 4:
        c["l2"] = "records (256 B per set) are far larger than the 126 MB L2; the 120 B of means per set are written once"
    if cfg == 0:
        c["l2"] = "latency measurement of single calls; not a throughput kernel"
    if extra:
        c.update(extra)
    return c


# ----------------------------------------------------------------------------------------------------------
# GPU workloads: each returns an object with step(), units (spectra per step on this rank), verify() -> (max_abs, rows),
# flops per spectrum, algorithmic bytes per spectrum, and e2e() -> dict
# ----------------------------------------------------------------------------------------------------------
class Workload(object):
    scaling = "weak"
    profile_key = "bands"

    def verify(self):
        return None, 0

    def e2e(self, args, barrier):
        return None

    def extra(self, args, dist, torch):
        return {}


def _oracle_rows(P, idx, version='5', lidf_type='campbell', geom=GEOM_DEG):
    from oracle import prosail_oracle as orc
    out = []
    for i in idx:
        kw = dict(a=P["lidf_a"][i]) if lidf_type == 'campbell' else dict(a=P["lidf_a"][i], b=P["lidf_b"][i])
        _, _, c = orc.prosail(P["N"][i], P["Cab"][i], P["Cxc"][i], P["Cbr"][i], P["Cw"][i], P["Cm"][i], P["lai"][i], P["hotspot"][i],
                              geom[0], geom[1], geom[2], P["soil_refl"][i], P["soil_moist"][i], version=version, lidf_type=lidf_type, **kw)
        out.append(c)
    return out


class LutBRF(Workload):
    """config 2: PROSAIL LUT, BRF plane."""

    def __init__(self, eng, capi, S, rank, dtype=np.float64):
        self.eng, self.capi, self.S, self.dt = eng, capi, S, dtype
        self.P = lhs(S, SEED + 1000 * rank)
        self.geom = np.deg2rad(np.array(GEOM_DEG))
        dev = {k: eng.to_device(v) for k, v in self.P.items()}
        self.leaf = {k: dev[k] for k in LEAF_NAMES}
        self.soil = {"reflectance": dev["soil_refl"], "moisture": dev["soil_moist"]}
        self.canopy = {"lidf_a": dev["lidf_a"], "lai": dev["lai"], "hotspot": dev["hotspot"], "iza": eng.to_device(self.geom[0:1]),
                       "vza": eng.to_device(self.geom[1:2]), "raa": eng.to_device(self.geom[2:3])}
        self.out = eng.empty((S, NB), dtype)
        self.units = S
        self.flops = NB * FLOPS_PER_BAND
        self.bytes = NB * np.dtype(dtype).itemsize
        self.kernel = "pb::k_bands<MODE_PROSAIL, planes = BRF>"

    def step(self):
        self.eng.prosail_device(self.S, self.leaf, self.soil, self.canopy, {self.capi.O_RSOT: self.out}, pitch=NB, version='5',
                                lidf_type='campbell', n_geom=1, dtype=self.dt)

    def verify(self):
        n = min(32, self.S // 2)
        idx = list(range(n)) + list(range(self.S - n, self.S))
        got = np.concatenate([self.out.download_rows(0, n), self.out.download_rows(self.S - n, n)])
        ref = np.stack([c["rsot"] for c in _oracle_rows(self.P, idx)])
        return float(np.max(np.abs(got.astype(np.float64) - ref))), len(idx)

    def e2e(self, args, barrier):
        eng, capi = self.eng, self.capi
        Se = min(args.e2e_samples, self.S)
        Pe = {k: pinned_copy(v[:Se]) for k, v in self.P.items()}      # the step's inputs live in page-locked host memory (allocated once)
        geom = self.geom
        from pyrism_b200.engine import PinnedPool
        obuf = {"rsot": PinnedPool.empty((Se, NB), force=True)}

        def spectra_step():
            r = eng.prosail(Pe["N"], Pe["Cab"], Pe["Cxc"], Pe["Cbr"], Pe["Cw"], Pe["Cm"], Pe["lai"], Pe["hotspot"], geom[0:1], geom[1:2],
                            geom[2:3], Pe["soil_refl"], Pe["soil_moist"], version='5', lidf_type='campbell', a=Pe["lidf_a"],
                            planes=(capi.O_RSOT,), out=obuf)
            return float(r["rsot"][Se - 1, 0, NB - 1])

        mbuf = {"means": PinnedPool.empty((Se, 1, len(eng.windows)), force=True)}      # allocated once, outside the timed region

        def means_step():
            r = eng.prosail(Pe["N"], Pe["Cab"], Pe["Cxc"], Pe["Cbr"], Pe["Cw"], Pe["Cm"], Pe["lai"], Pe["hotspot"], geom[0:1], geom[1:2],
                            geom[2:3], Pe["soil_refl"], Pe["soil_moist"], version='5', lidf_type='campbell', a=Pe["lidf_a"],
                            planes=(), mean_planes=(capi.O_RSOT,), windows_only=True, out=mbuf)
            return float(r["means"][Se - 1, 0, 0, 0])

        def timed(fn):
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                fn()
            barrier()
            return time.perf_counter() - t0
        dt = timed(spectra_step)
        dtm = timed(means_step)
        return dict(dt=dt, units=Se * args.e2e_steps, h2d=int(Se * len(NAMES) * 8 + 24), d2h=int(Se * NB * 8), samples=Se,
                    api="pyrism_b200.Engine.prosail (prosail_prosail_f64, on_host=1): numpy parameters in, pinned numpy BRF spectra out",
                    second=dict(dt=dtm, units=Se * args.e2e_steps, d2h=int(Se * 15 * 8),
                                api="the same call with windows_only=True: numpy parameters in, 15 L8/ASTER band means of the BRF out "
                                    "(120 B per spectrum: the product the LUT inversion consumes; not PCIe-bound)"))


class LeafOnly(Workload):
    """config 1: PROSPECT-5 leaf-only, n sets streamed through an output slab."""

    def __init__(self, eng, capi, n, rank):
        self.eng, self.capi = eng, capi
        self.n = n
        self.slab = min(n, 1000000)
        self.P = draw(1, n, BASE_SEED + 1 + 1000 * rank)
        self.dev = {k: eng.to_device(self.P[k]) for k in LEAF_NAMES}
        self.ks, self.kt = eng.empty((self.slab, NB)), eng.empty((self.slab, NB))
        self.units = n
        self.flops = NB * FLOPS_LEAF
        self.bytes = 2 * NB * 8
        self.kernel = "pb::k_bands<MODE_PROSPECT, planes = ks + kt>"

    def step(self):
        for s0 in range(0, self.n, self.slab):
            leaf = {k: self.dev[k].ptr + 8 * s0 for k in self.dev}
            self.eng.prospect_device(min(self.slab, self.n - s0), leaf, self.ks, self.kt)

    def verify(self):
        from oracle import prosail_oracle as orc
        last0 = ((self.n - 1) // self.slab) * self.slab          # the slab the last launch of a step wrote
        m = min(32, self.n - last0)
        got_r, got_t = self.ks.download_rows(0, m), self.kt.download_rows(0, m)
        worst = 0.0
        for j in range(m):
            i = last0 + j
            leaf = orc.prospect(*(self.P[k][i] for k in LEAF_NAMES))
            worst = max(worst, float(np.max(np.abs(got_r[j] - leaf["ks"]))), float(np.max(np.abs(got_t[j] - leaf["kt"]))))
        return worst, 2 * m

    def e2e(self, args, barrier):
        eng = self.eng
        Se = min(131072, self.n)
        Pe = {k: pinned_copy(v[:Se]) for k, v in self.P.items()}      # the step's inputs live in page-locked host memory (allocated once)

        from pyrism_b200.engine import PinnedPool
        obuf = {"ks": PinnedPool.empty((Se, NB), force=True), "kt": PinnedPool.empty((Se, NB), force=True)}   # allocated once, outside the timed region

        def fn():
            r = eng.prospect(Pe["N"], Pe["Cab"], Pe["Cxc"], Pe["Cbr"], Pe["Cw"], Pe["Cm"], version='5', out=obuf)
            return float(r["ks"][Se - 1, NB - 1])
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            fn()
        barrier()
        return dict(dt=time.perf_counter() - t0, units=Se * args.e2e_steps, h2d=int(Se * 6 * 8), d2h=int(Se * 2 * NB * 8), samples=Se,
                    api="pyrism_b200.Engine.prospect (prosail_prospect_f64, on_host=1): numpy leaf parameters in, ks + kt spectra out")


class MultiAngle(Workload):
    """config 3: n_total sets x 64 geometries, Verhoef; this rank owns a contiguous block of the sets (strong scaling)."""
    scaling = "strong"

    def __init__(self, eng, capi, n_total, rank, world):
        from pyrism_b200.lut import shard_bounds
        self.eng, self.capi = eng, capi
        self.G = 64
        Pall = draw(3, n_total, BASE_SEED + 3)                  # ONE hypercube for the whole job, sliced per rank
        b, e = shard_bounds(n_total, rank, world)
        self.P = {k: np.ascontiguousarray(v[b:e]) for k, v in Pall.items()}
        self.n = e - b
        self.slab = min(self.n, 12500)
        self.dev = {k: eng.to_device(v) for k, v in self.P.items()}
        gi, gv, gr = geometry_grid()
        self.gdeg = (gi, gv, gr)
        self.gd = {"iza": eng.to_device(np.deg2rad(gi)), "vza": eng.to_device(np.deg2rad(gv)), "raa": eng.to_device(np.deg2rad(gr))}
        self.out = eng.empty((self.slab * self.G, NB))
        self.units = self.n * self.G
        self.flops = NB * (FLOPS_SAIL + FLOPS_LEAF / 64.0)      # the reference too runs PROSPECT once per set and SAIL once per geometry
        self.bytes = NB * 8
        self.kernel = "pb::k_bands<MODE_PROSAIL, planes = BRF, geometry loop>"

    def _launch(self, s0, ns):
        sl = lambda k: self.dev[k].ptr + 8 * s0
        lf = {k: sl(k) for k in LEAF_NAMES}
        so = {"reflectance": sl("soil_refl"), "moisture": sl("soil_moist")}
        ca = dict(self.gd, lidf_a=sl("lidf_a"), lidf_b=sl("lidf_b"), lai=sl("lai"), hotspot=sl("hotspot"))
        self.eng.prosail_device(ns, lf, so, ca, {self.capi.O_RSOT: self.out}, lidf_type='verhoef', n_geom=self.G)

    def step(self):
        for s0 in range(0, self.n, self.slab):
            self._launch(s0, min(self.slab, self.n - s0))

    def verify(self):
        last0 = ((self.n - 1) // self.slab) * self.slab
        got = self.out.download_rows(0, self.G)                 # all 64 geometries of the first set of the last slab
        worst = 0.0
        for g in range(0, self.G, 3):
            c = _oracle_rows(self.P, [last0], lidf_type='verhoef', geom=(self.gdeg[0][g], self.gdeg[1][g], self.gdeg[2][g]))[0]
            worst = max(worst, float(np.max(np.abs(got[g] - c["rsot"]))))
        return worst, len(range(0, self.G, 3))

    def e2e(self, args, barrier):
        eng, capi = self.eng, self.capi
        Se = min(2048, self.n)
        Pe = {k: pinned_copy(v[:Se]) for k, v in self.P.items()}      # the step's inputs live in page-locked host memory (allocated once)
        gi, gv, gr = (np.deg2rad(x) for x in self.gdeg)
        from pyrism_b200.engine import PinnedPool
        obuf = {"rsot": PinnedPool.empty((Se * 64, NB), force=True)}            # allocated once, outside the timed region

        def fn():
            r = eng.prosail(Pe["N"], Pe["Cab"], Pe["Cxc"], Pe["Cbr"], Pe["Cw"], Pe["Cm"], Pe["lai"], Pe["hotspot"], gi, gv, gr,
                            Pe["soil_refl"], Pe["soil_moist"], version='5', lidf_type='verhoef', a=Pe["lidf_a"], b=Pe["lidf_b"],
                            planes=(capi.O_RSOT,), out=obuf)
            return float(r["rsot"][Se - 1, 63, NB - 1])
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            fn()
        barrier()
        return dict(dt=time.perf_counter() - t0, units=Se * 64 * args.e2e_steps, h2d=int(Se * len(NAMES3) * 8 + 3 * 64 * 8),
                    d2h=int(Se * 64 * NB * 8), samples=Se,
                    api="pyrism_b200.Engine.prosail (prosail_prosail_f64, on_host=1): numpy parameters + 64 geometries in, BRF spectra out")


class BandMeanLut(Workload):
    """config 4: band-mean LUT, n sets per GPU; means stay in HBM ([n, 15] fp64)."""

    def __init__(self, eng, capi, n, rank, world):
        self.eng, self.capi, self.n, self.rank, self.world = eng, capi, n, rank, world
        self.slab = min(n, 1250000)
        self.P = lhs(n, BASE_SEED + 4 + 1000 * rank)
        self.dev = {k: eng.to_device(v) for k, v in self.P.items()}
        geom = np.deg2rad(np.array(GEOM_DEG))
        self.geom = geom
        self.gd = {"iza": eng.to_device(geom[0:1]), "vza": eng.to_device(geom[1:2]), "raa": eng.to_device(geom[2:3])}
        self.nw = len(eng.windows)
        self.means = eng.empty((n, self.nw))
        self.units = n
        self.flops = NB_WINDOWS * FLOPS_PER_BAND                # only in-window wavelengths are part of this product
        self.bytes = self.nw * 8
        self.kernel = "pb::k_bands<MODE_PROSAIL, means = BRF, windows only>"

    def step(self):
        for s0 in range(0, self.n, self.slab):
            sl = lambda k: self.dev[k].ptr + 8 * s0
            lf = {k: sl(k) for k in LEAF_NAMES}
            so = {"reflectance": sl("soil_refl"), "moisture": sl("soil_moist")}
            ca = dict(self.gd, lidf_a=sl("lidf_a"), lai=sl("lai"), hotspot=sl("hotspot"))
            self.eng.prosail_device(min(self.slab, self.n - s0), lf, so, ca, None, mean_planes=(self.capi.O_RSOT,),
                                    means=self.means.ptr + s0 * self.nw * 8, windows_only=True)

    def verify(self):
        m = min(32, self.n // 2)
        idx = list(range(m)) + list(range(self.n - m, self.n))
        got = np.concatenate([self.means.download_rows(0, m), self.means.download_rows(self.n - m, m)])
        from oracle import prosail_oracle as orc
        ref = []
        for c in _oracle_rows(self.P, idx):
            ref.append(np.array([c["L8"]["BRF"]["B%d" % i] for i in range(2, 8)] + [c["ASTER"]["BRF"]["B%d" % i] for i in range(1, 10)]))
        return float(np.max(np.abs(got - np.stack(ref)))), len(idx)

    def e2e(self, args, barrier):
        eng, capi = self.eng, self.capi
        Se = min(1000000, self.n)
        Pe = {k: pinned_copy(v[:Se]) for k, v in self.P.items()}      # the step's inputs live in page-locked host memory (allocated once)
        geom = self.geom
        from pyrism_b200.engine import PinnedPool
        mbuf = {"means": PinnedPool.empty((Se, 1, self.nw), force=True)}              # allocated once, outside the timed region

        def fn():
            r = eng.prosail(Pe["N"], Pe["Cab"], Pe["Cxc"], Pe["Cbr"], Pe["Cw"], Pe["Cm"], Pe["lai"], Pe["hotspot"], geom[0:1], geom[1:2],
                            geom[2:3], Pe["soil_refl"], Pe["soil_moist"], version='5', lidf_type='campbell', a=Pe["lidf_a"],
                            planes=(), mean_planes=(capi.O_RSOT,), windows_only=True, out=mbuf)
            return float(r["means"][Se - 1, 0, 0, 0])
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            fn()
        barrier()
        return dict(dt=time.perf_counter() - t0, units=Se * args.e2e_steps, h2d=int(Se * len(NAMES) * 8 + 24), d2h=int(Se * 15 * 8), samples=Se,
                    api="pyrism_b200.Engine.prosail(windows_only=True) (prosail_prosail_f64, on_host=1): numpy parameters in, 15 band means out")

    def extra(self, args, dist, torch):
        """The two communication steps this path has: the final all-gather of the band-mean table (NCCL all_gather over
        NVLink, device to device) and one LUT inversion sharded over the ranks (arg-min fold of 12 B per observation and rank)."""
        res = {}
        eng = self.eng
        t_local = torch.as_tensor(self.means, device="cuda")                      # zero-copy view of the engine's buffer
        if dist is not None:
            world = self.world
            full = torch.empty((world * self.n, self.nw), dtype=torch.float64, device="cuda")
            for _ in range(2):
                dist.all_gather_into_tensor(full, t_local)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            e0.record()
            reps = 3
            for _ in range(reps):
                dist.all_gather_into_tensor(full, t_local)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            gbytes = world * self.n * self.nw * 8
            ok = bool(torch.equal(full[self.rank * self.n:(self.rank + 1) * self.n], t_local))
            res["gather_means"] = {"collective": "NCCL all_gather_into_tensor (torch.distributed), device to device over NVLink",
                                   "bytes_total": int(gbytes), "ms": float(ms.item()), "ranks": world,
                                   "bus_gbs_per_gpu": float(gbytes * (world - 1) / world / (ms.item() * 1e-3) / 1e9),
                                   "own_shard_intact": ok}
            del full
        # sharded inversion: observations = band means of 16384 held-out parameter sets (drawn once, identical on every rank)
        from pyrism_b200.inversion import invert_device
        Q = 16384
        Po = lhs(Q, BASE_SEED + 4 + 999999)
        obs = eng.prosail(Po["N"], Po["Cab"], Po["Cxc"], Po["Cbr"], Po["Cw"], Po["Cm"], Po["lai"], Po["hotspot"], self.geom[0:1], self.geom[1:2],
                          self.geom[2:3], Po["soil_refl"], Po["soil_moist"], version='5', lidf_type='campbell', a=Po["lidf_a"], planes=(),
                          mean_planes=(self.capi.O_RSOT,), windows_only=True)["means"][:, 0, 0, :].astype(np.float32)
        lut32 = t_local.to(torch.float32).contiguous()
        d_obs = eng.empty(obs.shape, np.float32).upload(obs)
        d_idx, d_cost = eng.empty((Q,), np.int64), eng.empty((Q,), np.float32)

        def inv():
            invert_device(eng, self.n, self.nw, lut32.data_ptr(), self.nw, Q, d_obs, self.nw, d_idx, d_cost, None, self.rank * self.n)
        inv()
        eng.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        eng.timer_begin()
        inv()
        ms_inv = eng.timer_end()
        idx, cost = d_idx.download(), d_cost.download()
        from pyrism_b200.inversion import last_stats
        st_inv = last_stats(eng)
        fold_ms = 0.0
        if dist is not None:
            t0 = time.perf_counter()
            ti = torch.from_numpy(idx).cuda()
            tc = torch.from_numpy(cost).cuda()
            gi = [torch.empty_like(ti) for _ in range(self.world)]
            gc = [torch.empty_like(tc) for _ in range(self.world)]
            dist.all_gather(gi, ti)
            dist.all_gather(gc, tc)
            from pyrism_b200.inversion import reduce_best
            idx, cost = reduce_best(np.stack([g.cpu().numpy() for g in gc]), np.stack([g.cpu().numpy() for g in gi]))
            fold_ms = (time.perf_counter() - t0) * 1e3
            t = torch.tensor([ms_inv], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_inv = float(t.item())
        res["inversion"] = {"observations": Q, "lut_rows_total": int(self.world * self.n), "features": self.nw,
                            "kernel_ms_max_over_ranks": float(ms_inv), "fold_ms": float(fold_ms),
                            "pairs_per_s": float(Q * self.world * self.n / (ms_inv * 1e-3)),
                            "median_cost": float(np.median(cost)), "all_found": bool(np.all(idx >= 0)),
                            "path_rank0": st_inv}
        return res


def d2h_probe(eng, mib=512, reps=3):
    """Plain device-to-host copy ceiling of this process (one cudaMemcpyAsync per copy into page-locked memory, through the same
    C-ABI helper the host API uses): the denominator of the end-to-end figure, measured in the same run.  GB/s."""
    from pyrism_b200.engine import PinnedPool
    n = (mib << 20) // 8
    host = PinnedPool.empty((n,), force=True)
    dev = eng.empty((n,))
    dev.download(out=host)
    t0 = time.perf_counter()
    for _ in range(reps):
        dev.download(out=host)
    dt = time.perf_counter() - t0
    dev.free()
    return n * 8 * reps / dt / 1e9


def pinned_copy(a):
    """A copy of `a` in page-locked host memory (prosail_host_alloc): the end-to-end steps upload their inputs from pinned memory, as
    the bench contract words it, so the H2D copies are asynchronous DMA transfers instead of staged pageable copies."""
    from pyrism_b200.engine import PinnedPool
    b = PinnedPool.empty(a.shape, a.dtype, force=True)
    b[...] = a
    return b


def readme_chain(args):
    """config 0: the README call chain through the drop-in classes; one 'step' = one chain."""
    import pyrism_b200 as pb

    def chain():
        leaf = pb.PROSPECT(N=1.5, Cab=35, Cxc=5, Cbr=0.15, Cw=0.003, Cm=0.0055)
        soil = pb.LSM(reflectance=3.14 / 4, moisture=0.15)
        can = pb.SAIL(iza=35, vza=30, raa=50, ks=leaf.ks, kt=leaf.kt, rho_surface=soil.ref, lidf_type='campbell', lai=3, hotspot=0.25)
        return can
    for _ in range(max(args.warmup, 3)):
        chain()
    eng = pb.get_engine(0)
    l0 = eng.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        can = chain()
    dt = time.perf_counter() - t0
    launches = eng.launch_count() - l0
    from oracle import prosail_oracle as orc
    _, _, c = orc.prosail(1.5, 35, 5, 0.15, 0.003, 0.0055, 3, 0.25, 35, 30, 50, 3.14 / 4, 0.15, lidf_type='campbell', a=57)
    par = float(np.max(np.abs(can.BRF.ref - c["rsot"])))
    v = args.steps / dt
    line = {"metric": METRIC, "value": v, "unit": "spectra/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(0, 1, 1), "clocks": None,
            "e2e": {"value": v, "unit": "spectra/s", "h2d_bytes_per_step": 11 * 8 + 2 * NB * 8 + NB * 8, "d2h_bytes_per_step": 14 * NB * 8,
                    "api": "pyrism_b200.PROSPECT / LSM / SAIL (the reference's own constructor chain); the value IS end to end"},
            "gpu_launches": int(launches), "roofline": None, "parity_max_abs": par, "parity_rows": 1,
            "device": eng.device_info()["name"]}
    if not args.no_cpu:
        kind = cpu_kind()
        cv, cdt = cpu_rate(kind, 0, 64, 1)
        line["cpu_baseline"] = {"value": cv, "unit": "spectra/s", "cores": 1, "kind": kind,
                                "sample": "64 README chains through %s on one core, %.1f s" % (CPU_WHAT[kind], cdt)}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[0, 1, 2, 3, 4], help="index into BASELINE.json configs (default 2)")
    ap.add_argument("--samples", type=int, default=0, help="parameter sets per GPU per step (0 = the configuration's own size)")
    ap.add_argument("--e2e-samples", type=int, default=262144, help="parameter sets per GPU per end-to-end step (config 2)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-samples", type=int, default=0, help="CPU baseline rows (0 = about 12 s of work)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="skip the secondary fp32-mode measurement (config 2)")
    ap.add_argument("--no-extra", action="store_true", help="config 4: skip the timed gather and the sharded inversion")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = args.config

    if args.impl == "reference":
        return run_reference(args, rank)
    if cfg == 0:
        if rank == 0:
            readme_chain(args)
        return

    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    import pyrism_b200 as pb
    from pyrism_b200 import _capi as capi
    eng = pb.Engine(local)

    if cfg == 1:
        wl = LeafOnly(eng, capi, args.samples or 10000000, rank)
    elif cfg == 2:
        wl = LutBRF(eng, capi, args.samples or 1000000, rank)
    elif cfg == 3:
        wl = MultiAngle(eng, capi, args.samples or 100000, rank, world)
    else:
        wl = BandMeanLut(eng, capi, args.samples or 12500000, rank, world)
    step = wl.step

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    fp64_peak = eng.measure_fp64_peak(200)       # live roofline denominator (MEASURED_PEAKS.json has no fp64 entry)

    sampler = ClockSampler(local) if rank == 0 else None     # started early: nvidia-smi takes a while to emit its first line
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    eng.profile(True)
    eng.profile_read()
    l0 = eng.launch_count()
    if sampler:
        sampler.begin()
    eng.timer_begin()
    for _ in range(args.steps):
        step()
    ms = eng.timer_end()
    barrier()
    clocks = sampler.stop() if sampler else None
    prof = eng.profile_read()
    eng.profile(False)
    launches = eng.launch_count() - l0
    units_total = wl.units
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        u = torch.tensor([float(wl.units)], dtype=torch.float64, device="cuda")
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        units_total = float(u.item())
    else:
        units_total = float(wl.units)
    value = units_total * args.steps / (ms * 1e-3)

    # ---- the buffer the timed loop wrote, against the oracle (rank 0) ------------------------------------------------
    parity, parity_rows = (None, 0)
    if rank == 0:
        parity, parity_rows = wl.verify()

    # ---- config 2: the same workload in the fp32 mode of the kernels (north star: <= 1e-5 absolute; a secondary number) ----
    fp32 = None
    if cfg == 2 and not args.no_fp32:
        S = wl.S
        w32 = LutBRF.__new__(LutBRF)
        w32.__dict__.update(wl.__dict__)
        w32.dt = np.float32
        w32.out = eng.empty((S, NB), np.float32)
        for _ in range(3):
            w32.step()
        barrier()
        n32 = max(3, args.steps // 3)
        eng.timer_begin()
        for _ in range(n32):
            w32.step()
        ms32 = eng.timer_end()
        barrier()
        if dist is not None:
            t = torch.tensor([ms32], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms32 = float(t.item())
        p32 = None
        if rank == 0:
            got = w32.out.download_rows(0, 16).astype(np.float64)
            ref = np.stack([c["rsot"] for c in _oracle_rows(wl.P, range(16))])
            p32 = float(np.max(np.abs(got - ref)))
        fp32 = {"value": world * S * n32 / (ms32 * 1e-3), "unit": "spectra/s", "steps": n32, "ms_per_step": ms32 / n32, "dtype": "f32",
                "tolerance": "<= 1e-5 absolute vs the float64 reference (tests/test_gpu_fp32.py)", "parity_max_abs": p32,
                "hbm_write_gbs": world * S * n32 * NB * 4 / (ms32 * 1e-3) / 1e9}
        w32.out.free()

    # ---- end to end through the host-array API ----------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        r = wl.e2e(args, barrier)
        if r is not None:
            def agg(dt):
                if dist is None:
                    return dt
                t = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t.item())
            dt = agg(r["dt"])
            e2e = {"value": world * r["units"] / dt, "unit": "spectra/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                   "samples_per_gpu": int(r["samples"]), "steps": args.e2e_steps, "api": r["api"],
                   "d2h_gbs": world * r["d2h"] * args.e2e_steps / dt / 1e9}
            # the ceiling of that figure: this rank's plain D2H copy rate into pinned memory, all ranks copying at the same time
            barrier()
            probe = d2h_probe(eng)
            if dist is not None:
                t = torch.tensor([probe], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                probe = float(t.item())
            e2e["d2h_probe_gbs"] = probe
            e2e["d2h_frac_of_probe"] = e2e["d2h_gbs"] / probe
            e2e["d2h_probe_note"] = ("sum over ranks of a plain cudaMemcpyAsync D2H of 512 MiB into page-locked memory, all ranks at once; "
                                     "tools/d2h_probe.cu is the standalone form (profiles/r2g_d2h_probe_8gpu.txt)")
            if "second" in r:
                s2 = r["second"]
                dt2 = agg(s2["dt"])
                e2e["band_means"] = {"value": world * s2["units"] / dt2, "unit": "spectra/s", "d2h_bytes_per_step": s2["d2h"], "api": s2["api"]}

    extra = {}
    if cfg == 4 and not args.no_extra:
        extra = wl.extra(args, dist, torch)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -------------------------------------------------------------------------
    k_ms, k_n = prof["bands"]
    k_avg = k_ms / max(k_n, 1)
    spectra_per_launch = wl.units * args.steps / max(k_n, 1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ach_tf = spectra_per_launch * wl.flops / (k_avg * 1e-3) / 1e12
    ach_gbs = spectra_per_launch * wl.bytes / (k_avg * 1e-3) / 1e9
    roofline = {"bound": "fp64", "kernel": wl.kernel, "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": ach_tf / fp64_peak,
                "peak_source": "DFMA microbenchmark measured in this run (prosail_measure_fp64_peak); MEASURED_PEAKS.json has no fp64 entry",
                "flops_per_spectrum": wl.flops, "kernel_ms_avg": k_avg, "kernel_launches": int(k_n),
                "spectra_per_launch": spectra_per_launch,
                "kernel_share_of_step": k_ms / ms if world == 1 else None,
                "traffic": spectra_per_launch * NCU_DRAM_BYTES_PER_SPECTRUM if cfg == 2 else None,
                "traffic_source": NCU_DRAM_SOURCE if cfg == 2 else None,
                # lower bounds of one launch from each roofline; the larger one binds
                "t_fp64_min_ms": spectra_per_launch * wl.flops / (fp64_peak * 1e12) * 1e3,
                "t_hbm_min_ms": spectra_per_launch * wl.bytes / (hbm_peak * 1e9) * 1e3,
                "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                        "algorithmic_bytes_per_spectrum": wl.bytes}}

    cpu = None
    if not args.no_cpu and world == 1:
        cores = os.cpu_count() or 1
        kind = cpu_kind()
        per_core = {"reference": 1200, "port": 4000}[kind]
        if cfg == 3:
            per_core = max(2, per_core // 48)
        if cfg == 1:
            per_core *= 3
        n_cpu = args.cpu_samples or cores * per_core
        v, dt = cpu_rate(kind, cfg, n_cpu, cores)
        cpu = {"value": v, "unit": "spectra/s", "cores": cores, "kind": kind,
               "sample": "%d parameter sets (%d spectra) of the same workload through %s, %d processes, %.1f s"
                         % (n_cpu, n_cpu * cpu_units_per_row(cfg), CPU_WHAT[kind], cores, dt)}
        if kind == "reference":          # the port as a second figure (it drops the reference's namedtuple churn and is several times faster)
            n_p = max(cores, n_cpu * 2)
            vp, dtp = cpu_rate("port", cfg, n_p, cores)
            cpu["port"] = {"value": vp, "unit": "spectra/s", "cores": cores, "kind": "port",
                           "sample": "%d parameter sets through %s, %.1f s" % (n_p, CPU_WHAT["port"], dtp)}

    line = {"metric": METRIC, "value": value, "unit": "spectra/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(cfg, wl.units // (64 if cfg == 3 else 1), world), "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "parity_max_abs": parity, "parity_rows": parity_rows,
            "parity_note": "max |GPU - oracle/prosail_oracle.py| over rows of the buffer the timed loop wrote; bar 1e-9",
            "device": eng.device_info()["name"]}
    if fp32 is not None:
        line["fp32_mode"] = fp32
    line.update(extra)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
