#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Benchmark of the batched PROSAIL hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--samples S]

Workload (BASELINE.json configs[2], the configuration the metric "PROSAIL spectra/s (2101 bands, fp64)" is quoted
on): a PROSAIL look-up table of S = 1e6 Latin-hypercube parameter sets per GPU (N, Cab, Cxc, Cbr, Cw, Cm, lai,
hotspot, Campbell mean leaf angle, soil brightness, soil moisture; SURVEY.md 8(d) bounds), PROSPECT-5, one
geometry (35/30/50 deg), output = the full 2101-band BRF spectrum (rsot) in fp64.  One "step" = one pass of
prologue + fused PROSPECT->4SAIL kernel over the S parameter sets of this rank.  Multi-GPU: every rank owns its own
LHS shard (weak scaling, no data-path collective).

Printed JSON (rank 0): see the keys in main().  `value` = spectra/s with parameters resident in HBM and the output
written to HBM; `e2e` = the same through the host-array API (pageable numpy parameters in, pinned numpy spectra
out, H2D + D2H inside the timed region); `roofline` = the fused kernel against the measured FP64 (DFMA) peak of the
device, with the HBM-write view beside it; `cpu_baseline` = the CPU oracle port on the host cores.
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
FLOPS_PER_BAND = 234            # PROSPECT 83 + 4SAIL 151, div/sqrt/exp/E1/pow counted as ONE flop each
BYTES_PER_SPECTRUM = NB * 8     # one fp64 output plane
NCU_DRAM_BYTES_PER_SPECTRUM = (4.361098e9 + 80.665344e6) / 262144   # measured, see profiles/r1k_k_bands_prosail_brf_f64.txt
SEED = 20261019 + 2             # config index 2
GEOM_DEG = (35.0, 30.0, 50.0)

# SURVEY.md 8(d) bounds: N Cab Cxc Cbr Cw Cm lai hotspot campbell_a soil_refl soil_moist
LO = np.array([1.0, 5.0, 1.0, 0.0, 0.001, 0.001, 0.1, 0.01, 20.0, 0.3, 0.0])
HI = np.array([3.0, 90.0, 25.0, 1.0, 0.05, 0.02, 8.0, 0.5, 80.0, 1.2, 1.0])
NAMES = ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm", "lai", "hotspot", "lidf_a", "soil_refl", "soil_moist")


def lhs(n, seed):
    from scipy.stats import qmc
    x = qmc.LatinHypercube(d=len(LO), seed=seed).random(n)
    x = qmc.scale(x, LO, HI)
    return {k: np.ascontiguousarray(x[:, i]) for i, k in enumerate(NAMES)}


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
# CPU arm: the oracle port on the host cores (the reference is pure Python and cannot travel to the GPU box)
# ----------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import prosail_oracle as orc
    rows, = args
    acc = 0.0
    for r in rows:
        p = dict(zip(NAMES, r))
        leaf, soil, can = orc.prosail(p["N"], p["Cab"], p["Cxc"], p["Cbr"], p["Cw"], p["Cm"], p["lai"], p["hotspot"], GEOM_DEG[0],
                                      GEOM_DEG[1], GEOM_DEG[2], p["soil_refl"], p["soil_moist"], lidf_type="campbell", a=p["lidf_a"])
        acc += float(can["rsot"][0])
    return acc


def cpu_rate(n_samples, cores, pool=None):
    """Spectra/s of the oracle port (PROSPECT -> LSM -> SAIL incl. band means, one sample per call) on `cores` processes."""
    import multiprocessing as mp
    P = lhs(n_samples, SEED + 777)
    rows = np.stack([P[k] for k in NAMES], axis=1)
    chunks = [rows[i::cores] for i in range(cores)]
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(cores)
        pool.map(_cpu_worker, [(c[:2],) for c in chunks])      # import + table load outside the timed region
    t0 = time.perf_counter()
    pool.map(_cpu_worker, [(c,) for c in chunks])
    dt = time.perf_counter() - t0
    if own:
        pool.close()
    return n_samples / dt, dt


def run_reference(args, rank):
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_step = cores * 400
    pool = mp.get_context("fork").Pool(cores)
    tiny = np.stack([lhs(2, 1)[k] for k in NAMES], axis=1)
    pool.map(_cpu_worker, [(tiny,)] * cores)                   # imports and table load outside the timed steps
    for _ in range(args.warmup):
        cpu_rate(per_step, cores, pool)
    t = 0.0
    for _ in range(args.steps):
        _, dt = cpu_rate(per_step, cores, pool)
        t += dt
    pool.close()
    v = per_step * args.steps / t
    line = {"impl": "reference", "metric": "PROSAIL spectra/s (2101 bands, fp64)", "value": v, "unit": "spectra/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(per_step, args.gpus),
            "cpu_baseline": {"value": v, "unit": "spectra/s", "cores": cores, "kind": "port",
                             "sample": "%d LHS parameter sets per step through oracle/prosail_oracle.py (numpy/scipy restatement of "
                                       "pyrism PROSPECT->LSM->SAIL incl. L8/ASTER means), one process per core" % per_step},
            "e2e": {"value": v, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(samples, gpus):
    return {"workload": "BASELINE configs[2]: PROSAIL LUT, Latin-hypercube parameter sets x 2101 bands, fp64, PROSPECT-5 + "
                        "LSM soil + 4SAIL campbell, geometry 35/30/50 deg, output BRF spectrum",
            "samples_per_gpu": int(samples), "bands": NB, "geometries": 1, "output_planes": ["rsot"],
            "parallelism": "dp%d (parameter-set shards, no collective)" % gpus,
            "l2": "output (samples x 16808 B) and records are far larger than the 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=int, default=1000000, help="parameter sets per GPU per step")
    ap.add_argument("--e2e-samples", type=int, default=262144, help="parameter sets per GPU per end-to-end step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-samples", type=int, default=0, help="CPU baseline sample (0 = 4000 per core)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="skip the secondary fp32-mode measurement")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return run_reference(args, rank)

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

    S = args.samples
    P = lhs(S, SEED + 1000 * rank)
    geom = np.deg2rad(np.array(GEOM_DEG))
    dev = {k: eng.to_device(v) for k, v in P.items()}
    leaf = {k: dev[k] for k in ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm")}
    soil = {"reflectance": dev["soil_refl"], "moisture": dev["soil_moist"]}
    canopy = {"lidf_a": dev["lidf_a"], "lai": dev["lai"], "hotspot": dev["hotspot"], "iza": eng.to_device(geom[0:1]),
              "vza": eng.to_device(geom[1:2]), "raa": eng.to_device(geom[2:3])}
    out = eng.empty((S, NB))

    def step():
        eng.prosail_device(S, leaf, soil, canopy, {capi.O_RSOT: out}, pitch=NB, version='5', lidf_type='campbell', n_geom=1)

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
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * S * args.steps / (ms * 1e-3)

    # ---- the same workload in the fp32 mode of the kernels (north star: <= 1e-5 absolute; a secondary number) ------
    fp32 = None
    if not args.no_fp32:
        out32 = eng.empty((S, NB), np.float32)

        def step32():
            eng.prosail_device(S, leaf, soil, canopy, {capi.O_RSOT: out32}, pitch=NB, version='5', lidf_type='campbell', n_geom=1,
                               dtype=np.float32)
        for _ in range(3):
            step32()
        barrier()
        n32 = max(3, args.steps // 3)
        eng.timer_begin()
        for _ in range(n32):
            step32()
        ms32 = eng.timer_end()
        barrier()
        if dist is not None:
            t = torch.tensor([ms32], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms32 = float(t.item())
        fp32 = {"value": world * S * n32 / (ms32 * 1e-3), "unit": "spectra/s", "steps": n32, "ms_per_step": ms32 / n32, "dtype": "f32",
                "tolerance": "<= 1e-5 absolute vs the float64 reference (tests/test_gpu_fp32.py)",
                "hbm_write_gbs": world * S * n32 * NB * 4 / (ms32 * 1e-3) / 1e9}
        out32.free()

    # ---- end to end through the host-array API ----------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        Se = min(args.e2e_samples, S)
        Pe = {k: v[:Se].copy() for k, v in P.items()}
        from pyrism_b200.engine import PinnedPool
        obuf = {"rsot": PinnedPool.empty((Se, NB), force=True)}

        def e2e_step():
            r = eng.prosail(Pe["N"], Pe["Cab"], Pe["Cxc"], Pe["Cbr"], Pe["Cw"], Pe["Cm"], Pe["lai"], Pe["hotspot"], geom[0:1], geom[1:2],
                            geom[2:3], Pe["soil_refl"], Pe["soil_moist"], version='5', lidf_type='campbell', a=Pe["lidf_a"],
                            planes=(capi.O_RSOT,), out=obuf)
            return float(r["rsot"][Se - 1, 0, NB - 1])

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * Se * args.e2e_steps / dt, "unit": "spectra/s", "h2d_bytes_per_step": int(Se * len(NAMES) * 8 + 24),
               "d2h_bytes_per_step": int(Se * NB * 8), "samples_per_gpu": int(Se), "steps": args.e2e_steps,
               "numa_node": eng.numa_node,
               "api": "pyrism_b200.Engine.prosail (prosail_prosail_f64, on_host=1): numpy parameters in, pinned numpy BRF spectra out"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -------------------------------------------------------------------------
    k_ms, k_n = prof["bands"]
    k_avg = k_ms / max(k_n, 1)
    spectra_per_launch = S * args.steps / max(k_n, 1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ach_tf = spectra_per_launch * NB * FLOPS_PER_BAND / (k_avg * 1e-3) / 1e12
    ach_gbs = spectra_per_launch * BYTES_PER_SPECTRUM / (k_avg * 1e-3) / 1e9
    roofline = {"bound": "fp64", "kernel": "pb::k_bands<MODE_PROSAIL>", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": ach_tf / fp64_peak,
                "peak_source": "DFMA microbenchmark measured in this run (prosail_measure_fp64_peak); MEASURED_PEAKS.json has no fp64 entry",
                "flops_per_band": FLOPS_PER_BAND, "kernel_ms_avg": k_avg, "kernel_launches": int(k_n),
                "kernel_share_of_step": k_ms / ms if world == 1 else None,
                # DRAM bytes per launch: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel
                # (profiles/r1k_k_bands_prosail_brf_f64.txt: 4.442 GB for 262144 spectra = 16944 B per spectrum), scaled to this launch
                "traffic": spectra_per_launch * NCU_DRAM_BYTES_PER_SPECTRUM,
                # lower bounds of one launch from each roofline; the larger one binds (fp64 by ~5x)
                "t_fp64_min_ms": spectra_per_launch * NB * FLOPS_PER_BAND / (fp64_peak * 1e12) * 1e3,
                "t_hbm_min_ms": spectra_per_launch * BYTES_PER_SPECTRUM / (hbm_peak * 1e9) * 1e3,
                "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                        "algorithmic_bytes_per_spectrum": BYTES_PER_SPECTRUM}}

    cpu = None
    if not args.no_cpu and world == 1:
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_samples or cores * 4000
        v, dt = cpu_rate(n_cpu, cores)
        cpu = {"value": v, "unit": "spectra/s", "cores": cores, "kind": "port",
               "sample": "%d LHS parameter sets of the same workload through oracle/prosail_oracle.py (numpy/scipy restatement of pyrism "
                         "PROSPECT->LSM->SAIL incl. L8/ASTER means), %d processes, %.1f s" % (n_cpu, cores, dt)}

    line = {"metric": "PROSAIL spectra/s (2101 bands, fp64)", "value": value, "unit": "spectra/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(S, world), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "fp32_mode": fp32,
            "device": eng.device_info()["name"]}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
