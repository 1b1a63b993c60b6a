#!/usr/bin/env python
"""Parity sweep on a GPU box: N Latin-hypercube parameter sets (random geometry, campbell and verhoef, PROSPECT-5 and D)
through the CUDA path in fp64 and in fp32 mode, against the CPU oracle evaluated on the host cores.  Prints the worst absolute
error per quantity.  usage: python tools/parity_sweep.py [N]   (test infrastructure; imports oracle/)"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
QUANT = ('rsot', 'rddt', 'rsdt', 'rdot', 'rso', 'rdd', 'tdd', 'rsd', 'tsd', 'rdo', 'tdo', 'ke')
PL = {'rsot': 0, 'rddt': 1, 'rsdt': 2, 'rdot': 3, 'rso': 4, 'rdd': 5, 'tdd': 6, 'rsd': 7, 'tsd': 8, 'rdo': 9, 'tdo': 10, 'ke': 11, 'leaf_r': 12, 'leaf_t': 13, 'rho': 14}


def _oracle(args):
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import prosail_oracle as orc
    rows, version, lidf = args
    out = []
    for r in rows:
        (N, Cab, Cxc, Cbr, Cw, Cm, Can, lai, hot, a, b, sr, sm, iza, vza, raa) = r
        leaf, soil, can = orc.prosail(N, Cab, Cxc, Cbr, Cw, Cm, lai, hot, iza, vza, raa, sr, sm, Can=Can if version == 'D' else 0, version=version,
                                      lidf_type=lidf, a=a, b=b)
        out.append(np.stack([can[q] * np.ones(2101) for q in QUANT] + [leaf['ks'], leaf['kt'], soil['ref']]))
    return np.stack(out)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    import pyrism_b200 as pb
    eng = pb.get_engine(0)
    rng = np.random.default_rng(20261019 + 99)
    cores = os.cpu_count() or 1
    pool = mp.get_context("fork").Pool(cores)
    t0 = time.time()
    for version in ('5', 'D'):
        for lidf in ('campbell', 'verhoef'):
            m = n // 4
            P = np.stack([rng.uniform(1, 3, m), rng.uniform(5, 90, m), rng.uniform(1, 25, m), rng.uniform(0, 1, m), rng.uniform(0.001, 0.05, m),
                          rng.uniform(0.001, 0.02, m), rng.uniform(0.1, 8, m), rng.uniform(0.1, 8, m), rng.uniform(0.01, 0.5, m),
                          rng.uniform(20, 80, m) if lidf == 'campbell' else rng.uniform(-0.6, 0.6, m),
                          np.zeros(m) if lidf == 'campbell' else rng.uniform(-0.35, 0.35, m), rng.uniform(0.3, 1.2, m), rng.uniform(0, 1, m),
                          rng.uniform(0, 70, m), rng.uniform(0, 60, m), rng.uniform(0, 180, m)], axis=1)
            ref = np.concatenate(pool.map(_oracle, [(P[i::cores], version, lidf) for i in range(cores)]))
            order = np.concatenate([np.arange(m)[i::cores] for i in range(cores)])
            P = P[order]
            for dt in (np.float64, np.float32):
                r = eng.prosail(P[:, 0], P[:, 1], P[:, 2], P[:, 3], P[:, 4], P[:, 5], P[:, 7], P[:, 8], np.radians(P[:, 13]), np.radians(P[:, 14]),
                                np.radians(P[:, 15]), P[:, 11], P[:, 12], Can=P[:, 6], version=version, lidf_type=lidf, a=P[:, 9], b=P[:, 10],
                                planes=tuple(range(15)), geom_per_sample=True, dtype=dt)
                worst = {q: float(np.max(np.abs(r[q][:, 0].astype(np.float64) - ref[:, PL[q]]))) for q in PL}
                print("PROSPECT-%s %-8s %-7s n=%d  " % (version, lidf, np.dtype(dt).name, m) + " ".join("%s %.1e" % kv for kv in worst.items())
                      + "   MAX %.2e" % max(worst.values()))
                sys.stdout.flush()
    print("elapsed %.0f s on %d host cores" % (time.time() - t0, cores))


if __name__ == "__main__":
    main()
