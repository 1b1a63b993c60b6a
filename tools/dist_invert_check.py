#!/usr/bin/env python
"""Multi-GPU check of the sharded inversion (run under torchrun, one rank per GPU, NCCL): every rank builds its shard of a band-mean
LUT on its own GPU, all ranks invert the same observations, the (cost, index) pairs are folded over NCCL; rank 0 also builds the
whole LUT alone and compares.  Also exercises lut.gather_means over NCCL.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_invert_check.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pyrism_b200 as pb
from pyrism_b200 import inversion, lut

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = pb.Engine(local)
n = 400003
names = ("N", "Cab", "Cxc", "Cbr", "Cw", "Cm", "lai", "hotspot", "lidf_a", "soil_reflectance", "soil_moisture")
P = lut.latin_hypercube(n, names, 20261019 + 77)                      # the same global LUT definition on every rank
shard = inversion.BandLutInverter(eng, P, (35.0, 30.0, 50.0), rank=rank, world=world)
rng = np.random.default_rng(5)                                        # same observations on every rank
pick = rng.integers(0, n, 3000)
means_local = shard.features.astype(np.float64)
allm = lut.gather_means(means_local, n)                               # NCCL all_gather of the band-mean table
obs = (allm[pick] * (1 + 0.01 * rng.standard_normal((3000, allm.shape[1])))).astype(np.float32)
idx, cost, best = shard.invert(obs)                                   # folds over the NCCL group
ok = True
if rank == 0:
    full = inversion.BandLutInverter(eng, P, (35.0, 30.0, 50.0))
    ref_idx, ref_cost = full.invert_local(obs)
    ok = np.array_equal(idx, ref_idx) and np.array_equal(cost, ref_cost) and np.array_equal(allm.astype(np.float32), full.features)
    print("sharded inversion over %d GPUs == single-GPU inversion: %s   (LUT %d x %d, %d observations; best lai[0:3] = %s)"
          % (world, ok, n, allm.shape[1], obs.shape[0], np.round(best["lai"][:3], 4)))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
