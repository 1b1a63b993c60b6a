"""world_size-2 CPU (gloo) test of the only multi-rank logic the path has: contiguous sharding of the parameter-set
index and the optional final gather of band-mean tables.  The data path itself has no collective."""
import os
import socket

import numpy as np
import pytest


def _worker(rank, world, port, n, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyrism_b200 import lut
    params = {"x": np.arange(n, dtype=np.float64), "scalar": 3.0}
    local, (b, e) = lut.shard(params, rank, world)
    assert local["scalar"] == 3.0 and np.array_equal(local["x"], np.arange(b, e))
    # stand-in for the per-rank LUT kernel: "means" derived from the shard's own parameter rows
    means = np.stack([local["x"] * 2.0, local["x"] + 0.5], axis=1).reshape(e - b, 1, 2)
    full = lut.gather_means(means, n)
    np.save(os.path.join(out_dir, "full%d.npy" % rank), full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 64])
def test_shard_and_gather_world2(tmp_path, n):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    x = np.arange(n, dtype=np.float64)
    want = np.stack([x * 2.0, x + 0.5], axis=1).reshape(n, 1, 2)
    for r in range(2):
        assert np.array_equal(np.load(os.path.join(str(tmp_path), "full%d.npy" % r)), want)
