"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: static sharding of vertices + gather of compact records."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, n_vertices, n_ant, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nuradiomc_b200.distributed import gather_compact, shard_vertices
    idx = shard_vertices(n_vertices, world, rank)
    # stand-in for the device pass: per-pair records that encode the global pair index
    pairs = (idx[:, None] * n_ant + np.arange(n_ant)[None, :]).ravel()
    local = {"n_sol": torch.tensor(pairs % 3, dtype=torch.int32), "C0": torch.tensor(np.stack([pairs * 0.5, pairs * 2.0], 1))}
    full = gather_compact(local)
    if rank == 0:
        ret["n_sol"] = full["n_sol"].numpy()
        ret["C0"] = full["C0"].numpy()
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    n_vertices, n_ant = 1001, 4
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29517, n_vertices, n_ant, ret), nprocs=2, join=True)
    pairs = np.arange(n_vertices * n_ant)
    assert np.array_equal(ret["n_sol"], pairs % 3)
    assert np.array_equal(ret["C0"], np.stack([pairs * 0.5, pairs * 2.0], 1))
