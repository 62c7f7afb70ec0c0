"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: static sharding of vertices, the gatherv of per-pair records and
the gather of a compact (CSR) result with unequal row counts (`nuradiomc_b200.distributed`, the "nccl" variant's host logic)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _fake_compact(pairs):
    """stand-in for the device pass: a compact result whose rows encode (global pair, slot)"""
    n_sol = (pairs % 3).astype(np.int32)
    so = np.concatenate([[0], np.cumsum(n_sol)]).astype(np.int64)
    rows_pair = np.repeat(pairs, n_sol)
    slot = np.arange(so[-1]) - np.repeat(so[:-1], n_sol)
    cap = 2 * len(pairs) + 5                                   # per-slot arrays are allocated with capacity, not with the row count
    C0 = np.full(cap, np.nan); C0[:so[-1]] = rows_pair * 10.0 + slot
    vec = np.full((cap, 3), np.nan); vec[:so[-1]] = np.stack([rows_pair, slot, rows_pair + slot], 1)
    return {"n_sol": torch.tensor(n_sol), "status": torch.zeros(len(pairs), dtype=torch.int32), "sol_offset": torch.tensor(so),
            "C0": torch.tensor(C0), "launch_vector": torch.tensor(vec)}


def _worker(rank, world, port, n_vertices, n_ant, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nuradiomc_b200.distributed import gather_compact, gather_compact_result, shard_vertices
    idx = shard_vertices(n_vertices, world, rank)
    pairs = (idx[:, None] * n_ant + np.arange(n_ant)[None, :]).ravel()
    local = {"n_sol": torch.tensor(pairs % 3, dtype=torch.int32), "C0": torch.tensor(np.stack([pairs * 0.5, pairs * 2.0], 1))}
    full = gather_compact(local)
    g = gather_compact_result(_fake_compact(pairs), len(pairs))
    if rank == 0:
        ret["n_sol"] = full["n_sol"].numpy()
        ret["C0"] = full["C0"].numpy()
        ret["compact"] = {k: v.numpy() for k, v in g.items() if not k.startswith("_")}
        ret["row_counts"] = g["_row_counts"]
    else:
        assert full is None and g is None
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
    # the gathered compact result equals the compact result of the whole job
    one = {k: v.numpy() for k, v in _fake_compact(pairs).items()}
    g = ret["compact"]
    n_rows = int(one["sol_offset"][-1])
    assert sum(ret["row_counts"]) == n_rows
    assert np.array_equal(g["n_sol"], one["n_sol"]) and np.array_equal(g["sol_offset"], one["sol_offset"])
    assert np.array_equal(g["C0"][:n_rows], one["C0"][:n_rows])
    assert np.array_equal(g["launch_vector"][:n_rows], one["launch_vector"][:n_rows])


def test_shard_bounds_cover_everything():
    from nuradiomc_b200.distributed import shard_bounds
    for n, w in ((10, 3), (1_000_000, 8), (5, 8), (0, 2)):
        b = [shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
