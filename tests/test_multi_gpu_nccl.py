"""2-GPU tests (skipped on boxes with one GPU): vertices sharded over two ranks, each traces its block on its own device in the
compact (CSR) layout, and the results are gathered on rank 0
  * "p2p":  fused -- the kernels store the rows straight into rank 0's HBM through NVLink peer mappings (`P2PGather`),
  * "nccl": grouped ncclSend / ncclRecv of the filled rows (`gather_compact_result`),
both of which must equal the single-device compact trace bit for bit; plus the round-1 padded gather (`gather_compact`)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, cylinder

pytestmark = pytest.mark.gpu

KEYS = ("n_sol", "solution_type", "C0", "travel_time", "launch_vector", "attenuation_sparse")
ROW_KEYS = ("solution_type", "reflection", "reflection_case", "C0", "C1", "path_length", "travel_time", "launch_vector", "receive_vector",
            "reflection_angle", "attenuation_sparse")


def _worker(rank, world, port, V, A, ff, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from nuradiomc_b200.SignalProp import propagation
    from nuradiomc_b200.distributed import P2PGather, gather_compact, gather_compact_result, shard_bounds
    from nuradiomc_b200.utilities import medium
    rt = propagation.get_propagation_module("analytic")(medium.get_ice_model("southpole_2015"), attenuation_model="SP1",
                                                         n_frequencies_integration=10, device=rank)
    lo, hi = shard_bounds(V.shape[0], world, rank)
    dv = torch.tensor(np.ascontiguousarray(V[lo:hi].T), device=f"cuda:{rank}")
    da = torch.tensor(np.ascontiguousarray(A.T), device=f"cuda:{rank}")
    n_local = (hi - lo) * A.shape[0]
    kw = dict(outer=True, frequency=ff, max_detector_freq=0.6, attenuation="sparse")
    # round-1 padded gather
    res = rt.trace_batch_device(dv, da, **kw)
    full = gather_compact({k: res[k] for k in KEYS})
    # compact + NCCL gatherv
    resc = rt.trace_batch_device(dv, da, compact=True, **kw)
    g = gather_compact_result(dict(resc), n_local)
    # compact + fused peer-memory gather
    Fs = len(resc.frequencies_sparse)
    pg = P2PGather(rt, n_local, names=ROW_KEYS, Fs=Fs)
    for _ in range(2):          # twice: the block is reused from step to step
        pg.trace(dv, da, **kw)
        pg.finish()
    # the same gather with the rows pushed by the copy engines, chunk-pipelined
    pd = P2PGather(rt, n_local, names=ROW_KEYS, Fs=Fs)
    for _ in range(2):
        pd.trace_pushed(dv, da, n_chunks=3, **kw)
        pd.finish()
    if rank == 0:
        ret["p2p_dma"] = {k: v.cpu().numpy() for k, v in pd.compacted().items()}
    pd.close()
    # records only: the factors stay in the HBM of the rank that computed them, in LOCAL rows
    pr = P2PGather(rt, n_local, names=[k for k in ROW_KEYS if k != "attenuation_sparse"], Fs=Fs)
    loc = pr.trace(dv, da, **kw)
    pr.finish()
    ret[f"local_att_{rank}"] = loc["attenuation_sparse"][:loc.n_rows()].cpu().numpy()
    if rank == 0:
        ret["p2p_records"] = {k: v.cpu().numpy() for k, v in pr.compacted().items()}
    pr.close()
    if rank == 0:
        for k in KEYS:
            ret[k] = full[k].cpu().numpy()
        ret["nccl"] = {k: v.cpu().numpy() for k, v in g.items() if not k.startswith("_")}
        ret["p2p"] = {k: v.cpu().numpy() for k, v in pg.compacted().items()}
        ret["p2p_raw"] = {k: v.cpu().numpy() for k, v in pg.arrays.items() if k in ("n_sol", "sol_offset", "C0")}
        ret["row_base"] = list(pg.row_base)
    else:
        assert full is None and g is None
    pg.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_trace_and_gather_equal_single_device():
    ff = np.fft.rfftfreq(128, 0.5)
    V, A = cylinder(81, 5001, 6000, -2700), np.array([[0, 0, -150.], [1500, 0, -160.], [0, -1500, -145.]])
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29527, V, A, ff, ret), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    from nuradiomc_b200.SignalProp import propagation
    from nuradiomc_b200.utilities import medium
    rt = propagation.get_propagation_module("analytic")(medium.get_ice_model("southpole_2015"), attenuation_model="SP1",
                                                         n_frequencies_integration=10)
    one = rt.trace_batch(V, A, outer=True, frequency=ff, max_detector_freq=0.6, attenuation="sparse")
    for k in KEYS:
        np.testing.assert_array_equal(ret[k], one[k], err_msg=k)
    onec = rt.trace_batch(V, A, outer=True, frequency=ff, max_detector_freq=0.6, attenuation="sparse", compact=True)
    n_rows = int(onec["sol_offset"][-1])
    assert n_rows > 1000
    for tag in ("nccl", "p2p", "p2p_dma"):
        g = ret[tag]
        np.testing.assert_array_equal(g["n_sol"], onec["n_sol"], err_msg=tag)
        np.testing.assert_array_equal(g["sol_offset"], onec["sol_offset"], err_msg=tag)
        for k in ROW_KEYS:
            np.testing.assert_array_equal(g[k][:n_rows], onec[k][:n_rows], err_msg=f"{tag} {k}")
    g = ret["p2p_records"]
    for k in ROW_KEYS:
        if k != "attenuation_sparse":
            np.testing.assert_array_equal(g[k][:n_rows], onec[k][:n_rows], err_msg=f"p2p records {k}")
    np.testing.assert_array_equal(np.concatenate([ret["local_att_0"], ret["local_att_1"]]), onec["attenuation_sparse"][:n_rows])
    # the un-compacted peer block: rank 1's rows start at its segment base and are addressed by sol_offset / n_sol
    raw, n_pairs0 = ret["p2p_raw"], (5001 // 2 + 1) * 3
    assert raw["sol_offset"][n_pairs0] == ret["row_base"][1]
    i = n_pairs0 + int(np.argmax(raw["n_sol"][n_pairs0:] > 0))
    assert raw["C0"][raw["sol_offset"][i]] == onec["C0"][onec["sol_offset"][i]]
