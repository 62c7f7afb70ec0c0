"""2-GPU test (skipped on boxes with one GPU): vertices sharded over two ranks, each traces its block on its own device,
the compact records are gathered over NCCL (`gather_compact`) and must equal the single-device trace bit for bit."""
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


def _worker(rank, world, port, V, A, ff, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from nuradiomc_b200.SignalProp import propagation
    from nuradiomc_b200.distributed import gather_compact, shard_bounds
    from nuradiomc_b200.utilities import medium
    rt = propagation.get_propagation_module("analytic")(medium.get_ice_model("southpole_2015"), attenuation_model="SP1",
                                                         n_frequencies_integration=10, device=rank)
    lo, hi = shard_bounds(V.shape[0], world, rank)
    dv = torch.tensor(np.ascontiguousarray(V[lo:hi].T), device=f"cuda:{rank}")
    da = torch.tensor(np.ascontiguousarray(A.T), device=f"cuda:{rank}")
    res = rt.trace_batch_device(dv, da, outer=True, frequency=ff, max_detector_freq=0.6, attenuation="sparse")
    full = gather_compact({k: res[k] for k in KEYS})
    if rank == 0:
        for k in KEYS:
            ret[k] = full[k].cpu().numpy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_trace_and_nccl_gather_equal_single_device():
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
