"""Device-resident throughput of the other BASELINE.json configs (informational; bench.py times cfg5):
python scratch/bench_configs.py"""
import sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from nuradiomc_b200.SignalProp import propagation
from nuradiomc_b200.utilities import medium
from conftest import cylinder
from test_gpu_parity import RNOG
prop = propagation.get_propagation_module("analytic")
ff512 = np.fft.rfftfreq(1022, 0.2)
from conftest import t05_points
cfgs = {
    "cfg1": dict(ice="southpole_simple", att="SP1", nr=0, V=t05_points(0, -3000.), A=np.array([[0, 0, -100.]]),
                 kw=dict(frequency=np.linspace(0, 0.5, 129), attenuation="dense"), nfreq=100),
    "cfg2": dict(ice="southpole_2015", att=None, nr=0, V=cylinder(2, 1_000_000, 4000, -2700),
                 A=np.array([[10, 10, -190.], [10, -10, -190.], [-10, -10, -190.], [-10, 10, -190.]]), kw={}),
    "cfg3 (2e5 of the 1e6 vertices)": dict(ice="greenland_simple", att="GL1", nr=0, V=cylinder(3, 200_000, 4000, -2700), A=RNOG,
                 kw=dict(frequency=ff512, max_detector_freq=1.2, attenuation="sparse")),
    "cfg4": dict(ice="mooresbay_simple", att=None, nr=1, V=cylinder(4, 1_000_000, 1000, -500),
                 A=np.array([[-3, 0, -1.], [0, 3, -1.], [3, 0, -1.], [0, -3, -1.], [3, 3, -5.], [3, -3, -5.], [-3, -3, -5.], [-3, 3, -5.]]), kw={}),
    "cfg4 + MB1 attenuation (2e5 vertices)": dict(ice="mooresbay_simple", att="MB1", nr=1, V=cylinder(4, 200_000, 1000, -500),
                 A=np.array([[-3, 0, -1.], [0, 3, -1.], [3, 0, -1.], [0, -3, -1.], [3, 3, -5.], [3, -3, -5.], [-3, -3, -5.], [-3, 3, -5.]]),
                 kw=dict(frequency=ff512, max_detector_freq=1.2, attenuation="sparse")),
}
for name, c in cfgs.items():
    rt = prop(medium.get_ice_model(c["ice"]), attenuation_model=c["att"], n_reflections=c["nr"], n_frequencies_integration=c.get("nfreq", 25))
    dv = torch.tensor(np.ascontiguousarray(c["V"].T), device="cuda:0"); da = torch.tensor(np.ascontiguousarray(c["A"].T), device="cuda:0")
    out = None
    for _ in range(3):
        out = rt.trace_batch_device(dv, da, outer=True, out=out, **c["kw"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 5
    for _ in range(K):
        out = rt.trace_batch_device(dv, da, outer=True, out=out, **c["kw"])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    n = len(c["V"]) * len(c["A"])
    print(json.dumps({"config": name, "pairs": n, "solutions_per_pair": float(out["n_sol"].sum()) / n, "ms_per_step": ms, "pairs_per_s": n / ms * 1e3}))
