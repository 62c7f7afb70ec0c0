import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from nuradiomc_b200.SignalProp import propagation
from nuradiomc_b200.utilities import medium
from conftest import cylinder
from test_gpu_parity import RNOG
rt = propagation.get_propagation_module("analytic")(medium.get_ice_model("greenland_simple"), attenuation_model="GL1", n_frequencies_integration=25)
ff = np.fft.rfftfreq(1022, 0.2)
V = cylinder(3, 200_000, 4000, -2700)
dv = torch.tensor(np.ascontiguousarray(V.T), device="cuda:0"); da = torch.tensor(np.ascontiguousarray(RNOG.T), device="cuda:0")
for _ in range(2):
    res = rt.trace_batch_device(dv, da, outer=True, frequency=ff, max_detector_freq=1.2, attenuation="sparse", sync_stats=True)
st = res.stats
print("cfg3 slice: pairs %d solutions %d  ms solve %.2f att %.2f total %.2f -> %.3e pairs/s" % (st["n_pairs"], st["n_solutions"], st["ms_solve"], st["ms_attenuation"], st["ms_total"], st["n_pairs"]/st["ms_total"]*1e3))
