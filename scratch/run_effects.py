"""drives K_apply_effects and K_focusing at a realistic size (for ncu captures and a timing line): python scratch/run_effects.py [n_vertices]"""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from nuradiomc_b200.SignalProp import propagation
from nuradiomc_b200.utilities import medium
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
cfg = bench.CONFIGS["cfg5"]
V, A, ff = bench.workload(nv, "cfg5")
rt = propagation.get_propagation_module("analytic")(medium.get_ice_model(cfg["ice"]), attenuation_model="SP1", n_frequencies_integration=25)
dv, da = torch.tensor(V, device="cuda"), torch.tensor(A, device="cuda")
res = rt.trace_batch_device(dv, da, outer=True, compact=True, frequency=ff, max_detector_freq=1.2, attenuation="sparse")
n = res.n_rows()
F = len(ff)
spec = torch.randn(n, 3, F, dtype=torch.complex128, device="cuda")
out = {"rows": n, "F": F}
for name, fn in (("apply_effects_sparse", lambda: rt.apply_propagation_effects_batch(spec, reflection_angle=res["reflection_angle"][:n], attenuation_sparse=res["attenuation_sparse"][:n])),
                 ("focusing", lambda: rt.focusing_batch(dv, da, res, outer=True))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    out[name + "_ms"] = e0.elapsed_time(e1) / 5
out["apply_effects_GBs"] = n * 3 * F * 16 * 2 / (out["apply_effects_sparse_ms"] * 1e-3) / 1e9
print(json.dumps(out))
