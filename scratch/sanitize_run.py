"""small end-to-end exercise of every kernel path for compute-sanitizer (memcheck / racecheck)"""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from nuradiomc_b200.SignalProp import propagation
from nuradiomc_b200.utilities import medium
from conftest import cylinder
prop = propagation.get_propagation_module("analytic")
ff = np.fft.rfftfreq(64, 0.5)
for ice, att, nr in (("southpole_2015", "SP1", 0), ("greenland_simple", "GL1", 0), ("greenland_simple", "GL3", 0), ("mooresbay_simple", "MB1", 1)):
    rt = prop(medium.get_ice_model(ice), attenuation_model=att, n_reflections=nr, n_frequencies_integration=8)
    V, A = cylinder(1, int(__import__("os").environ.get("SAN_N", "700")), 3000 if nr == 0 else 800, -2500 if nr == 0 else -500), np.array([[0, 0, -5.], [300, 0, -150.]])
    axes = np.random.default_rng(2).normal(size=(len(V), 3))
    for chunk in (0, 333):
        rt.set_chunk_pairs(chunk)
        r1 = rt.trace_batch(V, A, outer=True, frequency=ff, attenuation="both")
        r2 = rt.trace_batch(V, A, outer=True, frequency=ff, attenuation="both", compact=True, pinned=True)
        r3 = rt.trace_batch(V, A, outer=True, frequency=ff, attenuation="dense", shower_axis=axes, delta_C_cut=0.7)
        dv = torch.tensor(np.ascontiguousarray(V.T), device="cuda:0"); da = torch.tensor(np.ascontiguousarray(A.T), device="cuda:0")
        r4 = rt.trace_batch_device(dv, da, outer=True, frequency=ff, attenuation="both", sync_stats=True)
        if nr == 0:
            r5 = rt.trace_batch_device(dv, da, outer=True, frequency=ff, attenuation="dense", compact=True, sync_stats=True)
            n = r5.n_rows()
            spec = torch.randn((n, 3, len(ff)), dtype=torch.complex128, device="cuda:0")
            rt.apply_propagation_effects_batch(spec, reflection_angle=r5["reflection_angle"][:n].contiguous(), reflection=r5["reflection"][:n].contiguous(),
                                               attenuation=r5["attenuation"][:n].contiguous())
    rt.set_start_and_end_point(V[0], A[0]); rt.find_solutions()
    print(ice, att, nr, "ok", int(r1["n_sol"].sum()), r2["C0"].shape)
torch.cuda.synchronize()
print("done")
