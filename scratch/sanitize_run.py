"""small end-to-end exercise of every kernel path for compute-sanitizer (memcheck / racecheck / synccheck)"""
import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from nuradiomc_b200.SignalProp import propagation
from nuradiomc_b200.utilities import medium
from nuradiomc_b200.distributed import P2PGather
from conftest import cylinder
prop = propagation.get_propagation_module("analytic")
ff = np.fft.rfftfreq(64, 0.05)       # 0 .. 10 GHz on 33 bins: GL1 poles inside the paths' range (item / fine kernels)
N = int(os.environ.get("SAN_N", "700"))
for small in (False, True):
    if small:
        os.environ.pop("NRMC_NO_SMALL_PATH", None)
    else:
        os.environ["NRMC_NO_SMALL_PATH"] = "1"
    for ice, att, nr in (("southpole_2015", "SP1", 0), ("greenland_simple", "GL1", 0), ("greenland_simple", "GL3", 0), ("mooresbay_simple", "MB1", 1),
                         ("greenland_simple", "GL2", 0)):
        rt = prop(medium.get_ice_model(ice), attenuation_model=att, n_reflections=nr, n_frequencies_integration=8)
        V, A = cylinder(1, N, 3000 if nr == 0 else 800, -2700 if nr == 0 else -500), np.array([[0, 0, -5.], [300, 0, -150.]])
        axes = np.random.default_rng(2).normal(size=(len(V), 3))
        for chunk in ((0,) if small else (0, 333)):
            rt.set_chunk_pairs(chunk)
            r1 = rt.trace_batch(V, A, outer=True, frequency=ff, max_detector_freq=1.0, attenuation="both")
            r2 = rt.trace_batch(V, A, outer=True, frequency=ff, max_detector_freq=1.0, attenuation="both", compact=True, pinned=True)
            r3 = rt.trace_batch(V, A, outer=True, frequency=ff, attenuation="dense", shower_axis=axes, delta_C_cut=0.7)
            r6 = rt.trace_batch(V, A, outer=True, frequency=np.array([0.3]), attenuation="both")          # a single frequency
            if att in ("MB1", "GL2"):      # K_att_sep (sparse output), incl. a GL2 frequency whose 1 m floor is crossed on the path (fall-back list)
                r7 = rt.trace_batch(V, A, outer=True, frequency=np.array([0.0, 0.3, 1.5762, 1.9]), attenuation="sparse", compact=True)
            dv = torch.tensor(np.ascontiguousarray(V.T), device="cuda:0"); da = torch.tensor(np.ascontiguousarray(A.T), device="cuda:0")
            r4 = rt.trace_batch_device(dv, da, outer=True, frequency=ff, max_detector_freq=1.0, attenuation="both", sync_stats=True)
            foc = rt.focusing_batch(dv, da, r4, outer=True)
            r5 = rt.trace_batch_device(dv, da, outer=True, frequency=ff, max_detector_freq=1.0, attenuation="both" if nr else "sparse", compact=True, sync_stats=True)
            n = r5.n_rows()
            spec = torch.randn((n, 3, len(ff)), dtype=torch.complex128, device="cuda:0")
            if nr == 0:
                rt.apply_propagation_effects_batch(spec, reflection_angle=r5["reflection_angle"][:n].contiguous(), attenuation_sparse=r5["attenuation_sparse"][:n].contiguous())
            else:
                rt.apply_propagation_effects_batch(spec, reflection_angle=r5["reflection_angle"][:n].contiguous(), reflection=r5["reflection"][:n].contiguous(),
                                                   attenuation=r5["attenuation"][:n].contiguous())
        rt.set_chunk_pairs(0)
        if not small and nr == 0:      # the peer-memory gather, single process: rows stored through the mapped block
            pg = P2PGather(rt, len(V) * len(A), names=("C0", "travel_time", "launch_vector", "attenuation_sparse"), Fs=len(r5.frequencies_sparse))
            pg.trace(dv, da, outer=True, frequency=ff, max_detector_freq=1.0, attenuation="sparse"); pg.finish()
            pg.trace_pushed(dv, da, n_chunks=2, outer=True, frequency=ff, max_detector_freq=1.0, attenuation="sparse"); pg.finish()
            pg.close()
        rt.set_start_and_end_point(V[0], A[0]); rt.find_solutions()
        if rt.get_number_of_solutions():
            rt.get_attenuation(0, ff, 1.0)
        print("small" if small else "binned", ice, att, nr, "ok", int(r1["n_sol"].sum()), r2["C0"].shape, flush=True)
torch.cuda.synchronize()
print("done")
