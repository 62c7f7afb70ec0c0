"""latency of the scalar API and of small batches (the reference's production caller is scalar: simulation.py:173-210)"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from nuradiomc_b200.SignalProp import propagation
from nuradiomc_b200.utilities import medium
rt = propagation.get_propagation_module("analytic")(medium.get_ice_model("southpole_2015"), attenuation_model="SP1", n_frequencies_integration=25)
rng = np.random.default_rng(1)
N = 400
r = np.sqrt(rng.uniform(0, 4000.**2, N)); phi = rng.uniform(0, 2*np.pi, N)
X1 = np.array([r*np.cos(phi), r*np.sin(phi), rng.uniform(-2700, 0, N)]).T
x2 = np.array([10., 10., -190.])
ff = np.fft.rfftfreq(1022, 0.2)
out = {}
for i in range(20):
    rt.set_start_and_end_point(X1[i], x2); rt.find_solutions()
t = time.perf_counter()
for i in range(N):
    rt.set_start_and_end_point(X1[i], x2); rt.find_solutions()
out["scalar_find_solutions_us"] = (time.perf_counter() - t) / N * 1e6
t = time.perf_counter(); n = 0
for i in range(N):
    rt.set_start_and_end_point(X1[i], x2); rt.find_solutions()
    for iS in range(rt.get_number_of_solutions()):
        rt.get_path_length(iS); rt.get_travel_time(iS); rt.get_launch_vector(iS); n += 1
out["scalar_find_plus_properties_us"] = (time.perf_counter() - t) / N * 1e6
t = time.perf_counter()
for i in range(100):
    rt.set_start_and_end_point(X1[i], x2); rt.find_solutions()
    for iS in range(rt.get_number_of_solutions()):
        rt.get_attenuation(iS, ff, 1.2)
out["scalar_find_plus_attenuation_us"] = (time.perf_counter() - t) / 100 * 1e6
for nb in (1, 32, 1000, 4096):
    X = X1[np.arange(nb) % N]
    for _ in range(5):
        res = rt.trace_batch(X, x2[None], frequency=ff, max_detector_freq=1.2, attenuation="sparse")
    t = time.perf_counter()
    for _ in range(50):
        res = rt.trace_batch(X, x2[None], frequency=ff, max_detector_freq=1.2, attenuation="sparse")
    out[f"trace_batch_{nb}_pairs_us"] = (time.perf_counter() - t) / 50 * 1e6
    out[f"trace_batch_{nb}_device_ms_total"] = res.stats["ms_total"]
    out[f"trace_batch_{nb}_launches"] = res.stats["n_launches"]
print(json.dumps(out))
