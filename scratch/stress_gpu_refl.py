"""GPU stress with bottom reflections against the oracle"""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from nuradiomc_b200.SignalProp import propagation
from nuradiomc_b200.utilities import medium
from oracle.oracle import Oracle
from conftest import assert_parity
N = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
rng = np.random.default_rng(5)
for ice, nr in (("mooresbay_simple", 1), ("mooresbay_simple", 2), ("mooresbay_simple_2", 3)):
    zr = -np.exp(rng.uniform(np.log(0.5), np.log(570.), N)); ze = -np.exp(rng.uniform(np.log(0.5), np.log(575.), N))
    rho = np.exp(rng.uniform(np.log(0.01), np.log(6000.), N)); phi = rng.uniform(0, 2 * np.pi, N)
    X1 = np.stack([rho * np.cos(phi), rho * np.sin(phi), ze], 1); X2 = np.stack([np.zeros(N), np.zeros(N), zr], 1)
    res = propagation.get_propagation_module("analytic")(medium.get_ice_model(ice), n_reflections=nr).trace_batch(X1, X2)
    o = Oracle(ice, n_reflections=nr).trace(X1, X2, n_threads=16)
    bad = np.nonzero(res["n_sol"] != o["n_sol"])[0]
    print(ice, nr, "N", N, "count mismatches", len(bad), "hist", np.bincount(res["n_sol"]))
    assert_parity(res, o, exact_count=False)
    print("    parity of matching pairs ok")
