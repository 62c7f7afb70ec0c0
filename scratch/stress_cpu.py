"""CPU stress of the kernels' scalar maths (tests/cpu_harness) against the oracle on wide random geometry: counts must agree"""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scratch')
import numpy as np
from harness_cmp import harness
from oracle.oracle import Oracle
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 7
rng = np.random.default_rng(seed)
tot_bad = 0
for ice in ("southpole_2015", "greenland_simple", "mooresbay_simple", "southpole_simple"):
    zr = -np.exp(rng.uniform(np.log(0.5), np.log(3000.), N))
    ze = -np.exp(rng.uniform(np.log(0.5), np.log(3100.), N))
    rho = np.exp(rng.uniform(np.log(0.01), np.log(15000.), N))
    phi = rng.uniform(0, 2 * np.pi, N)
    X1 = np.stack([rho * np.cos(phi), rho * np.sin(phi), ze], 1)
    X2 = np.stack([np.zeros(N), np.zeros(N), zr], 1)
    h = harness(ice, 0, X1, X2)
    o = Oracle(ice).trace(X1, X2, n_threads=8)
    bad = np.nonzero(h["n_sol"] != o["n_sol"])[0]
    tot_bad += len(bad)
    print(ice, "N", N, "count mismatches", len(bad), "hist", np.bincount(h["n_sol"]))
    for i in bad[:5]:
        print("   ", X1[i], X2[i], "harness", h["n_sol"][i], h["C0"][i], "oracle", o["n_sol"][i], o["C0"][i])
print("TOTAL mismatches", tot_bad)
