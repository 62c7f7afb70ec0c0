"""CPU stress with bottom reflections (mooresbay, n_reflections = 2): kernel maths harness vs oracle, counts / order / values"""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scratch'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from harness_cmp import harness
from oracle.oracle import Oracle
from conftest import assert_parity
N = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
rng = np.random.default_rng(5)
for ice, nr in (("mooresbay_simple", 1), ("mooresbay_simple", 2), ("mooresbay_simple_2", 2)):
    zr = -np.exp(rng.uniform(np.log(0.5), np.log(570.), N))
    ze = -np.exp(rng.uniform(np.log(0.5), np.log(575.), N))
    rho = np.exp(rng.uniform(np.log(0.01), np.log(6000.), N))
    phi = rng.uniform(0, 2 * np.pi, N)
    X1 = np.stack([rho * np.cos(phi), rho * np.sin(phi), ze], 1)
    X2 = np.stack([np.zeros(N), np.zeros(N), zr], 1)
    h = harness(ice, nr, X1, X2)
    o = Oracle(ice, n_reflections=nr).trace(X1, X2, n_threads=8)
    bad = np.nonzero(h["n_sol"] != o["n_sol"])[0]
    print(ice, nr, "N", N, "count mismatches", len(bad), "hist", np.bincount(h["n_sol"]))
    for i in bad[:6]:
        print("   ", X1[i], X2[i], "harness", h["n_sol"][i], h["C0"][i], "oracle", o["n_sol"][i], o["C0"][i])
    assert_parity(h, o, exact_count=False)
    print("    parity of matching pairs ok")
