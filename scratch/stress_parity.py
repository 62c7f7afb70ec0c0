"""Large randomised parity run against the oracle (counts, types, C0, path, time): python scratch/stress_parity.py [N]"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from nuradiomc_b200.SignalProp import propagation
from nuradiomc_b200.utilities import medium
from oracle.oracle import Oracle
sys.path.insert(0, '/root/repo/tests')
from conftest import assert_parity
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rng = np.random.default_rng(123)
for ice in ("southpole_2015", "greenland_simple", "mooresbay_simple", "southpole_simple"):
    rt = propagation.get_propagation_module("analytic")(medium.get_ice_model(ice))
    # wide geometry: receivers from 0.5 m to 3000 m depth, emitters anywhere, distances from centimetres to 12 km
    zr = -np.exp(rng.uniform(np.log(0.5), np.log(3000.), N))
    ze = -np.exp(rng.uniform(np.log(0.5), np.log(3100.), N))
    rho = np.exp(rng.uniform(np.log(0.01), np.log(12000.), N))
    phi = rng.uniform(0, 2 * np.pi, N)
    X1 = np.stack([rho * np.cos(phi), rho * np.sin(phi), ze], 1)
    X2 = np.stack([np.zeros(N), np.zeros(N), zr], 1)
    t = time.time(); res = rt.trace_batch(X1, X2); tg = time.time() - t
    t = time.time(); ora = Oracle(ice).trace(X1, X2, n_threads=16); to = time.time() - t
    same = res["n_sol"] == ora["n_sol"]
    bad = np.nonzero(~same)[0]
    print(f"{ice}: N={N} gpu {tg:.2f}s oracle {to:.1f}s count mismatches {len(bad)}; n_sol hist {np.bincount(res['n_sol'])}")
    for i in bad[:10]:
        print("   mismatch", i, X1[i], X2[i], "gpu", res["n_sol"][i], res["C0"][i], "oracle", ora["n_sol"][i], ora["C0"][i])
    nm = assert_parity(res, ora, exact_count=False)
    print("   parity on the matching pairs ok")
