import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from nuradiomc_b200.SignalProp import propagation
from nuradiomc_b200.utilities import medium, attenuation
from oracle.oracle import Oracle
N=3000
rng = np.random.default_rng(321)
ff = np.fft.rfftfreq(256, 0.25)
# consume the rng as stress_att does for the first config
for _ in range(1):
    rng.uniform(size=N); rng.uniform(size=N); rng.uniform(size=N); rng.uniform(size=N)
zmax, rmax = 2900., 9000.
zr = -np.exp(rng.uniform(np.log(0.5), np.log(zmax), N)); ze = -np.exp(rng.uniform(np.log(0.5), np.log(zmax), N))
rho = np.exp(rng.uniform(np.log(0.1), np.log(rmax), N)); phi = rng.uniform(0, 2 * np.pi, N)
X1 = np.stack([rho * np.cos(phi), rho * np.sin(phi), ze], 1); X2 = np.stack([np.zeros(N), np.zeros(N), zr], 1)
rt = propagation.get_propagation_module("analytic")(medium.get_ice_model("greenland_simple"), attenuation_model="GL1", n_frequencies_integration=20)
res = rt.trace_batch(X1, X2, frequency=ff, attenuation="both")
ora = Oracle("greenland_simple", attenuation_model="GL1", n_freq=20, tight=True).trace(X1, X2, ff, None, n_threads=16)
a, b = res["attenuation_sparse"], ora["attenuation_sparse"]
with np.errstate(invalid="ignore", divide="ignore"):
    rel = np.where(b > 1e-3, np.abs(a - b) / b, 0)
rel = np.nan_to_num(rel)
for _ in range(6):
    i, s, j = np.unravel_index(np.argmax(rel), rel.shape)
    f = res.frequencies_sparse[j]
    L = attenuation.get_attenuation_length(np.array([X1[i, 2], X2[i, 2]]), np.array([f, f]), "GL1")
    print(f"rel {rel[i,s,j]:.2e} pair {i} slot {s} type {res['solution_type'][i,s]} f={f:.3f} GHz z1={X1[i,2]:.1f} z2={X2[i,2]:.1f} rho={rho[i]:.1f} path={res['path_length'][i,s]:.2f} factor={b[i,s,j]:.3e} L(z1),L(z2)={L}")
    rel[i] = 0
print("---- dense")
a, b = res["attenuation"], ora["attenuation"]
with np.errstate(invalid="ignore", divide="ignore"):
    rel = np.nan_to_num(np.where(b > 1e-3, np.abs(a - b) / b, 0))
sp = res.frequencies_sparse
for _ in range(4):
    i, s, j = np.unravel_index(np.argmax(rel), rel.shape)
    k = np.searchsorted(sp, ff[j]) - 1
    print(f"rel {rel[i,s,j]:.2e} pair {i} slot {s} f={ff[j]:.4f} dense ours {a[i,s,j]:.6e} oracle {b[i,s,j]:.6e}; sparse neighbours f={sp[k]:.4f},{sp[k+1]:.4f} ours {res['attenuation_sparse'][i,s,k]:.6e},{res['attenuation_sparse'][i,s,k+1]:.6e} oracle {ora['attenuation_sparse'][i,s,k]:.6e},{ora['attenuation_sparse'][i,s,k+1]:.6e} path {res['path_length'][i,s]:.3f}")
    rel[i] = 0
