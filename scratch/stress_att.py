"""Attenuation stress on wide random geometry against the tight oracle, every model: python scratch/stress_att.py [N]"""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from nuradiomc_b200.SignalProp import propagation
from nuradiomc_b200.utilities import medium, attenuation
from oracle.oracle import Oracle
from conftest import assert_attenuation_parity
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
rng = np.random.default_rng(321)
ff = np.fft.rfftfreq(256, 0.25)     # 0 .. 2 GHz
for ice, model, nr, zmax, rmax in (("southpole_2015", "SP1", 0, 2800., 9000.), ("greenland_simple", "GL1", 0, 2900., 9000.),
                                   ("greenland_simple", "GL2", 0, 2900., 9000.), ("greenland_simple", "GL3", 0, 2900., 6000.),
                                   ("mooresbay_simple", "MB1", 1, 570., 3000.), ("southpole_simple", "SP1", 0, 2800., 9000.)):
    zr = -np.exp(rng.uniform(np.log(0.5), np.log(zmax), N))
    ze = -np.exp(rng.uniform(np.log(0.5), np.log(zmax), N))
    rho = np.exp(rng.uniform(np.log(0.1), np.log(rmax), N))
    phi = rng.uniform(0, 2 * np.pi, N)
    X1 = np.stack([rho * np.cos(phi), rho * np.sin(phi), ze], 1)
    X2 = np.stack([np.zeros(N), np.zeros(N), zr], 1)
    rt = propagation.get_propagation_module("analytic")(medium.get_ice_model(ice), attenuation_model=model, n_reflections=nr,
                                                         n_frequencies_integration=20)
    for fmax in (None, 0.8):
        res = rt.trace_batch(X1, X2, frequency=ff, max_detector_freq=fmax, attenuation="both")
        t = time.time()
        ora = Oracle(ice, attenuation_model=model, n_reflections=nr, n_freq=20, tight=True,
                     gl3_table=attenuation.gl3_parameters() if model == "GL3" else None).trace(X1, X2, ff, fmax, n_threads=16)
        same = res["n_sol"] == ora["n_sol"]
        a, b = res["attenuation"][same], ora["attenuation"][same]
        big = b > 1e-3
        rel = np.nanmax(np.abs(a - b)[big] / b[big]) if big.any() else 0
        ab = np.nanmax(np.abs(a - b)[~big & np.isfinite(b)]) if (~big & np.isfinite(b)).any() else 0
        print(f"{ice} {model} n_refl={nr} fmax={fmax}: N={N} count mismatch {(~same).sum()} max rel dev {rel:.2e} max abs (small bins) {ab:.2e}  oracle {time.time()-t:.1f}s", flush=True)
        assert_attenuation_parity(a, b, atol=2e-7 if model == "GL1" else 1e-7)   # GL1 above ~1 GHz: factors ~1e-6 next to the pole of 1/(A - s_f)
