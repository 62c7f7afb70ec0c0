"""first-contact GPU check: parity vs oracle on several configs + rough timings (scratch tool, not a test)"""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nuradiomc_b200.SignalProp import propagation
from nuradiomc_b200.SignalProp.analyticraytracing import measure_fp64_peak
from nuradiomc_b200.utilities import medium
from oracle.oracle import Oracle

def cyl(seed, n, rmax, zmin):
    rng = np.random.default_rng(seed)
    r = np.sqrt(rng.uniform(0, rmax**2, n)); ph = rng.uniform(0, 2*np.pi, n); z = rng.uniform(zmin, 0, n)
    return np.array([r*np.cos(ph), r*np.sin(ph), z]).T

def compare(res, ora, label, att=False):
    ok = res['n_sol'] == ora['n_sol']
    line = f"{label}: N={len(ok)} count_mismatch={int((~ok).sum())}"
    m = ok
    line += f" type_mis={int((res['solution_type'][m] != ora['type'][m]).sum())}"
    def rel(a, b):
        d = np.abs(a-b)/np.maximum(np.abs(b), 1e-300)
        return np.nanmax(d) if np.isfinite(d).any() else 0.
    for k in ['C0', 'path_length', 'travel_time']:
        line += f" {k}={rel(res[k][m], ora[k][m]):.1e}"
    line += f" launch={np.nanmax(np.abs(res['launch_vector'][m]-ora['launch'][m])):.1e} recv={np.nanmax(np.abs(res['receive_vector'][m]-ora['receive'][m])):.1e}"
    ra, rb = res['reflection_angle'][m], ora['reflection_angle'][m]
    line += f" refl_nanmis={int((np.isnan(ra) != np.isnan(rb)).sum())}"
    if att:
        for k in ['attenuation', 'attenuation_sparse']:
            if k in res and k in ora:
                a, b = res[k][m], ora[k][m]
                big = b > 1e-3
                line += f" {k}: rel(>1e-3)={np.nanmax(np.abs(a-b)[big]/b[big]):.1e} abs={np.nanmax(np.abs(a-b)):.1e} nanmis={int((np.isnan(a)!=np.isnan(b)).sum())}"
    print(line, flush=True)

prop = propagation.get_propagation_module('analytic')
print(torch.cuda.get_device_name(0))
t, clk = measure_fp64_peak(0, 2.0)
print(f"FP64 FMA peak measured: {t:.2f} TFLOP/s (nominal clock {clk} MHz)", flush=True)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
# 1. no attenuation configs
for ice, n_refl, rmax, zmin, ant in [('southpole_2015',0,4000,-2700,[10,10,-190.]), ('southpole_simple',0,3000,-3000,[0,0,-5.]),
                                    ('greenland_simple',0,4000,-2700,[1.5,11,-2.]), ('mooresbay_simple',1,1000,-500,[3,3,-5.]),
                                    ('mooresbay_simple',2,1000,-570,[-3,0,-1.])]:
    X1 = cyl(7+n_refl, N, rmax, zmin); X2 = np.repeat([ant], N, 0)
    rt = prop(medium.get_ice_model(ice), n_reflections=n_refl)
    res = rt.trace_batch(X1, X2)
    ora = Oracle(ice, n_reflections=n_refl).trace(X1, X2)
    compare(res, ora, f"{ice} refl={n_refl}")
# 2. attenuation
Na = 400
for ice, model, n_refl, rmax, zmin, ant, ff, fmax, nfreq in [
        ('southpole_2015','SP1',0,6000,-2700,[0,0,-150.], np.fft.rfftfreq(1022,0.2), 1.2, 25),
        ('southpole_simple','SP1',0,3000,-3000,[0,0,-100.], np.linspace(0,0.5,129), None, 100),
        ('greenland_simple','GL1',0,4000,-2700,[0,20,-97.], np.fft.rfftfreq(1022,0.2), 1.2, 25),
        ('greenland_simple','GL1',0,4000,-2700,[1.5,11,-2.], np.fft.rfftfreq(1022,0.2), 1.2, 25),
        ('greenland_simple','GL2',0,4000,-2700,[0,20,-97.], np.fft.rfftfreq(256,0.5), None, 25),
        ('mooresbay_simple','MB1',1,1000,-500,[3,3,-5.], np.fft.rfftfreq(256,0.5), None, 25),
        ('mooresbay_simple','MB1',2,1000,-570,[3,3,-5.], np.fft.rfftfreq(256,0.5), None, 25)]:
    X1 = cyl(11+n_refl, Na, rmax, zmin); X2 = np.repeat([ant], Na, 0)
    rt = prop(medium.get_ice_model(ice), attenuation_model=model, n_reflections=n_refl, n_frequencies_integration=nfreq)
    res = rt.trace_batch(X1, X2, frequency=ff, max_detector_freq=fmax, attenuation='both')
    ora = Oracle(ice, attenuation_model=model, n_reflections=n_refl, n_freq=nfreq, tight=True).trace(X1, X2, ff, fmax)
    compare(res, ora, f"{ice} {model} refl={n_refl}", att=True)
# 3. timing (device resident, cfg5-like)
dev = torch.device('cuda:0')
rt = prop(medium.get_ice_model('southpole_2015'), attenuation_model='SP1', n_frequencies_integration=25)
ff = np.fft.rfftfreq(1022, 0.2)
for Nv in [100000, 1000000]:
    V = torch.tensor(cyl(5, Nv, 6000, -2700).T.copy(), device=dev)
    A = torch.tensor(np.array([[x, y, z] for x in (-3000,-1500,0,1500,3000) for y in (-3000,-1500,0,1500,3000) for z in (-145,-150,-155,-160.)][:10]).T.copy(), device=dev)
    for label, kw in [('no-att', dict()), ('SP1 sparse', dict(frequency=ff, max_detector_freq=1.2, attenuation='sparse'))]:
        out = None
        for rep in range(3):
            out = rt.trace_batch_device(V, A, outer=True, out=out, sync_stats=True, **kw)
        st = out.stats
        npairs = Nv*A.shape[1]
        print(f"timing {label}: pairs={npairs} ms_total={st['ms_total']:.2f} solve={st['ms_solve']:.2f} att={st['ms_attenuation']:.2f} -> {npairs/st['ms_total']*1e3:.3e} pairs/s; nsol={int(out['n_sol'].sum())}", flush=True)
