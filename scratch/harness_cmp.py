import sys, ctypes as C, time
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle.oracle import Oracle, ICE_MODELS
H = C.CDLL('/root/repo/tests/cpu_harness/libharness.so')
def harness(ice, n_refl, X1, X2):
    n_ice, dn, z0, zr = ICE_MODELS[ice]
    if zr is None: n_refl = 0
    N = len(X1); S = 2+4*n_refl; K1 = n_refl+1
    out = {"n_sol": np.zeros(N, np.int32), "status": np.zeros(N, np.int32), "type": np.zeros((N,S), np.int8),
           "reflection": np.zeros((N,S), np.int8), "reflection_case": np.zeros((N,S), np.int8),
           "C0": np.zeros((N,S)), "C1": np.zeros((N,S)), "path_length": np.zeros((N,S)), "travel_time": np.zeros((N,S)),
           "launch": np.zeros((N,S,3)), "receive": np.zeros((N,S,3)), "reflection_angle": np.zeros((N,S,K1))}
    X1 = np.ascontiguousarray(X1, float); X2 = np.ascontiguousarray(X2, float)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    H.harness_trace(C.c_double(n_ice), C.c_double(dn), C.c_double(z0), C.c_double(zr or 0.), C.c_int(n_refl), C.c_int64(N), p(X1), p(X2),
        *[p(out[k]) for k in ["n_sol","status","type","reflection","reflection_case","C0","C1","path_length","travel_time","launch","receive","reflection_angle"]])
    return out
def compare(a, b, label):
    ok = a['n_sol']==b['n_sol']
    print(label, 'N', len(ok), 'count mismatch', (~ok).sum(), np.nonzero(~ok)[0][:10])
    m = ok
    def rel(x,y):
        d = np.abs(x-y)/np.maximum(np.abs(y),1e-300); return np.nanmax(d) if np.isfinite(d).any() else 0
    for k in ['type','reflection','reflection_case']:
        mm = (a[k][m]!=b[k][m]).sum()
        if mm: print('  ',k,'mismatch', mm)
    print('   rel: ' + ' '.join(f"{k}={rel(a[k][m], b[k][m]):.2e}" for k in ['C0','path_length','travel_time']),
          'C1 abs %.2e'%np.nanmax(np.abs(a['C1'][m]-b['C1'][m])),
          'launch %.2e receive %.2e'%(np.nanmax(np.abs(a['launch'][m]-b['launch'][m])), np.nanmax(np.abs(a['receive'][m]-b['receive'][m]))),
          'refl_angle %.2e nanmis %d'%(np.nanmax(np.abs(a['reflection_angle'][m]-b['reflection_angle'][m])) if np.isfinite(b['reflection_angle'][m]).any() else 0, (np.isnan(a['reflection_angle'][m])!=np.isnan(b['reflection_angle'][m])).sum()))
    for k in ['C0','C1','path_length','travel_time','launch','receive']:
        nm = (np.isnan(a[k][m])!=np.isnan(b[k][m])).sum()
        if nm: print('   nan mismatch', k, nm)
def cyl(seed, n, rmax, zmin):
    rng = np.random.default_rng(seed)
    r = np.sqrt(rng.uniform(0, rmax**2, n)); ph = rng.uniform(0, 2*np.pi, n); z = rng.uniform(zmin, 0, n)
    return np.array([r*np.cos(ph), r*np.sin(ph), z]).T
if __name__ == '__main__':
    N = int(sys.argv[1]) if len(sys.argv)>1 else 2000
    for ice, n_refl, rmax, zmin, ant in [('southpole_2015',0,4000,-2700,[10,10,-190.]), ('southpole_simple',0,3000,-3000,[0,0,-5.]),
                                        ('greenland_simple',0,4000,-2700,[0,20,-97.]), ('greenland_simple',0,4000,-2700,[1.5,11,-2.]),
                                        ('mooresbay_simple',1,1000,-500,[3,3,-5.]), ('mooresbay_simple',2,1000,-570,[-3,0,-1.]), ('southpole_2015',0,6000,-2700,[0,0,-150.])]:
        X1 = cyl(hash(ice)%1000+n_refl, N, rmax, zmin); X2 = np.repeat([ant], N, 0)
        t=time.time(); h = harness(ice, n_refl, X1, X2); th=time.time()-t
        t=time.time(); o = Oracle(ice, n_reflections=n_refl).trace(X1, X2, n_threads=8); to=time.time()-t
        compare(h, o, f'{ice} refl={n_refl} ant={ant} (harness {th:.2f}s oracle {to:.2f}s)')
