"""CPU emulation of the GL1 quadrature schemes of K_att (scheme 0) and K_att_gl1 (scheme 3) against the tight oracle:\n g++ -O2 -fPIC -shared -std=c++17 -o /tmp/libgl1.so scratch/gl1_emul.cpp && python scratch/gl1_emul.py"""
import sys, ctypes as C, numpy as np, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle.oracle import Oracle, ICE_MODELS
from conftest import cylinder
from test_gpu_parity import RNOG
L=C.CDLL('/tmp/libgl1.so')
def emul(X1,X2,sp,scheme,margin=60.0,ratio=1.0,n1=4,n0=2):
    n_ice,dn,z0,_=ICE_MODELS["greenland_simple"]
    N=len(X1); Fs=len(sp); fa=np.ascontiguousarray(0.55*(sp/1e-3-75.0))
    out=np.zeros((N,2,Fs)); ns=np.zeros(N,np.int32); work=(C.c_long*5)(); work[3]=n1; work[4]=n0
    p=lambda a:a.ctypes.data_as(C.c_void_p)
    L.gl1_emul(C.c_double(n_ice),C.c_double(dn),C.c_double(z0),C.c_int64(N),p(np.ascontiguousarray(X1)),p(np.ascontiguousarray(X2)),C.c_int(Fs),p(fa),C.c_int(scheme),C.c_double(margin),C.c_double(ratio),p(out),p(ns),work)
    return out,ns,(work[0],work[1],work[2])
def compare(name,X1,X2,ff,fmax,nfreq):
    o=Oracle("greenland_simple",attenuation_model="GL1",n_freq=nfreq,tight=True)
    t=time.time(); ora=o.trace(X1,X2,ff,fmax,n_threads=8,dense=True); sp=ora["frequencies_sparse"]; b=ora["attenuation_sparse"]
    for scheme,margin,ratio,n1,n0 in CASES:
        a,ns,work=emul(X1,X2,sp,scheme,margin,ratio,n1,n0)
        same=ns==ora["n_sol"]
        aa,bb=a[same],b[same]
        with np.errstate(invalid="ignore",divide="ignore"):
            big=bb>1e-3
            rel=np.nanmax(np.abs(aa-bb)[big]/bb[big]); ab=np.nanmax(np.abs(aa-bb)[~big&np.isfinite(bb)])
        # dense bins: np.interp of the sparse factors (py:1077-1078), f = 0 bin = 1
        m=ff>0
        ad=np.ones(aa.shape[:-1]+(len(ff),)); 
        idx=np.clip(np.searchsorted(sp,ff[m],side="right")-1,0,len(sp)-2); t=(ff[m]-sp[idx])/(sp[idx+1]-sp[idx]); t=np.clip(t,0,1)
        ad[...,m]=aa[...,idx]*(1-t)+aa[...,idx+1]*t
        bd=ora["attenuation"][same]
        with np.errstate(invalid="ignore",divide="ignore"):
            bigd=bd>1e-3; reld=np.nanmax(np.abs(ad-bd)[bigd]/bd[bigd])
        print(f"   dense max rel {reld:.2e}",end=" ")
        print(f"{name} scheme {scheme} margin {margin} ratio {ratio} n1 {n1} n0 {n0}: count mismatch {(~same).sum()} max rel {rel:.2e} max abs small {ab:.2e} work/solution {work[0]/max(ns.sum(),1):.0f} redo {work[1]}/{work[2]}",flush=True)
CASES=((0,60.,1.,4,2),(3,10.,.4,3,1))
_OLD=((2,10.,0.4,4,1),(2,10.,0.4,3,1),(2,10.,0.3,3,1),(2,15.,0.4,3,1),(2,10.,0.5,3,2),(2,10.,0.35,4,2))

ff=np.fft.rfftfreq(1022,0.2)
V=cylinder(63,1500,4000,-2700); A=RNOG[[0,8,13,21]]
X1,X2=np.repeat(V,len(A),0),np.tile(A,(len(V),1))
rng=np.random.default_rng(64); n=3000
ze,zr=-np.exp(rng.uniform(np.log(0.5),np.log(2900.),n)),-np.exp(rng.uniform(np.log(0.5),np.log(2900.),n))
rho,phi=np.exp(rng.uniform(np.log(0.1),np.log(9000.),n)),rng.uniform(0,2*np.pi,n)
W1,W2=np.stack([rho*np.cos(phi),rho*np.sin(phi),ze],1),np.stack([np.zeros(n),np.zeros(n),zr],1)
compare("test cfg3",X1[:2000],X2[:2000],ff,1.2,25)
compare("test wide None",W1[:2000],W2[:2000],ff,None,20)
compare("test wide 0.8",W1[:2000],W2[:2000],ff,0.8,20)
