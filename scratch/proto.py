"""numpy prototype of the kernel math (range function per mode/branch)"""
import numpy as np
class Ice:
    def __init__(s, n_ice, dn, z0, zr=None):
        s.n_ice, s.dn, s.z0, s.zr = n_ice, dn, z0, zr
        s.ns = n_ice - dn
    def n(s, z): return s.n_ice - s.dn*np.exp(z/s.z0)

def coeffs(refl, case, turned):
    k = refl
    if k == 0: return (2, -1, -1, 0) if turned else (0, -1, 1, 0)
    if case == 1: return (2*k+2, -1, -1, -2*k) if turned else (2*k, -1, 1, -2*k)
    return (2*k, 1, -1, -2*k) if turned else (2*k-2, 1, 1, -2*k)

def R(ice, z1, z2, beta, refl, case, turned):
    """total horizontal range for Snell invariant beta"""
    aT, a1, a2, ar = coeffs(refl, case, turned)
    b2 = beta*beta
    c = ice.n_ice**2 - b2; rc = np.sqrt(c)
    def k1(n): return rc*np.sqrt(np.maximum(n*n-b2, 0)) + ice.n_ice*n - b2
    n1, n2 = ice.n(z1), ice.n(z2)
    KT = np.where(beta <= ice.ns, k1(ice.ns), ice.dn*beta)
    lnP = a1*np.log(k1(n1)) + a2*np.log(k1(n2)) + aT*np.log(KT)
    lin = a1*z1 + a2*z2
    if refl:
        nr = ice.n(ice.zr); lnP = lnP + ar*np.log(k1(nr)); lin = lin + ar*ice.zr
    return beta/rc*(lin - ice.z0*lnP)
