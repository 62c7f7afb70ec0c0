"""CPU emulation of the GL1 pole-expansion scheme (K_att_gl1 v2) against the tight oracle on a cfg3 sample:
I_j = sum_q w_q / (A_q - s_j) = (1 / A_half) * 2 / sqrt(a^2 - 1) * sum'_k (-r)^k M_k,  a = (A_mid - s_j) / A_half > 1, r = a - sqrt(a^2 - 1),
M_k = sum_q w_q T_k(x_q), x_q = (A_q - A_mid) / A_half.   python scratch/gl1_pole_emul.py [n_vertices] [K] [nodes per panel] [sub-panels]"""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, bench
from oracle.oracle import Oracle
nvert = int(sys.argv[1]) if len(sys.argv) > 1 else 300
K = int(sys.argv[2]) if len(sys.argv) > 2 else 16
NQ = int(sys.argv[3]) if len(sys.argv) > 3 else 16
SUB = int(sys.argv[4]) if len(sys.argv) > 4 else 2
XPMIN = float(sys.argv[5]) if len(sys.argv) > 5 else 1.3
cfg = bench.CONFIGS["cfg3"]
V, A, ff = bench.workload(nvert, "cfg3")
X1, X2 = bench.pairs_of(V, A, nvert * 24)
import os
if os.environ.get("WIDE"):
    rng = np.random.default_rng(64); n = nvert * 24
    ze, zr = -np.exp(rng.uniform(np.log(0.5), np.log(2900.), n)), -np.exp(rng.uniform(np.log(0.5), np.log(2900.), n))
    rho, phi = np.exp(rng.uniform(np.log(0.1), np.log(9000.), n)), rng.uniform(0, 2 * np.pi, n)
    X1, X2 = np.stack([rho * np.cos(phi), rho * np.sin(phi), ze], 1), np.stack([np.zeros(n), np.zeros(n), zr], 1)
o = Oracle("greenland_simple", attenuation_model="GL1", n_freq=25, tight=True)
out = o.trace(X1, X2, ff, 1.2, dense=False)
fs = out["frequencies_sparse"]; s = 0.55 * (fs * 1e3 - 75)
n_ice, dn, z0 = 1.78, 0.51, 37.25
fit = [1.16052586e+03, 6.87257150e-02, -9.82378264e-05, -3.50628312e-07, -2.21040482e-10, -3.63912864e-14]
Afun = lambda z: np.maximum(sum(c * z ** p for p, c in enumerate(fit)), 100.)
xg, wg = np.polynomial.legendre.leggauss(NQ)
zz_ = np.linspace(-3300, 0, 33001); DADZ = np.abs(np.gradient(Afun(zz_), zz_)).max(); print("max |dA/dz|", DADZ)
stats = dict(easy=0, floor=0, invisible=0, hard=0, items=0)
worst = 0.0; worst_floor = 0.0; n_cmp = 0; sols = 0; sol_with_hard = 0
errs = []
for i in range(len(X1)):
    for k in range(out["n_sol"][i]):
        sols += 1
        z1, z2 = min(X1[i, 2], X2[i, 2]), max(X1[i, 2], X2[i, 2])
        t = out["type"][i, k]; beta = 1 / out["C0"][i, k]; delta = n_ice - beta
        zv = z0 * np.log(delta / dn)
        turned = t >= 2
        uT = np.sqrt(max(zv, 0.0)) if t == 3 else 0.0
        u2, u1 = np.sqrt(max(zv - z2, 0)), np.sqrt(max(zv - z1, 0))
        panels = ([(uT, u2, 2.0)] if turned and u2 > uT else []) + ([(u2, u1, 1.0)] if u1 > u2 else [])
        us, ws = [], []
        for lo, hi, mult in panels:
            edges = np.linspace(lo, hi, SUB + 1)
            for a_, b_ in zip(edges[:-1], edges[1:]):
                h = 0.5 * (b_ - a_); u = 0.5 * (a_ + b_) + h * xg
                em = -np.expm1(-u * u / z0); n = beta + delta * em
                us.append(u); ws.append(mult * h * wg * 2 * u * n / np.sqrt(delta * em * (n + beta)))
        u = np.concatenate(us); w = np.concatenate(ws)
        Aq = Afun(zv - u * u)
        Alo, Ahi = Aq.min(), Aq.max()
        # true range over the path (the kernel: end points + interior extrema of the polynomial)
        zz = np.linspace(z1, min(zv, 0) if turned else z2, 200); Alo = min(Alo, Afun(zz).min()); Ahi = max(Ahi, Afun(zz).max())
        Amid, Ahalf = 0.5 * (Ahi + Alo), max(0.5 * (Ahi - Alo), 1e-3 * 0.5 * (Ahi + Alo))
        x = (Aq - Amid) / Ahalf
        M = np.zeros(K); t0, t1 = w.copy(), w * x; M[0], M[1] = t0.sum(), t1.sum()
        for kk in range(2, K):
            t0, t1 = t1, 2 * x * t1 - t0; M[kk] = t1.sum()
        S = M[0]
        truth = out["attenuation_sparse"][i, k]
        any_hard = False
        for j in range(len(s)):
            stats["items"] += 1
            if Alo - s[j] >= max(1.0, (XPMIN - 1) * Ahalf):
                a = (Amid - s[j]) / Ahalf; sq = np.sqrt(a * a - 1); r = a - sq
                ck = 2 / sq * (-r) ** np.arange(K); ck[0] *= 0.5
                I = float(ck @ M) / Ahalf
                stats["easy"] += 1
                fac = np.exp(-I)
                if truth[j] > 1e-3:
                    e = abs(fac / truth[j] - 1); worst = max(worst, e); n_cmp += 1; errs.append(e)
                else:
                    worst_floor = max(worst_floor, abs(fac - truth[j]))
            elif s[j] >= Ahi - 1.0:
                stats["floor"] += 1
                worst_floor = max(worst_floor, abs(np.exp(-S) - truth[j])) if truth[j] <= 1e-3 else worst_floor
                if truth[j] > 1e-3: worst = max(worst, abs(np.exp(-S) / truth[j] - 1))
            elif min((s[j] + 1 - Alo) / DADZ, (min(zv, 0) if turned else z2) - z1) >= 20 or S / max(Ahi - s[j], 1.0) >= 20:
                stats["invisible"] += 1
                worst_floor = max(worst_floor, truth[j])
            else:
                stats["hard"] += 1; any_hard = True
        sol_with_hard += any_hard
errs = np.array(errs)
print(f"K={K} NQ={NQ} SUB={SUB} XPMIN={XPMIN}: solutions {sols}, items {stats['items']}: easy {stats['easy']/stats['items']:.3f} floor {stats['floor']/stats['items']:.3f} "
      f"invisible {stats['invisible']/stats['items']:.3f} hard {stats['hard']/stats['items']:.4f} (solutions with hard items {sol_with_hard/sols:.3f})")
print(f"easy items compared (truth > 1e-3): {n_cmp}; worst relative error {worst:.2e}; 99.9% quantile {np.quantile(errs, 0.999):.2e}; worst abs error on small/invisible bins {worst_floor:.2e}")
