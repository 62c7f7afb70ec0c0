"""
CPU check of the mathematics behind the SP1 attenuation kernel (K_att_sp1, nuradiomc_b200/csrc/nrmc_rt.cu): the attenuation
exponent  I(f) = sum_q c_q exp(p_q ln f)  over quadrature nodes q (c_q = ds-weight / L(z_q, 1 GHz), p_q the slope of
ln(1/L) in ln f at the node's depth, attenuation.py:170-192) is evaluated from 8 frequency-independent Chebyshev moments
per band,  I(f) = exp(p_ref w) sum_k eps_k I_k(r w) M_k,  M_k = sum_q c_q T_k((p_q - p_ref) / r),  w = ln f,  |r w| <= 0.9.
Restated here in numpy with the kernel's three-term recurrence and the host code's Bessel series; truncation must stay
below 2.5e-7 relative (the kernel's own budget; the parity tolerance on the attenuation factor is 1e-4).
"""
import math

import numpy as np

K = 8


def bessel_table(z):
    """eps_k I_k(z), k < K, by the ascending series -- the loop of nrmc_rt_set_frequencies (SP1 tables)"""
    hz = 0.5 * abs(z)
    out = np.zeros(K)
    for k in range(K):
        term = 1.0
        for i in range(1, k + 1):
            term *= hz / i
        total = 0.0
        for m in range(40):
            total += term
            term *= hz * hz / ((m + 1.0) * (m + 1.0 + k))
        out[k] = (1.0 if k == 0 else 2.0) * (-total if (z < 0 and k & 1) else total)
    return out


def chebyshev_moments(c, x):
    """M_k = sum_q c_q T_k(x_q) with t_{k+1} = 2 x t_k - t_{k-1}: sp1_node"""
    M = np.zeros(K)
    t0, t1 = c.copy(), c * x
    M[0], M[1] = t0.sum(), t1.sum()
    for k in range(2, K):
        t0, t1 = t1, 2.0 * x * t1 - t0
        M[k] = t1.sum()
    return M


def test_bessel_series_against_scipy():
    from scipy.special import iv
    for z in np.linspace(-0.9, 0.9, 19):
        ref = np.array([(1 if k == 0 else 2) * iv(k, z) for k in range(K)])
        np.testing.assert_allclose(bessel_table(z), ref, rtol=1e-14, atol=1e-300)


def test_moment_form_reproduces_the_direct_sum():
    # SP1 slopes over the temperature range of South Pole ice (attenuation.py:141-142, :176-185)
    B0, B1, B2 = (-6.74890, 0.026709, -0.000884), (-6.22121, -0.070927, -0.001773), (-4.09468, -0.002213, -0.000332)
    b = lambda B, T: B[0] + B[1] * T + B[2] * T * T
    rng = np.random.default_rng(3)
    freqs = np.concatenate([np.linspace(2.5 / 511, 1.2, 25), np.linspace(1.2 + 2.5 / 511, 2.5, 12)])     # cfg3 / cfg5 grid
    for band, sel, pref in ((0, freqs < 1.0, 0.24), (1, freqs >= 1.0, 1.75)):
        w = np.log(freqs[sel])
        r = 0.9 / np.abs(w).max()
        worst = 0.0
        for _ in range(200):
            depth = np.sort(rng.uniform(0, 2700, 24))
            T = 1.83415e-09 * depth ** 3 - 1.59061e-08 * depth ** 2 + 0.00267687 * depth - 51.0696
            p = (b(B1, T) - b(B0, T)) / 9.210340371976182 if band == 0 else (b(B2, T) - b(B1, T)) / 1.1505720275988207   # ln 1e4, ln 3.16
            x = (p - pref) / r
            if np.abs(x).max() > 1.0:        # outside the series' band: the kernel hands such paths to the generic kernel
                continue
            c = rng.uniform(0.1, 3.0, 24) * np.exp(b(B1, T))
            M = chebyshev_moments(c, x)
            for wj in w:
                direct = np.sum(c * np.exp(p * wj))
                series = math.exp(pref * wj) * float(bessel_table(r * wj) @ M)
                worst = max(worst, abs(series / direct - 1.0))
        assert 0 < worst < 2.5e-7, (band, worst)
