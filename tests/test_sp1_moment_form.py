"""
CPU check of the mathematics behind the SP1 attenuation kernel (K_att_sp1, nuradiomc_b200/csrc/nrmc_rt.cu).  The attenuation
exponent of a path is  I(f) = sum_q w_q / L(z_q, f)  over quadrature nodes q with ds-weights w_q, and for SP1
1/L(z, f) = exp(b1(T) + p(T) ln f): b1 and the slope p are quadratics in the ice temperature T, T a monotone cubic in depth
(attenuation.py:141-142, :170-192).  The kernel expands, per integration frequency f_j, g_j(tau) = 1/L(T(tau), f_j) in
Chebyshev polynomials of the normalised temperature tau in [-1, 1] (surface ... 2800 m):
    I(f_j) = sum_k A_jk M_k,   M_k = sum_q w_q T_k(tau_q)     (12 frequency-independent moments, three-term recurrence)
with A_jk from a 64-point Chebyshev-Gauss rule on the host (nrmc_rt_set_frequencies).  Restated here in numpy; the truncation
must stay below 3e-7 relative (the kernel's budget; the parity tolerance on the attenuation factor is 1e-4).
"""
import numpy as np

K = 12
NC = 64
DEPTH_MAX = 2800.0
B0, B1, B2 = (-6.74890, 0.026709, -0.000884), (-6.22121, -0.070927, -0.001773), (-4.09468, -0.002213, -0.000332)
TC = (1.83415e-09, -1.59061e-08, 0.00267687, -51.0696)


def temperature(depth):
    return ((TC[0] * depth + TC[1]) * depth + TC[2]) * depth + TC[3]


def inv_length(T, f):
    """1 / L of attenuation.py:170-192 as a function of the ice temperature (f in GHz)"""
    b = lambda B: B[0] + B[1] * T + B[2] * T * T
    w = np.log(f)
    slope = (b(B2) - b(B1)) / 1.1505720275988207 if f >= 1.0 else (b(B1) - b(B0)) / 9.210340371976182
    return np.exp(b(B1) + slope * w)


def chebyshev_table(f, n_coef):
    """A_k of g(tau) = 1/L(T(tau), f): the host loop of nrmc_rt_set_frequencies"""
    T0, T1 = temperature(0.0), temperature(DEPTH_MAX)
    theta = np.pi * (np.arange(NC) + 0.5) / NC
    g = inv_length(0.5 * (T0 + T1) + 0.5 * (T1 - T0) * np.cos(theta), f)
    return np.array([(1.0 if k == 0 else 2.0) * np.sum(g * np.cos(k * theta)) / NC for k in range(n_coef)])


def chebyshev_moments(w, tau):
    """M_k = sum_q w_q T_k(tau_q) with t_{k+1} = 2 tau t_k - t_{k-1}: sp1_node"""
    M = np.zeros(K)
    t0, t1 = w.copy(), w * tau
    M[0], M[1] = t0.sum(), t1.sum()
    for k in range(2, K):
        t0, t1 = t1, 2.0 * tau * t1 - t0
        M[k] = t1.sum()
    return M


def test_attenuation_length_restatement_matches_the_oracle():
    """the temperature form above is the reference's SP1 model (checked through the C oracle's get_attenuation_length)"""
    from oracle import oracle
    oracle.build()
    o = oracle.Oracle("southpole_2015", attenuation_model="SP1")
    z = -np.linspace(1.0, 2790.0, 40)
    for f in (0.005, 0.3, 0.99, 1.0, 1.7, 2.5):
        L = np.array([o.attenuation_length(zz, f) for zz in z])
        np.testing.assert_allclose(1.0 / inv_length(temperature(-z), f), L, rtol=1e-12)


def test_moment_form_reproduces_the_direct_sum():
    rng = np.random.default_rng(3)
    grids = {"cfg3/cfg5": np.concatenate([np.linspace(2.5 / 511, 1.2, 25), np.linspace(1.2 + 2.5 / 511, 2.5, 12)]),
             "cfg1": np.linspace(0.5 / 128, 0.5, 100), "1 MHz .. 3 GHz": np.geomspace(1e-3, 3.0, 40)}
    T0, T1 = temperature(0.0), temperature(DEPTH_MAX)
    for name, freqs in grids.items():
        worst = 0.0
        for f in freqs:
            A = chebyshev_table(f, K + 6)
            assert np.abs(A[K:]).sum() <= 3e-7 * abs(A[0]), (name, f)          # the host's own acceptance test
            assert inv_length(np.linspace(T0, T1, 200), f).max() < 1.0           # the 1 m floor is out of reach
            for _ in range(20):
                depth = rng.uniform(0, DEPTH_MAX, 24)
                w = rng.uniform(0.1, 30.0, 24)
                tau = (temperature(depth) - 0.5 * (T0 + T1)) / (0.5 * (T1 - T0))
                assert np.abs(tau).max() <= 1.0
                direct = np.sum(w * inv_length(temperature(depth), f))
                series = float(A[:K] @ chebyshev_moments(w, tau))
                worst = max(worst, abs(series / direct - 1.0))
        assert 0 < worst < 3e-7, (name, worst)
        print(name, "worst relative truncation", worst)


# ---------------------------------------------------------------------------------------------------------------
# table exponentials of K_att_sp1 (sp1_emit / one_minus_exp_neg_tab in nrmc_rt.cu), restated in numpy
# ---------------------------------------------------------------------------------------------------------------
MAGIC = 6755399441055744.0          # 1.5 * 2^52: adding it rounds to the nearest integer, which sits in the low word
TAB = np.exp2(np.arange(256) / 256.0)


def _reduce(x):
    """n = round(x 256 / ln 2), r = x - n ln 2 / 256  (|r| <= ln 2 / 512), as the two FMAs of the kernel"""
    t = x * 369.3299304675746 + MAGIC
    n = (t - MAGIC).astype(np.int64)
    r = x - n * 0.0027076061740622863
    return n, r


def exp_table(x):
    """emit: exp(x) = 2^k 2^(i/256) (1 + r + r^2/2), x <= 0 clamped at -700"""
    x = np.where(x > -700.0, x, -700.0)
    n, r = _reduce(x)
    return np.ldexp(TAB[n & 255] * (1.0 + r * (1.0 + 0.5 * r)), (n >> 8).astype(np.int64))


def one_minus_exp_neg_table(y):
    """node geometry: 1 - exp(-y) = (1 - A) - A q, A = 2^k 2^(i/256), q = r + r^2/2 + r^3/6"""
    n, r = _reduce(-y)
    A = np.ldexp(TAB[n & 255], (n >> 8).astype(np.int64))
    q = r + r * r * (0.5 + r / 6.0)
    return (1.0 - A) - A * q


def test_table_exponential_of_the_emit():
    """truncation r^3/6 <= 4.1e-10 relative on the attenuation factor (tolerance 1e-4), over the whole exponent range"""
    x = -np.concatenate([np.linspace(0, 40, 400001), np.logspace(-12, 2.8, 20001), [0.0, 699.9, 700.0, 1e4]])
    ref = np.exp(np.where(x > -700.0, x, -700.0))
    assert np.max(np.abs(exp_table(x) / ref - 1.0)) < 4.2e-10
    assert exp_table(np.array([0.0]))[0] == 1.0


def test_table_one_minus_exp_of_the_node_geometry():
    """relative accuracy 1e-10 of 1 - exp(-u^2/z0) including u -> 0, where it IS -q (the quadrature weights need 1e-7)"""
    y = np.concatenate([np.logspace(-14, 1.7, 200001), np.linspace(0, 45, 100001)[1:]])
    ref = -np.expm1(-y)
    assert np.max(np.abs(one_minus_exp_neg_table(y) / ref - 1.0)) < 1.2e-10
    assert one_minus_exp_neg_table(np.array([0.0]))[0] == 0.0
