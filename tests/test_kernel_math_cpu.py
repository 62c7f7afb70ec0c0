"""
The kernels' scalar maths (nuradiomc_b200/csrc/nrmc_math.cuh) compiled for the host by tests/cpu_harness and compared
with the oracle and the reference's goldens.  This validates the algorithm (root bracketing on Snell's invariant,
closed-form properties) on machines without a GPU; the GPU tests then check the same code as compiled by nvcc.
The harness is test-only: the product never runs this code on the CPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, assert_parity, cylinder, load_golden, golden_config, t05_points

HDIR = os.path.join(ROOT, "tests", "cpu_harness")


@pytest.fixture(scope="module")
def harness(oracle_mod):
    so, src = os.path.join(HDIR, "libharness.so"), os.path.join(HDIR, "harness.cpp")
    hdr = os.path.join(ROOT, "nuradiomc_b200", "csrc", "nrmc_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-fPIC", "-shared", "-x", "c++", "-std=c++17", "-ffp-contract=off",
                               "-o", so, src])
    H = C.CDLL(so)

    def run(ice, n_refl, X1, X2):
        n_ice, dn, z0, zr = oracle_mod.ICE_MODELS[ice]
        if zr is None:
            n_refl = 0
        X1 = np.ascontiguousarray(np.atleast_2d(X1), float)
        X2 = np.atleast_2d(np.asarray(X2, float))
        if len(X2) == 1:
            X2 = np.repeat(X2, len(X1), 0)
        X2 = np.ascontiguousarray(X2)
        N, S, K1 = len(X1), 2 + 4 * n_refl, n_refl + 1
        out = {"n_sol": np.zeros(N, np.int32), "status": np.zeros(N, np.int32), "type": np.zeros((N, S), np.int8),
               "reflection": np.zeros((N, S), np.int8), "reflection_case": np.zeros((N, S), np.int8),
               "C0": np.zeros((N, S)), "C1": np.zeros((N, S)), "path_length": np.zeros((N, S)),
               "travel_time": np.zeros((N, S)), "launch": np.zeros((N, S, 3)), "receive": np.zeros((N, S, 3)),
               "reflection_angle": np.zeros((N, S, K1))}
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        H.harness_trace(C.c_double(n_ice), C.c_double(dn), C.c_double(z0), C.c_double(zr or 0.), C.c_int(n_refl), C.c_int64(N),
                        p(X1), p(X2), *[p(out[k]) for k in ("n_sol", "status", "type", "reflection", "reflection_case", "C0", "C1",
                                                           "path_length", "travel_time", "launch", "receive", "reflection_angle")])
        return out

    def focusing(ice, n_refl, X1, X2, res, limit):
        n_ice, dn, z0, zr = oracle_mod.ICE_MODELS[ice]
        if zr is None:
            n_refl = 0
        X1 = np.ascontiguousarray(np.atleast_2d(X1), float)
        X2 = np.ascontiguousarray(np.atleast_2d(X2), float)
        p = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
        out = np.zeros_like(res["C0"])
        H.harness_focusing(C.c_double(n_ice), C.c_double(dn), C.c_double(z0), C.c_double(zr or 0.), C.c_int(n_refl), C.c_int64(len(X1)),
                           p(X1), p(X2), p(res["n_sol"]), p(res["C0"]), p(res["reflection"]), p(res["reflection_case"]),
                           p(res["path_length"]), C.c_double(limit), p(out))
        return out
    run.focusing = focusing
    return run


def _assert_parity(a, b, exact_count=True):
    return assert_parity(a, b, exact_count)


def test_T05_T06_goldens_through_kernel_math(harness):
    ref = np.load(os.path.join(GOLDEN, "reference_C0.npy"))
    out = harness("southpole_simple", 0, t05_points(10, -3000.), [0, 0, -5.])
    assert np.array_equal(out["n_sol"], np.count_nonzero(ref, axis=1))
    np.testing.assert_allclose(np.nan_to_num(out["C0"]), ref, rtol=1e-6)
    ref = np.load(os.path.join(GOLDEN, "reference_C0_MooresBay.npy"))
    out = harness("mooresbay_simple", 2, t05_points(10, -500.), [0, 0, -5.])
    assert np.array_equal(out["n_sol"], np.count_nonzero(ref, axis=1))
    np.testing.assert_allclose(np.nan_to_num(out["C0"]), ref, rtol=1e-6)


@pytest.mark.parametrize("ice,n_refl,rmax,zmin,ant", [
    ("southpole_2015", 0, 4000, -2700, [10, 10, -190.]), ("southpole_simple", 0, 3000, -3000, [0, 0, -5.]),
    ("greenland_simple", 0, 4000, -2700, [0, 20, -97.]), ("greenland_simple", 0, 4000, -2700, [1.5, 11, -2.]),
    ("mooresbay_simple", 1, 1000, -500, [3, 3, -5.]), ("mooresbay_simple", 2, 1000, -570, [-3, 0, -1.]),
    ("ARA_2022", 0, 6000, -2700, [0, 0, -150.]), ("mooresbay_simple_2", 1, 600, -570, [0, 0, -300.])])
def test_kernel_math_vs_oracle(harness, oracle_mod, ice, n_refl, rmax, zmin, ant):
    X1 = cylinder(42 + n_refl, 1500, rmax, zmin)
    a = harness(ice, n_refl, X1, ant)
    b = oracle_mod.Oracle(ice, n_reflections=n_refl).trace(X1, np.array(ant))
    _assert_parity(a, b)
    assert (b["status"] == 0).all()   # the oracle's dense scan never saw more than 2 roots per mode


@pytest.mark.parametrize("name", ["sp_simple_T05", "sp2015_cfg2", "mooresbay_cfg4", "mooresbay_T06"])
def test_kernel_math_vs_python_reference_fixture(harness, name):
    g = load_golden(name)
    c = golden_config(g)
    a = harness(c["ice"], c["n_reflections"], g["X1"], g["X2"])
    same = a["n_sol"] == g["n_sol"]
    assert np.array_equal(a["n_sol"][~same], g["arbiter_n"][~same])   # SURVEY.md F6 protocol
    K1 = a["reflection_angle"].shape[2]
    b = {k: g[k] for k in ("n_sol", "type", "reflection", "reflection_case", "C0", "path_length", "travel_time", "launch", "receive")}
    b["reflection_angle"] = g["reflection_angle"][:, :, :K1]
    _assert_parity(a, b, exact_count=False)


def test_edge_cases(harness, oracle_mod):
    """swapped points, equal depths, receiver at the surface / in air, points below the reflector, vertical pairs"""
    o = oracle_mod.Oracle("southpole_2015")
    X1 = np.array([[0, 0, -50.], [100, 50, -300.], [0, 0, -100.], [0, 0, -100.], [5, 5, -20.], [0, 0, -1000.]])
    X2 = np.array([[200, 0, -300.], [0, 0, -300.], [100, 0, 0.], [50, 0, 1.], [5.001, 5, -10.], [3000, 0, -1000.]])
    a, b = harness("southpole_2015", 0, X1, X2), o.trace(X1, X2)
    assert list(a["n_sol"]) == list(b["n_sol"]) == [2, 2, 1, 0, 2, 2]
    # receiver exactly at the surface: direct and reflected branch coincide; the type is a rounding coin flip in the
    # reference (rho == y_turn analytically, py:1392), the kernel reports 'reflected' -- compare the rest only
    np.testing.assert_allclose(a["C0"][2, 0], b["C0"][2, 0], rtol=1e-6)
    np.testing.assert_allclose(a["path_length"][2, 0], b["path_length"][2, 0], rtol=1e-6)
    assert a["type"][2, 0] == 3
    keep = np.array([0, 1, 3, 4, 5])
    _assert_parity({k: v[keep] for k, v in a.items()}, {k: v[keep] for k, v in b.items() if isinstance(v, np.ndarray) and v.ndim and len(v) == 6})
    assert a["status"][3] == 1                       # in air
    a = harness("mooresbay_simple", 1, [[0, 0, -600.]], [[10, 0, -5.]])
    assert a["n_sol"][0] == 0 and a["status"][0] == 2   # below the reflective layer
    # reciprocity: exchanging start and end point keeps C0 / path / time and exchanges the vectors
    X1, X2 = cylinder(1, 200, 2000, -2000), cylinder(2, 200, 2000, -300)
    a, b = harness("southpole_2015", 0, X1, X2), harness("southpole_2015", 0, X2, X1)
    assert np.array_equal(a["n_sol"], b["n_sol"])
    np.testing.assert_allclose(a["C0"], b["C0"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(a["path_length"], b["path_length"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(a["launch"], b["receive"], atol=1e-12, equal_nan=True)


def test_near_horizontal_rays_against_mpmath(harness, oracle_mod):
    """Where the reference's closed form is ill-conditioned (apex just above two nearly equal depths in deep ice) the
    kernel maths is checked against 40-digit quadrature of ds = n / sqrt(n^2 - beta^2) dz instead of the oracle."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    ice = "mooresbay_simple_2"
    n_ice, dn, z0, zr = oracle_mod.ICE_MODELS[ice]
    X1 = np.array([[0., 0., -300.0192512042436]])
    X2 = np.array([[574.0, 0., -300.]])
    a = harness(ice, 0, X1, X2)
    assert a["n_sol"][0] == 2 and a["type"][0, 0] == 2
    beta = 1 / mp.mpf(float(a["C0"][0, 0]))
    n = lambda z: mp.mpf(n_ice) - mp.mpf(dn) * mp.e ** (z / mp.mpf(z0))
    zt = mp.mpf(z0) * mp.log((mp.mpf(n_ice) - beta) / mp.mpf(dn))
    ds = lambda z: n(z) / mp.sqrt(n(z) ** 2 - beta ** 2)
    path = mp.quad(ds, [X1[0, 2], zt]) + mp.quad(ds, [X2[0, 2], zt])
    ctime = mp.quad(lambda z: ds(z) * n(z), [X1[0, 2], zt]) + mp.quad(lambda z: ds(z) * n(z), [X2[0, 2], zt])
    assert abs(float(path) - a["path_length"][0, 0]) / float(path) < 1e-8
    assert abs(float(ctime) / 0.299792458 - a["travel_time"][0, 0]) / (float(ctime) / 0.299792458) < 1e-8
    b = oracle_mod.Oracle(ice).trace(X1, X2)
    assert abs(b["path_length"][0, 0] - float(path)) / float(path) > 1e-7   # documents the reference's own loss of digits


FAR_FIELD_HUMP = {   # pairs ~10 km apart at a few hundred metres depth: the range curve has a sharp hump next to the beta = n(z2)
    # junction (all three junction values far below rho); an early-exit heuristic of the maximum search once lost them
    "greenland_simple": ([[-9180.4207706, 1164.20989001, -262.20068663], [-8629.38431682, -7236.22619227, -514.73425153],
                          [-6384.5831188, -8814.6125283, -604.84198019]],
                         [[0, 0, -404.48022707], [0, 0, -241.39690073], [0, 0, -218.40253047]]),
    "mooresbay_simple": ([[-5563.01357923, 8648.34810806, -318.91790404], [7929.90411936, 8613.46890284, -340.82411275]],
                         [[0, 0, -303.14734287], [0, 0, -341.08483289]]),
}


@pytest.mark.parametrize("ice", sorted(FAR_FIELD_HUMP))
def test_far_field_sharp_hump_regression(harness, oracle_mod, ice):
    X1, X2 = (np.array(a, float) for a in FAR_FIELD_HUMP[ice])
    out = harness(ice, 0, X1, X2)
    ora = oracle_mod.Oracle(ice).trace(X1, X2)
    assert list(out["n_sol"]) == [2] * len(X1) == list(ora["n_sol"])
    _assert_parity(out, ora)


def test_wide_geometry_stress_counts(harness, oracle_mod):
    """log-uniform depths (0.5 m - 3 km) and distances (1 cm - 15 km): solution counts must equal the oracle's; where they do
    not, the oracle's roots must be a subset of ours (its scan window in log C0 misses rays within 1e-10 of horizontal in
    numerically homogeneous deep ice -- SURVEY.md 8(c) protocol: 'oracle missed a root')"""
    rng = np.random.default_rng(11)
    N = 40000
    for ice in ("southpole_2015", "greenland_simple"):
        zr = -np.exp(rng.uniform(np.log(0.5), np.log(3000.), N))
        ze = -np.exp(rng.uniform(np.log(0.5), np.log(3100.), N))
        rho = np.exp(rng.uniform(np.log(0.01), np.log(15000.), N))
        phi = rng.uniform(0, 2 * np.pi, N)
        X1 = np.stack([rho * np.cos(phi), rho * np.sin(phi), ze], 1)
        X2 = np.stack([np.zeros(N), np.zeros(N), zr], 1)
        out = harness(ice, 0, X1, X2)
        ora = oracle_mod.Oracle(ice).trace(X1, X2, n_threads=8)
        bad = np.nonzero(out["n_sol"] != ora["n_sol"])[0]
        assert len(bad) <= 2
        for i in bad:
            assert out["n_sol"][i] > ora["n_sol"][i]
            for c in ora["C0"][i][:ora["n_sol"][i]]:
                assert np.nanmin(np.abs(out["C0"][i] - c)) < 1e-6 * c
        _assert_parity(out, ora, exact_count=False)


@pytest.mark.parametrize("tag", ["sp", "sp_nolimit", "mb"])
def test_focusing_vs_reference_and_own_difference_quotient(harness, tag):
    """focusing_factor (nrmc_math.cuh; exact d launch / d z_receiver from the closed-form dR/dbeta) against
    (a) the reference's get_focusing (tests/golden/focusing.npz): 5e-3 = the reference's own root-finding noise on its 1 cm
        difference quotient (see test_oracle_golden.py::test_focusing_restatement_vs_reference);
    (b) the difference quotient of the kernel maths' own traces (roots to 1e-16), which converges linearly in dz to the exact
        derivative: 2e-3 at the reference's 1 cm, 2e-4 at 1 mm."""
    g = load_golden("focusing")
    ice, n_refl, limit = str(g[f"{tag}_ice"]), int(g[f"{tag}_n_reflections"]), float(g[f"{tag}_limit"])
    X1, X2, ref = g[f"{tag}_X1"], g[f"{tag}_X2"], g[f"{tag}_focusing"]
    a = harness(ice, n_refl, X1, X2)
    f = harness.focusing(ice, n_refl, X1, X2, a, limit)
    S = a["C0"].shape[1]
    filled = np.arange(S)[None, :] < a["n_sol"][:, None]
    assert np.isnan(f[~filled]).all() and np.isfinite(f[filled]).all()
    ok = filled & (a["n_sol"] == g[f"{tag}_n_sol"])[:, None] & (g[f"{tag}_n_sol"] == g[f"{tag}_n_sol_displaced"])[:, None]
    assert ok.sum() >= 0.98 * np.isfinite(ref).sum()
    np.testing.assert_allclose(f[ok], ref[ok], rtol=5e-3)
    n_ice, dn, z0, _ = __import__("oracle.oracle", fromlist=["x"]).ICE_MODELS[ice]
    n_of = lambda z: n_ice - dn * np.exp(z / z0)
    for dz, tol in ((-1e-2, 3e-3), (-1e-3, 3e-4)):
        X2b = X2.copy()
        X2b[:, 2] += dz
        b = harness(ice, n_refl, X1, X2b)
        same = filled & (a["n_sol"] == b["n_sol"])[:, None] & (a["reflection"] == b["reflection"]) & (a["reflection_case"] == b["reflection_case"])
        with np.errstate(invalid="ignore", divide="ignore"):
            la, lb = np.arccos(a["launch"][..., 2]), np.arccos(b["launch"][..., 2])
            D, rho = a["path_length"], np.linalg.norm((X2 - X1)[:, :2], axis=1)[:, None]
            fd = np.sqrt(D / np.sin(np.arccos(-a["receive"][..., 2])) * np.abs((lb - la) / dz)) * np.sqrt(D * np.sin(la) / rho)
        fd = np.minimum(fd, limit) * np.sqrt(n_of(X1[:, 2]) / n_of(X2[:, 2]))[:, None]
        clipped = np.isclose(np.maximum(fd, f), limit * np.sqrt(n_of(X1[:, 2]) / n_of(X2[:, 2]))[:, None], rtol=1e-2)   # near the limit either may clip first
        sel = same & ~clipped
        assert sel.sum() > 0.8 * filled.sum()
        np.testing.assert_allclose(f[sel], fd[sel], rtol=tol)


def test_newton_evaluations_per_root(harness, oracle_mod):
    """Starting points and termination of the bracketed Newton solver (nrmc_math.cuh::solve_bracket / solve_piece): a warp of
    K_roots waits for its slowest lane, so what matters is the TAIL of the evaluation counts.  cfg5 geometry: direct rays converge
    in 2.5 evaluations on average; turning rays (refracted / reflected), started from the vertex-parabola / chord blend, in about 3
    with fewer than 4 % of the solves above 4 (14 % with the secant start of round 1)."""
    import ctypes as C
    H = C.CDLL(os.path.join(HDIR, "libharness.so"))
    rng = np.random.default_rng(5)
    N = 60000
    r, phi, z = np.sqrt(rng.uniform(0, 6000. ** 2, N)), rng.uniform(0, 2 * np.pi, N), rng.uniform(-2700, 0, N)
    st = rng.integers(0, 25, N)
    X1 = np.ascontiguousarray(np.array([r * np.cos(phi), r * np.sin(phi), z]).T)
    X2 = np.ascontiguousarray(np.array([(st % 5 - 2) * 1500., (st // 5 - 2) * 1500., -145. - 5 * rng.integers(0, 4, N)]).T)
    n_ice, dn, z0, _ = oracle_mod.ICE_MODELS["southpole_2015"]
    ev, pc = np.zeros(2 * N, np.int32), np.zeros(2 * N, np.int8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    H.harness_solver_evaluations(C.c_double(n_ice), C.c_double(dn), C.c_double(z0), C.c_int64(N), p(X1), p(X2), p(ev), p(pc))
    direct, turning = ev[(pc == 0) | (pc == 1)], ev[pc >= 2]
    assert len(direct) > 20000 and len(turning) > 20000
    assert direct.mean() < 2.7 and (direct > 4).mean() < 0.02
    assert turning.mean() < 3.3 and (turning > 4).mean() < 0.04
    assert ev.max() <= 12
