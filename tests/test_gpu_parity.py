"""
GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI (include/nrmc_rt.h) via
the Python `ray_tracing` class, against the oracle on seeded inputs, against the committed golden fixtures, and -- at the
BASELINE sizes -- through size-independent properties.  Tolerances are BASELINE.json's: solution count and type
bit-exact; angles 1e-6 rad; path length / travel time 1e-6 relative; attenuation 1e-4 relative.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_attenuation_parity, assert_parity, cylinder, golden_config, load_golden, t05_points

pytestmark = pytest.mark.gpu

RNOG = np.array([[0, 20, -97], [0, 20, -96], [0, 20, -95], [0, 20, -94], [0, 20, -93], [0, 20, -92], [0, 20, -80],
                 [0, 20, -60], [0, 20, -40], [-17.3, -10, -96], [-17.3, -10, -95], [-17.3, -10, -94], [1.5, 11, -2],
                 [0, 11, -2], [-1.5, 11, -2], [-10.276, -4.2, -2], [-9.526, -5.5, -2], [-8.776, -6.8, -2],
                 [8.776, -6.8, -2], [9.526, -5.5, -2], [10.276, -4.2, -2], [17.3, -10, -96], [17.3, -10, -95],
                 [17.3, -10, -94]], float)


@pytest.fixture(scope="module")
def make():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from nuradiomc_b200.SignalProp import propagation
    from nuradiomc_b200.utilities import medium
    prop = propagation.get_propagation_module("analytic")

    def _make(ice, **kw):
        return prop(medium.get_ice_model(ice), **kw)
    return _make


# ---------------------------------------------------------------------------------------------------------------
# reference goldens through the kernels
# ---------------------------------------------------------------------------------------------------------------
def test_T05_golden_through_kernel(make):
    """T05unit_test_C0_SP.py: reference_C0.pkl, identical counts and slot order"""
    ref = np.load(os.path.join(GOLDEN, "reference_C0.npy"))
    res = make("southpole_simple").trace_batch(t05_points(10, -3000.), np.array([[0, 0, -5.]]))
    assert np.array_equal(res["n_sol"], np.count_nonzero(ref, axis=1))
    np.testing.assert_allclose(np.nan_to_num(res["C0"]), ref, rtol=1e-6)


def test_T06_golden_through_kernel(make):
    """T06unit_test_C0_mooresbay.py: reference_C0_MooresBay.pkl (n_reflections=2, up to 10 solutions)"""
    ref = np.load(os.path.join(GOLDEN, "reference_C0_MooresBay.npy"))
    res = make("mooresbay_simple", n_reflections=2).trace_batch(t05_points(10, -500.), np.array([[0, 0, -5.]]))
    assert np.array_equal(res["n_sol"], np.count_nonzero(ref, axis=1))
    np.testing.assert_allclose(np.nan_to_num(res["C0"]), ref, rtol=1e-6)


@pytest.mark.parametrize("name", ["sp_simple_T05", "sp2015_cfg2", "mooresbay_cfg4", "mooresbay_T06", "sp1_cfg1",
                                  "greenland_cfg3", "mooresbay_cfg4_MB1", "sp2015_cfg5_SP1", "greenland_GL2", "greenland_GL3",
                                  "mooresbay_GL3", "greenland_cfg3_large", "sp2015_cfg5_large", "greenland_GL2_large",
                                  "mooresbay_MB1_direct"])
def test_python_reference_fixtures(make, name):
    """fixtures produced by the reference's own Python path (tests/golden/make_golden.py)"""
    g = load_golden(name)
    c = golden_config(g)
    rt = make(c["ice"], attenuation_model=c["attenuation_model"], n_reflections=c["n_reflections"],
              n_frequencies_integration=c["n_freq"])
    res = rt.trace_batch(g["X1"], g["X2"], frequency=c["frequencies"], max_detector_freq=c["max_detector_freq"])
    same = res["n_sol"] == g["n_sol"]
    assert np.array_equal(res["n_sol"][~same], g["arbiter_n"][~same])   # the arbiter decides
    assert_parity(res, g, exact_count=False)
    if c["attenuation_model"]:
        # types 1,3 and 2 alike against the reference integrand at quad(epsrel=1e-11) ...
        assert_attenuation_parity(res["attenuation"][same], g["attenuation_tight"][same])
        # ... and within the reference's own tolerance of its stock epsrel=1e-2 result (T01test_python_vs_cpp.py:86-90)
        np.testing.assert_allclose(res["attenuation"][same], g["attenuation"][same], rtol=1e-2, atol=1e-3, equal_nan=True)


# ---------------------------------------------------------------------------------------------------------------
# oracle on seeded inputs, one case per BASELINE config
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ice,n_refl,rmax,zmin,antennas", [
    ("southpole_2015", 0, 4000, -2700, [[10, 10, -190.], [10, -10, -190.], [-10, -10, -190.], [-10, 10, -190.]]),   # cfg2
    ("southpole_simple", 0, 3000, -3000, [[0, 0, -100.]]),                                                          # cfg1
    ("greenland_simple", 0, 4000, -2700, RNOG[[0, 6, 8, 9, 12, 17, 22]].tolist()),                                  # cfg3
    ("mooresbay_simple", 1, 1000, -500, [[-3, 0, -1.], [3, 3, -5.]]),                                               # cfg4
    ("mooresbay_simple", 2, 1000, -570, [[0, 0, -5.]]),
    ("southpole_2015", 0, 6000, -2700, [[0, 0, -145.], [1500, -3000, -160.]]),                                      # cfg5
    ("ARA_2022", 0, 3000, -1500, [[0, 0, -50.]]),
])
def test_solutions_vs_oracle(make, oracle_mod, ice, n_refl, rmax, zmin, antennas):
    V = cylinder(100 + n_refl, 4000 // len(antennas), rmax, zmin)
    A = np.array(antennas, float)
    res = make(ice, n_reflections=n_refl).trace_batch(V, A, outer=True)
    X1, X2 = np.repeat(V, len(A), axis=0), np.tile(A, (len(V), 1))
    ora = oracle_mod.Oracle(ice, n_reflections=n_refl).trace(X1, X2)
    assert (ora["status"] == 0).all()
    assert_parity(res, ora)
    assert (res["status"] == 0).all()


@pytest.mark.parametrize("ice,model,n_refl,rmax,zmin,ant,freqs,fmax,n_freq", [
    ("southpole_simple", "SP1", 0, 3000, -3000, [0, 0, -100.], np.linspace(0, 0.5, 129), None, 100),               # cfg1
    ("greenland_simple", "GL1", 0, 4000, -2700, [0, 20, -97.], np.fft.rfftfreq(1022, 0.2), 1.2, 25),                # cfg3 deep
    ("greenland_simple", "GL1", 0, 4000, -2700, [1.5, 11, -2.], np.fft.rfftfreq(1022, 0.2), 1.2, 25),               # cfg3 LPDA
    ("mooresbay_simple", "MB1", 1, 1000, -500, [3, 3, -5.], np.fft.rfftfreq(256, 0.5), None, 25),                   # cfg4
    ("southpole_2015", "SP1", 0, 6000, -2700, [1500, 1500, -150.], np.fft.rfftfreq(1022, 0.2), 1.2, 25),            # cfg5
    ("greenland_simple", "GL2", 0, 3000, -2500, [0, 0, -100.], np.fft.rfftfreq(256, 0.5), None, 20),
    ("mooresbay_simple", "MB1", 2, 800, -570, [0, 0, -5.], np.fft.rfftfreq(256, 0.5), 0.6, 10),
])
def test_attenuation_vs_tight_oracle(make, oracle_mod, ice, model, n_refl, rmax, zmin, ant, freqs, fmax, n_freq):
    from nuradiomc_b200.utilities import attenuation
    X1 = cylinder(200 + n_refl, 500, rmax, zmin)
    X2 = np.repeat([ant], len(X1), axis=0)
    rt = make(ice, attenuation_model=model, n_reflections=n_refl, n_frequencies_integration=n_freq)
    res = rt.trace_batch(X1, X2, frequency=freqs, max_detector_freq=fmax, attenuation="both")
    gl3 = attenuation.gl3_parameters() if model == "GL3" else None
    ora = oracle_mod.Oracle(ice, attenuation_model=model, n_reflections=n_refl, n_freq=n_freq, tight=True,
                            gl3_table=gl3).trace(X1, X2, freqs, fmax)
    assert_parity(res, ora)
    np.testing.assert_allclose(res.frequencies_sparse, ora["frequencies_sparse"], rtol=1e-15)
    assert_attenuation_parity(res["attenuation_sparse"], ora["attenuation_sparse"])
    assert_attenuation_parity(res["attenuation"], ora["attenuation"])
    assert (res["attenuation"][:, :, freqs <= 0][res["n_sol"] > 0][:, 0] == 1.0).all()   # the 0 Hz bin is 1 (py:1077)


def test_gl3_discretised_path_attenuation_vs_oracle(make, oracle_mod):
    """GL3: the reference defines the result by its 10 m midpoint discretisation (analyticraytracing.py:998-1064); the kernel
    reproduces it cell for cell.  Seeded pairs against the oracle (tight window quadrature), deep and shallow receivers."""
    from nuradiomc_b200.utilities import attenuation
    ff = np.fft.rfftfreq(256, 0.5)
    V = cylinder(31, 300, 3500, -2900)
    for ant in ([0, 20, -97.], [1.5, 11, -2.], [0, 0, -1500.]):
        X2 = np.repeat([ant], len(V), 0)
        res = make("greenland_simple", attenuation_model="GL3", n_frequencies_integration=12).trace_batch(V, X2, frequency=ff)
        ora = oracle_mod.Oracle("greenland_simple", attenuation_model="GL3", n_freq=12, tight=True,
                                gl3_table=attenuation.gl3_parameters()).trace(V, X2, ff, None)
        assert np.array_equal(res["n_sol"], ora["n_sol"])
        assert_attenuation_parity(res["attenuation"], ora["attenuation"], rtol=1e-6, atol=1e-9)


def test_attenuation_length_models(make, oracle_mod):
    """get_attenuation_length(z, f, model) on the device vs the oracle's restatement of attenuation.py:145-262"""
    from nuradiomc_b200.utilities import attenuation
    rng = np.random.default_rng(0)
    z = rng.uniform(-3000, 0, 2000)
    f = rng.uniform(0.01, 2.5, 2000)
    for model in ("SP1", "GL1", "MB1", "GL2", "GL3"):
        if model == "MB1":
            z = np.maximum(z, -576.)   # the MB1 depth factor is only defined inside the 576 m thick shelf (attenuation.py:239-240)
        o = oracle_mod.Oracle("southpole_2015", attenuation_model=model,
                              gl3_table=attenuation.gl3_parameters() if model == "GL3" else None)
        ref = np.array([o.attenuation_length(a, b) for a, b in zip(z, f)])
        got = attenuation.get_attenuation_length(z, f, model)
        np.testing.assert_allclose(got, ref, rtol=1e-12, err_msg=model)
    assert attenuation.get_attenuation_length(1.0, 0.3, "SP1") == np.inf          # air (attenuation.py:256-257)
    assert attenuation.get_attenuation_length(-2900., 2.4, "GL1") == 1.0          # 1 m floor (:252-255)


# ---------------------------------------------------------------------------------------------------------------
# the reference's scalar API
# ---------------------------------------------------------------------------------------------------------------
def test_scalar_api_matches_batch_and_reference_semantics(make, oracle_mod):
    rt = make("southpole_2015", attenuation_model="SP1", n_frequencies_integration=25)
    ff = np.fft.rfftfreq(256, 0.5)
    X1 = cylinder(5, 40, 3000, -2000)
    x2 = np.array([10., 10., -190.])
    batch = rt.trace_batch(X1, x2[None, :], frequency=ff, max_detector_freq=0.8)
    ora = oracle_mod.Oracle("southpole_2015", attenuation_model="SP1", n_freq=25).trace(X1, x2, ff, 0.8)
    for i, x1 in enumerate(X1):
        rt.set_start_and_end_point(x1, x2)
        rt.find_solutions()
        n = rt.get_number_of_solutions()
        assert n == ora["n_sol"][i] == batch["n_sol"][i]
        assert rt.has_solution() == (n > 0)
        assert rt.get_results() == batch.solutions(i)
        for iS in range(n):
            assert rt.get_solution_type(iS) == ora["type"][i, iS]
            assert rt.get_results()[iS]["reflection"] == 0
            np.testing.assert_allclose(rt.get_launch_vector(iS), ora["launch"][i, iS], atol=1e-6)
            np.testing.assert_allclose(rt.get_receive_vector(iS), ora["receive"][i, iS], atol=1e-6)
            np.testing.assert_allclose(rt.get_path_length(iS), ora["path_length"][i, iS], rtol=1e-6)
            np.testing.assert_allclose(rt.get_travel_time(iS), ora["travel_time"][i, iS], rtol=1e-6)
            ra = rt.get_reflection_angle(iS)
            if ora["type"][i, iS] == 3:
                np.testing.assert_allclose(float(ra), ora["reflection_angle"][i, iS, 0], atol=1e-6)
            else:
                assert ra[()] is None
            att = rt.get_attenuation(iS, ff, 0.8)
            np.testing.assert_array_equal(att, batch["attenuation"][i, iS])
            assert att.shape == ff.shape and att[0] == 1.0 and (att[1:] > 0).all() and (att <= 1).all()
            out = rt.get_raytracing_output(iS)
            assert out["ray_tracing_solution_type"] == ora["type"][i, iS] and out["focusing_factor"] == 1
        with pytest.raises(IndexError):
            rt.get_launch_vector(n)
        with pytest.raises(IndexError):
            rt.get_attenuation(n, ff)


def test_scalar_api_with_bottom_reflections(make, oracle_mod):
    rt = make("mooresbay_simple", n_reflections=1)
    o = oracle_mod.Oracle("mooresbay_simple", n_reflections=1)
    for x1 in cylinder(9, 25, 800, -500):
        x2 = np.array([3., 3., -5.])
        rt.set_start_and_end_point(x1, x2)
        rt.find_solutions()
        ora = o.trace(x1[None, :], x2)
        assert rt.get_number_of_solutions() == ora["n_sol"][0]
        for iS, r in enumerate(rt.get_results()):
            assert (r["reflection"], r["reflection_case"]) == (ora["reflection"][0, iS], ora["reflection_case"][0, iS])
            ra = np.atleast_1d(rt.get_reflection_angle(iS))
            assert len(ra) == r["reflection"] + 1
            ref = ora["reflection_angle"][0, iS, :len(ra)]
            for a, b in zip(ra, ref):
                assert (a is None and np.isnan(b)) or abs(a - b) < 1e-6
    with pytest.raises(AttributeError):
        rt.set_start_and_end_point([0, 0, -580.], [0, 0, -5.])


def test_prepare_batch_serves_scalar_loop(make):
    """the simulation-loop hook: pre-trace all (shower, channel) pairs, then the scalar calls are lookups"""
    rt = make("greenland_simple")
    V, A = cylinder(3, 30, 2000, -1500), RNOG[[0, 12]]
    res = rt.prepare_batch(V, A, outer=True)
    launches_before = res.stats["n_launches"]
    assert launches_before >= 1
    for i, v in enumerate(V):
        for j, a in enumerate(A):
            rt.set_start_and_end_point(v, a)
            rt.find_solutions()
            assert rt._batch_index == i * len(A) + j
            assert rt.get_number_of_solutions() == res["n_sol"][i * len(A) + j]
            for iS in range(rt.get_number_of_solutions()):
                assert rt.get_travel_time(iS) == res["travel_time"][i * len(A) + j, iS]


# ---------------------------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------------------------
def test_edge_cases(make, oracle_mod):
    rt = make("southpole_2015")
    # empty batch
    res = rt.trace_batch(np.zeros((0, 3)), np.zeros((0, 3)))
    assert res["n_sol"].shape == (0,) and res["C0"].shape == (0, 2)
    # swapped points, equal depths, receiver at the surface, receiver in air, tiny horizontal offset, deep horizontal, NaN
    X1 = np.array([[0, 0, -50.], [100, 50, -300.], [0, 0, -100.], [0, 0, -100.], [5, 5, -20.], [0, 0, -1000.], [np.nan, 0, -10.]])
    X2 = np.array([[200, 0, -300.], [0, 0, -300.], [100, 0, 0.], [50, 0, 1.], [5.001, 5, -10.], [3000, 0, -1000.], [1, 1, -1.]])
    res = rt.trace_batch(X1, X2)
    assert list(res["n_sol"]) == [2, 2, 1, 0, 2, 2, 0]
    assert list(res["status"]) == [0, 0, 0, 1, 0, 0, 4]
    assert np.isnan(res["C0"][3]).all() and (res["solution_type"][3] == 0).all()
    ora = oracle_mod.Oracle("southpole_2015").trace(X1[[0, 1, 4, 5]], X2[[0, 1, 4, 5]])
    assert_parity({k: v[[0, 1, 4, 5]] for k, v in res.items()}, ora)
    # physical bound where the reference's closed form breaks down: path >= straight line (reference: 2998.04 m)
    assert (res["path_length"][5] >= 3000.0).all() and (res["path_length"][5] < 3001.0).any()
    # single pair / ragged tail (N not a multiple of the block size)
    for n in (1, 31, 129, 1000):
        V = cylinder(n, n, 3000, -2000)
        r1 = rt.trace_batch(V, np.array([[0, 0, -100.]]))
        r2 = rt.trace_batch(V[::-1].copy(), np.array([[0, 0, -100.]]))
        assert np.array_equal(r1["n_sol"], r2["n_sol"][::-1])
        np.testing.assert_array_equal(r1["C0"], r2["C0"][::-1])
    # points below the reflective layer
    rtm = make("mooresbay_simple", n_reflections=1)
    res = rtm.trace_batch(np.array([[0, 0, -600.], [0, 0, -100.]]), np.array([[10, 0, -5.], [10, 0, -5.]]))
    assert list(res["status"]) == [2, 0] and res["n_sol"][0] == 0 and res["n_sol"][1] > 0


def test_pair_mode_equals_outer_mode_and_pinned_buffers(make):
    rt = make("southpole_2015", attenuation_model="SP1", n_frequencies_integration=10)
    ff = np.fft.rfftfreq(128, 0.5)
    V, A = cylinder(8, 300, 4000, -2700), np.array([[0, 0, -150.], [1500, 0, -160.], [0, -1500, -145.]])
    r_outer = rt.trace_batch(V, A, outer=True, frequency=ff, attenuation="both", pinned=True)
    r_pairs = rt.trace_batch(np.repeat(V, 3, axis=0), np.tile(A, (300, 1)), frequency=ff, attenuation="both")
    for k in r_pairs:
        np.testing.assert_array_equal(r_outer[k], r_pairs[k], err_msg=k)
    assert r_outer.stats["h2d_bytes"] < r_pairs.stats["h2d_bytes"]


@pytest.mark.parametrize("ice,n_refl,model,zmin,rmax", [("southpole_2015", 0, "SP1", -2700, 6000), ("mooresbay_simple", 1, "MB1", -500, 1000)])
def test_compact_layout_equals_padded_layout(make, ice, n_refl, model, zmin, rmax):
    """per-solution (CSR) rows == the filled slots of the padded layout, bit for bit; several chunks, buffer reuse"""
    rt = make(ice, attenuation_model=model, n_reflections=n_refl, n_frequencies_integration=10)
    ff = np.fft.rfftfreq(64, 0.5)
    V, A = cylinder(21, 3001, rmax, zmin), np.array([[0, 0, -5.], [400, 0, -160.], [0, -1500, -145.]])
    pad = rt.trace_batch(V, A, outer=True, frequency=ff, attenuation="both")
    cmp_ = None
    for _ in range(2):   # second pass reuses the pinned capacity buffers
        cmp_ = rt.trace_batch(V, A, outer=True, frequency=ff, attenuation="both", compact=True, pinned=True, out=cmp_)
    N, S = pad["n_sol"].shape[0], pad["C0"].shape[1]
    np.testing.assert_array_equal(cmp_["n_sol"], pad["n_sol"])
    np.testing.assert_array_equal(cmp_["status"], pad["status"])
    np.testing.assert_array_equal(cmp_["sol_offset"], np.concatenate([[0], np.cumsum(pad["n_sol"])]))
    filled = np.arange(S)[None, :] < pad["n_sol"][:, None]
    assert cmp_["C0"].shape[0] == filled.sum() == cmp_.stats["n_solutions"]
    for k in pad:
        if k in ("n_sol", "status"):
            continue
        np.testing.assert_array_equal(cmp_[k], pad[k][filled], err_msg=k)
    assert cmp_.stats["d2h_bytes"] < pad.stats["d2h_bytes"]
    assert cmp_.solutions(7) == pad.solutions(7)
    with pytest.raises(RuntimeError, match="capacity"):
        rt.trace_batch(V, A, outer=True, compact=True, row_capacity=10)
    # device-resident compact layout (the binned solver assigns the rows itself), several chunks
    import torch
    dv = torch.tensor(np.ascontiguousarray(V.T), device="cuda:0")
    da = torch.tensor(np.ascontiguousarray(A.T), device="cuda:0")
    for chunk in (0, 1500):
        rt.set_chunk_pairs(chunk)
        dev = rt.trace_batch_device(dv, da, outer=True, frequency=ff, attenuation="both", compact=True)
        n_rows = dev.n_rows()
        assert n_rows == filled.sum()
        for k in cmp_:
            got = dev[k].cpu().numpy()
            np.testing.assert_array_equal(got if k in ("n_sol", "status", "sol_offset") else got[:n_rows], cmp_[k], err_msg=f"device {k}")
        only_dense = rt.trace_batch_device(dv, da, outer=True, frequency=ff, attenuation="dense", compact=True)
        np.testing.assert_array_equal(only_dense["attenuation"][:n_rows].cpu().numpy(), cmp_["attenuation"])
        with pytest.raises(RuntimeError, match="capacity"):
            rt.trace_batch_device(dv, da, outer=True, compact=True, row_capacity=10, sync_stats=True)
        host_chunked = rt.trace_batch(V, A, outer=True, frequency=ff, attenuation="both", compact=True)
        for k in cmp_:
            np.testing.assert_array_equal(host_chunked[k], cmp_[k], err_msg=f"host chunk {chunk} {k}")
    rt.set_chunk_pairs(0)
    empty = rt.trace_batch(np.zeros((0, 3)), np.zeros((0, 3)), compact=True)
    assert list(empty["sol_offset"]) == [0] and empty["C0"].shape[0] == 0


def test_device_resident_path_equals_host_path(make):
    import torch
    rt = make("southpole_2015", attenuation_model="SP1", n_frequencies_integration=25)
    ff = np.fft.rfftfreq(1022, 0.2)
    V, A = cylinder(12, 5000, 6000, -2700), np.array([[0, 0, -150.], [1500, 0, -160.]])
    host = rt.trace_batch(V, A, outer=True, frequency=ff, max_detector_freq=1.2, attenuation="both")
    dv = torch.tensor(np.ascontiguousarray(V.T), device="cuda:0")
    da = torch.tensor(np.ascontiguousarray(A.T), device="cuda:0")
    dev = rt.trace_batch_device(dv, da, outer=True, frequency=ff, max_detector_freq=1.2, attenuation="both", sync_stats=True)
    for k in host:
        np.testing.assert_array_equal(host[k], dev[k].cpu().numpy(), err_msg=k)
    assert dev.stats["n_solutions"] == int(host["n_sol"].sum())


def test_results_do_not_depend_on_chunking(make):
    """the device path (2^24-pair chunks) and the host pipeline (two streams, per-chunk scan for the compact layout) cut
    the batch at arbitrary places: force tiny chunks and compare bit for bit with the single-chunk result"""
    import torch
    ff = np.fft.rfftfreq(128, 0.5)
    V, A = cylinder(71, 1300, 6000, -2700), np.array([[0, 0, -150.], [1500, 0, -160.], [0, -1500, -145.]])
    rt = make("southpole_2015", attenuation_model="SP1", n_frequencies_integration=10)
    one = rt.trace_batch(V, A, outer=True, frequency=ff, max_detector_freq=0.6, attenuation="both")
    one_c = rt.trace_batch(V, A, outer=True, frequency=ff, max_detector_freq=0.6, attenuation="both", compact=True)
    dv = torch.tensor(np.ascontiguousarray(V.T), device="cuda:0")
    da = torch.tensor(np.ascontiguousarray(A.T), device="cuda:0")
    for chunk in (300, 999, 1024):
        rt.set_chunk_pairs(chunk)
        many = rt.trace_batch(V, A, outer=True, frequency=ff, max_detector_freq=0.6, attenuation="both")
        assert many.stats["n_chunks"] >= 3
        for k in one:
            np.testing.assert_array_equal(many[k], one[k], err_msg=f"host {chunk} {k}")
        many_c = rt.trace_batch(V, A, outer=True, frequency=ff, max_detector_freq=0.6, attenuation="both", compact=True)
        for k in one_c:
            np.testing.assert_array_equal(many_c[k], one_c[k], err_msg=f"compact {chunk} {k}")
        dev = rt.trace_batch_device(dv, da, outer=True, frequency=ff, max_detector_freq=0.6, attenuation="both", sync_stats=True)
        assert dev.stats["n_chunks"] >= 3
        for k in one:
            np.testing.assert_array_equal(dev[k].cpu().numpy(), one[k], err_msg=f"device {chunk} {k}")
    rt.set_chunk_pairs(0)
    # pair mode with bottom reflections (generic solver + generic attenuation kernel)
    rtm = make("mooresbay_simple", attenuation_model="MB1", n_reflections=1, n_frequencies_integration=8)
    Vm = cylinder(72, 700, 900, -550)
    Am = np.repeat([[3, 3, -5.]], len(Vm), 0)
    one = rtm.trace_batch(Vm, Am, frequency=ff)
    rtm.set_chunk_pairs(128)
    many = rtm.trace_batch(Vm, Am, frequency=ff)
    for k in one:
        np.testing.assert_array_equal(many[k], one[k], err_msg=f"mooresbay {k}")


def test_sp1_series_fallback_path(make, oracle_mod):
    """frequencies far below the band where the SP1 moment series converges (|ln f| large) push solutions to the generic
    quadrature kernel: same tolerance against the tight oracle"""
    ff = np.linspace(0, 0.02, 65)          # first bin 0.3 MHz: |ln f| = 8.1
    V = cylinder(73, 400, 4000, -2700)
    X2 = np.repeat([[0, 0, -150.]], len(V), 0)
    rt = make("southpole_2015", attenuation_model="SP1", n_frequencies_integration=20)
    res = rt.trace_batch(V, X2, frequency=ff)
    ora = oracle_mod.Oracle("southpole_2015", attenuation_model="SP1", n_freq=20, tight=True).trace(V, X2, ff, None)
    assert np.array_equal(res["n_sol"], ora["n_sol"])
    assert_attenuation_parity(res["attenuation"], ora["attenuation"])


# ---------------------------------------------------------------------------------------------------------------
# the step after the trace: propagation effects on spectra
# ---------------------------------------------------------------------------------------------------------------
class _Field:
    """minimal ElectricField stand-in (the members apply_propagation_effects uses)"""

    def __init__(self, spec, ff, sr):
        self.spec, self.ff, self.sr, self.params = np.array(spec), ff, sr, {}

    def get_sampling_rate(self): return self.sr
    def get_frequency_spectrum(self): return self.spec
    def get_frequencies(self): return self.ff
    def set_frequency_spectrum(self, spec, sr): self.spec = np.array(spec)
    def __setitem__(self, k, v): self.params[getattr(k, "name", str(k))] = v


@pytest.mark.parametrize("tag,ice,att,n_refl", [("sp", "southpole_2015", "SP1", 0), ("mb", "mooresbay_simple", "MB1", 1)])
def test_propagation_effects_kernel(make, oracle_mod, tag, ice, att, n_refl):
    """K_apply_effects against the numpy restatement (bit-level: 1e-12) and the reference's own outputs (its attenuation
    carries quad(epsrel=1e-2) noise: 1e-2 band, as the reference's T01 test); sparse on-the-fly interpolation == dense"""
    from oracle import propagation_effects as pe
    g = load_golden("propagation_effects")
    ff = g[f"{tag}_frequencies"]
    rt = make(ice, attenuation_model=att, n_reflections=n_refl, n_frequencies_integration=12)
    X1, X2 = g[f"{tag}_X1"], g[f"{tag}_X2"]
    res = rt.trace_batch(X1, X2, frequency=ff, max_detector_freq=float(ff.max()), attenuation="both", compact=True)
    assert np.array_equal(res["n_sol"], g[f"{tag}_n_sol"])
    S = g[f"{tag}_spec_in"].shape[1]
    filled = np.arange(S)[None, :] < res["n_sol"][:, None]
    spec_in, ref = g[f"{tag}_spec_in"][filled], g[f"{tag}_spec_out"][filled]
    out, r_t, r_p = rt.apply_propagation_effects_batch(spec_in, reflection_angle=res["reflection_angle"], reflection=res["reflection"],
                                                       attenuation=res["attenuation"], return_coefficients=True)
    n_ice, dn, z0, _ = oracle_mod.ICE_MODELS[ice]
    med = rt._medium
    for r in range(len(spec_in)):
        k = int(res["reflection"][r])
        exp, et, ep = pe.apply_propagation_effects(spec_in[r], res["attenuation"][r], res["reflection_angle"][r, :k + 1], k,
                                                   n_ice - dn * np.exp(-0.01 / z0), med.reflection_coefficient, med.reflection_phase_shift)
        np.testing.assert_allclose(out[r], exp, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose([r_t[r], r_p[r]], [et, ep], rtol=1e-12)
        np.testing.assert_allclose(out[r], ref[r], rtol=1e-2, atol=1e-3 * np.abs(ref[r]).max())
    # the coefficients the reference left on its field objects (the last surface reflection's, py:2993-2994)
    ref_t, ref_p = g[f"{tag}_r_theta"][filled], g[f"{tag}_r_phi"][filled]
    has = ~np.isnan(ref_t)
    assert has.sum() > 0
    np.testing.assert_allclose(r_t[has], ref_t[has], rtol=1e-9)
    np.testing.assert_allclose(r_p[has], ref_p[has], rtol=1e-9)
    if n_refl == 0:
        out2 = rt.apply_propagation_effects_batch(spec_in, reflection_angle=res["reflection_angle"], reflection=res["reflection"],
                                                  attenuation_sparse=res["attenuation_sparse"])
        np.testing.assert_allclose(out2, out, rtol=1e-13, atol=1e-300)
        # scalar API, as simulation.py uses it
        i = int(np.nonzero(res["n_sol"] == 2)[0][0])
        rt.set_start_and_end_point(X1[i], X2[i])
        rt.find_solutions()
        r0 = int(res["sol_offset"][i])
        for iS in range(2):
            f = _Field(spec_in[r0 + iS], ff, 2.0)
            rt.apply_propagation_effects(f, iS)
            np.testing.assert_allclose(f.spec, out[r0 + iS], rtol=1e-12, atol=1e-14)
    else:
        with pytest.raises(RuntimeError, match="dense"):
            rt.apply_propagation_effects_batch(spec_in, attenuation_sparse=res["attenuation_sparse"])


def test_gl1_thread_per_solution_kernel_equals_generic_kernel(make, oracle_mod, monkeypatch):
    """K_att_gl1 (thread per solution, graded sub-panels towards the deep end of the path, solutions with a visible frequency
    within 10 m of the pole handed to the generic kernel) against the generic warp-per-solution kernel (NRMC_GL1_GENERIC=1) and
    the tight oracle, on the cfg3 set-up and on wide random geometry: 1e-4 relative on factors above 1e-3, 2e-7 absolute below"""
    ff = np.fft.rfftfreq(1022, 0.2)
    V = cylinder(63, 1500, 4000, -2700)
    A = RNOG[[0, 8, 13, 21]]
    X1, X2 = np.repeat(V, len(A), 0), np.tile(A, (len(V), 1))
    rng = np.random.default_rng(64)
    n = 3000
    ze, zr = -np.exp(rng.uniform(np.log(0.5), np.log(2900.), n)), -np.exp(rng.uniform(np.log(0.5), np.log(2900.), n))
    rho, phi = np.exp(rng.uniform(np.log(0.1), np.log(9000.), n)), rng.uniform(0, 2 * np.pi, n)
    W1, W2 = np.stack([rho * np.cos(phi), rho * np.sin(phi), ze], 1), np.stack([np.zeros(n), np.zeros(n), zr], 1)
    for P1, P2, fmax, nf in ((X1, X2, 1.2, 25), (W1, W2, None, 20), (W1, W2, 0.8, 20)):
        fast = make("greenland_simple", attenuation_model="GL1", n_frequencies_integration=nf).trace_batch(
            P1, P2, frequency=ff, max_detector_freq=fmax, attenuation="both")
        monkeypatch.setenv("NRMC_GL1_GENERIC", "1")
        generic = make("greenland_simple", attenuation_model="GL1", n_frequencies_integration=nf).trace_batch(
            P1, P2, frequency=ff, max_detector_freq=fmax, attenuation="both")
        monkeypatch.delenv("NRMC_GL1_GENERIC")
        assert np.array_equal(fast["n_sol"], generic["n_sol"])
        for k in ("attenuation_sparse", "attenuation"):
            assert np.array_equal(np.isnan(fast[k]), np.isnan(generic[k]))
            assert_attenuation_parity(fast[k], generic[k], atol=2e-7)
        ora = oracle_mod.Oracle("greenland_simple", attenuation_model="GL1", n_freq=nf, tight=True).trace(P1[:2000], P2[:2000], ff, fmax)
        same = fast["n_sol"][:2000] == ora["n_sol"]
        assert_attenuation_parity(fast["attenuation"][:2000][same], ora["attenuation"][same], atol=2e-7)
        assert_attenuation_parity(fast["attenuation_sparse"][:2000][same], ora["attenuation_sparse"][same], atol=2e-7)


from test_kernel_math_cpu import harness  # noqa: E402,F401  (fixture: host build of nrmc_math.cuh, test-only)


@pytest.mark.parametrize("ice,n_refl,rmax,zmin", [("southpole_2015", 0, 6000, -2700), ("greenland_simple", 0, 4000, -2700),
                                                   ("mooresbay_simple", 1, 1000, -500)])
def test_device_maths_equals_host_build(make, harness, ice, n_refl, rmax, zmin):
    """The device evaluates 1/x, rsqrt and log of the solver loop with its own fast paths (hardware seed + Newton steps,
    nrmc_math.cuh); the host build of the same header uses libm.  Same algorithm, so the results must agree to rounding:
    identical counts and types, C0 to 1e-12, path length and travel time to 1e-11 relative (the root itself is only defined to
    |R - rho| <= 1e-10 m), vectors to 1e-11."""
    V = cylinder(61, 4000, rmax, zmin)
    A = np.array([[0, 0, -150.], [1500, -1500, -5.], [30, 40, -60.]])
    X1, X2 = np.repeat(V, len(A), 0), np.tile(A, (len(V), 1))
    res = make(ice, n_reflections=n_refl).trace_batch(X1, X2)
    h = harness(ice, n_refl, X1, X2)
    assert np.array_equal(res["n_sol"], h["n_sol"])
    assert np.array_equal(res["solution_type"], h["type"]) and np.array_equal(res["reflection"], h["reflection"])
    S = h["C0"].shape[1]
    filled = np.arange(S)[None, :] < h["n_sol"][:, None]
    assert filled.sum() > 5000
    np.testing.assert_allclose(res["C0"][filled], h["C0"][filled], rtol=1e-12)
    np.testing.assert_allclose(res["path_length"][filled], h["path_length"][filled], rtol=1e-11)
    np.testing.assert_allclose(res["travel_time"][filled], h["travel_time"][filled], rtol=1e-11)
    np.testing.assert_allclose(res["launch_vector"][filled], h["launch"][filled], atol=1e-11)
    np.testing.assert_allclose(res["receive_vector"][filled], h["receive"][filled], atol=1e-11)


@pytest.mark.parametrize("tag", ["sp", "sp_nolimit", "mb"])
def test_focusing_kernel(make, oracle_mod, tag):
    """K_focusing (get_focusing, analyticraytracing.py:2778-2888) through the C ABI: against the reference's own values
    (tests/golden/focusing.npz, 5e-3 = the reference's root-finding noise on its 1 cm difference quotient), against the oracle
    restatement (same quotient on the oracle's traces: 2e-3, its truncation error), padded == compact == device-resident,
    and the scalar API"""
    import torch
    from oracle.focusing import get_focusing
    g = load_golden("focusing")
    ice, n_refl, limit = str(g[f"{tag}_ice"]), int(g[f"{tag}_n_reflections"]), float(g[f"{tag}_limit"])
    X1, X2, ref = g[f"{tag}_X1"], g[f"{tag}_X2"], g[f"{tag}_focusing"]
    rt = make(ice, n_reflections=n_refl)
    res = rt.trace_batch(X1, X2)
    f = rt.focusing_batch(X1, X2, res, limit=limit)
    S = res["C0"].shape[1]
    filled = np.arange(S)[None, :] < res["n_sol"][:, None]
    assert f.shape == res["C0"].shape and np.isnan(f[~filled]).all() and np.isfinite(f[filled]).all()
    n_of = lambda z: rt._medium.get_index_of_refraction(np.array([0, 0, 1.])[None, :] * np.asarray(z)[:, None])
    assert (f[filled] <= (limit * np.sqrt(n_of(X1[:, 2]) / n_of(X2[:, 2])))[:, None].repeat(S, 1)[filled] * (1 + 1e-12)).all()
    ok = filled & (res["n_sol"] == g[f"{tag}_n_sol"])[:, None] & (g[f"{tag}_n_sol"] == g[f"{tag}_n_sol_displaced"])[:, None]
    assert ok.sum() >= 0.98 * np.isfinite(ref).sum()
    np.testing.assert_allclose(f[ok], ref[ok], rtol=5e-3)
    o = oracle_mod.Oracle(ice, n_reflections=n_refl)
    fo, comparable = get_focusing(o, X1, X2, limit=limit)
    sel = comparable & filled
    near_limit = np.isclose(np.maximum(fo, f), (limit * np.sqrt(n_of(X1[:, 2]) / n_of(X2[:, 2])))[:, None], rtol=1e-2)
    sel &= ~near_limit                       # either side may clip first within the quotient's truncation error
    assert sel.sum() > 0.8 * filled.sum()
    np.testing.assert_allclose(f[sel], fo[sel], rtol=2e-3)
    # compact rows and device-resident tensors give the same numbers
    resc = rt.trace_batch(X1, X2, compact=True)
    fc = rt.focusing_batch(X1, X2, resc, limit=limit)
    np.testing.assert_array_equal(fc, f[filled])
    v = torch.as_tensor(np.ascontiguousarray(X1.T)).cuda()
    a = torch.as_tensor(np.ascontiguousarray(X2.T)).cuda()
    resd = rt.trace_batch_device(v, a, compact=True)
    fd = rt.focusing_batch(v, a, resd, limit=limit)
    np.testing.assert_array_equal(fd[:resd.n_rows()].cpu().numpy(), fc)
    # scalar API (default limit 2) and the HDF5 output parameter (:2905-2935)
    i = int(np.nonzero(res["n_sol"] == res["n_sol"].max())[0][0])
    rt.set_start_and_end_point(X1[i], X2[i])
    rt.find_solutions()
    f2 = rt.focusing_batch(X1[i:i + 1], X2[i:i + 1], {k: res[k][i:i + 1] for k in res}, limit=2.0)[0]
    for iS in range(rt.get_number_of_solutions()):
        assert rt.get_focusing(iS) == f2[iS]
        assert rt.get_raytracing_output(iS)["focusing_factor"] == 1
    rt.set_config({"propagation": {"attenuate_ice": False, "focusing": True, "focusing_limit": 1.5, "birefringence": False}})
    f15 = rt.focusing_batch(X1[i:i + 1], X2[i:i + 1], {k: res[k][i:i + 1] for k in res})[0]
    assert rt.get_raytracing_output(0)["focusing_factor"] == f15[0]
    with pytest.raises(IndexError):
        rt.get_focusing(rt.get_number_of_solutions())
    # pre-traced event group with focusing on: the scalar loop's get_focusing / get_raytracing_output are look-ups
    rt2 = make(ice, n_reflections=n_refl, config={"propagation": {"attenuate_ice": False, "focusing": True, "focusing_limit": 1.5,
                                                                  "birefringence": False}})
    pre = rt2.prepare_batch(X1[i:i + 4], X2[i:i + 4])
    np.testing.assert_array_equal(pre["focusing_factor"][0], f15)
    rt2.set_start_and_end_point(X1[i], X2[i])
    rt2.find_solutions()
    assert 1.5 in rt2._foc_cache
    for iS in range(rt2.get_number_of_solutions()):
        assert rt2.get_raytracing_output(iS)["focusing_factor"] == f15[iS]


def test_propagation_effects_with_focusing(make, oracle_mod):
    """config['propagation']['focusing']: eTheta and ePhi scaled by get_focusing, eR untouched (analyticraytracing.py:3012-3015)"""
    from oracle import propagation_effects as pe
    g = load_golden("propagation_effects")
    ff, X1, X2 = g["sp_frequencies"], g["sp_X1"], g["sp_X2"]
    cfg = {"propagation": {"attenuate_ice": True, "focusing": True, "focusing_limit": 2, "birefringence": False}}
    rt = make("southpole_2015", attenuation_model="SP1", n_frequencies_integration=12, config=cfg)
    res = rt.trace_batch(X1, X2, frequency=ff, max_detector_freq=float(ff.max()), attenuation="dense", compact=True)
    foc = rt.focusing_batch(X1, X2, res)
    S = g["sp_spec_in"].shape[1]
    spec_in = g["sp_spec_in"][np.arange(S)[None, :] < res["n_sol"][:, None]]
    out = rt.apply_propagation_effects_batch(spec_in, reflection_angle=res["reflection_angle"], reflection=res["reflection"],
                                             attenuation=res["attenuation"], focusing=foc)
    n_ice, dn, z0, _ = oracle_mod.ICE_MODELS["southpole_2015"]
    for r in range(len(spec_in)):
        exp, _, _ = pe.apply_propagation_effects(spec_in[r], res["attenuation"][r], res["reflection_angle"][r, :1], 0,
                                                 n_ice - dn * np.exp(-0.01 / z0), focusing=foc[r])
        np.testing.assert_allclose(out[r], exp, rtol=1e-12, atol=1e-14)
    i = int(np.nonzero(res["n_sol"] == 2)[0][0])
    rt.set_start_and_end_point(X1[i], X2[i])
    rt.find_solutions()
    r0 = int(res["sol_offset"][i])
    for iS in range(2):
        fld = _Field(spec_in[r0 + iS], ff, 2.0)
        rt.apply_propagation_effects(fld, iS)
        np.testing.assert_allclose(fld.spec, out[r0 + iS], rtol=1e-12, atol=1e-14)


# ---------------------------------------------------------------------------------------------------------------
# the callers either side of the path
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ice,att,n_refl,rmax,zmin", [("southpole_2015", "SP1", 0, 3000, -2500), ("mooresbay_simple", "MB1", 1, 800, -500)])
def test_viewing_angle_cut(make, ice, att, n_refl, rmax, zmin):
    """simulation.py:175-208: viewing angle = angle(-shower axis, launch vector); solutions further than delta_C_cut from the
    Cherenkov cone keep their geometry but get no attenuation; everything else is unchanged by the cut"""
    rt = make(ice, attenuation_model=att, n_reflections=n_refl, n_frequencies_integration=8)
    ff = np.fft.rfftfreq(64, 0.5)
    V, A = cylinder(51, 400, rmax, zmin), np.array([[0, 0, -5.], [200, 0, -150.]])
    rng = np.random.default_rng(52)
    axes = rng.normal(size=(len(V), 3))
    cut = np.deg2rad(40.)
    plain = rt.trace_batch(V, A, outer=True, frequency=ff, attenuation="both")
    res = rt.trace_batch(V, A, outer=True, frequency=ff, attenuation="both", shower_axis=axes, delta_C_cut=cut)
    S = plain["C0"].shape[1]
    filled = np.arange(S)[None, :] < plain["n_sol"][:, None]
    ax = np.repeat(axes, len(A), axis=0)
    cosv = np.einsum("nk,nsk->ns", ax, plain["launch_vector"]) / np.linalg.norm(ax, axis=1)[:, None]
    va = np.arccos(np.clip(cosv, -1, 1))
    np.testing.assert_allclose(res["viewing_angle"][filled], va[filled], atol=1e-9)
    assert np.isnan(res["viewing_angle"][~filled]).all()
    n_vertex = np.repeat(rt._medium.get_index_of_refraction(V), len(A))
    keep = np.abs(va - np.arccos(1. / n_vertex)[:, None]) <= cut
    margin = np.abs(np.abs(va - np.arccos(1. / n_vertex)[:, None]) - cut) > 1e-9
    # the cut is a second instantiation of the solver kernel (K_roots<true>): same source, but ptxas may contract and order the
    # arithmetic differently, so floating-point outputs agree to rounding (1e-13), integers exactly
    for k in plain:
        if k.startswith("attenuation"):
            sel = filled & keep & margin
            np.testing.assert_allclose(res[k][sel], plain[k][sel], rtol=1e-12, atol=1e-300, err_msg=k)
            assert np.isnan(res[k][filled & ~keep & margin]).all(), k
        elif np.issubdtype(np.asarray(plain[k]).dtype, np.floating):
            np.testing.assert_allclose(res[k], plain[k], rtol=1e-13, atol=1e-13, equal_nan=True, err_msg=k)
        else:
            np.testing.assert_array_equal(res[k], plain[k], err_msg=k)
    assert 0.05 < keep[filled].mean() < 0.95
    from nuradiomc_b200 import simulation
    np.testing.assert_array_equal(simulation.cherenkov_mask(res, rt._medium, V, len(A), cut)[filled & margin], keep[filled & margin])


def test_simulation_hooks_and_hdf5_datasets(make):
    """pretrace of an event group serves the scalar loop of simulation.py:155-210 from the cache; the HDF5 ray-tracing
    datasets (output_writer_hdf5.py:267-294) round-trip through set_solution (analyticraytracing.py:2092-2116)"""
    from nuradiomc_b200 import simulation
    rt = make("mooresbay_simple", n_reflections=1)
    V, A = cylinder(61, 12, 800, -500), np.array([[3, 3, -5.], [-3, 0, -1.], [0, 3, -1.]])
    res = simulation.pretrace_event_group(rt, V, A)
    ds = simulation.raytracing_datasets(res, len(V), len(A))
    S = rt.get_number_of_raytracing_solutions()
    assert ds["travel_times"].shape == (12, 3, S) and ds["launch_vectors"].shape == (12, 3, S, 3)
    fresh = make("mooresbay_simple", n_reflections=1)
    for i in range(len(V)):
        for j in range(len(A)):
            rt.set_start_and_end_point(V[i], A[j])
            rt.find_solutions()
            assert rt._batch_index == i * len(A) + j          # served from the batch, no kernel launch
            n = rt.get_number_of_solutions()
            assert np.isnan(ds["ray_tracing_C0"][i, j, n:]).all() and not np.isnan(ds["ray_tracing_C0"][i, j, :n]).any()
            for iS in range(n):
                assert rt.get_travel_time(iS) == ds["travel_times"][i, j, iS]
                assert rt.get_path_length(iS) == ds["travel_distances"][i, j, iS]
                np.testing.assert_array_equal(rt.get_launch_vector(iS), ds["launch_vectors"][i, j, iS])
                assert rt.get_raytracing_output(iS)["ray_tracing_C0"] == ds["ray_tracing_C0"][i, j, iS]
            if n and (i + j) % 5 == 0:       # reload: a fresh propagator fed from the stored datasets
                fresh.set_start_and_end_point(V[i], A[j])
                fresh.set_solution(simulation.solutions_from_datasets(ds, i, j))
                assert fresh.get_number_of_solutions() == n
                for iS in range(n):
                    assert fresh.get_solution_type(iS) == rt.get_solution_type(iS)
                    np.testing.assert_allclose(fresh.get_receive_vector(iS), rt.get_receive_vector(iS), atol=1e-12)
                    assert fresh.get_travel_time(iS) == pytest.approx(rt.get_travel_time(iS), rel=1e-12)
    fresh.set_start_and_end_point(V[0], A[0])
    fresh.set_solution({"ray_tracing_C0": np.array([7.7]), "ray_tracing_C1": np.array([0.]), "ray_tracing_solution_type": np.array([1])})
    with pytest.raises(AttributeError):
        fresh.get_launch_vector(0)


def test_lookup_table_vs_oracle(make, oracle_mod):
    """create_lookup_table.py:62-107 as one batched pass: travel time per solution type on an (r, z) grid"""
    from nuradiomc_b200 import lookup_table
    tab = lookup_table.create_lookup_table(100., r_min=10, r_max=2010, z_min=2000, z_max=50, d_r=100, d_z=130, ice_model="greenland_simple")
    hdr, t = tab["header"], tab["antenna_100.0"]
    x_pos, z_pos = np.arange(10, 2010, 100.), np.arange(-2000, -50, 130.)
    assert t["direct"].shape == (len(x_pos), len(z_pos)) and hdr["d_x"] == 100 and hdr["z_min"] == -2000
    V = np.array([[x, 0, z] for x in x_pos for z in z_pos])
    ora = oracle_mod.Oracle("greenland_simple").trace(V, np.repeat([[0, 0, -100.]], len(V), 0))
    for typ, name in ((1, "direct"), (2, "refracted"), (3, "reflected")):
        exp = np.zeros(len(V))
        for s in range(2):
            m = (s < ora["n_sol"]) & (ora["type"][:, s] == typ)
            exp[m] = ora["travel_time"][m, s]
        np.testing.assert_allclose(t[name].ravel(), exp, rtol=1e-6)
        assert (exp > 0).any()


# ---------------------------------------------------------------------------------------------------------------
# BASELINE sizes: size-independent properties
# ---------------------------------------------------------------------------------------------------------------
def test_full_size_properties_cfg2(make):
    """cfg2: 1e6 vertices x 4 antennas, southpole_2015, no attenuation -- invariants instead of the (too slow) oracle"""
    import torch
    rt = make("southpole_2015")
    V = cylinder(2, 1_000_000, 4000, -2700)
    A = np.array([[10, 10, -190.], [10, -10, -190.], [-10, -10, -190.], [-10, 10, -190.]])
    dv = torch.tensor(np.ascontiguousarray(V.T), device="cuda:0")
    da = torch.tensor(np.ascontiguousarray(A.T), device="cuda:0")
    res = rt.trace_batch_device(dv, da, outer=True)
    torch.cuda.synchronize()
    n_sol = res["n_sol"].cpu().numpy()
    assert set(np.unique(n_sol)) <= {0, 2}                     # 0 or 2 solutions per pair (unimodal range curve)
    frac2 = (n_sol == 2).mean()
    assert 0.6 < frac2 < 0.95                                   # ~81 % of cfg2 pairs have two solutions
    C0 = res["C0"].cpu().numpy()
    typ = res["solution_type"].cpu().numpy()
    has = n_sol == 2
    assert (C0[has, 0] < C0[has, 1]).all() and (C0[has] > 1 / 1.78).all()      # sorted by C0 (py:1547), C0 > 1/n_ice
    assert np.isnan(C0[~has]).all() and (typ[~has] == 0).all() and np.isin(typ[has], (1, 2, 3)).all()
    # path >= straight line; n_min * path <= c * t <= n_ice * path
    X1, X2 = np.repeat(V, 4, axis=0), np.tile(A, (len(V), 1))
    dist = np.linalg.norm(X1 - X2, axis=1)
    path, time = res["path_length"].cpu().numpy(), res["travel_time"].cpu().numpy()
    assert (path[has] >= dist[has, None] * (1 - 1e-12)).all()
    ct = time[has] * 0.299792458
    assert (ct <= 1.78 * path[has] * (1 + 1e-12)).all() and (ct >= (1.78 - 0.423) * path[has]).all()
    # launch / receive vectors are unit vectors; direct rays arrive from below, reflected / refracted from above
    lv, rv = res["launch_vector"].cpu().numpy(), res["receive_vector"].cpu().numpy()
    np.testing.assert_allclose(np.linalg.norm(lv[has], axis=-1), 1.0, atol=1e-12)
    np.testing.assert_allclose(np.linalg.norm(rv[has], axis=-1), 1.0, atol=1e-12)
    # determinism: a second pass is bit-identical
    res2 = rt.trace_batch_device(dv, da, outer=True)
    torch.cuda.synchronize()
    assert torch.equal(res["C0"].nan_to_num(), res2["C0"].nan_to_num()) and torch.equal(res["n_sol"], res2["n_sol"])
    # reciprocity on a slice: exchanging start and end point exchanges the vectors, keeps C0 / path / time
    sl = slice(0, 200_000)
    fwd = rt.trace_batch(X1[sl], X2[sl])
    bwd = rt.trace_batch(X2[sl], X1[sl])
    assert np.array_equal(fwd["n_sol"], bwd["n_sol"])
    np.testing.assert_allclose(fwd["C0"], bwd["C0"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(fwd["travel_time"], bwd["travel_time"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(fwd["launch_vector"], bwd["receive_vector"], atol=1e-12, equal_nan=True)


def test_full_size_properties_attenuation_cfg5_slice(make, oracle_mod):
    """cfg5 set-up (southpole_2015, SP1, 37 integration frequencies, sparse output) on 1e5 vertices x 20 channels:
    invariants at size + oracle parity on a random subsample"""
    import torch
    rt = make("southpole_2015", attenuation_model="SP1", n_frequencies_integration=25)
    ff = np.fft.rfftfreq(1022, 0.2)
    V = cylinder(5, 100_000, 6000, -2700)
    A = np.array([[x, y, z] for x in (-1500., 1500.) for y in (-3000., 0.) for z in (-145., -150., -155., -160., -100.)])
    dv = torch.tensor(np.ascontiguousarray(V.T), device="cuda:0")
    da = torch.tensor(np.ascontiguousarray(A.T), device="cuda:0")
    res = rt.trace_batch_device(dv, da, outer=True, frequency=ff, max_detector_freq=1.2, attenuation="sparse")
    torch.cuda.synchronize()
    att = res["attenuation_sparse"].cpu().numpy()
    n_sol = res["n_sol"].cpu().numpy()
    assert att.shape == (2_000_000, 2, 37)
    has = n_sol == 2
    assert np.isnan(att[~has]).all()
    a = att[has]
    assert (a > 0).all() and (a <= 1).all()
    assert (np.diff(a[..., :25], axis=-1) <= 1e-15).all()       # SP1: attenuation grows with frequency
    path = res["path_length"].cpu().numpy()[has]
    # exp(-path / L_min) <= factor: SP1 attenuation lengths stay above 200 m below 1.2 GHz
    assert (a[..., :25] >= np.exp(-path / 200.)[..., None]).all()
    idx = np.random.default_rng(1).choice(len(n_sol), 300, replace=False)
    X1, X2 = V[idx // len(A)], A[idx % len(A)]
    ora = oracle_mod.Oracle("southpole_2015", attenuation_model="SP1", n_freq=25).trace(X1, X2, ff, 1.2, dense=False)
    sub = {k: v.cpu().numpy()[idx] for k, v in res.items()}
    assert_parity(sub, ora)
    assert_attenuation_parity(sub["attenuation_sparse"], ora["attenuation_sparse"])


def test_full_size_properties_cfg3_gl1_dense(make, oracle_mod):
    """cfg3 set-up (greenland_simple, 24-channel RNO-G station, GL1, 512-bin dense output) on 5e4 vertices = 1.2e6 pairs:
    invariants at size + oracle parity (tight quadrature) on a random subsample"""
    import torch
    rt = make("greenland_simple", attenuation_model="GL1", n_frequencies_integration=25)
    ff = np.fft.rfftfreq(1022, 0.2)
    V = cylinder(3, 50_000, 4000, -2700)
    dv = torch.tensor(np.ascontiguousarray(V.T), device="cuda:0")
    da = torch.tensor(np.ascontiguousarray(RNOG.T), device="cuda:0")
    res = rt.trace_batch_device(dv, da, outer=True, frequency=ff, max_detector_freq=1.2, attenuation="both")
    torch.cuda.synchronize()
    n_sol = res["n_sol"].cpu().numpy()
    assert set(np.unique(n_sol)) <= {0, 2}
    has = torch.tensor(n_sol == 2, device="cuda:0")
    dense = res["attenuation"]
    assert dense.shape == (1_200_000, 2, 512)
    assert bool(torch.isnan(dense[~has]).all())
    d = dense[has]
    assert bool((d[..., 0] == 1).all())                                  # the 0 Hz bin keeps factor 1 (py:1077)
    assert bool(((d >= 0) & (d <= 1)).all())                           # above ~2 GHz the GL1 length hits its 1 m floor: exp(-km) = 0.0
    assert bool((d[..., 1:245].diff(dim=-1) <= 1e-15).all())              # GL1: attenuation grows with frequency (detector band)
    idx = np.random.default_rng(3).choice(len(n_sol), 200, replace=False)
    X1, X2 = V[idx // len(RNOG)], RNOG[idx % len(RNOG)]
    ora = oracle_mod.Oracle("greenland_simple", attenuation_model="GL1", n_freq=25, tight=True).trace(X1, X2, ff, 1.2)
    sub = {k: v[torch.tensor(idx, device="cuda:0")].cpu().numpy() for k, v in res.items()}
    assert_parity(sub, ora)
    assert_attenuation_parity(sub["attenuation"], ora["attenuation"])


def test_full_size_properties_cfg4_bottom_reflections(make, oracle_mod):
    """cfg4 set-up (mooresbay_simple, n_reflections = 1, 8 channels) on 1e5 vertices = 8e5 pairs, up to 6 solutions"""
    import torch
    rt = make("mooresbay_simple", n_reflections=1)
    V = cylinder(4, 100_000, 1000, -500)
    A = np.array([[-3, 0, -1.], [0, 3, -1.], [3, 0, -1.], [0, -3, -1.], [3, 3, -5.], [3, -3, -5.], [-3, -3, -5.], [-3, 3, -5.]])
    dv = torch.tensor(np.ascontiguousarray(V.T), device="cuda:0")
    da = torch.tensor(np.ascontiguousarray(A.T), device="cuda:0")
    res = {k: v.cpu().numpy() for k, v in rt.trace_batch_device(dv, da, outer=True).items()}
    n_sol, S = res["n_sol"], 6
    assert set(np.unique(n_sol)) <= {0, 2, 4, 6} and (n_sol == 6).mean() > 0.1 and (n_sol >= 4).mean() > 0.9
    filled = np.arange(S)[None, :] < n_sol[:, None]
    refl, case, C0 = res["reflection"], res["reflection_case"], res["C0"]
    assert np.isnan(C0[~filled]).all() and not np.isnan(C0[filled]).any()
    # result order of the reference: mode (reflection, case) first, then C0 ascending (py:2122-2125, :1547)
    key = refl.astype(np.int64) * 4 + case
    key[~filled] = 99
    assert (np.diff(key, axis=1) >= 0).all()
    same_mode = filled[:, 1:] & filled[:, :-1] & (key[:, 1:] == key[:, :-1])
    assert (C0[:, 1:][same_mode] > C0[:, :-1][same_mode]).all()
    # a bottom-reflected path is longer than the straight line to the mirror image of the receiver below the reflector
    X1, X2 = np.repeat(V, len(A), axis=0), np.tile(A, (len(V), 1))
    rho = np.hypot(X1[:, 0] - X2[:, 0], X1[:, 1] - X2[:, 1])
    image = np.sqrt(rho ** 2 + ((X1[:, 2] + 576.) + (X2[:, 2] + 576.)) ** 2)
    bounced = filled & (refl == 1)
    assert (res["path_length"][bounced] >= np.broadcast_to(image[:, None], bounced.shape)[bounced] * (1 - 1e-12)).all()
    idx = np.random.default_rng(4).choice(len(n_sol), 300, replace=False)
    ora = oracle_mod.Oracle("mooresbay_simple", n_reflections=1).trace(X1[idx], X2[idx])
    assert_parity({k: v[idx] for k, v in res.items()}, ora)


def test_ray_tracing_2D_facade(make, oracle_mod):
    """the 2-D interface of the bulk consumers (create_lookup_table.py:92-100): find_solutions / get_travel_time on [y, z] points"""
    from nuradiomc_b200.SignalProp.analyticraytracing import ray_tracing_2D
    from nuradiomc_b200.utilities import medium
    r2 = ray_tracing_2D(medium.get_ice_model("greenland_simple"))
    x1, x2 = [-500., -1000.], [0., -100.]
    sol = r2.find_solutions(x1, x2)
    ora = oracle_mod.Oracle("greenland_simple").trace(np.array([[-500., 0, -1000.]]), np.array([[0, 0, -100.]]))
    assert [s['type'] for s in sol] == list(ora["type"][0, :2]) and len(sol) == 2
    np.testing.assert_allclose([s['C0'] for s in sol], ora["C0"][0, :2], rtol=1e-6)
    # values of the reference's own 2-D class for this pair (run in the build container): 6108.0149791931 / 7037.4611398640 ns
    np.testing.assert_allclose([r2.get_travel_time(x1, x2, s['C0']) for s in sol], [6108.0149791931035, 7037.461139863954], rtol=1e-9)
    np.testing.assert_allclose(r2.get_path_length(x1, x2, sol[0]['C0']), ora["path_length"][0, 0], rtol=1e-6)
    assert r2.find_solutions([500., -1000.], [0., -100.]) == []          # receiver to the left: none, as the reference
    assert r2.get_travel_time(x1, x2, 7.7) is None
    rb = ray_tracing_2D(medium.get_ice_model("mooresbay_simple"), n_reflections=1)
    got = rb.find_solutions([0., -300.], [200., -5.], reflection=1, reflection_case=2)
    om = oracle_mod.Oracle("mooresbay_simple", n_reflections=1).trace(np.array([[0., 0, -300.]]), np.array([[200., 0, -5.]]))
    exp = [c for c, k, cs in zip(om["C0"][0], om["reflection"][0], om["reflection_case"][0]) if k == 1 and cs == 2]
    np.testing.assert_allclose([s['C0'] for s in got], exp, rtol=1e-6)
    batch = r2.find_solutions_batch(np.array([[-500., -1000.], [-50., -2000.]]), np.array([[0., -100.], [0., -100.]]))
    assert batch["mode"].shape == (2, 2) and batch["mode"][0].all()



def test_wide_geometry_stress_and_far_field_hump(make, oracle_mod):
    """log-uniform depths and distances (1 cm - 12 km) against the oracle, plus the far-field pairs whose range curve has a sharp
    hump next to the beta = n(z2) junction (regression: an early-exit heuristic of the maximum search once lost them)"""
    from test_kernel_math_cpu import FAR_FIELD_HUMP
    rng = np.random.default_rng(123)
    N = 60000
    for ice in ("greenland_simple", "mooresbay_simple"):
        rt = make(ice)
        zr = -np.exp(rng.uniform(np.log(0.5), np.log(3000.), N))
        ze = -np.exp(rng.uniform(np.log(0.5), np.log(3100.), N))
        rho = np.exp(rng.uniform(np.log(0.01), np.log(12000.), N))
        phi = rng.uniform(0, 2 * np.pi, N)
        X1 = np.concatenate([np.stack([rho * np.cos(phi), rho * np.sin(phi), ze], 1), np.array(FAR_FIELD_HUMP[ice][0], float)])
        X2 = np.concatenate([np.stack([np.zeros(N), np.zeros(N), zr], 1), np.array(FAR_FIELD_HUMP[ice][1], float)])
        res = rt.trace_batch(X1, X2)
        ora = oracle_mod.Oracle(ice).trace(X1, X2, n_threads=16)
        assert (res["n_sol"][N:] == 2).all()
        bad = np.nonzero(res["n_sol"] != ora["n_sol"])[0]
        assert len(bad) <= 2 and (res["n_sol"][bad] > ora["n_sol"][bad]).all()     # only roots the oracle's scan window misses
        assert_parity(res, ora, exact_count=False)


# ---------------------------------------------------------------------------------------------------------------
# regressions of the round-1 review
# ---------------------------------------------------------------------------------------------------------------
def test_get_attenuation_after_set_solution_subset(make):
    """set_solution with the first solution masked out (what the viewing-angle cut leaves in the HDF5 datasets): get_attenuation /
    the vectors of solution 0 must be those of the solution that was injected, not of the trace's slot 0"""
    ff = np.fft.rfftfreq(256, 0.5)
    rt = make("southpole_2015", attenuation_model="SP1", n_frequencies_integration=20)
    x1, x2 = np.array([400., 300., -900.]), np.array([10., 10., -190.])
    full = rt.trace_batch(x1[None], x2[None], frequency=ff, max_detector_freq=0.8)
    assert full["n_sol"][0] == 2 and not np.allclose(full["attenuation"][0, 0], full["attenuation"][0, 1])
    fresh = make("southpole_2015", attenuation_model="SP1", n_frequencies_integration=20)
    fresh.set_start_and_end_point(x1, x2)
    fresh.set_solution({"ray_tracing_C0": np.array([np.nan, full["C0"][0, 1]]), "ray_tracing_C1": np.array([np.nan, full["C1"][0, 1]]),
                        "ray_tracing_solution_type": np.array([0, full["solution_type"][0, 1]]),
                        "ray_tracing_reflection": np.array([0, 0]), "ray_tracing_reflection_case": np.array([0, 1])})
    assert fresh.get_number_of_solutions() == 1
    np.testing.assert_array_equal(fresh.get_attenuation(0, ff, 0.8), full["attenuation"][0, 1])
    np.testing.assert_array_equal(fresh.get_launch_vector(0), full["launch_vector"][0, 1])
    assert fresh.get_travel_time(0) == full["travel_time"][0, 1]
    # both injected, in reversed order
    fresh.set_start_and_end_point(x1, x2)
    fresh.set_solution({"ray_tracing_C0": full["C0"][0, ::-1], "ray_tracing_C1": full["C1"][0, ::-1],
                        "ray_tracing_solution_type": full["solution_type"][0, ::-1]})
    np.testing.assert_array_equal(fresh.get_attenuation(0, ff, 0.8), full["attenuation"][0, 1])
    np.testing.assert_array_equal(fresh.get_attenuation(1, ff, 0.8), full["attenuation"][0, 0])


def test_single_frequency_attenuation(make, oracle_mod):
    """one positive frequency: np.interp on one point returns the value itself (no read of a right neighbour)"""
    V = cylinder(77, 200, 3000, -2500)
    X2 = np.repeat([[0., 0., -150.]], len(V), 0)
    for model, ice, n_refl in (("SP1", "southpole_2015", 0), ("GL1", "greenland_simple", 0), ("MB1", "mooresbay_simple", 1)):
        if n_refl:
            V2, X22 = cylinder(78, 100, 800, -500), np.repeat([[3., 3., -5.]], 100, 0)
        else:
            V2, X22 = V, X2
        for ff in (np.array([0.3]), np.array([0., 0.3])):
            rt = make(ice, attenuation_model=model, n_reflections=n_refl, n_frequencies_integration=25)
            res = rt.trace_batch(V2, X22, frequency=ff, attenuation="both")
            ora = oracle_mod.Oracle(ice, attenuation_model=model, n_reflections=n_refl, n_freq=25, tight=True).trace(V2, X22, ff, None)
            assert np.array_equal(res["n_sol"], ora["n_sol"])
            assert res["attenuation_sparse"].shape[-1] == 1
            assert_attenuation_parity(res["attenuation"], ora["attenuation"])
            filled = np.arange(res["C0"].shape[1])[None, :] < res["n_sol"][:, None]
            assert np.isfinite(res["attenuation"][filled]).all() and np.isnan(res["attenuation"][~filled]).all()
            np.testing.assert_array_equal(res["attenuation"][..., -1], res["attenuation_sparse"][..., 0])
            # and through the effects kernel with the sparse factors (single-segment paths)
            if n_refl == 0:
                rows = res["attenuation_sparse"][filled]
                spec = np.ones((len(rows), 3, len(ff)), complex)
                out = rt.apply_propagation_effects_batch(spec, attenuation_sparse=rows)
                np.testing.assert_allclose(out[:, 0, -1].real, rows[:, 0], rtol=1e-15)


def test_prepare_batch_serves_attenuation_and_rejects_compact(make):
    ff = np.fft.rfftfreq(128, 0.5)
    rt = make("southpole_2015", attenuation_model="SP1", n_frequencies_integration=15)
    V, A = cylinder(91, 20, 2500, -2000), np.array([[0, 0, -150.], [30, 0, -100.]])
    with pytest.raises(ValueError):
        rt.prepare_batch(V, A, outer=True, compact=True)
    res = rt.prepare_batch(V, A, outer=True, frequency=ff, max_detector_freq=0.7)
    calls = []
    orig = rt.trace_batch
    rt.trace_batch = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    for i, v in enumerate(V):
        for j, a in enumerate(A):
            rt.set_start_and_end_point(v, a)
            rt.find_solutions()
            for iS in range(rt.get_number_of_solutions()):
                np.testing.assert_array_equal(rt.get_attenuation(iS, ff, 0.7), res["attenuation"][i * len(A) + j, iS])
    assert not calls, "the scalar loop after prepare_batch(frequency=...) must not trace again"
    rt.set_start_and_end_point(V[0], A[0])
    rt.find_solutions()
    if rt.get_number_of_solutions():
        rt.get_attenuation(0, ff, 0.5)            # another max_detector_freq: not in the batch, one trace
        assert len(calls) == 1


def test_device_call_followed_by_host_call_on_one_handle(make):
    """a device-resident call returns before its kernels ran; a host-memory call (own streams) right behind it, a second device call
    on another stream and a change of the frequency set must not disturb it (own scratch lane, event ordering)"""
    import torch
    ff, ff2 = np.fft.rfftfreq(256, 0.5), np.fft.rfftfreq(64, 1.0)
    rt = make("southpole_2015", attenuation_model="SP1", n_frequencies_integration=20)
    V, A = cylinder(92, 30000, 6000, -2700), np.array([[0, 0, -150.], [1500, 0, -160.]])
    ref = rt.trace_batch(V, A, outer=True, frequency=ff, max_detector_freq=0.8, attenuation="sparse")
    dv, da = torch.tensor(np.ascontiguousarray(V.T), device="cuda"), torch.tensor(np.ascontiguousarray(A.T), device="cuda")
    side = torch.cuda.Stream()
    for _ in range(3):
        d1 = rt.trace_batch_device(dv, da, outer=True, frequency=ff, max_detector_freq=0.8, attenuation="sparse")
        with torch.cuda.stream(side):
            d2 = rt.trace_batch_device(dv, da, outer=True, frequency=ff, max_detector_freq=0.8, attenuation="sparse")
        h = rt.trace_batch(V[:1100], A, outer=True, frequency=ff, max_detector_freq=0.8, attenuation="sparse")
        small = rt.trace_batch(V[:50], A, outer=True, frequency=ff2, attenuation="sparse")     # rewrites the frequency tables
        torch.cuda.synchronize()
        for d in (d1, d2):
            for k in ("n_sol", "C0", "travel_time", "attenuation_sparse"):
                np.testing.assert_array_equal(d[k].cpu().numpy(), ref[k], err_msg=k)
        for k in ("n_sol", "C0", "attenuation_sparse"):
            np.testing.assert_array_equal(h[k], ref[k][:2200], err_msg=k)
        assert small["attenuation_sparse"].shape[-1] == len(small.frequencies_sparse)


@pytest.mark.parametrize("tag", ["sp", "mb"])
def test_simulation_datasets_vs_reference(make, tag):
    """N3 against the REFERENCE: tests/golden/simulation_datasets.npz holds what the reference's own scalar loop
    (simulation.py:155-210: trace, viewing-angle cut, path length / travel time, get_raytracing_output) and its HDF5 writer
    (output_writer_hdf5.py:267-294) produce for an event group; one batched pre-trace + `raytracing_datasets` must reproduce
    every dataset, and the unchanged scalar loop over the pre-traced propagator must see the same numbers without tracing."""
    from nuradiomc_b200 import simulation
    g = load_golden("simulation_datasets")
    V, A, axes, cut = g[f"{tag}_vertices"], g[f"{tag}_channels"], g[f"{tag}_shower_axes"], float(g[f"{tag}_delta_C_cut"])
    rt = make(str(g[f"{tag}_ice"]), n_reflections=int(g[f"{tag}_n_reflections"]))
    res = simulation.pretrace_event_group(rt, V, A, shower_axes=axes, delta_C_cut=cut)
    np.testing.assert_array_equal(res["n_sol"].reshape(len(V), len(A)), g[f"{tag}_n_sol"])
    keep = simulation.cherenkov_mask(res, rt._medium, V, len(A), cut)
    ds = simulation.raytracing_datasets(res, len(V), len(A), keep=keep)
    ref_nan = np.isnan(g[f"{tag}_travel_times"])
    assert 0 < (~ref_nan).sum() < ref_nan.size
    for k in ("travel_times", "travel_distances", "ray_tracing_C0", "ray_tracing_C1"):
        assert np.array_equal(np.isnan(ds[k]), ref_nan), k
        np.testing.assert_allclose(ds[k][~ref_nan], g[f"{tag}_{k}"][~ref_nan], rtol=1e-6, err_msg=k)
    for k in ("ray_tracing_reflection", "ray_tracing_reflection_case", "ray_tracing_solution_type", "focusing_factor"):
        assert np.array_equal(np.isnan(ds[k]), ref_nan), k
        np.testing.assert_array_equal(ds[k][~ref_nan], g[f"{tag}_{k}"][~ref_nan], err_msg=k)
    for k in ("launch_vectors", "receive_vectors"):
        assert np.array_equal(np.isnan(ds[k][..., 0]), ref_nan), k
        np.testing.assert_allclose(ds[k][~ref_nan], g[f"{tag}_{k}"][~ref_nan], atol=1e-6, err_msg=k)
    np.testing.assert_allclose(res["viewing_angle"].reshape(ref_nan.shape)[~ref_nan], g[f"{tag}_viewing_angles"][~ref_nan], atol=1e-6)
    # the station group of the output file: one dataset per key, as the reference's writer lays them out
    group = {}
    simulation.write_station_group(group, ds)
    assert set(group) == set(ds) and group["launch_vectors"].shape == (len(V), len(A), ref_nan.shape[2], 3)
    simulation.write_station_group(group, ds)           # overwriting existing datasets
    # the reference's loop, unchanged, over the pre-traced propagator: no further trace
    calls = []
    orig = rt.trace_batch
    rt.trace_batch = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    for iSh in range(len(V)):
        n_index = rt._medium.get_index_of_refraction(V[iSh])
        for iCh in range(len(A)):
            rt.set_start_and_end_point(V[iSh], A[iCh])
            rt.find_solutions()
            if not rt.has_solution():
                assert g[f"{tag}_n_sol"][iSh, iCh] == 0
                continue
            for iS in range(rt.get_number_of_solutions()):
                lv = rt.get_launch_vector(iS)
                dC = np.arccos(np.clip(np.dot(-axes[iSh], lv), -1, 1)) - np.arccos(1. / n_index)
                if abs(dC) > cut:
                    assert ref_nan[iSh, iCh, iS]
                    continue
                assert rt.get_travel_time(iS) == ds["travel_times"][iSh, iCh, iS]
                assert rt.get_raytracing_output(iS)["ray_tracing_C0"] == ds["ray_tracing_C0"][iSh, iCh, iS]
    assert not calls


# ---------------------------------------------------------------------------------------------------------------
# the fused small-batch launch (K_small: the scalar API is a batch of one pair)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ice,model,n_refl,rmax,zmin,ant", [
    ("southpole_2015", "SP1", 0, 4000, -2700, [[10, 10, -190.], [0, 0, -2.]]),
    ("greenland_simple", "GL1", 0, 4000, -2700, [[0, 20, -97.], [1.5, 11, -2.]]),
    ("mooresbay_simple", "MB1", 1, 1000, -500, [[3, 3, -5.], [-3, 0, -1.]]),
    ("greenland_simple", "GL3", 0, 3000, -2700, [[0, 20, -97.]]),
    ("mooresbay_simple", None, 2, 1000, -570, [[0, 0, -5.]]),
])
def test_small_batch_path_equals_binned_pipeline(make, oracle_mod, monkeypatch, ice, model, n_refl, rmax, zmin, ant):
    """one cooperative launch (thread-per-pair solver + generic attenuation) against the production pipeline on the same pairs:
    identical solutions (same scalar maths), attenuation within 2e-5 (different quadrature layouts, both inside 1e-4 of the oracle);
    host and device entry points; viewing-angle cut"""
    import torch
    ff = np.fft.rfftfreq(256, 0.5)
    V, A = cylinder(300 + n_refl, 600, rmax, zmin), np.array(ant, float)
    axes = np.random.default_rng(5).normal(size=(len(V), 3))
    kw = dict(outer=True, shower_axis=axes, delta_C_cut=0.7)
    if model:
        kw.update(frequency=ff, max_detector_freq=0.8, attenuation="both")
    rt = make(ice, attenuation_model=model, n_reflections=n_refl, n_frequencies_integration=15)
    big = rt.trace_batch(V, A, **kw)
    assert big.stats["n_launches"] >= 3
    monkeypatch.delenv("NRMC_NO_SMALL_PATH")
    small = rt.trace_batch(V, A, **kw)
    assert small.stats["n_launches"] == 1 and small.stats["n_solutions"] == big.stats["n_solutions"] > 300
    for k in ("n_sol", "status", "solution_type", "reflection", "reflection_case"):
        np.testing.assert_array_equal(small[k], big[k], err_msg=k)
    for k in ("C0", "C1", "path_length", "travel_time", "launch_vector", "receive_vector", "reflection_angle", "viewing_angle"):
        np.testing.assert_allclose(small[k], big[k], rtol=1e-12, atol=1e-12, equal_nan=True, err_msg=k)
    if model:
        for k in ("attenuation", "attenuation_sparse"):
            assert np.array_equal(np.isnan(small[k]), np.isnan(big[k])), k
            m = np.isfinite(big[k]) & (big[k] > 1e-3)
            assert m.sum() > 1000 and np.max(np.abs(small[k][m] / big[k][m] - 1)) < 2e-5, k
        ora = oracle_mod.Oracle(ice, attenuation_model=model, n_reflections=n_refl, n_freq=15, tight=True,
                                gl3_table=__import__("nuradiomc_b200.utilities.attenuation", fromlist=["x"]).gl3_parameters() if model == "GL3" else None
                                ).trace(np.repeat(V, len(A), 0), np.tile(A, (len(V), 1)), ff, 0.8)
        kept = ~np.isnan(small["attenuation"][..., 0])          # solutions the viewing-angle cut kept
        assert_attenuation_parity(small["attenuation"][kept], ora["attenuation"][kept])
    # device entry point: the same launch on the caller's stream
    dv, da = torch.tensor(np.ascontiguousarray(V.T), device="cuda"), torch.tensor(np.ascontiguousarray(A.T), device="cuda")
    kd = {k: v for k, v in kw.items() if k not in ("shower_axis", "delta_C_cut")}
    dev = rt.trace_batch_device(dv, da, sync_stats=True, **kd)
    host = rt.trace_batch(V, A, **kd)
    assert dev.stats["n_launches"] == 1
    for k in host:
        np.testing.assert_array_equal(dev[k].cpu().numpy(), host[k], err_msg=k)


def test_small_batch_path_scalar_api(make, oracle_mod, monkeypatch):
    """the scalar API on its default path (a batch of one pair = one fused launch): reference semantics against the oracle"""
    monkeypatch.delenv("NRMC_NO_SMALL_PATH")
    rt = make("southpole_2015", attenuation_model="SP1", n_frequencies_integration=25)
    ff = np.fft.rfftfreq(256, 0.5)
    X1, x2 = cylinder(5, 60, 3000, -2000), np.array([10., 10., -190.])
    ora = oracle_mod.Oracle("southpole_2015", attenuation_model="SP1", n_freq=25).trace(X1, x2, ff, 0.8)
    for i, x1 in enumerate(X1):
        rt.set_start_and_end_point(x1, x2)
        rt.find_solutions()
        assert rt.get_number_of_solutions() == ora["n_sol"][i]
        for iS in range(rt.get_number_of_solutions()):
            assert rt.get_solution_type(iS) == ora["type"][i, iS]
            np.testing.assert_allclose(rt.get_launch_vector(iS), ora["launch"][i, iS], atol=1e-6)
            np.testing.assert_allclose(rt.get_path_length(iS), ora["path_length"][i, iS], rtol=1e-6)
            np.testing.assert_allclose(rt.get_travel_time(iS), ora["travel_time"][i, iS], rtol=1e-6)
            assert_attenuation_parity(rt.get_attenuation(iS, ff, 0.8), ora["attenuation"][i, iS])
    # no solution, swapped points, one point in air
    for x1, xb in (([3000., 0, -10.], [0, 0, -5.]), ([0, 0, -100.], [500., 0, -900.]), ([100., 0, 5.], [0, 0, -50.])):
        rt.set_start_and_end_point(x1, xb)
        rt.find_solutions()
        o = oracle_mod.Oracle("southpole_2015").trace(np.array([x1]), np.array([xb]))
        assert rt.get_number_of_solutions() == o["n_sol"][0]


@pytest.mark.parametrize("ice,model,n_refl,rmax,zmin,ant,fmax", [
    ("mooresbay_simple", "MB1", 1, 1000, -500, [[3, 3, -5.], [-3, 0, -1.]], None),          # cfg4 + MB1
    ("mooresbay_simple", "MB1", 2, 800, -570, [[0, 0, -5.]], 0.6),
    ("greenland_simple", "GL2", 0, 3000, -2500, [[0, 0, -100.], [10, 0, -2.]], None),       # bins above 1.58 GHz: negative length -> 1 m floor
])
def test_separable_models_thread_per_solution_kernel(make, oracle_mod, monkeypatch, ice, model, n_refl, rmax, zmin, ant, fmax):
    """K_att_sep (MB1 / GL2, sparse output, any number of bottom reflections) against the generic warp-per-solution kernel on
    the same solutions (same nodes and floor decisions: 1e-8) and against the tight oracle (1e-4); with a dense output of a
    bottom-reflected path the generic kernel must have been used (per-segment interpolation, py:1077-1086)"""
    ff = np.fft.rfftfreq(256, 0.25)          # 0 ... 2 GHz
    V, A = cylinder(700 + n_refl, 1500, rmax, zmin), np.array(ant, float)
    kw = dict(outer=True, frequency=ff, max_detector_freq=fmax, attenuation="sparse", compact=True)
    fast = make(ice, attenuation_model=model, n_reflections=n_refl, n_frequencies_integration=20).trace_batch(V, A, **kw)
    monkeypatch.setenv("NRMC_SEP_GENERIC", "1")
    rt_g = make(ice, attenuation_model=model, n_reflections=n_refl, n_frequencies_integration=20)
    gen = rt_g.trace_batch(V, A, **kw)
    monkeypatch.delenv("NRMC_SEP_GENERIC")
    n = int(fast["sol_offset"][-1])
    assert n == int(gen["sol_offset"][-1]) and n > 1000
    np.testing.assert_array_equal(fast["n_sol"], gen["n_sol"])
    a, b = fast["attenuation_sparse"][:n], gen["attenuation_sparse"][:n]
    assert np.isfinite(a).all() and np.isfinite(b).all()
    big = b > 1e-3
    assert big.sum() > 1000 and np.max(np.abs(a[big] / b[big] - 1)) < 1e-8
    assert np.max(np.abs(a - b)[~big]) < 1e-10 if (~big).any() else True
    if model == "GL2":
        assert (fast.frequencies_sparse > 1.6).any()      # floored bins on the whole path: exp(-path length)
    # tight oracle on the padded layout of a subset (pair mode)
    sel = np.arange(0, len(V), 5)
    X1 = np.repeat(V[sel], len(A), axis=0)
    X2 = np.tile(A, (len(sel), 1))
    pad = make(ice, attenuation_model=model, n_reflections=n_refl, n_frequencies_integration=20).trace_batch(
        X1, X2, frequency=ff, max_detector_freq=fmax, attenuation="sparse")
    ora = oracle_mod.Oracle(ice, attenuation_model=model, n_reflections=n_refl, n_freq=20, tight=True).trace(X1, X2, ff, fmax)
    np.testing.assert_array_equal(pad["n_sol"], ora["n_sol"])
    assert_attenuation_parity(pad["attenuation_sparse"], ora["attenuation_sparse"])


def test_separable_kernel_floor_crossed_on_the_path(make, oracle_mod, monkeypatch):
    """GL2 at 1.5762 GHz: the length (852 m - 540 m/GHz f) p0(z) crosses the 1 m floor of attenuation.py:252-255 ON the paths
    (p0 varies between 1.0 and 1.3 with depth).  K_att_sep hands such solutions to the generic kernel through its fall-back
    list; the other frequency of the same solutions stays on its fast path."""
    ff = np.array([0.0, 0.3, 1.5762])
    V, A = cylinder(911, 4000, 3000, -2500), np.array([[0, 0, -100.], [10, 0, -2.]])
    kw = dict(outer=True, frequency=ff, attenuation="sparse", compact=True)
    fast = make("greenland_simple", attenuation_model="GL2").trace_batch(V, A, **kw)
    monkeypatch.setenv("NRMC_SEP_GENERIC", "1")
    gen = make("greenland_simple", attenuation_model="GL2").trace_batch(V, A, **kw)
    monkeypatch.delenv("NRMC_SEP_GENERIC")
    n = int(fast["sol_offset"][-1])
    assert n == int(gen["sol_offset"][-1]) and n > 3000
    a, b = fast["attenuation_sparse"][:n], gen["attenuation_sparse"][:n]
    assert np.isfinite(a).all()
    np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-300)      # fast path or fall-back (then the generic kernel wrote the row in both runs)
    # the floor really is crossed on some paths: the factor is neither exp(-path length) nor exp(-G / fa) there
    S = fast["path_length"][:n]
    assert (np.abs(a[:, 1] - np.exp(-S)) > 1e-6 * np.exp(-S)).sum() > 100
    sel = np.arange(0, len(V), 20)
    X1, X2 = np.repeat(V[sel], len(A), axis=0), np.tile(A, (len(sel), 1))
    pad = make("greenland_simple", attenuation_model="GL2").trace_batch(X1, X2, frequency=ff, attenuation="sparse")
    ora = oracle_mod.Oracle("greenland_simple", attenuation_model="GL2", tight=True).trace(X1, X2, ff, None)
    np.testing.assert_array_equal(pad["n_sol"], ora["n_sol"])
    assert_attenuation_parity(pad["attenuation_sparse"], ora["attenuation_sparse"])
