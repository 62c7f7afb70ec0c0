"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/nrmc_rt.h declares, the
Python class mirrors the reference's constructor / error behaviour, and nothing computes without a GPU."""
import ctypes
import logging
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_functions():
    txt = open(os.path.join(ROOT, "include", "nrmc_rt.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nrmc_rt_\w+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    from nuradiomc_b200 import _build, _lib
    lib_path = _build.build()
    assert os.path.exists(lib_path)
    lib = ctypes.CDLL(lib_path)
    declared = _declared_functions()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/nrmc_rt.h but not exported"
    assert set(_lib.EXPORTS) == set(declared)
    assert b"sm_100a" in _lib.load().nrmc_rt_version()


def test_sass_is_sm100a_and_uses_tma_bulk_copy():
    """the attenuation kernel stages its frequency tables with cp.async.bulk (SASS: UBLKCP), compiled for sm_100a"""
    import subprocess
    from nuradiomc_b200 import _build
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", _build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out
    assert "DFMA" in out


def test_registry_and_constants():
    from nuradiomc_b200.SignalProp import propagation
    from nuradiomc_b200.SignalProp.analyticraytracing import ray_tracing
    assert propagation.get_propagation_module("analytic") is ray_tracing
    assert propagation.solution_types == {1: 'direct', 2: 'refracted', 3: 'reflected'}
    assert propagation.solution_types_revert['reflected'] == 3
    with pytest.raises(NotImplementedError):
        propagation.get_propagation_module("does_not_exist")
    from nuradiomc_b200.utilities import attenuation
    assert attenuation.model_to_int == {"SP1": 1, "GL1": 2, "MB1": 3, "GL2": 4, "GL3": 5}
    assert attenuation.gl3_parameters().shape == (300, 3)


def test_constructor_semantics_match_reference():
    """propagation_base_class.py:86-133 precedence config > kwargs > defaults; error types of analyticraytracing.py"""
    from nuradiomc_b200.SignalProp.analyticraytracing import ray_tracing
    from nuradiomc_b200.utilities import medium
    ice = medium.get_ice_model("southpole_2015")
    r = ray_tracing(ice)
    assert (r._n_frequencies_integration, r._n_reflections, r._attenuation_model) == (100, 0, "SP1")
    assert r.get_number_of_raytracing_solutions() == 2
    assert r.get_config()["propagation"]["attenuate_ice"] is True
    cfg = {"propagation": {"n_freq": 25, "n_reflections": 0, "attenuation_model": "GL1", "focusing": False}}
    r = ray_tracing(ice, attenuation_model="SP1", n_frequencies_integration=7, config=cfg, log_level=logging.ERROR)
    assert (r._n_frequencies_integration, r._attenuation_model) == (25, "GL1")
    # reflections requested without a reflective layer -> silently 0 (base:128-133)
    assert ray_tracing(ice, n_reflections=2, log_level=logging.ERROR)._n_reflections == 0
    mb = medium.get_ice_model("mooresbay_simple")
    r = ray_tracing(mb, n_reflections=1)
    assert r.get_number_of_raytracing_solutions() == 6
    with pytest.raises(AttributeError):   # base:156-161
        r.set_start_and_end_point([0, 0, -600.], [10, 0, -5.])
    with pytest.raises(TypeError):
        ray_tracing(object())
    with pytest.raises(RuntimeError):
        ray_tracing(medium.uniform_ice())
    with pytest.raises(NotImplementedError):
        ray_tracing(ice, attenuation_model="XX9")
    names = [p["name"] for p in ray_tracing(ice).get_output_parameters()]
    assert names == ['ray_tracing_C0', 'ray_tracing_C1', 'focusing_factor', 'ray_tracing_reflection',
                     'ray_tracing_reflection_case', 'ray_tracing_solution_type']


def test_detector_max_frequency():
    """max_detector_frequency = max over stations of half the sampling rate of the first channel (base:64-80)"""
    from nuradiomc_b200.SignalProp.analyticraytracing import ray_tracing
    from nuradiomc_b200.utilities import medium

    class Det:
        def get_station_ids(self): return [11, 12]
        def get_channel_ids(self, s): return [0, 1]
        def get_sampling_frequency(self, s, c): return 2.4 if s == 11 else 1.0
    r = ray_tracing(medium.get_ice_model("greenland_simple"), detector=Det())
    assert r._max_detector_frequency == 1.2


def test_set_solution_roundtrip():
    from nuradiomc_b200.SignalProp.analyticraytracing import ray_tracing
    from nuradiomc_b200.utilities import medium
    r = ray_tracing(medium.get_ice_model("southpole_2015"))
    r.set_solution({'ray_tracing_C0': np.array([0.7, np.nan]), 'ray_tracing_C1': np.array([1.0, np.nan]),
                    'ray_tracing_solution_type': np.array([1, 0])})
    assert r.get_number_of_solutions() == 1 and r.get_results()[0]['reflection'] == 0
    assert r.get_solution_type(0) == 1
    with pytest.raises(IndexError):
        r.get_solution_type(1)


def test_no_cpu_fallback():
    """without a CUDA device every compute entry point fails loudly (NRMC_ERR_NO_DEVICE), it never computes on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nuradiomc_b200.SignalProp.analyticraytracing import ray_tracing
    from nuradiomc_b200.utilities import medium
    r = ray_tracing(medium.get_ice_model("southpole_2015"))
    r.set_start_and_end_point([0, 0, -500.], [100, 0, -100.])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        r.find_solutions()
    with pytest.raises(RuntimeError, match="no CUDA device"):
        r.trace_batch(np.zeros((4, 3)), np.zeros((4, 3)))


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under nuradiomc_b200/ may reference it"""
    for dp, _, files in os.walk(os.path.join(ROOT, "nuradiomc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower() or f == "nrmc_math.cuh" and "from oracle" not in txt, os.path.join(dp, f)


def test_shard_bounds():
    from nuradiomc_b200.distributed import shard_bounds, shard_vertices
    for n in (0, 1, 7, 1000003):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    idx = np.concatenate([shard_vertices(101, 4, r, permutation_seed=3) for r in range(4)])
    assert np.array_equal(np.sort(idx), np.arange(101))


def test_ctypes_structures_match_the_header(tmp_path):
    """the ctypes mirrors in nuradiomc_b200/_lib.py have the size and field offsets of the structs in include/nrmc_rt.h
    (compiled with gcc: a drift between the two is an ABI bug no parity test would localise)"""
    import ctypes as C
    import subprocess
    from nuradiomc_b200 import _lib
    pairs = {"nrmc_rt_config": _lib.Config, "nrmc_rt_input": _lib.Input, "nrmc_rt_output": _lib.Output,
             "nrmc_rt_stats": _lib.Stats, "nrmc_rt_effects": _lib.Effects, "nrmc_rt_focusing": _lib.Focusing}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "nrmc_rt.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).split("\n")
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out if l.strip()}
    for cname, cls in pairs.items():
        assert got[(cname, "size")] == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)


def test_pair_lookup_of_prepare_batch():
    """the vectorised key table of prepare_batch equals the per-pair construction find_solutions relies on"""
    from nuradiomc_b200.SignalProp.analyticraytracing import _pair_lookup
    rng = np.random.default_rng(5)
    X1, X2 = rng.normal(size=(40, 3)), rng.normal(size=(7, 3))
    X1[3] = X1[1]                                       # duplicate vertex: the later pair wins, as in a loop
    X1[5, 2] = -0.0                                     # -0.0 and 0.0 are different keys, as with tobytes()
    ref = {}
    for i in range(len(X1)):
        for j in range(len(X2)):
            ref[X1[i].tobytes() + X2[j].tobytes()] = i * len(X2) + j
    assert _pair_lookup(X1, X2, True) == ref
    ref = {X1[i].tobytes() + X2[0].tobytes(): i for i in range(len(X1))}
    assert _pair_lookup(X1, X2[:1], False) == ref
    Y2 = rng.normal(size=(40, 3))
    ref = {X1[i].tobytes() + Y2[i].tobytes(): i for i in range(len(X1))}
    got = _pair_lookup(X1, Y2, False)
    assert got == ref and all(isinstance(k, bytes) and len(k) == 48 for k in got)


def test_bench_workloads_and_roofline_inputs():
    """bench.py without a GPU: the BASELINE workloads have the sizes SURVEY.md 8(d) states, the reference-algorithm FLOP model
    gives the figures DESIGN.md quotes, and profiles/hw_counts.json holds executed-FLOP counts for every kernel the bench looks up"""
    import json
    import bench
    sizes = {"cfg1": (1000, 1), "cfg2": (1_000_000, 4), "cfg3": (1_000_000, 24), "cfg4": (1_000_000, 8), "cfg5": (1_000_000, 100)}
    for name, (nv, na) in sizes.items():
        cfg = bench.CONFIGS[name]
        assert cfg["n_vertices"] == nv and cfg["antennas"].shape == (na, 3), name
        V, A, _ = bench.workload(1000, name)
        assert V.shape == (3, 1000) and A.shape == (3, na) and (V[2] <= 0).all() and (A[2] < 0).all()
        assert np.array_equal(V, bench.workload(1000, name)[0])                       # seeded: the same vertices every time
    r = np.hypot(*bench.workload(20000, "cfg5")[0][:2])
    assert r.max() < 6000.0 and abs(np.mean(r ** 2) / 6000.0 ** 2 - 0.5) < 0.02       # uniform in the cylinder's cross section
    assert bench.reference_model_flops(bench.CONFIGS["cfg5"], 37) == (3050, 600, 31272)   # W_att(SP1, F = 37)
    assert bench.reference_model_flops(bench.CONFIGS["cfg2"], 0) == (3050, 600, 0)
    assert bench.reference_model_flops(bench.CONFIGS["cfg4"], 0)[0] == 50 + 3 * 30 * 100 * 2
    hw = json.load(open(os.path.join(ROOT, "profiles", "hw_counts.json")))
    for key in ("K_classify", "K_hump", "K_roots", "K_att_sp1", "cfg3:K_att_gl1", "cfg4:K_roots_m", "cfg4:K_classify_m", "cfg4mb1:K_att_sep"):
        assert hw[key]["fp64_flops_per_unit"] > 0 and hw[key]["dram_bytes_per_unit"] > 0, key
    X1, X2 = bench.pairs_of(*bench.workload(10, "cfg3")[:2], 50)
    assert X1.shape == X2.shape == (50, 3) and np.array_equal(X2[:24], bench.RNOG) and np.array_equal(X1[0], X1[23])
