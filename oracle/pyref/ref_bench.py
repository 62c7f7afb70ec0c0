"""
TEST / BENCH INFRASTRUCTURE -- times the UNMODIFIED NuRadioMC Python reference (scratch copy under ``baseline/_ref``, loaded
through ``ref_harness``) on host cores: one process per core (``multiprocessing``, spawn), every process runs the reference's
scalar API exactly as ``NuRadioMC/simulation/simulation.py:155-210`` drives it (set_start_and_end_point, find_solutions,
per solution: type, launch / receive vector, path length, travel time and -- where the configuration says so -- get_attenuation;
``analyticraytracing.py:1400-1547`` and ``:933-1089`` are where the time goes).  Two variants, as BASELINE.md section 3 plans:
``plain`` (compile_numba=False) and ``numba`` (the reference's own jit path, warm: the first pair of every process is untimed).
Only ``bench.py`` (cpu_baseline leg and ``--impl reference``) and scripts under ``profiles/`` call this; nothing in the product does.

The numba variant rebinds the module-level ray functions of the process it runs in (analyticraytracing.py:465-481), hence a
separate pool per variant.  With bottom reflections both variants need the F4 harness patch (ref_harness.py docstring).
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    sys.path.insert(0, _HERE)
    import ref_harness
    return os.path.isdir(os.path.join(ref_harness.REF_ROOT, "NuRadioMC"))


def _worker(args):
    cfg, X1, X2, numba, n_warm = args
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nrmc_numba_cache")
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    import warnings
    warnings.filterwarnings("ignore")
    sys.path.insert(0, _HERE)
    import logging
    import ref_harness as rh
    if numba and cfg.get("n_refl", 0) > 0:
        raise ValueError("the numba variant cannot carry the F4 harness patch: only without bottom reflections")
    ray, medium, _ = rh.load_reference(patch_f4=not numba)
    r = ray.ray_tracing(medium.get_ice_model(cfg["ice"]), attenuation_model=cfg.get("att") or "SP1",
                        n_frequencies_integration=cfg.get("n_freq"), n_reflections=cfg.get("n_refl", 0), use_cpp=False,
                        compile_numba=bool(numba), log_level=logging.ERROR)
    freqs, fmax, with_att = cfg.get("freqs"), cfg.get("fmax"), cfg.get("att") is not None

    def one(i):
        r.set_start_and_end_point(X1[i], X2[i])
        r.find_solutions()
        n = r.get_number_of_solutions()
        for iS in range(n):
            r.get_solution_type(iS)
            r.get_launch_vector(iS)
            r.get_receive_vector(iS)
            r.get_path_length(iS)
            r.get_travel_time(iS)
            if with_att:
                r.get_attenuation(iS, freqs, fmax)
        return n
    for i in range(min(n_warm, len(X1))):       # jit compilation / imports / first-call caches: untimed
        one(i)
    t0 = time.perf_counter()
    n_sol = 0
    for i in range(len(X1)):
        n_sol += one(i)
    return len(X1), n_sol, time.perf_counter() - t0


def run(cfg, X1, X2, n_procs=None, numba=False, n_warm=1):
    """
    cfg: dict(ice, att (None = no attenuation), n_refl, n_freq, freqs, fmax); X1, X2: (N, 3) pairs.  The pairs are split into
    equal contiguous chunks, one per process; the rate is N over the wall time of the slowest process (warm-up excluded).
    """
    n_procs = n_procs or len(os.sched_getaffinity(0))
    X1, X2 = np.asarray(X1, float).reshape(-1, 3), np.asarray(X2, float).reshape(-1, 3)
    n_procs = max(1, min(n_procs, len(X1)))
    bounds = np.linspace(0, len(X1), n_procs + 1).astype(int)
    jobs = [(cfg, X1[a:b], X2[a:b], numba, n_warm) for a, b in zip(bounds[:-1], bounds[1:])]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(n_procs) as pool:
        parts = pool.map(_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    slowest = max(p[2] for p in parts)
    n, n_sol = sum(p[0] for p in parts), sum(p[1] for p in parts)
    return {"value": n / slowest, "unit": "pairs/s", "cores": n_procs, "per_core": n / sum(p[2] for p in parts), "pairs": n,
            "solutions": n_sol, "seconds_slowest_process": slowest, "seconds_wall_with_startup": wall,
            "kind": "reference", "variant": "numba (warm)" if numba else "plain python"}
