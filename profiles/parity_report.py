#!/usr/bin/env python
"""
Parity report of SURVEY.md 8(c), written to profiles/r2_parity_report.json (run on the GPU box: python profiles/parity_report.py).

For every fixture produced by the reference's own Python path (tests/golden/*.npz, tests/golden/make_golden.py) and for seeded
samples of every BASELINE configuration against the CPU restatement (oracle/):
  * solution counts: exact-match rate, arbiter-resolved rate (count differs, the dense scan of the reference's own objective
    agrees with the kernel), unresolved (must be 0);
  * solution type / reflection / reflection_case: mismatches among the solutions of pairs with equal counts (must be 0);
  * per quantity and per solution type (1 direct, 2 refracted, 3 reflected): the largest deviation -- C0, path length, travel time
    relative; launch / receive vectors and reflection angles absolute (rad);
  * how many solutions need the widened path-length / travel-time band of tests/conftest.py::assert_parity (beyond 1e-6 relative);
  * attenuation: largest relative deviation on factors above 1e-3 from the reference integrand at quad(epsrel=1e-11) and the largest
    absolute one below, and the deviation from the reference's stock (epsrel=1e-2) result.
This is test infrastructure: it imports oracle/ and the fixtures; the product never does.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("NRMC_NO_SMALL_PATH", "1")          # the production kernels, also on the small fixtures

from conftest import GOLDEN, golden_config, load_golden          # noqa: E402
import bench                                                        # noqa: E402
from nuradiomc_b200.SignalProp import propagation                   # noqa: E402
from nuradiomc_b200.utilities import attenuation, medium            # noqa: E402
from oracle.oracle import Oracle                                    # noqa: E402

ALIAS = {"type": "solution_type", "launch": "launch_vector", "receive": "receive_vector"}


def compare(res, ref, arbiter_n=None, n_ice=1.78):
    get = lambda d, k: d[ALIAS[k]] if (k in ALIAS and ALIAS[k] in d) else d[k]
    N = len(ref["n_sol"])
    same = res["n_sol"] == ref["n_sol"]
    rep = {"pairs": int(N), "solutions_reference": int(ref["n_sol"].sum()), "solutions_kernel": int(res["n_sol"].sum()),
           "count_exact_match_rate": float(same.mean())}
    if arbiter_n is not None:
        resolved = (~same) & (res["n_sol"] == arbiter_n)
        rep["count_arbiter_resolved_rate"] = float(resolved.mean())
        rep["count_unresolved"] = int(((~same) & ~resolved).sum())
    else:
        rep["count_mismatches"] = int((~same).sum())
    S = ref["C0"].shape[1]
    filled = (np.arange(S)[None, :] < ref["n_sol"][:, None]) & same[:, None]
    typ = get(ref, "type")
    rep["type_mismatches"] = int((get(res, "type")[filled] != typ[filled]).sum())
    rep["reflection_mismatches"] = int((res["reflection"][filled] != ref["reflection"][filled]).sum() +
                                       (res["reflection_case"][filled] != ref["reflection_case"][filled]).sum())
    per_type = {}
    with np.errstate(invalid="ignore", divide="ignore"):
        tol = 1e-6 + 3e-9 / np.maximum(ref["C0"] * n_ice - 1, 1e-12)
    widened = 0
    for t in (1, 2, 3):
        m = filled & (typ == t)
        if not m.any():
            continue
        e = {"solutions": int(m.sum())}
        for k in ("C0", "path_length", "travel_time"):
            e[k + "_max_rel"] = float(np.nanmax(np.abs(res[k][m] / ref[k][m] - 1)))
        for k in ("launch", "receive"):
            e[k + "_max_abs"] = float(np.nanmax(np.abs(get(res, k)[m] - get(ref, k)[m])))
        ra, rb = res["reflection_angle"][m], ref["reflection_angle"][m]
        K1 = min(ra.shape[-1], rb.shape[-1])
        if np.isfinite(rb[..., :K1]).any():
            e["reflection_angle_max_abs"] = float(np.nanmax(np.abs(ra[..., :K1] - rb[..., :K1])))
        for k in ("path_length", "travel_time"):
            floor = 2e-5 if k == "path_length" else 2e-5 * n_ice / 0.299792458
            d = np.abs(res[k][m] - ref[k][m])
            e[k + "_beyond_1e-6"] = int((d > 1e-6 * np.abs(ref[k][m])).sum())
            e[k + "_beyond_widened_band"] = int((d > tol[m] * np.abs(ref[k][m]) + floor).sum())
            widened += e[k + "_beyond_1e-6"]
        per_type[str(t)] = e
    rep["per_solution_type"] = per_type
    rep["solutions_needing_the_widened_band"] = widened
    return rep, same


def attenuation_report(att, ref, same):
    a, b = att[same], ref[same]
    big = b > 1e-3
    small = ~big & np.isfinite(b)
    return {"bins_above_1e-3": int(big.sum()), "max_rel_above_1e-3": float(np.nanmax(np.abs(a - b)[big] / b[big])) if big.any() else None,
            "max_abs_below_1e-3": float(np.nanmax(np.abs(a - b)[small])) if small.any() else None,
            "nan_pattern_equal": bool(np.array_equal(np.isnan(a), np.isnan(b)))}


def main():
    prop = propagation.get_propagation_module("analytic")
    report = {"protocol": "SURVEY.md 8(c)", "fixtures_from_the_python_reference": {}, "seeded_samples_vs_oracle": {}}
    names = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f[:-4] not in ("propagation_effects", "focusing", "simulation_datasets"))
    for name in names:
        g = load_golden(name)
        c = golden_config(g)
        rt = prop(medium.get_ice_model(c["ice"]), attenuation_model=c["attenuation_model"], n_reflections=c["n_reflections"],
                  n_frequencies_integration=c["n_freq"])
        res = rt.trace_batch(g["X1"], g["X2"], frequency=c["frequencies"], max_detector_freq=c["max_detector_freq"])
        rep, same = compare(res, g, g["arbiter_n"])
        if c["attenuation_model"]:
            rep["attenuation_vs_tight_reference_integrand"] = attenuation_report(res["attenuation"], g["attenuation_tight"], same)
            stock = attenuation_report(res["attenuation"], g["attenuation"], same)
            rep["attenuation_vs_stock_reference_epsrel_1e-2"] = {"max_rel_above_1e-3": stock["max_rel_above_1e-3"]}
        rep["configuration"] = {k: (v if not isinstance(v, np.ndarray) else f"{len(v)} bins") for k, v in c.items()}
        report["fixtures_from_the_python_reference"][name] = rep
        print(name, rep["count_exact_match_rate"], rep.get("count_unresolved"), flush=True)
    for cname, n_vert in (("cfg1", 1000), ("cfg2", 25000), ("cfg3", 2500), ("cfg4", 6000), ("cfg4mb1", 1500), ("cfg5", 600)):
        cfg = bench.CONFIGS[cname]
        V, A, ff = bench.workload(n_vert, cname)
        X1, X2 = bench.pairs_of(V, A, n_vert * A.shape[1])
        rt = prop(medium.get_ice_model(cfg["ice"]), attenuation_model=cfg["att"] or "SP1", n_reflections=cfg["n_refl"],
                  n_frequencies_integration=cfg["n_freq"])
        kw = dict(frequency=ff, max_detector_freq=cfg["fmax"], attenuation="both") if cfg["att"] else {}
        res = rt.trace_batch(X1, X2, **kw)
        o = Oracle(cfg["ice"], attenuation_model=cfg["att"], n_reflections=cfg["n_refl"], n_freq=cfg["n_freq"] or 100, tight=True)
        ora = o.trace(X1, X2, ff if cfg["att"] else None, cfg["fmax"])
        rep, same = compare(res, ora)
        if cfg["att"]:
            rep["attenuation_sparse_vs_tight_oracle"] = attenuation_report(res["attenuation_sparse"], ora["attenuation_sparse"], same)
            rep["attenuation_dense_vs_tight_oracle"] = attenuation_report(res["attenuation"], ora["attenuation"], same)
        report["seeded_samples_vs_oracle"][cname] = rep
        print(cname, rep["pairs"], rep["count_exact_match_rate"], flush=True)
    tot = {"pairs": 0, "solutions": 0, "unresolved": 0, "type_mismatches": 0, "needing_widened_band": 0}
    for grp in ("fixtures_from_the_python_reference", "seeded_samples_vs_oracle"):
        for r in report[grp].values():
            tot["pairs"] += r["pairs"]; tot["solutions"] += r["solutions_kernel"]
            tot["unresolved"] += r.get("count_unresolved", r.get("count_mismatches", 0))
            tot["type_mismatches"] += r["type_mismatches"] + r["reflection_mismatches"]
            tot["needing_widened_band"] += r["solutions_needing_the_widened_band"]
    report["totals"] = tot
    out = os.path.join(ROOT, "gpurun_out" if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else "profiles", "r2_parity_report.json")
    json.dump(report, open(out, "w"), indent=1)
    print(json.dumps(tot))


if __name__ == "__main__":
    main()
