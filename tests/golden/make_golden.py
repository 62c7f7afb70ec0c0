"""
Generates the committed golden fixtures tests/golden/*.npz by running the UNMODIFIED NuRadioMC Python reference
(scratch copy baseline/_ref, see oracle/pyref/ref_harness.py) in the build container.

    python tests/golden/make_golden.py [case ...]

Every fixture holds the inputs (X1, X2, frequencies, configuration) and the reference's outputs through its public
scalar API (find_solutions / get_solution_type / get_launch_vector / get_receive_vector / get_path_length /
get_travel_time / get_reflection_angle / get_attenuation), plus
  * `attenuation_tight`: same reference integrand integrated with quad(epsrel=1e-11), and
  * `arbiter_C0`: for pairs where the reference's count differs from the oracle's, the roots found by a dense
    scan of the reference's *own* objective (SURVEY.md F6 protocol).
The reference's golden pickles of T05/T06 are copied as plain arrays too (they are data, 1000x2 and 1000x10 f64).
"""
import os
import pickle
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle", "pyref"))
sys.path.insert(0, REPO)

RNOG = np.array([[0, 20, -97], [0, 20, -96], [0, 20, -95], [0, 20, -94], [0, 20, -93], [0, 20, -92], [0, 20, -80],
                 [0, 20, -60], [0, 20, -40], [-17.3, -10, -96], [-17.3, -10, -95], [-17.3, -10, -94], [1.5, 11, -2],
                 [0, 11, -2], [-1.5, 11, -2], [-10.276, -4.2, -2], [-9.526, -5.5, -2], [-8.776, -6.8, -2],
                 [8.776, -6.8, -2], [9.526, -5.5, -2], [10.276, -4.2, -2], [17.3, -10, -96], [17.3, -10, -95],
                 [17.3, -10, -94]], float)


def cylinder(seed, n, rmax, zmin, rmin=0.0):
    """uniform-in-cylinder vertices, NuRadioMC/EvtGen/generator.py:613-618"""
    rng = np.random.default_rng(seed)
    r = np.sqrt(rng.uniform(rmin ** 2, rmax ** 2, n))
    phi = rng.uniform(0, 2 * np.pi, n)
    z = rng.uniform(zmin, 0, n)
    return np.array([r * np.cos(phi), r * np.sin(phi), z]).T


def t05_points(seed, zmax, n=1000):
    """T05/T06/T01 vertex distribution (legacy numpy RNG), T05unit_test_C0_SP.py:15-26"""
    np.random.seed(seed)
    rr = np.random.triangular(50., 3000., 3000., n)
    ph = np.random.uniform(0, 2 * np.pi, n)
    xx, yy = rr * np.cos(ph), rr * np.sin(ph)
    zz = np.random.uniform(0., zmax, n)
    return np.array([xx, yy, zz]).T


def pairs(vertices, antennas):
    V = np.repeat(vertices, len(antennas), axis=0)
    A = np.tile(antennas, (len(vertices), 1))
    return V, A


def cases():
    c = {}
    X1 = t05_points(10, -3000.)[:300]
    c["sp_simple_T05"] = dict(ice="southpole_simple", att=None, n_refl=0, n_freq=100, X1=X1,
                              X2=np.repeat([[0, 0, -5.]], len(X1), 0))
    V, A = pairs(cylinder(2, 150, 4000., -2700.), np.array([[10, 10, -190.], [-10, 10, -190.]]))
    c["sp2015_cfg2"] = dict(ice="southpole_2015", att=None, n_refl=0, n_freq=100, X1=V, X2=A)
    X1 = t05_points(0, -3000.)[:36]
    c["sp1_cfg1"] = dict(ice="southpole_simple", att="SP1", n_refl=0, n_freq=100, X1=X1,
                         X2=np.repeat([[0, 0, -100.]], len(X1), 0), freqs=np.linspace(0, 0.5, 129), fmax=None)
    V, A = pairs(cylinder(3, 24, 4000., -2700.), RNOG[[0, 8, 13]])
    c["greenland_cfg3"] = dict(ice="greenland_simple", att="GL1", n_refl=0, n_freq=25, X1=V, X2=A,
                               freqs=np.fft.rfftfreq(1022, 0.2), fmax=1.2)
    V, A = pairs(cylinder(4, 60, 1000., -500.), np.array([[3, 3, -5.], [-3, 0, -1.]]))
    c["mooresbay_cfg4"] = dict(ice="mooresbay_simple", att=None, n_refl=1, n_freq=25, X1=V, X2=A)
    V, A = pairs(cylinder(14, 12, 1000., -500.), np.array([[3, 3, -5.]]))
    c["mooresbay_cfg4_MB1"] = dict(ice="mooresbay_simple", att="MB1", n_refl=1, n_freq=25, X1=V, X2=A,
                                   freqs=np.fft.rfftfreq(256, 0.5), fmax=None)
    X1 = t05_points(10, -500.)[:80]
    c["mooresbay_T06"] = dict(ice="mooresbay_simple", att=None, n_refl=2, n_freq=100, X1=X1,
                              X2=np.repeat([[0, 0, -5.]], len(X1), 0))
    V, A = pairs(cylinder(5, 20, 6000., -2700.), np.array([[0, 0, -150.], [1500, -1500, -160.]]))
    c["sp2015_cfg5_SP1"] = dict(ice="southpole_2015", att="SP1", n_refl=0, n_freq=25, X1=V, X2=A,
                                freqs=np.fft.rfftfreq(1022, 0.2), fmax=1.2)
    V, A = pairs(cylinder(6, 16, 3000., -2500.), np.array([[0, 0, -100.]]))
    c["greenland_GL2"] = dict(ice="greenland_simple", att="GL2", n_refl=0, n_freq=20, X1=V, X2=A,
                              freqs=np.fft.rfftfreq(256, 0.5), fmax=None)
    V, A = pairs(cylinder(7, 24, 3000., -2700.), RNOG[[0, 13]])
    c["greenland_GL3"] = dict(ice="greenland_simple", att="GL3", n_refl=0, n_freq=15, X1=V, X2=A,
                              freqs=np.fft.rfftfreq(256, 0.5), fmax=None)
    # round 2: >= 1e3 pairs of the two attenuation benchmarks (cfg3: all 24 RNO-G channels; cfg5: 5 of the 25 stations x 4 depths)
    V, A = pairs(cylinder(33, 42, 4000., -2700.), RNOG)
    c["greenland_cfg3_large"] = dict(ice="greenland_simple", att="GL1", n_refl=0, n_freq=25, X1=V, X2=A,
                                     freqs=np.fft.rfftfreq(1022, 0.2), fmax=1.2)
    g5 = np.array([[x, y, z] for x, y in ((0., 0.), (1500., -1500.), (-3000., 3000.), (3000., 0.), (-1500., -3000.))
                   for z in (-145., -150., -155., -160.)])
    V, A = pairs(cylinder(55, 50, 6000., -2700.), g5)
    c["sp2015_cfg5_large"] = dict(ice="southpole_2015", att="SP1", n_refl=0, n_freq=25, X1=V, X2=A,
                                  freqs=np.fft.rfftfreq(1022, 0.2), fmax=1.2)
    # round 2: the separable models through K_att_sep (n_reflections = 0: sparse factors from the thread-per-solution kernel, dense by
    # interpolation), pinned to the reference directly
    V, A = pairs(cylinder(66, 150, 3000., -2500.), np.array([[0, 0, -100.], [10, 0, -2.]]))
    c["greenland_GL2_large"] = dict(ice="greenland_simple", att="GL2", n_refl=0, n_freq=20, X1=V, X2=A,
                                    freqs=np.fft.rfftfreq(256, 0.25), fmax=None)
    V, A = pairs(cylinder(67, 110, 1000., -550.), np.array([[3, 3, -5.], [-3, 0, -1.]]))
    c["mooresbay_MB1_direct"] = dict(ice="mooresbay_simple", att="MB1", n_refl=0, n_freq=25, X1=V, X2=A,
                                     freqs=np.fft.rfftfreq(256, 0.5), fmax=0.6)
    V, A = pairs(cylinder(8, 10, 800., -550.), np.array([[3, 3, -5.]]))
    c["mooresbay_GL3"] = dict(ice="mooresbay_simple", att="GL3", n_refl=1, n_freq=10, X1=V, X2=A,
                              freqs=np.fft.rfftfreq(128, 0.5), fmax=None)
    return c


def _work(args):
    name, cfg, lo, hi = args
    import ref_harness as rh
    r = rh.make_tracer(cfg["ice"], attenuation_model=cfg["att"] or "SP1", n_freq=cfg["n_freq"], n_reflections=cfg["n_refl"])
    with_att = cfg["att"] is not None
    t = time.time()
    out = rh.trace_pairs(r, cfg["X1"][lo:hi], cfg["X2"][lo:hi], freqs=cfg.get("freqs"), max_detector_freq=cfg.get("fmax"),
                         with_attenuation=with_att, tight_attenuation=with_att)
    out["_seconds"] = time.time() - t
    return name, lo, out


def _arbiter(args):
    cfg, i = args
    import ref_harness as rh
    r = rh.make_tracer(cfg["ice"], n_reflections=cfg["n_refl"])
    res = []
    for md in range(1 + 2 * cfg["n_refl"]):
        refl = 0 if md == 0 else (md - 1) // 2 + 1
        case = 1 if md == 0 else (md - 1) % 2 + 1
        res.extend(list(rh.arbiter_roots(r, cfg["X1"][i], cfg["X2"][i], refl, case, n_scan=28001, lo=-12.0, hi=16.0)))
    return i, res


def main(names):
    all_cases = cases()
    names = names or list(all_cases)
    pool = Pool(8)
    for name in names:
        cfg = all_cases[name]
        N = len(cfg["X1"])
        chunk = max(1, N // 32)
        jobs = [(name, cfg, lo, min(N, lo + chunk)) for lo in range(0, N, chunk)]
        t0 = time.time()
        parts = sorted(pool.map(_work, jobs), key=lambda p: p[1])
        out = {}
        for k in parts[0][2]:
            if k == "_seconds":
                out["ref_cpu_seconds"] = np.array(sum(p[2][k] for p in parts))
            else:
                out[k] = np.concatenate([p[2][k] for p in parts], axis=0)
        # F6 arbiter where the oracle disagrees on the count
        from oracle.oracle import Oracle
        o = Oracle(cfg["ice"], n_reflections=cfg["n_refl"])
        oo = o.trace(cfg["X1"], cfg["X2"])
        bad = np.nonzero(oo["n_sol"] != out["n_sol"])[0]
        arb = np.full((N, out["C0"].shape[1]), np.nan)
        arb_n = np.full(N, -1, np.int32)
        for i, roots in pool.map(_arbiter, [(cfg, int(i)) for i in bad]):
            arb_n[i] = len(roots)
            arb[i, :min(len(roots), arb.shape[1])] = roots[:arb.shape[1]]
        out["arbiter_C0"], out["arbiter_n"] = arb, arb_n
        meta = dict(ice=cfg["ice"], attenuation_model=cfg["att"] or "", n_reflections=cfg["n_refl"], n_freq=cfg["n_freq"],
                    max_detector_freq=np.nan if cfg.get("fmax") is None else cfg["fmax"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), X1=cfg["X1"], X2=cfg["X2"],
                            frequencies=cfg.get("freqs", np.zeros(0)), **meta, **out)
        print(f"{name}: N={N} solutions={int(out['n_sol'].sum())} count-mismatch-vs-oracle={len(bad)} "
              f"ref_cpu={float(out['ref_cpu_seconds']):.1f}s wall={time.time() - t0:.1f}s", flush=True)
    # the reference's own goldens (data): T05 / T06
    for fn in ("reference_C0.pkl", "reference_C0_MooresBay.pkl"):
        src = os.path.join("/root/reference/NuRadioMC/test/SignalProp", fn)
        if os.path.exists(src):
            with open(src, "rb") as f:
                arr = pickle.load(f, encoding="latin1")
            np.save(os.path.join(HERE, fn.replace(".pkl", ".npy")), np.asarray(arr, float))


def make_effects_golden():
    """fixtures of ray_tracing.apply_propagation_effects (analyticraytracing.py:2937-3033): random spectra in, spectra out"""
    import ref_harness as rh
    rng = np.random.default_rng(77)
    out = {}
    for tag, ice, att, n_refl, V, A in (
            ("sp", "southpole_2015", "SP1", 0, cylinder(41, 10, 2500., -2000.), np.array([[0, 0, -150.]])),
            ("mb", "mooresbay_simple", "MB1", 1, cylinder(42, 5, 600., -400.), np.array([[3, 3, -5.]]))):
        r = rh.make_tracer(ice, attenuation_model=att, n_freq=12, n_reflections=n_refl)
        S = r.get_number_of_raytracing_solutions()
        n_samples, sr = 128, 2.0
        ff = np.fft.rfftfreq(n_samples, 1. / sr)
        spec_in = rng.normal(size=(len(V), S, 3, len(ff))) + 1j * rng.normal(size=(len(V), S, 3, len(ff)))
        spec_out = np.full_like(spec_in, np.nan)
        r_theta = np.full((len(V), S), np.nan, complex)
        r_phi = np.full((len(V), S), np.nan, complex)
        n_sol = np.zeros(len(V), np.int32)
        for i in range(len(V)):
            o, rt, rp = rh.apply_effects(r, V[i], A[0], spec_in[i], ff, sr)
            n_sol[i] = len(o)
            if len(o) == 0:
                continue
            spec_out[i, :len(o)] = o
            r_theta[i, :len(o)], r_phi[i, :len(o)] = rt, rp
        out.update({f"{tag}_X1": V, f"{tag}_X2": np.repeat(A, len(V), 0), f"{tag}_frequencies": ff, f"{tag}_spec_in": spec_in,
                    f"{tag}_spec_out": spec_out, f"{tag}_r_theta": r_theta, f"{tag}_r_phi": r_phi, f"{tag}_n_sol": n_sol})
    # Fresnel coefficients straight from the reference's helpers (NuRadioReco/utilities/geometryUtilities.py:211-263)
    from NuRadioReco.utilities import geometryUtilities as gu
    ang = np.linspace(0.01, np.pi / 2 - 0.01, 60)
    n1 = 1.3587
    out["fresnel_angle"], out["fresnel_n1"] = ang, np.array(n1)
    out["fresnel_r_p"] = np.array([gu.get_fresnel_r_p(a, n_2=1., n_1=n1) for a in ang], complex)
    out["fresnel_r_s"] = np.array([gu.get_fresnel_r_s(a, n_2=1., n_1=n1) for a in ang], complex)
    np.savez_compressed(os.path.join(HERE, "propagation_effects.npz"), **out)
    print("propagation_effects: solutions", int(out["sp_n_sol"].sum()), int(out["mb_n_sol"].sum()))


def _focusing_work(args):
    ice, n_refl, X1, X2, limit = args
    import ref_harness as rh
    r = rh.make_tracer(ice, n_reflections=n_refl)
    # get_focusing builds its second tracer without compile_numba=False (analyticraytracing.py:2834-2835), which would switch
    # the whole module to the numba functions; give it the plain one it would otherwise create
    r._r1 = rh.make_tracer(ice, n_reflections=n_refl)
    S = r.get_number_of_raytracing_solutions()
    foc = np.full((len(X1), S), np.nan)
    n_sol = np.zeros(len(X1), np.int32)
    n_sol_displaced = np.zeros(len(X1), np.int32)
    C0 = np.full((len(X1), S), np.nan)
    for i in range(len(X1)):
        r.set_start_and_end_point(X1[i], X2[i])
        r.find_solutions()
        n_sol[i] = r.get_number_of_solutions()
        for iS in range(n_sol[i]):
            C0[i, iS] = r.get_results()[iS]["C0"]
            foc[i, iS] = r.get_focusing(iS, limit=limit)
        n_sol_displaced[i] = r._r1.get_number_of_solutions() if n_sol[i] else 0
    return foc, n_sol, n_sol_displaced, C0


def make_focusing_golden():
    """fixtures of ray_tracing.get_focusing (analyticraytracing.py:2778-2888, numerical branch, dz = -1 cm): the reference's
    own values; `n_sol_displaced` is the solution count of its second trace (receiver moved by dz) -- where it differs from
    n_sol the reference returns 1 or pairs unrelated rays (:2840,:2863)."""
    out = {}
    pool = Pool(8)
    V, A = pairs(cylinder(51, 150, 4000., -2700.), np.array([[10, 10, -190.], [0, 0, -2.]]))
    Vs = cylinder(52, 60, 2000., -300.)       # emitters above the receiver: the reference swaps the points (:2072-2077)
    X1 = np.concatenate([V, Vs])
    X2 = np.concatenate([A, np.repeat([[0, 0, -800.]], len(Vs), 0)])
    Vm, Am = pairs(cylinder(53, 60, 1000., -500.), np.array([[3, 3, -5.], [-3, 0, -1.]]))
    for tag, ice, n_refl, P1, P2, limit in (("sp", "southpole_2015", 0, X1, X2, 2.0), ("sp_nolimit", "southpole_2015", 0, X1[:120], X2[:120], 1e9),
                                            ("mb", "mooresbay_simple", 1, Vm, Am, 2.0)):
        N = len(P1)
        chunk = max(1, N // 32)
        parts = pool.map(_focusing_work, [(ice, n_refl, P1[lo:lo + chunk], P2[lo:lo + chunk], limit) for lo in range(0, N, chunk)])
        foc, n_sol, n_disp, C0 = (np.concatenate([p[k] for p in parts]) for k in range(4))
        out.update({f"{tag}_X1": P1, f"{tag}_X2": P2, f"{tag}_focusing": foc, f"{tag}_n_sol": n_sol, f"{tag}_n_sol_displaced": n_disp,
                    f"{tag}_C0": C0, f"{tag}_limit": np.array(limit), f"{tag}_ice": ice, f"{tag}_n_reflections": n_refl})
        print(f"focusing {tag}: N={N} solutions={int(n_sol.sum())} displaced-count-differs={int((n_sol != n_disp).sum())}", flush=True)
    np.savez_compressed(os.path.join(HERE, "focusing.npz"), **out)


def _n3_work(args):
    """the ray-tracing part of the reference's per-(shower, channel) loop (NuRadioMC/simulation/simulation.py:155-210) with the
    reference propagator, and the per-station HDF5 datasets as output_writer_hdf5.py:267-294 assembles them"""
    ice, n_refl, V, axes, A, delta_C_cut = args
    import ref_harness as rh
    sys.path.insert(0, os.path.join(REPO, "oracle", "pyref", "stubs"))
    from radiotools import helper as hp
    r = rh.make_tracer(ice, n_reflections=n_refl)
    nS = r.get_number_of_raytracing_solutions()
    nSh, nCh = len(V), len(A)
    ds = {k: np.full((nSh, nCh, nS), np.nan) for k in ("travel_times", "travel_distances", "ray_tracing_C0", "ray_tracing_C1",
                                                        "ray_tracing_reflection", "ray_tracing_reflection_case",
                                                        "ray_tracing_solution_type", "focusing_factor", "viewing_angles")}
    ds["launch_vectors"] = np.full((nSh, nCh, nS, 3), np.nan)
    ds["receive_vectors"] = np.full((nSh, nCh, nS, 3), np.nan)
    n_sol = np.zeros((nSh, nCh), np.int32)
    medium = r._medium
    for iSh in range(nSh):
        x1 = V[iSh]
        shower_direction = -1 * axes[iSh]                                   # simulation.py:175
        cherenkov_angle = np.arccos(1. / medium.get_index_of_refraction(x1))
        for iCh in range(nCh):
            r.set_start_and_end_point(x1, A[iCh])
            r.find_solutions()
            if not r.has_solution():
                continue
            n = r.get_number_of_solutions()
            n_sol[iSh, iCh] = n
            viewing = np.array([np.arccos(np.clip(np.dot(shower_direction, r.get_launch_vector(iS)) / np.linalg.norm(shower_direction), -1, 1))
                                for iS in range(n)])                        # hp.get_angle (:191)
            delta_Cs = viewing - cherenkov_angle
            if min(np.abs(delta_Cs)) > delta_C_cut:                         # :195-197
                continue
            for iS in range(n):
                if np.abs(delta_Cs[iS]) > delta_C_cut:                      # :204-206
                    continue
                ds["travel_distances"][iSh, iCh, iS] = r.get_path_length(iS)
                ds["travel_times"][iSh, iCh, iS] = r.get_travel_time(iS)
                ds["viewing_angles"][iSh, iCh, iS] = viewing[iS]
                ds["launch_vectors"][iSh, iCh, iS] = r.get_launch_vector(iS)
                zen, az = hp.cartesian_to_spherical(*r.get_receive_vector(iS))      # simulation.py stores zenith / azimuth ...
                ds["receive_vectors"][iSh, iCh, iS] = hp.spherical_to_cartesian(zen, az)   # ... output_writer_hdf5.py:289-290
                for key, value in r.get_raytracing_output(iS).items():      # efp.raytracing_solution (output_writer_hdf5.py:283-286)
                    ds[key][iSh, iCh, iS] = value
    ds["n_sol"] = n_sol
    return ds


def make_n3_golden():
    """fixture of the caller either side of the path: the reference's scalar loop + HDF5 dataset layout"""
    rng = np.random.default_rng(66)
    out = {}
    pool = Pool(8)
    for tag, ice, n_refl, V, A in (
            ("sp", "southpole_2015", 0, cylinder(61, 48, 3000., -2500.), np.array([[0, 0, -150.], [10, 10, -190.], [-8, 3, -60.], [3, -9, -2.]])),
            ("mb", "mooresbay_simple", 1, cylinder(62, 24, 800., -500.), np.array([[3, 3, -5.], [-3, 0, -1.], [0, 3, -1.]]))):
        axes = rng.normal(size=(len(V), 3))
        axes /= np.linalg.norm(axes, axis=1)[:, None]
        cut = 40. * np.pi / 180.                                            # config_default.yaml speedup.delta_C_cut
        chunk = 6
        parts = pool.map(_n3_work, [(ice, n_refl, V[lo:lo + chunk], axes[lo:lo + chunk], A, cut) for lo in range(0, len(V), chunk)])
        for k in parts[0]:
            out[f"{tag}_{k}"] = np.concatenate([p[k] for p in parts], axis=0)
        out.update({f"{tag}_vertices": V, f"{tag}_shower_axes": axes, f"{tag}_channels": A, f"{tag}_delta_C_cut": np.array(cut),
                    f"{tag}_ice": ice, f"{tag}_n_reflections": n_refl})
        kept = int(np.isfinite(out[f"{tag}_travel_times"]).sum())
        print(f"n3 {tag}: {len(V)} showers x {len(A)} channels, solutions {int(out[f'{tag}_n_sol'].sum())}, kept after the viewing-angle cut {kept}", flush=True)
    np.savez_compressed(os.path.join(HERE, "simulation_datasets.npz"), **out)


if __name__ == "__main__":
    if sys.argv[1:] == ["n3"]:
        make_n3_golden()
    elif sys.argv[1:] == ["effects"]:
        make_effects_golden()
    elif sys.argv[1:] == ["focusing"]:
        make_focusing_golden()
    else:
        main(sys.argv[1:])
