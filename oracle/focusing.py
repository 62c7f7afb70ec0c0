"""
TEST INFRASTRUCTURE -- numpy restatement of ray_tracing.get_focusing (NuRadioMC/SignalProp/analyticraytracing.py:2778-2888,
numerical branch; the analytic branch of the reference raises AttributeError at :831 and cannot be an oracle) on top of the
CPU oracle's traces.  Pinned by tests/golden/focusing.npz (the reference's own get_focusing, tests/golden/make_golden.py
focusing).  Only tests/ may import this module.
"""
import numpy as np


def get_focusing(oracle, X1, X2, dz=-0.01, limit=2.0, trace=None):
    """
    X1: emitters (N,3); X2: receivers (N,3).  Returns (focusing (N,S), comparable (N,S) bool).
    As the reference: trace the pair again with the RECEIVER moved by dz (:2812-2839) and difference the launch zenith angles
    of the solutions with the same index iS (:2840-2848); azimuthal term (:2850-2855); limit (:2865-2867); index ratio
    (:2869-2876).  `comparable` is False where the displaced trace has a different number of solutions or a different mode
    in slot iS: the reference then returns 1 or compares unrelated rays (:2857-2862), an artefact no kernel should copy.
    """
    X1 = np.atleast_2d(np.asarray(X1, float))
    X2 = np.atleast_2d(np.asarray(X2, float))
    if len(X2) == 1 and len(X1) > 1:
        X2 = np.repeat(X2, len(X1), 0)
    a = trace if trace is not None else oracle.trace(X1, X2)
    X2b = X2.copy()
    X2b[:, 2] += dz
    b = oracle.trace(X1, X2b)
    cfg = oracle.cfg
    n_of = lambda z: cfg.n_ice - cfg.delta_n * np.exp(z / cfg.z_0)   # medium_base.py:254-277 (in ice)
    S = a["C0"].shape[1]
    with np.errstate(invalid="ignore", divide="ignore"):
        rec = -a["receive"]
        rec_ang = np.arccos(rec[..., 2] / np.linalg.norm(rec, axis=-1))                      # :2806-2808
        lau_ang = np.arccos(a["launch"][..., 2] / np.linalg.norm(a["launch"], axis=-1))      # :2809-2810
        lau_ang1 = np.arccos(b["launch"][..., 2] / np.linalg.norm(b["launch"], axis=-1))     # :2841-2842
        distance = a["path_length"]
        f = np.sqrt(distance / np.sin(rec_ang) * np.abs((lau_ang1 - lau_ang) / dz))           # :2848
        radius = np.linalg.norm(X2 - X1, axis=1)[:, None]
        sin_theta = np.linalg.norm((X2 - X1)[:, :2], axis=1)[:, None] / radius
        f = f * np.sqrt(distance * np.sin(lau_ang) / (radius * sin_theta))                    # :2851-2855
    slot = np.arange(S)[None, :]
    missing = (slot < a["n_sol"][:, None]) & (slot >= b["n_sol"][:, None])
    f[missing] = 1.0                                                                          # :2863-2864
    f = np.where(f > limit, limit, f)                                                         # :2866-2868
    f = f * np.sqrt(n_of(X1[:, 2]) / n_of(X2[:, 2]))[:, None]                                 # :2870-2876
    f[slot >= a["n_sol"][:, None]] = np.nan
    comparable = (slot < a["n_sol"][:, None]) & (a["n_sol"] == b["n_sol"])[:, None] & (a["reflection"] == b["reflection"]) \
        & (a["reflection_case"] == b["reflection_case"]) & (a["type"] == b["type"])
    return f, comparable
