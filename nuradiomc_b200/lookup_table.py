"""
Travel-time lookup tables for the vertex reconstructors: the bulk consumer
NuRadioReco/modules/neutrinoVertexReconstructor/create_lookup_table.py:62-107 traces ~3.7e6 (r, z) grid points to one
antenna depth, one scalar call at a time.  Here the whole grid is one batched pass.

Same dictionary layout as the reference script (header + 'antenna_<depth>' -> 'direct' / 'refracted' / 'reflected'
arrays of shape (len(x_pos), len(z_pos)), 0 where no solution of that type exists); travel times are the analytic ones
(the script calls the numerical ray_tracing_2D.get_travel_time, which agrees to the 1e-12 level, T04 of the reference).
The horizontal distance is used as a distance (the script's literal argument order puts the receiver to the LEFT of the
emitter, for which the reference's Python 2-D solver finds no solution at all).
"""
import pickle

import numpy as np


def create_lookup_table(antenna_depth, r_min=10., r_max=5000., z_min=3000., z_max=50., d_r=2., d_z=2.,
                        ice_model="greenland_simple", device=0, propagator=None):
    from nuradiomc_b200.SignalProp import propagation
    from nuradiomc_b200.utilities import medium
    x_pos = np.arange(r_min, r_max, d_r)
    z_pos = np.arange(-z_min, -z_max, d_z)
    if propagator is None:
        propagator = propagation.get_propagation_module("analytic")(medium.get_ice_model(ice_model), device=device)
    V = np.zeros((len(x_pos) * len(z_pos), 3))
    V[:, 0] = np.repeat(x_pos, len(z_pos))
    V[:, 2] = np.tile(z_pos, len(x_pos))
    res = propagator.trace_batch(V, np.array([[0., 0., -1. * antenna_depth]]), outputs=("n_sol", "solution_type", "travel_time"))
    tables = {name: np.zeros(len(V)) for name in ("direct", "refracted", "reflected")}
    S = res["travel_time"].shape[1]
    for s in range(S):      # in solution order: a later solution of the same type overwrites an earlier one, as the script's loop
        ok = s < res["n_sol"]
        for t, name in ((1, "direct"), (2, "refracted"), (3, "reflected")):
            m = ok & (res["solution_type"][:, s] == t)
            tables[name][m] = res["travel_time"][m, s]
    name = "antenna_{}".format(antenna_depth)
    return {
        "header": {"x_min": r_min, "x_max": r_max, "d_x": d_r, "z_min": -z_min, "z_max": -z_max, "d_z": d_z},
        name: {k: v.reshape(len(x_pos), len(z_pos)) for k, v in tables.items()},
    }


def write_lookup_table(table, output_path=".", antenna_depth=None):
    """pickle file named as the reference script names it (create_lookup_table.py:106)"""
    if antenna_depth is None:
        antenna_depth = float([k for k in table if k != "header"][0].split("_", 1)[1])
    fn = "{}/lookup_table_greenland_{:.0f}.p".format(output_path, antenna_depth)
    with open(fn, "wb") as f:
        pickle.dump(table, f)
    return fn
