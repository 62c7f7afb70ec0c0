"""
Batched hooks for the NuRadioMC simulation loop.

The reference traces one (shower, channel) pair at a time inside `calculate_sim_efield`
(NuRadioMC/simulation/simulation.py:155-210): set_start_and_end_point, find_solutions, viewing-angle cut, path length,
travel time, then `apply_propagation_effects`.  Here all pairs of an event group are traced in ONE device pass ahead of
that loop; the loop's scalar calls then hit the propagator's cache (`ray_tracing.prepare_batch`).  The per-station
ray-tracing datasets of the HDF5 output (NuRadioMC/simulation/output_writer_hdf5.py:267-294) come straight from the
batch result, and `solutions_from_datasets` + `ray_tracing.set_solution` reload them (speedup.redo_raytracing = False,
the TODO at simulation.py:178-180).
"""
import numpy as np

RT_DATASETS = ("ray_tracing_C0", "ray_tracing_C1", "ray_tracing_reflection", "ray_tracing_reflection_case",
               "ray_tracing_solution_type", "focusing_factor")   # ray_tracing.get_output_parameters (analyticraytracing.py:2895-2903)


def pretrace_event_group(propagator, vertices, channel_positions, shower_axes=None, delta_C_cut=None, frequency=None,
                         max_detector_freq=None, attenuation="dense"):
    """
    Trace every (shower, channel) pair of an event group in one pass and arm the propagator's cache, so that the unchanged
    scalar loop of simulation.py:155-210 does no further ray tracing.

    vertices: (nSh, 3) shower vertices; channel_positions: (nCh, 3) absolute antenna positions;
    shower_axes: (nSh, 3) shower axes as stored on the showers (the loop uses the propagation direction, -axis,
    simulation.py:175); delta_C_cut: config['speedup']['delta_C_cut'] [rad].
    Returns the BatchResult (shower-major: pair = i_shower * nCh + i_channel).  With config['propagation']['focusing'] on,
    the result also holds "focusing_factor" (N, S) (`ray_tracing.focusing_batch`), so that the loop's get_focusing /
    get_raytracing_output calls (analyticraytracing.py:2913-2916, :3012-3015) launch nothing either; pass it to
    `raytracing_datasets(..., focusing=res["focusing_factor"])` for the HDF5 output.
    """
    kw = {}
    if shower_axes is not None:
        kw["shower_axis"] = -np.asarray(shower_axes, dtype=np.float64).reshape(-1, 3)
        kw["delta_C_cut"] = delta_C_cut
    return propagator.prepare_batch(vertices, channel_positions, outer=True, frequency=frequency,
                                    max_detector_freq=max_detector_freq, attenuation=attenuation, **kw)


def cherenkov_mask(result, medium, vertices, n_channels, delta_C_cut):
    """(N, S) bool: solutions the reference's viewing-angle cut keeps (simulation.py:187-208).  `result` must come from a
    trace with shower axes (it then holds "viewing_angle")."""
    n_index = np.asarray(medium.get_index_of_refraction(np.asarray(vertices, dtype=np.float64).reshape(-1, 3)))
    cherenkov = np.repeat(np.arccos(1. / n_index), n_channels)
    with np.errstate(invalid="ignore"):
        return np.abs(result["viewing_angle"] - cherenkov[:, None]) <= delta_C_cut


def raytracing_datasets(result, n_showers, n_channels, keep=None, focusing=None):
    """
    The per-station ray-tracing datasets of the HDF5 output file, shapes as output_writer_hdf5.py:267-294 builds them
    per shower and stacks them: travel_times, travel_distances (nSh, nCh, nS); launch_vectors, receive_vectors
    (nSh, nCh, nS, 3); the propagator's output parameters (nSh, nCh, nS).  Entries without solution are NaN
    (:272-275).  keep: optional (N, S) mask (e.g. `cherenkov_mask`): solutions the simulation skipped are NaN as well.
    focusing: optional (N, S) factors of `ray_tracing.focusing_batch` (config['propagation']['focusing']); 1 otherwise.
    """
    if getattr(result, "compact", False):
        raise ValueError("raytracing_datasets needs the padded layout (compact=False)")
    S = result["C0"].shape[1]
    filled = np.arange(S)[None, :] < np.asarray(result["n_sol"])[:, None]
    if keep is not None:
        filled &= np.asarray(keep, dtype=bool)

    def grid(a, fill=np.nan, dtype=np.float64):
        a = np.array(a, dtype=dtype)
        a[~filled] = fill
        return a.reshape((n_showers, n_channels) + a.shape[1:])
    ds = {
        "travel_times": grid(result["travel_time"]),
        "travel_distances": grid(result["path_length"]),
        "launch_vectors": grid(result["launch_vector"]),
        "receive_vectors": grid(result["receive_vector"]),
        "ray_tracing_C0": grid(result["C0"]),
        "ray_tracing_C1": grid(result["C1"]),
        "ray_tracing_reflection": grid(result["reflection"]),
        "ray_tracing_reflection_case": grid(result["reflection_case"]),
        "ray_tracing_solution_type": grid(result["solution_type"]),
        # get_raytracing_output (analyticraytracing.py:2905-2935): 1 unless focusing is enabled
        "focusing_factor": grid(np.ones_like(result["C0"]) if focusing is None else focusing),
    }
    return ds


def solutions_from_datasets(datasets, i_shower, i_channel):
    """the dict `ray_tracing.set_solution` expects (analyticraytracing.py:2092-2116) for one (shower, channel) entry"""
    return {k: datasets[k][i_shower, i_channel] for k in RT_DATASETS if k in datasets}


def write_station_group(h5group, datasets):
    """write the datasets into an (h5py) group, one dataset per key, as output_writer_hdf5.py does for a station group"""
    for k, v in datasets.items():
        if k in h5group:
            del h5group[k]
        h5group[k] = v
