"""ctypes binding of the C ABI in include/nrmc_rt.h.  Fails loudly when the CUDA library is missing."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NRMC_RT_LIB", os.path.join(HERE, "libnrmc_rt.so"))   # override: A/B builds of the same source

NRMC_OK = 0
ERRORS = {-1: "invalid argument", -2: "CUDA error", -3: "no CUDA device", -4: "unsupported configuration",
          -5: "no frequencies set", -6: "row capacity of the compact output arrays exceeded"}
MEMORY_HOST, MEMORY_DEVICE = 0, 1
PAIR_IN_AIR, PAIR_BELOW_REFLECTOR, PAIR_NONFINITE = 1, 2, 4


class Config(C.Structure):
    _fields_ = [("n_ice", C.c_double), ("delta_n", C.c_double), ("z_0", C.c_double), ("reflection_z", C.c_double),
                ("attenuation_model", C.c_int32), ("n_reflections", C.c_int32), ("n_frequencies_integration", C.c_int32),
                ("device", C.c_int32), ("gl3_table", C.c_void_p), ("gl3_rows", C.c_int32), ("reserved", C.c_int32)]


class Input(C.Structure):
    _fields_ = [("n_vertices", C.c_int64), ("vx", C.c_void_p), ("vy", C.c_void_p), ("vz", C.c_void_p),
                ("n_antennas", C.c_int64), ("ax", C.c_void_p), ("ay", C.c_void_p), ("az", C.c_void_p),
                ("outer", C.c_int32), ("memory", C.c_int32), ("sx", C.c_void_p), ("sy", C.c_void_p), ("sz", C.c_void_p),
                ("delta_C_cut", C.c_double)]


OUTPUT_FIELDS = ("n_sol", "status", "solution_type", "reflection", "reflection_case", "C0", "C1", "path_length",
                 "travel_time", "launch_vector", "receive_vector", "reflection_angle", "attenuation_sparse", "attenuation",
                 "viewing_angle")


class Output(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in OUTPUT_FIELDS] + [("compact", C.c_int32), ("reserved", C.c_int32),
                                                          ("sol_offset", C.c_void_p), ("row_capacity", C.c_int64),
                                                          ("row_base", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("n_pairs", C.c_int64), ("n_solutions", C.c_int64), ("ms_solve", C.c_float), ("ms_attenuation", C.c_float),
                ("ms_total", C.c_float), ("n_launches", C.c_int32), ("n_chunks", C.c_int32), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("ms_kernel", C.c_float * 6)]


class Effects(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_freq", C.c_int32), ("reserved", C.c_int32), ("spectrum", C.c_void_p),
                ("attenuation", C.c_void_p), ("attenuation_sparse", C.c_void_p), ("reflection_angle", C.c_void_p),
                ("reflection", C.c_void_p), ("reflection_coefficient", C.c_double), ("reflection_phase_shift", C.c_double),
                ("r_theta", C.c_void_p), ("r_phi", C.c_void_p), ("focusing", C.c_void_p)]


class Focusing(C.Structure):
    _fields_ = [("n_sol", C.c_void_p), ("C0", C.c_void_p), ("reflection", C.c_void_p), ("reflection_case", C.c_void_p),
                ("path_length", C.c_void_p), ("sol_offset", C.c_void_p), ("limit", C.c_double), ("focusing", C.c_void_p)]


EXPORTS = ("nrmc_rt_create", "nrmc_rt_destroy", "nrmc_rt_last_error", "nrmc_rt_max_solutions", "nrmc_rt_set_frequencies",
           "nrmc_rt_get_sparse_frequencies", "nrmc_rt_trace", "nrmc_rt_set_chunk_pairs", "nrmc_rt_host_alloc", "nrmc_rt_host_free",
           "nrmc_rt_peer_alloc", "nrmc_rt_peer_open", "nrmc_rt_peer_close", "nrmc_rt_peer_free", "nrmc_rt_copy_async",
           "nrmc_rt_attenuation_length", "nrmc_rt_apply_propagation_effects", "nrmc_rt_focusing_factor", "nrmc_rt_measure_fp64_peak", "nrmc_rt_device_count", "nrmc_rt_version")

_lib = None


def load():
    """Load libnrmc_rt.so.  No fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). nuradiomc_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.nrmc_rt_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    lib.nrmc_rt_create.restype = C.c_int
    lib.nrmc_rt_destroy.argtypes = [C.c_void_p]
    lib.nrmc_rt_destroy.restype = None
    lib.nrmc_rt_last_error.argtypes = [C.c_void_p]
    lib.nrmc_rt_last_error.restype = C.c_char_p
    lib.nrmc_rt_max_solutions.argtypes = [C.c_void_p]
    lib.nrmc_rt_max_solutions.restype = C.c_int
    lib.nrmc_rt_set_frequencies.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double]
    lib.nrmc_rt_set_frequencies.restype = C.c_int
    lib.nrmc_rt_get_sparse_frequencies.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    lib.nrmc_rt_get_sparse_frequencies.restype = C.c_int
    lib.nrmc_rt_trace.argtypes = [C.c_void_p, C.POINTER(Input), C.POINTER(Output), C.c_void_p, C.POINTER(Stats)]
    lib.nrmc_rt_trace.restype = C.c_int
    lib.nrmc_rt_set_chunk_pairs.argtypes = [C.c_void_p, C.c_int64]
    lib.nrmc_rt_set_chunk_pairs.restype = C.c_int
    lib.nrmc_rt_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_uint64]
    lib.nrmc_rt_host_alloc.restype = C.c_int
    lib.nrmc_rt_host_free.argtypes = [C.c_void_p]
    lib.nrmc_rt_host_free.restype = C.c_int
    lib.nrmc_rt_peer_alloc.argtypes = [C.c_int32, C.c_uint64, C.POINTER(C.c_void_p), C.c_char_p]
    lib.nrmc_rt_peer_alloc.restype = C.c_int
    lib.nrmc_rt_peer_open.argtypes = [C.c_int32, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.nrmc_rt_peer_open.restype = C.c_int
    lib.nrmc_rt_peer_close.argtypes = [C.c_void_p]
    lib.nrmc_rt_peer_close.restype = C.c_int
    lib.nrmc_rt_peer_free.argtypes = [C.c_void_p]
    lib.nrmc_rt_peer_free.restype = C.c_int
    lib.nrmc_rt_copy_async.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.nrmc_rt_copy_async.restype = C.c_int
    lib.nrmc_rt_attenuation_length.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.nrmc_rt_attenuation_length.restype = C.c_int
    lib.nrmc_rt_apply_propagation_effects.argtypes = [C.c_void_p, C.POINTER(Effects), C.c_void_p]
    lib.nrmc_rt_apply_propagation_effects.restype = C.c_int
    lib.nrmc_rt_focusing_factor.argtypes = [C.c_void_p, C.POINTER(Input), C.POINTER(Focusing), C.c_void_p]
    lib.nrmc_rt_focusing_factor.restype = C.c_int
    lib.nrmc_rt_measure_fp64_peak.argtypes = [C.c_int32, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.nrmc_rt_measure_fp64_peak.restype = C.c_int
    lib.nrmc_rt_device_count.argtypes = []
    lib.nrmc_rt_device_count.restype = C.c_int
    lib.nrmc_rt_version.argtypes = []
    lib.nrmc_rt_version.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc, handle=None, what=""):
    if rc >= 0:
        return rc
    msg = ERRORS.get(rc, f"error {rc}")
    if handle:
        detail = load().nrmc_rt_last_error(handle)
        if detail:
            msg += ": " + detail.decode()
    raise RuntimeError(f"nrmc_rt {what} failed: {msg}")
