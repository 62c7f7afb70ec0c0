"""
TEST INFRASTRUCTURE — ctypes front end of the CPU oracle (oracle/nrmc_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
the product (nuradiomc_b200) never does.  See the header of nrmc_oracle.c for what is restated and how it is pinned.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libnrmc_oracle.so")

# ice-model constants, NuRadioMC/utilities/medium.py:57-154 (n_ice, delta_n, z_0 [m], reflective layer z [m] or None)
ICE_MODELS = {
    "southpole_simple": (1.78, 0.426, 71.0, None),
    "southpole_2015": (1.78, 0.423, 77.0, None),
    "ARAsim_southpole": (1.78, 0.43, 75.75757575757576, None),
    "ARA_2022": (1.78, 0.454, 49.5049505, None),
    "mooresbay_simple": (1.78, 0.46, 34.5, -576.0),
    "mooresbay_simple_2": (1.78, 0.481, 37.0, -576.0),
    "greenland_simple": (1.78, 0.51, 37.25, None),
}
MODEL_TO_INT = {None: 0, "SP1": 1, "GL1": 2, "MB1": 3, "GL2": 4, "GL3": 5}  # NuRadioMC/utilities/attenuation.py:14


class _Cfg(C.Structure):
    _fields_ = [("n_ice", C.c_double), ("delta_n", C.c_double), ("z_0", C.c_double), ("reflection", C.c_double),
                ("att_model", C.c_int32), ("n_reflections", C.c_int32), ("n_freq", C.c_int32), ("quad_mode", C.c_int32),
                ("scan_n", C.c_int32), ("scan_lo", C.c_double), ("scan_hi", C.c_double),
                ("gl3", C.c_void_p), ("gl3_rows", C.c_int32), ("pad_", C.c_int32)]


class _Out(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("n_sol", "type", "reflection", "reflection_case", "C0", "C1", "path_length",
                                          "travel_time", "launch", "receive", "reflection_angle", "attenuation",
                                          "attenuation_sparse", "status")]


def build(force=False):
    """Compile oracle/nrmc_oracle.c -> oracle/libnrmc_oracle.so (gcc, see oracle/Makefile)."""
    src = os.path.join(_HERE, "nrmc_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.ora_trace.restype = C.c_int64
        _lib.ora_trace.argtypes = [C.POINTER(_Cfg), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                   C.c_double, C.POINTER(_Out), C.c_int32]
        _lib.ora_sparse_frequencies.restype = C.c_int
        _lib.ora_sparse_frequencies.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        _lib.ora_attenuation_length.restype = C.c_double
        _lib.ora_attenuation_length.argtypes = [C.POINTER(_Cfg), C.c_double, C.c_double]
        _lib.ora_delta_y.restype = C.c_double
        _lib.ora_delta_y.argtypes = [C.POINTER(_Cfg), C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _lib.ora_find_roots_2d.restype = C.c_int
        _lib.ora_find_roots_2d.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib.ora_max_threads.restype = C.c_int
    return _lib


class Oracle:
    def __init__(self, ice="southpole_simple", attenuation_model=None, n_reflections=0, n_freq=100, tight=True,
                 scan=None, gl3_table=None):
        if isinstance(ice, str):
            ice = ICE_MODELS[ice]
        n_ice, delta_n, z_0, refl = ice
        if n_reflections and refl is None:
            n_reflections = 0  # propagation_base_class.py:128-133
        self.cfg = _Cfg()
        self.cfg.n_ice, self.cfg.delta_n, self.cfg.z_0 = n_ice, delta_n, z_0
        self.cfg.reflection = float("nan") if refl is None else refl
        self.cfg.att_model = MODEL_TO_INT[attenuation_model]
        self.cfg.n_reflections = n_reflections
        self.cfg.n_freq = n_freq
        self.cfg.quad_mode = 1 if tight else 0
        if scan is not None:
            self.cfg.scan_n, self.cfg.scan_lo, self.cfg.scan_hi = scan
        self._gl3 = None
        if gl3_table is not None:
            self._gl3 = np.ascontiguousarray(gl3_table, dtype=np.float64)
            self.cfg.gl3 = self._gl3.ctypes.data
            self.cfg.gl3_rows = self._gl3.shape[0]
        self.n_reflections = n_reflections
        self.S = 2 + 4 * n_reflections

    def sparse_frequencies(self, frequency, max_detector_freq=None):
        frequency = np.ascontiguousarray(frequency, np.float64)
        buf = np.zeros(self.cfg.n_freq + self.cfg.n_freq // 2 + 4)
        n = lib().ora_sparse_frequencies(self.cfg.n_freq, frequency.ctypes.data, len(frequency),
                                         float("nan") if max_detector_freq is None else max_detector_freq, buf.ctypes.data)
        return buf[:n].copy()

    def attenuation_length(self, z, f):
        return lib().ora_attenuation_length(C.byref(self.cfg), float(z), float(f))

    def delta_y(self, logC0, x1, x2, reflection=0, reflection_case=1):
        x1 = np.ascontiguousarray(x1, np.float64)
        x2 = np.ascontiguousarray(x2, np.float64)
        return lib().ora_delta_y(C.byref(self.cfg), float(logC0), x1.ctypes.data, x2.ctypes.data, reflection, reflection_case)

    def trace(self, X1, X2, frequency=None, max_detector_freq=None, n_threads=0, dense=True):
        """N pairs X1[N,3], X2[N,3] (X2 may be a single point) -> dict of SoA arrays (NaN/0 padded, S = 2+4*n_refl)."""
        X1 = np.ascontiguousarray(np.atleast_2d(X1), np.float64)
        X2 = np.atleast_2d(np.asarray(X2, np.float64))
        if X2.shape[0] == 1 and X1.shape[0] > 1:
            X2 = np.repeat(X2, X1.shape[0], axis=0)
        X2 = np.ascontiguousarray(X2)
        N, S, K1 = X1.shape[0], self.S, self.n_reflections + 1
        out = {
            "n_sol": np.zeros(N, np.int32), "status": np.zeros(N, np.int32),
            "type": np.zeros((N, S), np.int8), "reflection": np.zeros((N, S), np.int8),
            "reflection_case": np.zeros((N, S), np.int8),
            "C0": np.full((N, S), np.nan), "C1": np.full((N, S), np.nan),
            "path_length": np.full((N, S), np.nan), "travel_time": np.full((N, S), np.nan),
            "launch": np.full((N, S, 3), np.nan), "receive": np.full((N, S, 3), np.nan),
            "reflection_angle": np.full((N, S, K1), np.nan),
        }
        nf = 0
        fptr = None
        if frequency is not None and self.cfg.att_model > 0:
            frequency = np.ascontiguousarray(frequency, np.float64)
            nf = len(frequency)
            fptr = frequency.ctypes.data
            sp = self.sparse_frequencies(frequency, max_detector_freq)
            out["frequencies_sparse"] = sp
            if dense:
                out["attenuation"] = np.full((N, S, nf), np.nan)
            out["attenuation_sparse"] = np.full((N, S, len(sp)), np.nan)
        o = _Out()
        for k, _ in _Out._fields_:
            setattr(o, k, out[k].ctypes.data if k in out else None)
        ne = lib().ora_trace(C.byref(self.cfg), N, X1.ctypes.data, X2.ctypes.data, fptr, nf,
                             float("nan") if max_detector_freq is None else max_detector_freq, C.byref(o), n_threads)
        out["n_evaluations"] = ne
        return out


def max_threads():
    return lib().ora_max_threads()
