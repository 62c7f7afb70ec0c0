"""
Attenuation-length models of the reference (NuRadioMC/utilities/attenuation.py:14,145-262) evaluated on the GPU.

`model_to_int` is the integer code the reference already uses to talk to its native code (attenuation.py:14).
`get_attenuation_length(z, frequency, model)` has the reference's signature and runs the same device functions the
ray-tracing kernel uses (nrmc_rt_attenuation_length in include/nrmc_rt.h; replaces wrapper.pyx:32-33).
"""
import ctypes as C
import os

import numpy as np

from nuradiomc_b200 import _lib

model_to_int = {"SP1": 1, "GL1": 2, "MB1": 3, "GL2": 4, "GL3": 5}

_gl3 = None


def gl3_parameters():
    """300 x (depth [m], slope, offset) table of the GL3 model (data file of the reference, attenuation.py:16-19)."""
    global _gl3
    if _gl3 is None:
        _gl3 = np.ascontiguousarray(np.genfromtxt(os.path.join(os.path.dirname(__file__), "data", "GL3_params.csv"),
                                                  delimiter=","), dtype=np.float64)
    return _gl3


_handles = {}


def get_attenuation_length(z, frequency, model):
    if model not in model_to_int:
        raise NotImplementedError("attenuation model {} is not implemented.".format(model))
    from nuradiomc_b200.SignalProp.analyticraytracing import _make_handle
    if model not in _handles:
        _handles[model] = _make_handle(1.78, 0.4, 70.0, None, model_to_int[model], 0, 100, 0)
    h = _handles[model]
    scalar = np.ndim(z) == 0 and np.ndim(frequency) == 0
    zz, ff = np.broadcast_arrays(np.asarray(z, dtype=np.float64), np.asarray(frequency, dtype=np.float64))
    zz = np.ascontiguousarray(zz).ravel()
    ff = np.ascontiguousarray(ff).ravel()
    out = np.empty_like(zz)
    lib = _lib.load()
    _lib.check(lib.nrmc_rt_attenuation_length(h.ptr, zz.ctypes.data, ff.ctypes.data, zz.size, out.ctypes.data), h.ptr,
               "attenuation_length")
    if scalar:
        return float(out[0])
    return out.reshape(np.broadcast(np.asarray(z), np.asarray(frequency)).shape)
