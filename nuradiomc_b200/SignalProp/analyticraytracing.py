"""
Analytic ray tracer -- the reference's `ray_tracing` class (NuRadioMC/SignalProp/analyticraytracing.py:1932-3052)
re-implemented on top of the batched sm_100a library (include/nrmc_rt.h).

* Same constructor, same stateful scalar API (set_start_and_end_point / find_solutions / get_* ...), same exceptions.
  A scalar call is a batch of one pair on the GPU (or a lookup in a batch traced ahead with `prepare_batch`).
* New: `trace_batch` / `trace_batch_device` -- every (vertex, antenna) pair in one device pass, SoA results.

There is no CPU implementation behind this class: without libnrmc_rt.so and a CUDA device every compute call raises.
"""
import ctypes as C
import logging

import numpy as np

from nuradiomc_b200 import _lib
from nuradiomc_b200.SignalProp.propagation import solution_types, solution_types_revert  # noqa: F401 (re-export, as the reference)
from nuradiomc_b200.SignalProp.propagation_base_class import ray_tracing_base
from nuradiomc_b200.utilities import attenuation as attenuation_util
from nuradiomc_b200.utilities import units

logger = logging.getLogger("NuRadioMC.analytic_ray_tracing")

speed_of_light = units.speed_of_light
cpp_available = False     # kept for source compatibility with the reference module; the native path here is CUDA
numba_available = False
cuda_available = True     # resolved lazily: the library is loaded on first use


def _stats_dict(st):
    d = {k: getattr(st, k) for k, _ in _lib.Stats._fields_ if k != "ms_kernel"}
    d["ms_kernel"] = dict(zip(("classify", "hump", "roots", "attenuation_main", "attenuation_other"), list(st.ms_kernel)[:5]))
    return d


class _Handle:
    """owns one nrmc_rt_t"""

    def __init__(self, ptr):
        self.ptr = ptr
        self.freq_key = None
        self.sparse = None

    def __del__(self):
        try:
            if self.ptr:
                _lib.load().nrmc_rt_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


def _make_handle(n_ice, delta_n, z_0, reflection_z, att_model_int, n_reflections, n_freq, device):
    lib = _lib.load()
    cfg = _lib.Config()
    cfg.n_ice, cfg.delta_n, cfg.z_0 = float(n_ice), float(delta_n), float(z_0)
    cfg.reflection_z = float("nan") if reflection_z is None else float(reflection_z)
    cfg.attenuation_model = int(att_model_int)
    cfg.n_reflections = int(n_reflections)
    cfg.n_frequencies_integration = int(n_freq)
    cfg.device = int(device)
    keep = None
    if att_model_int == attenuation_util.model_to_int["GL3"]:
        keep = attenuation_util.gl3_parameters()
        cfg.gl3_table = keep.ctypes.data
        cfg.gl3_rows = keep.shape[0]
    ptr = C.c_void_p()
    _lib.check(lib.nrmc_rt_create(C.byref(cfg), C.byref(ptr)), None, "create")
    return _Handle(ptr)


class PinnedArray:
    """numpy view of page-locked host memory (nrmc_rt_host_alloc) for fast host<->device copies"""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        nbytes = max(int(np.prod(self.shape)) * self.dtype.itemsize, 1)
        self._ptr = C.c_void_p()
        _lib.check(_lib.load().nrmc_rt_host_alloc(C.byref(self._ptr), nbytes), None, "host_alloc")
        buf = (C.c_char * nbytes).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def __del__(self):
        try:
            if self._ptr:
                self.array = None
                _lib.load().nrmc_rt_host_free(self._ptr)
                self._ptr = None
        except Exception:
            pass


_OUT_SPECS = {  # name -> (dtype, trailing shape as function of (S, K1, Fs, F))
    "n_sol": (np.int32, lambda S, K1, Fs, F: ()),
    "status": (np.int32, lambda S, K1, Fs, F: ()),
    "solution_type": (np.int8, lambda S, K1, Fs, F: (S,)),
    "reflection": (np.int8, lambda S, K1, Fs, F: (S,)),
    "reflection_case": (np.int8, lambda S, K1, Fs, F: (S,)),
    "C0": (np.float64, lambda S, K1, Fs, F: (S,)),
    "C1": (np.float64, lambda S, K1, Fs, F: (S,)),
    "path_length": (np.float64, lambda S, K1, Fs, F: (S,)),
    "travel_time": (np.float64, lambda S, K1, Fs, F: (S,)),
    "launch_vector": (np.float64, lambda S, K1, Fs, F: (S, 3)),
    "receive_vector": (np.float64, lambda S, K1, Fs, F: (S, 3)),
    "reflection_angle": (np.float64, lambda S, K1, Fs, F: (S, K1)),
    "attenuation_sparse": (np.float64, lambda S, K1, Fs, F: (S, Fs)),
    "attenuation": (np.float64, lambda S, K1, Fs, F: (S, F)),
    "viewing_angle": (np.float64, lambda S, K1, Fs, F: (S,)),
}
DEFAULT_OUTPUTS = ("n_sol", "status", "solution_type", "reflection", "reflection_case", "C0", "C1", "path_length",
                   "travel_time", "launch_vector", "receive_vector", "reflection_angle")


def _pair_lookup(X1, X2, outer):
    """{48-byte key of (x1, x2): pair index} for `prepare_batch`: the key is what `find_solutions` forms from the current points
    (x1.tobytes() + x2.tobytes()).  Built without a Python loop over the pairs (an event group has 1e5 of them); for duplicate
    pairs the last index wins, as in a loop."""
    X1 = np.ascontiguousarray(X1, dtype=np.float64).reshape(-1, 3)
    X2 = np.ascontiguousarray(X2, dtype=np.float64).reshape(-1, 3)
    if outer:
        pairs = np.concatenate([np.repeat(X1, X2.shape[0], axis=0), np.tile(X2, (X1.shape[0], 1))], axis=1)
    else:
        X2b = X2 if X2.shape[0] == X1.shape[0] else np.repeat(X2, X1.shape[0], axis=0)
        pairs = np.concatenate([X1, X2b], axis=1)
    keys = np.ascontiguousarray(pairs).view(np.dtype((np.void, 48))).ravel().tolist()
    return dict(zip(keys, range(len(keys))))


class BatchResult(dict):
    """SoA results of `trace_batch`: arrays keyed by the field names of nrmc_rt_output, plus `.stats`,
    `.frequencies_sparse`."""
    stats = None
    frequencies_sparse = None
    compact = False
    row_base = 0
    freq_key = None
    _keep = None
    _full = None

    def n_rows(self):
        """number of per-solution rows of a compact result (synchronises a device-resident result)"""
        return int(self["sol_offset"][-1]) - int(getattr(self, "row_base", 0))

    def rows(self, i):
        """index of the per-slot arrays for the solutions of pair i: (i, slice) padded, slice of rows compact"""
        if self.compact:
            return slice(int(self["sol_offset"][i]), int(self["sol_offset"][i + 1]))
        return (i, slice(0, int(self["n_sol"][i])))

    def solutions(self, i):
        """results of pair i in the reference's `get_results()` format"""
        r = self.rows(i)
        return [{'type': int(t), 'C0': float(c0), 'C1': float(c1), 'reflection': int(k), 'reflection_case': int(rc)}
                for t, c0, c1, k, rc in zip(self["solution_type"][r], self["C0"][r], self["C1"][r], self["reflection"][r],
                                            self["reflection_case"][r])]


class ray_tracing(ray_tracing_base):
    """
    utility class to get ray tracing solutions in 3D for two arbitrary points x1 and x2
    (drop-in for NuRadioMC.SignalProp.analyticraytracing.ray_tracing, :1932)
    """

    def __init__(self, medium, attenuation_model=None, log_level=logging.NOTSET,
                 n_frequencies_integration=None, n_reflections=None, config=None,
                 detector=None, ray_tracing_2D_kwards={},
                 use_cpp=None, compile_numba=None, device=0):
        """
        Same parameters as the reference (:1938-1998).  `use_cpp` / `compile_numba` select between the reference's CPU
        back ends and are accepted for call compatibility; the computation always runs on CUDA device `device`.
        """
        self.__logger = logging.getLogger('NuRadioMC.ray_tracing')
        self.__logger.setLevel(log_level)

        if not all(hasattr(medium, a) for a in ("n_ice", "delta_n", "z_0")) or getattr(medium, "z_shift", 0) != 0:
            # the reference checks isinstance(medium, IceModelSimple) (:2002-2005); any object exposing the
            # exponential-profile parameters (this package's or the reference's IceModelSimple) is accepted here
            self.__logger.error("The analytic raytracer can only handle ice model of the type 'IceModelSimple'")
            raise TypeError("The analytic raytracer can only handle ice model of the type 'IceModelSimple'")

        super().__init__(medium=medium, attenuation_model=attenuation_model, log_level=log_level,
                         n_frequencies_integration=n_frequencies_integration, n_reflections=n_reflections,
                         config=config, detector=detector)
        self.set_config(config=config)

        if medium.delta_n == 0:   # uniform_ice (:433-437)
            msg = ('Analytic raytracer does not work with a uniform ice model. '
                   'Abort.... ! Use direct raytracing or a non-uniform ice model instead.')
            self.__logger.error(msg)
            raise RuntimeError(msg)
        if not hasattr(self._medium, "reflection"):
            self._medium.reflection = None
        if self._attenuation_model not in attenuation_util.model_to_int:   # (:449-450)
            raise NotImplementedError("attenuation model {} is not implemented".format(self._attenuation_model))
        if use_cpp:
            self.__logger.warning("use_cpp=True requests the reference's C++/GSL extension; this implementation runs "
                                  "the CUDA library instead")
        self.use_cpp = False
        self._device = device
        self._handle = None
        self._swap = None
        self._x1 = None
        self._x2 = None
        self._cache = None          # results of the current pair (arrays with leading dimension S)
        self._att_cache = {}
        self._foc_cache = {}
        self._batch = None          # (lookup dict, BatchResult) from prepare_batch
        self._batch_index = None

    # ------------------------------------------------------------------------------------------------------
    # native handle
    # ------------------------------------------------------------------------------------------------------
    def _h(self):
        if self._handle is None:
            m = self._medium
            self._handle = _make_handle(m.n_ice, m.delta_n, m.z_0, getattr(m, "reflection", None),
                                        attenuation_util.model_to_int[self._attenuation_model], self._n_reflections,
                                        self._n_frequencies_integration, self._device)
        return self._handle

    def set_chunk_pairs(self, pairs):
        """pairs per internal chunk of the batched calls (0 = automatic); results do not depend on it"""
        h = self._h()
        _lib.check(_lib.load().nrmc_rt_set_chunk_pairs(h.ptr, int(pairs)), h.ptr, "set_chunk_pairs")

    @staticmethod
    def _freq_key(frequency, max_detector_freq):
        frequency = np.ascontiguousarray(frequency, dtype=np.float64)
        return (frequency.tobytes(), None if max_detector_freq is None else float(max_detector_freq))

    def _set_frequencies(self, frequency, max_detector_freq):
        h = self._h()
        frequency = np.ascontiguousarray(frequency, dtype=np.float64)
        key = self._freq_key(frequency, max_detector_freq)
        if h.freq_key != key:
            lib = _lib.load()
            fmax = float("nan") if max_detector_freq is None else float(max_detector_freq)
            Fs = _lib.check(lib.nrmc_rt_set_frequencies(h.ptr, frequency.ctypes.data, len(frequency), fmax), h.ptr,
                            "set_frequencies")
            sp = np.empty(Fs)
            lib.nrmc_rt_get_sparse_frequencies(h.ptr, sp.ctypes.data, Fs)
            h.freq_key, h.sparse, h.n_out = key, sp, len(frequency)
        return h.sparse

    # ------------------------------------------------------------------------------------------------------
    # batched entry points (new)
    # ------------------------------------------------------------------------------------------------------
    def trace_batch(self, X1, X2, frequency=None, max_detector_freq=None, outer=False, outputs=None,
                    attenuation="dense", pinned=False, out=None, compact=False, row_capacity=None, shower_axis=None,
                    delta_C_cut=None):
        """
        Trace all pairs in one device pass (host arrays in, host arrays out).

        X1: (Nv, 3) start points (vertices); X2: (Na, 3) end points (antennas).
        outer=False: Na == Nv (or Na == 1), pair i = (X1[i], X2[i]);  outer=True: all Nv x Na pairs, vertex-major.
        frequency / max_detector_freq: as in `get_attenuation`; if given, attenuation factors are computed
        (`attenuation` = "dense": on `frequency`; "sparse": at the integration frequencies; "both").
        outputs: iterable of field names (default: everything but attenuation).  pinned=True allocates page-locked
        result arrays; `out` may pass a previous BatchResult of the same shape to reuse its buffers.
        compact=True: per-solution (CSR) layout -- every per-slot array has one row per EXISTING solution instead of
        S slots per pair; the rows of pair i are res["sol_offset"][i] : res["sol_offset"][i+1] (slot order).  Empty
        slots are neither stored nor copied from the device.  `row_capacity` bounds the rows allocated (default N*S).
        shower_axis: (Nv, 3) propagation direction of the shower at every vertex (simulation.py:175 uses -shower.get_axis());
        adds the output "viewing_angle" and, with delta_C_cut [rad], applies the reference's viewing-angle cut
        (simulation.py:195-208): solutions further than delta_C_cut from the Cherenkov cone get no attenuation (NaN).
        """
        X1 = np.asarray(X1, dtype=np.float64).reshape(-1, 3)
        X2 = np.asarray(X2, dtype=np.float64).reshape(-1, 3)
        if not outer and X2.shape[0] == 1 and X1.shape[0] != 1:
            X2 = np.repeat(X2, X1.shape[0], axis=0)
        if not outer and X1.shape[0] != X2.shape[0]:
            raise ValueError("X1 and X2 must have the same number of points unless outer=True")
        v = np.ascontiguousarray(X1.T)
        a = np.ascontiguousarray(X2.T)
        N = X1.shape[0] * X2.shape[0] if outer else X1.shape[0]
        h = self._h()
        names = list(outputs) if outputs is not None else list(DEFAULT_OUTPUTS)
        Fs = F = 0
        if frequency is not None:
            sp = self._set_frequencies(frequency, max_detector_freq)
            Fs, F = len(sp), len(frequency)
            if attenuation in ("dense", "both") and "attenuation" not in names:
                names.append("attenuation")
            if attenuation in ("sparse", "both") and "attenuation_sparse" not in names:
                names.append("attenuation_sparse")
        elif any(n in ("attenuation", "attenuation_sparse") for n in names):
            raise ValueError("attenuation outputs need `frequency`")
        S, K1 = self.get_number_of_raytracing_solutions(), self._n_reflections + 1
        sax = None
        if shower_axis is not None:
            sax = np.ascontiguousarray(np.asarray(shower_axis, dtype=np.float64).reshape(-1, 3).T)
            if sax.shape[1] != X1.shape[0]:
                raise ValueError("shower_axis needs one direction per start point")
            if "viewing_angle" not in names:
                names.append("viewing_angle")
        res = out if out is not None else BatchResult()
        keep = []
        o = _lib.Output()
        rows = N * S if row_capacity is None else int(row_capacity)
        full = getattr(res, "_full", None) or {}
        if compact:
            if "n_sol" not in names:
                names.insert(0, "n_sol")
            names.append("sol_offset")
        for name in names:
            if name == "sol_offset":
                dtype, shape = np.int64, (N + 1,)
            else:
                dtype, trail = _OUT_SPECS[name]
                per_slot = len(trail(S, K1, Fs, F)) > 0
                shape = ((rows,) + trail(S, K1, Fs, F)[1:]) if (compact and per_slot) else (N,) + trail(S, K1, Fs, F)
            if name in full and full[name].shape == shape:
                arr = full[name]
            elif name in res and res[name].shape == shape:
                arr = res[name]
            elif pinned:
                pa = PinnedArray(shape, dtype)
                keep.append(pa)
                arr = pa.array
            else:
                arr = np.empty(shape, dtype=dtype)
            res[name] = arr
            full[name] = arr
            setattr(o, name, arr.ctypes.data)
        if keep:
            res._keep = (res._keep or []) + keep
        o.compact, o.row_capacity = int(bool(compact)), rows
        inp = _lib.Input()
        inp.n_vertices, inp.vx, inp.vy, inp.vz = X1.shape[0], v[0].ctypes.data, v[1].ctypes.data, v[2].ctypes.data
        inp.n_antennas, inp.ax, inp.ay, inp.az = X2.shape[0], a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data
        inp.outer, inp.memory = int(bool(outer)), _lib.MEMORY_HOST
        if sax is not None:
            inp.sx, inp.sy, inp.sz = sax[0].ctypes.data, sax[1].ctypes.data, sax[2].ctypes.data
            inp.delta_C_cut = float(np.pi) if delta_C_cut is None else float(delta_C_cut)
        st = _lib.Stats()
        _lib.check(_lib.load().nrmc_rt_trace(h.ptr, C.byref(inp), C.byref(o), None, C.byref(st)), h.ptr, "trace")
        res.stats = _stats_dict(st)
        res.frequencies_sparse = h.sparse if frequency is not None else None
        res.freq_key = None if frequency is None else self._freq_key(frequency, max_detector_freq)
        res.compact = bool(compact)
        if compact:     # expose the filled rows; the capacity-sized buffers are kept for reuse through `out=`
            res._full = full
            n_rows = int(res["sol_offset"][N])
            for name in names:
                if name not in ("n_sol", "status", "sol_offset"):
                    res[name] = full[name][:n_rows]
        return res

    def trace_batch_device(self, v, a, frequency=None, max_detector_freq=None, outer=False, outputs=None,
                           attenuation="sparse", out=None, sync_stats=False, compact=False, row_capacity=None,
                           out_ptrs=None, row_base=0):
        """
        Device-resident variant: `v` (3, Nv) and `a` (3, Na) are contiguous float64 CUDA torch tensors (SoA); the results
        are CUDA torch tensors, the kernels are enqueued on torch's current stream.  Used by bench.py for the
        HBM-resident number and by callers that keep the next stage on the GPU.
        compact=True: per-solution rows as in `trace_batch`; the per-slot tensors keep
        their capacity (`row_capacity`, default N*S) and only the first `res["sol_offset"][N]` rows are defined -- reading
        that number is the caller's synchronisation point (`res.n_rows()`).
        out_ptrs / row_base (compact only): {output name: raw device pointer} of per-solution arrays that live elsewhere --
        typically the gathering GPU's arrays mapped through NVLink (`nuradiomc_b200.distributed.P2PGather`) -- addressed as
        array[row_base + local row]; those outputs are not allocated here and do not appear in the result dict.
        """
        import torch
        assert v.is_cuda and a.is_cuda and v.dtype == torch.float64 and a.dtype == torch.float64
        assert v.dim() == 2 and v.shape[0] == 3 and a.dim() == 2 and a.shape[0] == 3 and v.is_contiguous() and a.is_contiguous()
        Nv, Na = v.shape[1], a.shape[1]
        N = Nv * Na if outer else Nv
        h = self._h()
        names = list(outputs) if outputs is not None else list(DEFAULT_OUTPUTS)
        Fs = F = 0
        if frequency is not None:
            sp = self._set_frequencies(frequency, max_detector_freq)
            Fs, F = len(sp), len(frequency)
            if attenuation in ("dense", "both") and "attenuation" not in names:
                names.append("attenuation")
            if attenuation in ("sparse", "both") and "attenuation_sparse" not in names:
                names.append("attenuation_sparse")
        S, K1 = self.get_number_of_raytracing_solutions(), self._n_reflections + 1
        res = out if out is not None else BatchResult()
        o = _lib.Output()
        tdt = {np.int32: torch.int32, np.int8: torch.int8, np.float64: torch.float64}
        rows = N * S if row_capacity is None else int(row_capacity)
        if compact and "n_sol" not in names:
            names.insert(0, "n_sol")
        out_ptrs = dict(out_ptrs or {})
        if out_ptrs and not compact:
            raise ValueError("out_ptrs / row_base need the compact layout")
        for name in list(out_ptrs):
            if name not in names:
                names.append(name)
        for name in names:
            dtype, trail = _OUT_SPECS[name]
            per_slot = len(trail(S, K1, Fs, F)) > 0
            if name in out_ptrs:
                if not per_slot:
                    raise ValueError("per-pair outputs stay local (the kernels read them back); gather them with a copy")
                setattr(o, name, int(out_ptrs[name]))
                res.pop(name, None)
                continue
            shape = ((rows,) + trail(S, K1, Fs, F)[1:]) if (compact and per_slot) else (N,) + trail(S, K1, Fs, F)
            if not (name in res and tuple(res[name].shape) == shape):
                res[name] = torch.empty(shape, dtype=tdt[dtype], device=v.device)
            ptr = res[name].data_ptr()
            if compact and per_slot and row_base:
                # the kernels address every per-solution array as array[row_base + local row]: arrays that stay local (not in
                # out_ptrs) are handed over shifted back by row_base rows, so that local row r lands in res[name][r]
                ptr -= int(row_base) * int(np.prod(shape[1:], dtype=np.int64)) * res[name].element_size()
            setattr(o, name, ptr)
        if compact:
            if not ("sol_offset" in res and tuple(res["sol_offset"].shape) == (N + 1,)):
                res["sol_offset"] = torch.empty(N + 1, dtype=torch.int64, device=v.device)
            o.compact, o.sol_offset, o.row_capacity, o.row_base = 1, res["sol_offset"].data_ptr(), rows, int(row_base)
        res.compact = bool(compact)
        res.row_base = int(row_base)
        inp = _lib.Input()
        inp.n_vertices, inp.vx, inp.vy, inp.vz = Nv, v[0].data_ptr(), v[1].data_ptr(), v[2].data_ptr()
        inp.n_antennas, inp.ax, inp.ay, inp.az = Na, a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr()
        inp.outer, inp.memory = int(bool(outer)), _lib.MEMORY_DEVICE
        stream = torch.cuda.current_stream(v.device).cuda_stream
        st = _lib.Stats()
        _lib.check(_lib.load().nrmc_rt_trace(h.ptr, C.byref(inp), C.byref(o), C.c_void_p(stream),
                                             C.byref(st) if sync_stats else None), h.ptr, "trace")
        res.stats = _stats_dict(st) if sync_stats else None
        res.frequencies_sparse = h.sparse if frequency is not None else None
        return res

    def focusing_batch(self, X1, X2, result, outer=False, limit=None):
        """
        Focusing factor of every solution of a batch result (`get_focusing`, reference :2778-2888, numerical branch) in one
        kernel: the derivative of the launch angle with respect to the receiver depth is exact (closed-form dR/dbeta) where
        the reference differences two traces 1 cm apart.  X2 are the receivers.  `result`: what `trace_batch` /
        `trace_batch_device` returned for the same points (padded or compact; needs n_sol, C0, path_length and, with
        bottom reflections, reflection / reflection_case).  Host result -> numpy array, device result -> CUDA tensor
        (X1, X2 then as in `trace_batch_device`: (3, N) tensors); shape (N, S) or (rows,).
        limit: maximum amplification, default config['propagation']['focusing_limit'] (2).
        """
        import torch
        h = self._h()
        dev = torch.device("cuda", self._device)
        if limit is None:
            limit = float(self._config['propagation'].get('focusing_limit', 2))
        on_host = isinstance(result["C0"], np.ndarray)
        if on_host:
            X1 = np.asarray(X1, dtype=np.float64).reshape(-1, 3)
            X2 = np.asarray(X2, dtype=np.float64).reshape(-1, 3)
            if not outer and X2.shape[0] == 1 and X1.shape[0] != 1:
                X2 = np.repeat(X2, X1.shape[0], axis=0)
            v = torch.as_tensor(np.ascontiguousarray(X1.T)).to(dev)
            a = torch.as_tensor(np.ascontiguousarray(X2.T)).to(dev)
        else:
            v, a = X1, X2
            assert v.is_cuda and a.is_cuda and v.dtype == torch.float64 and a.dtype == torch.float64
            assert v.dim() == 2 and v.shape[0] == 3 and a.dim() == 2 and a.shape[0] == 3 and v.is_contiguous() and a.is_contiguous()
        Nv, Na = v.shape[1], a.shape[1]
        if not outer and Nv != Na:
            raise ValueError("X1 and X2 must have the same number of points unless outer=True")

        def dev_tensor(name, required=True):
            x = result.get(name)
            if x is None:
                if required:
                    raise ValueError(f"focusing_batch needs the output '{name}' of the trace")
                return None
            if isinstance(x, np.ndarray):
                return torch.as_tensor(np.ascontiguousarray(x)).to(dev)
            return x.contiguous()
        compact = bool(getattr(result, "compact", False))
        need_modes = self._n_reflections > 0
        t = {k: dev_tensor(k) for k in ("n_sol", "C0", "path_length")}
        t["reflection"], t["reflection_case"] = dev_tensor("reflection", need_modes), dev_tensor("reflection_case", need_modes)
        t["sol_offset"] = dev_tensor("sol_offset") if compact else None
        foc = torch.empty_like(t["C0"])
        fo = _lib.Focusing()
        for k in ("n_sol", "C0", "path_length", "reflection", "reflection_case", "sol_offset"):
            setattr(fo, k, t[k].data_ptr() if t[k] is not None else None)
        fo.limit, fo.focusing = float(limit), foc.data_ptr()
        inp = _lib.Input()
        inp.n_vertices, inp.vx, inp.vy, inp.vz = Nv, v[0].data_ptr(), v[1].data_ptr(), v[2].data_ptr()
        inp.n_antennas, inp.ax, inp.ay, inp.az = Na, a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr()
        inp.outer, inp.memory = int(bool(outer)), _lib.MEMORY_DEVICE
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().nrmc_rt_focusing_factor(h.ptr, C.byref(inp), C.byref(fo), C.c_void_p(stream)), h.ptr, "focusing_factor")
        return foc.cpu().numpy() if on_host else foc

    def apply_propagation_effects_batch(self, spectra, reflection_angle=None, reflection=None, attenuation=None,
                                        attenuation_sparse=None, return_coefficients=False, focusing=None):
        """
        Batched `apply_propagation_effects` (reference :2937-3033, in-ice branch; birefringence is out of scope): one row
        per ray-tracing solution.  focusing: (R,) factors of `focusing_batch` on eTheta and ePhi (:3012-3015), optional.

        spectra: (R, 3, F) complex128 -- eR, eTheta, ePhi spectra; a CUDA torch tensor is modified in place (and
        returned), a numpy array is copied to the device and the result is returned as a new numpy array.
        reflection_angle: (R, n_reflections+1) as returned by `trace_batch` (NaN = no surface reflection on that segment);
        reflection: (R,) number of bottom reflections; attenuation: (R, F) factors on the spectra's frequency grid, or
        attenuation_sparse: (R, Fs) factors at the integration frequencies (single-segment paths; interpolated on the fly
        with the tables of the last `frequency` passed to trace_batch*).
        """
        import torch
        h = self._h()
        dev = torch.device("cuda", self._device)
        is_np = isinstance(spectra, np.ndarray)

        def dev_tensor(x, dtype):
            if x is None:
                return None
            if isinstance(x, np.ndarray):
                return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(dev)
            assert x.is_cuda and x.dtype == dtype and x.is_contiguous()
            return x
        spec = dev_tensor(spectra, torch.complex128)
        if spec.dim() != 3 or spec.shape[1] != 3:
            raise ValueError("spectra must have the shape (rows, 3, n_frequencies)")
        R, F = spec.shape[0], spec.shape[2]
        ang = dev_tensor(reflection_angle, torch.float64)
        refl = dev_tensor(reflection, torch.int8)
        att = dev_tensor(attenuation, torch.float64)
        att_sp = dev_tensor(attenuation_sparse, torch.float64)
        foc = dev_tensor(focusing, torch.float64)
        if foc is not None and tuple(foc.shape) != (spec.shape[0],):
            raise ValueError("focusing must have one factor per row")
        K1 = self._n_reflections + 1
        if ang is not None and tuple(ang.shape) != (R, K1):
            raise ValueError(f"reflection_angle must have the shape ({R}, {K1})")
        if att is not None and tuple(att.shape) != (R, F):
            raise ValueError(f"attenuation must have the shape ({R}, {F})")
        fx = _lib.Effects()
        fx.n_rows, fx.n_freq = R, F
        fx.spectrum = spec.data_ptr()
        fx.attenuation = att.data_ptr() if att is not None else None
        fx.attenuation_sparse = att_sp.data_ptr() if att_sp is not None else None
        fx.reflection_angle = ang.data_ptr() if ang is not None else None
        fx.reflection = refl.data_ptr() if refl is not None else None
        fx.focusing = foc.data_ptr() if foc is not None else None
        rc, ph = getattr(self._medium, "reflection_coefficient", None), getattr(self._medium, "reflection_phase_shift", None)
        fx.reflection_coefficient = 1.0 if rc is None else float(rc)
        fx.reflection_phase_shift = 0.0 if ph is None else float(ph)
        r_t = r_p = None
        if return_coefficients:
            r_t = torch.empty(R, dtype=torch.complex128, device=dev)
            r_p = torch.empty(R, dtype=torch.complex128, device=dev)
            fx.r_theta, fx.r_phi = r_t.data_ptr(), r_p.data_ptr()
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().nrmc_rt_apply_propagation_effects(h.ptr, C.byref(fx), C.c_void_p(stream)), h.ptr,
                   "apply_propagation_effects")
        out = spec.cpu().numpy() if is_np else spec
        if return_coefficients:
            return (out, r_t.cpu().numpy(), r_p.cpu().numpy()) if is_np else (out, r_t, r_p)
        return out

    def apply_propagation_effects(self, efield, i_solution):
        """
        Apply propagation effects to the electric field (reference :2937-3033): attenuation, Fresnel coefficients of
        surface reflections, bottom-reflection amplitude and phase.  `efield` is a NuRadioReco ElectricField or any object
        with get_frequency_spectrum / get_frequencies / get_sampling_rate / set_frequency_spectrum (and item assignment).
        """
        self._check(i_solution)
        self._need_cache()
        if self._config['propagation'].get('birefringence', False):
            raise NotImplementedError("birefringence is outside the scope of nuradiomc_b200 (SURVEY.md section 8a)")
        if self._X2[2] > 0 or self._X1[2] > 0:
            raise NotImplementedError("air/ice transmission is experimental in the reference (:2971) and not provided")
        spec = np.array(efield.get_frequency_spectrum(), dtype=np.complex128)
        ff = np.asarray(efield.get_frequencies(), dtype=np.float64)
        att = None
        if self._config['propagation']['attenuate_ice']:
            max_freq = np.max(ff) if self._max_detector_frequency is None else self._max_detector_frequency
            att = self.get_attenuation(i_solution, ff, max_freq)[None, :]
        k = self._results[i_solution]['reflection']
        foc = None
        if self._config['propagation'].get('focusing', False):     # :3012-3015
            foc = np.array([self.get_focusing(i_solution, limit=float(self._config['propagation']['focusing_limit']))])
        out, r_t, r_p = self.apply_propagation_effects_batch(
            spec[None], reflection_angle=np.ascontiguousarray(self._cache["reflection_angle"][i_solution][None, :]),
            reflection=np.array([k], np.int8), attenuation=att, return_coefficients=True, focusing=foc)
        if not np.all(np.isnan(self._cache["reflection_angle"][i_solution][:k + 1])):
            try:    # the reference stores the coefficients on the field object (:2993-2994)
                try:
                    from NuRadioReco.framework.parameters import electricFieldParameters as efp
                    efield[efp.reflection_coefficient_theta], efield[efp.reflection_coefficient_phi] = r_t[0], r_p[0]
                except ImportError:
                    efield["reflection_coefficient_theta"], efield["reflection_coefficient_phi"] = r_t[0], r_p[0]
            except (TypeError, KeyError, AttributeError):
                pass
        efield.set_frequency_spectrum(out[0], efield.get_sampling_rate())
        return efield

    def prepare_batch(self, X1, X2, outer=False, **kwargs):
        """
        Trace many pairs ahead of a scalar loop (e.g. all (shower, channel) pairs of an event group in
        NuRadioMC/simulation/simulation.py:155-210).  Afterwards `set_start_and_end_point(x1, x2)` +
        `find_solutions()` on one of these pairs is a cache lookup instead of a kernel launch.
        """
        if kwargs.get("compact", False):
            raise ValueError("prepare_batch serves the scalar API from the padded layout: compact=True is not supported")
        X1 = np.asarray(X1, dtype=np.float64).reshape(-1, 3)
        X2 = np.asarray(X2, dtype=np.float64).reshape(-1, 3)
        res = self.trace_batch(X1, X2, outer=outer, **kwargs)
        if self._config['propagation'].get('focusing', False):
            # the loop will ask for get_focusing / get_raytracing_output of every solution (:2913-2916, :3012-3015): one launch now
            limit = float(self._config['propagation'].get('focusing_limit', 2))
            res["focusing_factor"] = self.focusing_batch(X1, X2, res, outer=outer, limit=limit)
            res.focusing_limit = limit
        self._batch = (_pair_lookup(X1, X2, outer), res)
        return res

    # ------------------------------------------------------------------------------------------------------
    # the reference's scalar API
    # ------------------------------------------------------------------------------------------------------
    def reset_solutions(self):
        super().reset_solutions()
        self._x1 = None
        self._x2 = None
        self._swap = None
        self._cache = None
        self._att_cache = {}
        self._foc_cache = {}
        self._batch_index = None

    def set_start_and_end_point(self, x1, x2):
        super().set_start_and_end_point(x1, x2)
        # 2-D frame as exposed by the reference (:2072-2089): deeper point first, rho along +x
        self._swap = bool(self._X2[2] < self._X1[2])
        A, B = (self._X2, self._X1) if self._swap else (self._X1, self._X2)
        self._x1 = np.array([A[0], A[2]])
        self._x2 = np.array([A[0] + np.hypot(B[0] - A[0], B[1] - A[1]), B[2]])

    def set_solution(self, raytracing_results):
        """Read an already calculated raytracing solution from the input array (:2092-2116)"""
        results = []
        C0s = raytracing_results['ray_tracing_C0']
        for i in range(len(C0s)):
            if not np.isnan(C0s[i]):
                if 'ray_tracing_reflection' in raytracing_results.keys():
                    reflection = raytracing_results['ray_tracing_reflection'][i]
                    reflection_case = raytracing_results['ray_tracing_reflection_case'][i]
                else:
                    reflection = 0
                    reflection_case = 0
                results.append({'type': raytracing_results['ray_tracing_solution_type'][i], 'C0': C0s[i],
                                'C1': raytracing_results['ray_tracing_C1'][i], 'reflection': reflection,
                                'reflection_case': reflection_case})
        self._results = results

    def find_solutions(self):
        """find all solutions between x1 and x2 (:2118-2130)"""
        if self._X1 is None:
            raise AttributeError("set_start_and_end_point has to be called before find_solutions")
        src, idx = None, None
        if self._batch is not None:
            idx = self._batch[0].get(self._X1.tobytes() + self._X2.tobytes())
            if idx is not None:
                src = self._batch[1]
        if src is None:
            src, idx = self._trace_one(), 0
        self._batch_index = idx if (self._batch is not None and src is self._batch[1]) else None
        n = int(src["n_sol"][idx])
        self._cache = {k: np.array(src[k][idx]) for k in DEFAULT_OUTPUTS if k in src}
        self._results = [{'type': int(self._cache["solution_type"][s]), 'C0': float(self._cache["C0"][s]),
                          'C1': float(self._cache["C1"][s]), 'reflection': int(self._cache["reflection"][s]),
                          'reflection_case': int(self._cache["reflection_case"][s])} for s in range(n)]
        if "focusing_factor" in src:     # pre-traced with focusing on: get_focusing(limit = the configured one) is a lookup
            self._foc_cache[float(getattr(src, "focusing_limit", 2.))] = np.array(src["focusing_factor"][idx])

    def _trace_one(self):
        """the current pair through the C ABI with buffers and ctypes structures that are set up once per propagator: the scalar API
        is a batch of one pair, and at ~70 us of device time per call the Python-side set-up of `trace_batch` (a dozen numpy
        allocations, three structures) would cost as much again"""
        ctx = getattr(self, "_one_ctx", None)
        if ctx is None:
            S, K1 = self.get_number_of_raytracing_solutions(), self._n_reflections + 1
            out = {}
            o = _lib.Output()
            for name in DEFAULT_OUTPUTS:
                dtype, trail = _OUT_SPECS[name]
                out[name] = np.empty((1,) + trail(S, K1, 0, 0), dtype=dtype)
                setattr(o, name, out[name].ctypes.data)
            o.compact, o.row_capacity = 0, S
            pts = np.zeros((6, 1))
            inp = _lib.Input()
            inp.n_vertices, inp.vx, inp.vy, inp.vz = 1, pts[0].ctypes.data, pts[1].ctypes.data, pts[2].ctypes.data
            inp.n_antennas, inp.ax, inp.ay, inp.az = 1, pts[3].ctypes.data, pts[4].ctypes.data, pts[5].ctypes.data
            inp.outer, inp.memory = 0, _lib.MEMORY_HOST
            ctx = self._one_ctx = dict(out=out, o=o, pts=pts, inp=inp, fn=_lib.load().nrmc_rt_trace, h=self._h())
        ctx["pts"][:3, 0] = self._X1
        ctx["pts"][3:, 0] = self._X2
        rc = ctx["fn"](ctx["h"].ptr, C.byref(ctx["inp"]), C.byref(ctx["o"]), None, None)
        if rc < 0:
            _lib.check(rc, ctx["h"].ptr, "trace")
        return ctx["out"]

    def _check(self, iS):
        n = self.get_number_of_solutions()
        if iS >= n:
            self.__logger.error("solution number {:d} requested but only {:d} solutions exist".format(iS + 1, n))
            raise IndexError

    def _need_cache(self):
        if self._cache is None and self._X1 is not None and self._results:
            self._rebuild_cache_from_results()
        if self._cache is None:
            raise AttributeError("find_solutions (or set_start_and_end_point + set_solution) has to be called first")

    def _rebuild_cache_from_results(self):
        """solutions injected with `set_solution` (reference :2092-2116; speedup.redo_raytracing = False) carry C0 / type
        only: trace the pair once and keep, in the injected order, the solutions that were injected"""
        src = self.trace_batch(self._X1[None, :], self._X2[None, :])
        n = int(src["n_sol"][0])
        pick = []
        for r in self._results:
            cand = [s for s in range(n) if int(src["reflection"][0, s]) == int(r['reflection'])
                    and abs(src["C0"][0, s] - r['C0']) <= 1e-6 * abs(r['C0'])]
            if not cand:
                raise AttributeError("the injected ray-tracing solution (C0 = {}) does not belong to this pair of points".format(r['C0']))
            pick.append(cand[0])
        self._cache = {k: np.array(src[k][0][pick]) for k in DEFAULT_OUTPUTS if k in src and np.ndim(src[k][0]) > 0}

    def get_solution_type(self, iS):
        self._check(iS)
        if self._cache is None:
            return self._results[iS]['type']
        return int(self._cache["solution_type"][iS])

    def get_launch_vector(self, iS):
        self._check(iS)
        self._need_cache()
        return np.array(self._cache["launch_vector"][iS])

    def get_receive_vector(self, iS):
        self._check(iS)
        self._need_cache()
        return np.array(self._cache["receive_vector"][iS])

    def get_reflection_angle(self, iS):
        """angle of reflection at the surface: None for direct/refracted rays; array (one entry per path segment)
        when bottom reflections are simulated (:2626-2648, np.squeeze of a per-segment list :1237)"""
        self._check(iS)
        self._need_cache()
        k = self._results[iS]['reflection']
        ang = self._cache["reflection_angle"][iS][:k + 1]
        vals = [None if np.isnan(a) else float(a) for a in ang]
        return np.squeeze(vals)

    def get_path_length(self, iS, analytic=True):
        self._check(iS)
        self._need_cache()
        return float(self._cache["path_length"][iS])

    def get_travel_time(self, iS, analytic=True):
        self._check(iS)
        self._need_cache()
        return float(self._cache["travel_time"][iS])

    def get_attenuation(self, iS, frequency, max_detector_freq=None):
        """fraction of the signal that reaches the observer per frequency (:2744-2776)"""
        self._check(iS)
        frequency = np.asarray(frequency, dtype=np.float64)
        key = self._freq_key(frequency, max_detector_freq)
        if key not in self._att_cache:
            src, idx = None, 0
            if self._batch is not None and self._batch_index is not None and self._batch[1].freq_key == key \
                    and "attenuation" in self._batch[1]:
                src, idx = self._batch[1], self._batch_index      # pre-traced with the same frequencies: a lookup
            else:
                src = self.trace_batch(self._X1[None, :], self._X2[None, :], frequency=frequency, max_detector_freq=max_detector_freq,
                                       outputs=("n_sol", "C0", "reflection", "reflection_case"), attenuation="dense")
            # rows in the trace's slot order, with the identity of each slot: _results may hold a subset in another order
            # (set_solution after a viewing-angle cut), and the reference looks the solution up by its C0 / reflection / case
            n = int(src["n_sol"][idx])
            self._att_cache[key] = (np.array(src["attenuation"][idx][:n]), np.array(src["C0"][idx][:n]),
                                    np.array(src["reflection"][idx][:n]), np.array(src["reflection_case"][idx][:n]))
        att, C0s, refl, case = self._att_cache[key]
        r = self._results[iS]
        cand = [s for s in range(len(C0s)) if int(refl[s]) == int(r['reflection']) and abs(C0s[s] - r['C0']) <= 1e-6 * abs(r['C0'])
                and (int(r['reflection']) == 0 or int(r.get('reflection_case', case[s])) in (0, int(case[s])))]
        if not cand:
            raise AttributeError("the ray-tracing solution (C0 = {}) does not belong to this pair of points".format(r['C0']))
        return np.array(att[cand[0]])

    def get_focusing(self, iS, dz=-1. * units.cm, limit=2., analytic=False):
        """
        gain of the signal at the receiver due to the focusing effect (reference :2778-2888).  `dz` is accepted for
        compatibility: the kernel differentiates exactly (the dz -> 0 limit of the reference's difference quotient).
        `analytic=True` raises AttributeError in the reference (:831, `self.n`); both values give the same result here.
        """
        self._check(iS)
        self._need_cache()
        if self._X1[2] > 0 or self._X2[2] > 0:
            raise NotImplementedError("air/ice transmission is experimental in the reference (:2877-2886) and not provided")
        key = float(limit)
        if key not in self._foc_cache:
            n, S = len(self._results), self.get_number_of_raytracing_solutions()

            def slots(name, dtype, fill):      # one padded [1, S] row (the cache holds n rows after set_solution)
                a = np.full((1, S), fill, dtype=dtype)
                a[0, :n] = self._cache[name][:n]
                return a
            res = {"n_sol": np.array([n], np.int32), "C0": slots("C0", np.float64, np.nan),
                   "path_length": slots("path_length", np.float64, np.nan), "reflection": slots("reflection", np.int8, 0),
                   "reflection_case": slots("reflection_case", np.int8, 0)}
            self._foc_cache[key] = self.focusing_batch(self._X1[None, :], self._X2[None, :], res, limit=limit)[0]
        return float(self._foc_cache[key][iS])

    def get_path(self, iS, n_points=1000):
        """not on the hot path (path sampling is plotting support) -- not provided"""
        raise NotImplementedError("get_path (path sampling for plotting) is outside the scope of nuradiomc_b200")

    def get_output_parameters(self):
        return [
            {'name': 'ray_tracing_C0', 'ndim': 1},
            {'name': 'ray_tracing_C1', 'ndim': 1},
            {'name': 'focusing_factor', 'ndim': 1},
            {'name': 'ray_tracing_reflection', 'ndim': 1},
            {'name': 'ray_tracing_reflection_case', 'ndim': 1},
            {'name': 'ray_tracing_solution_type', 'ndim': 1}
        ]

    def get_raytracing_output(self, i_solution):
        focusing = 1
        if self._config['propagation']['focusing']:     # :2913-2916
            focusing = self.get_focusing(i_solution, limit=float(self._config['propagation']['focusing_limit']))
        return {
            'ray_tracing_C0': self.get_results()[i_solution]['C0'],
            'ray_tracing_C1': self.get_results()[i_solution]['C1'],
            'ray_tracing_reflection': self.get_results()[i_solution]['reflection'],
            'ray_tracing_reflection_case': self.get_results()[i_solution]['reflection_case'],
            'ray_tracing_solution_type': self.get_solution_type(i_solution),
            'focusing_factor': focusing
        }

    def set_config(self, config):
        """default config as the reference (:3035-3052)"""
        if config is None:
            self._config = {'propagation': {}}
            self._config['propagation']['attenuate_ice'] = True
            self._config['propagation']['focusing_limit'] = 2
            self._config['propagation']['focusing'] = False
            self._config['propagation']['birefringence'] = False
        else:
            self._config = config


class ray_tracing_2D:
    """
    The 2-D interface the bulk consumers call directly (reference :372-1929; NuRadioReco's vertex reconstructors,
    create_lookup_table.py:62-107, channelTimeOffsetCalculator.py:109-120): points are [y, z] in the propagation plane.
    A thin facade over `ray_tracing.trace_batch`: every call is one pair on the GPU; for bulk work use
    `find_solutions_batch`, which serves millions of pairs in one pass.

    Like the reference's solver, rays propagate towards +y: a receiver to the left of the emitter has no solution.
    """

    def __init__(self, medium, attenuation_model=None, log_level=logging.NOTSET, n_frequencies_integration=25,
                 use_optimized_start_values=False, overwrite_speedup=None, use_cpp=None, compile_numba=None, n_reflections=None,
                 device=0):
        self.medium = medium
        self.attenuation_model = attenuation_model or "SP1"
        self._n_reflections = 0 if (n_reflections is None or getattr(medium, "reflection", None) is None) else int(n_reflections)
        self._tracer = {}
        self._kw = dict(attenuation_model=self.attenuation_model, log_level=log_level,
                        n_frequencies_integration=n_frequencies_integration, device=device)

    def _rt(self, reflection):
        n = max(self._n_reflections, int(reflection))
        if n not in self._tracer:
            self._tracer[n] = ray_tracing(self.medium, n_reflections=n, **self._kw)
        return self._tracer[n]

    @staticmethod
    def _xyz(x):
        x = np.asarray(x, dtype=np.float64)
        return np.stack([x[..., 0], np.zeros_like(x[..., 0]), x[..., 1]], axis=-1)

    def find_solutions_batch(self, x1, x2, reflection=0, reflection_case=1, outputs=None):
        """all pairs x1[i] -> x2[i] ((N, 2) arrays [y, z]) in one device pass: the padded SoA result of `trace_batch`,
        restricted to the requested (reflection, reflection_case) mode through the mask `res["mode"]`"""
        x1, x2 = np.atleast_2d(np.asarray(x1, float)), np.atleast_2d(np.asarray(x2, float))
        res = self._rt(reflection).trace_batch(self._xyz(x1), self._xyz(x2), outputs=outputs)
        S = res["C0"].shape[1]
        filled = np.arange(S)[None, :] < res["n_sol"][:, None]
        forward = (np.broadcast_to(x2[:, 0], res["n_sol"].shape) >= np.broadcast_to(x1[:, 0], res["n_sol"].shape))[:, None]
        if reflection == 0:
            res["mode"] = filled & forward & (res["reflection"] == 0)
        else:
            res["mode"] = filled & forward & (res["reflection"] == reflection) & (res["reflection_case"] == reflection_case)
        return res

    def find_solutions(self, x1, x2, plot=False, reflection=0, reflection_case=1):
        """list of {'type', 'C0', 'C1', 'reflection', 'reflection_case'} sorted by C0 (reference :1400-1547)"""
        res = self.find_solutions_batch(np.asarray(x1, float)[None], np.asarray(x2, float)[None], reflection, reflection_case)
        out = []
        for s in np.nonzero(res["mode"][0])[0]:
            out.append({'type': int(res["solution_type"][0, s]), 'C0': float(res["C0"][0, s]), 'C1': float(res["C1"][0, s]),
                        'reflection': int(reflection), 'reflection_case': int(reflection_case)})
        return sorted(out, key=lambda d: d['C0'])

    def _lookup(self, x1, x2, C_0, key, reflection, reflection_case):
        res = self.find_solutions_batch(np.asarray(x1, float)[None], np.asarray(x2, float)[None], reflection, reflection_case)
        cand = np.nonzero(res["mode"][0])[0]
        if len(cand) == 0:
            return None
        s = cand[np.argmin(np.abs(res["C0"][0, cand] - C_0))]
        if abs(res["C0"][0, s] - C_0) > 1e-6 * abs(C_0):
            return None          # not a solution of this pair of points
        return res[key][0, s]

    def get_travel_time_analytic(self, x1, x2, C_0, reflection=0, reflection_case=1):
        v = self._lookup(x1, x2, C_0, "travel_time", reflection, reflection_case)
        return None if v is None else float(v)

    def get_path_length_analytic(self, x1, x2, C_0, reflection=0, reflection_case=1):
        v = self._lookup(x1, x2, C_0, "path_length", reflection, reflection_case)
        return None if v is None else float(v)

    # the reference's numerical versions (:519-599) integrate the same quantities with quad; the closed forms agree to 1e-12
    get_travel_time = get_travel_time_analytic
    get_path_length = get_path_length_analytic

    def get_launch_angle(self, x1, C_0, reflection=0, reflection_case=1, x2=None):
        """launch zenith angle (:1195); needs the end point as well here because the ray is identified through the trace"""
        if x2 is None:
            raise TypeError("nuradiomc_b200.ray_tracing_2D.get_launch_angle needs x2 (the ray is looked up by its end points)")
        v = self._lookup(x1, x2, C_0, "launch_vector", reflection, reflection_case)
        return None if v is None else float(np.arctan2(v[0], v[2]))

    def get_receive_angle(self, x1, x2, C_0, reflection=0, reflection_case=1):
        v = self._lookup(x1, x2, C_0, "receive_vector", reflection, reflection_case)
        return None if v is None else float(np.arctan2(-v[0], v[2]))


def measure_fp64_peak(device=0, seconds=1.0):
    """measured FP64 FMA peak [TFLOP/s] of the device (roofline denominator), and the nominal SM clock [MHz]"""
    t, clk = C.c_double(), C.c_double()
    _lib.check(_lib.load().nrmc_rt_measure_fp64_peak(device, seconds, C.byref(t), C.byref(clk)), None, "fp64_peak")
    return t.value, clk.value
