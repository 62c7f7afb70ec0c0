#!/usr/bin/env python
"""
bench.py -- benchmark of the B200-native analytic ray tracer.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg5] [--vertices NV]

Metric (BASELINE.json): ray-trace pairs/s (vertex x antenna pairs fully processed: solutions, type, C0/C1, launch/receive
vectors, path length, travel time AND -- where the configuration has them -- the attenuation factors), whole job over N GPUs.

Workloads (BASELINE.json configs[0..4] = cfg1..cfg5, inputs of SURVEY.md 8(d)); the default and the one the metric is quoted on
is cfg5: 1e6 vertices uniform in a cylinder r < 6 km, z in [-2700, 0] m (seed 5) x 100 channels (5 x 5 stations on a 1.5 km
grid, 4 channels per station at -145/-150/-155/-160 m) = 1e8 pairs, southpole_2015 ice, SP1 attenuation on the 512-bin
0-2.5 GHz grid with max_detector_freq = 1.2 GHz and n_freq = 25 -> 37 integration frequencies, sparse attenuation output.
A "step" is one pass of the hot path over all pairs.  N > 1: the vertices are sharded over the ranks (one process per GPU;
strong scaling: the total stays fixed).

  value           inputs resident in HBM when the timed region starts, outputs left in the HBM of the GPU that computed them.
  value_gathered  the same step with the compact result rows GATHERED on rank 0 inside the timed region (`gathered`):
                  "p2p"  -- the kernels store the rows straight into rank 0's HBM through NVLink peer mappings (fused, no collective),
                  "nccl" -- grouped ncclSend/ncclRecv of the filled rows after the kernels (the baseline variant);
                  "records+sparse" = every per-solution output incl. the 37 factors, "records" = without the factors.
  e2e             the same pass through the public API with HOST buffers: numpy in, pinned numpy out, H2D and D2H copies
                  inside the timed region (on a slice of the workload sized to the host memory, see config.e2e_vertices),
                  next to a pure D2H copy of the same bytes into the same kind of buffers (the ceiling of the box).
--impl reference times the reference's own CPU implementation on all host cores on a bounded sample of the same workload:
the unmodified Python reference (baseline/_ref through oracle/pyref/ref_bench.py: multiprocessing, one process per core, the
reference's numba path warm and its plain path) when that tree is present, and the C oracle port (oracle/) always.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RNOG = np.array([[0, 20, -97], [0, 20, -96], [0, 20, -95], [0, 20, -94], [0, 20, -93], [0, 20, -92], [0, 20, -80],
                 [0, 20, -60], [0, 20, -40], [-17.3, -10, -96], [-17.3, -10, -95], [-17.3, -10, -94], [1.5, 11, -2],
                 [0, 11, -2], [-1.5, 11, -2], [-10.276, -4.2, -2], [-9.526, -5.5, -2], [-8.776, -6.8, -2],
                 [8.776, -6.8, -2], [9.526, -5.5, -2], [10.276, -4.2, -2], [17.3, -10, -96], [17.3, -10, -95],
                 [17.3, -10, -94]], float)

# SURVEY.md 8(d).  n_freq None: the class default (100).  att_out: which attenuation output the configuration asks for.
CONFIGS = {
    "cfg1": dict(ice="southpole_simple", att="SP1", n_refl=0, n_freq=100, n_vertices=1000, seed=0, dist="T01", rmax=3000., zmin=-3000.,
                 antennas=np.array([[0., 0., -100.]]), freqs=np.linspace(0, 0.5, 129), fmax=None, att_out="dense",
                 text="cfg1 (BASELINE.json configs[0]): 1e3 vertices (T01 distribution, seed 0) to one antenna at -100 m; southpole_simple; SP1; "
                      "129 bins 0-0.5 GHz, n_freq 100; dense attenuation output"),
    "cfg2": dict(ice="southpole_2015", att=None, n_refl=0, n_freq=25, n_vertices=1_000_000, seed=2, dist="cyl", rmax=4000., zmin=-2700.,
                 antennas=np.array([[10., 10., -190.], [10., -10., -190.], [-10., -10., -190.], [-10., 10., -190.]]), freqs=None, fmax=None,
                 att_out=None, text="cfg2 (configs[1]): 1e6 vertices in r<4 km, z in [-2700,0] m x 4 ARA-like antennas at -190 m = 4e6 pairs; "
                                    "southpole_2015; solutions + travel time, no attenuation"),
    "cfg3": dict(ice="greenland_simple", att="GL1", n_refl=0, n_freq=25, n_vertices=1_000_000, seed=3, dist="cyl", rmax=4000., zmin=-2700.,
                 antennas=RNOG, freqs=np.fft.rfftfreq(1022, 0.2), fmax=1.2, att_out="dense",
                 text="cfg3 (configs[2]): 1e6 vertices in r<4 km, z in [-2700,0] m x 24 RNO-G channels = 2.4e7 pairs; greenland_simple; GL1; "
                      "512-bin grid 0-2.5 GHz, max_detector_freq 1.2 GHz, n_freq 25 -> 37 integration frequencies; DENSE 512-bin attenuation output"),
    "cfg4": dict(ice="mooresbay_simple", att=None, n_refl=1, n_freq=25, n_vertices=1_000_000, seed=4, dist="cyl", rmax=1000., zmin=-500.,
                 antennas=np.array([[-3., 0, -1], [0, 3, -1], [3, 0, -1], [0, -3, -1], [3, 3, -5], [3, -3, -5], [-3, -3, -5], [-3, 3, -5]], float),
                 freqs=None, fmax=None, att_out=None,
                 text="cfg4 (configs[3]): 1e6 vertices in r<1 km, z in [-500,0] m x 8 channels = 8e6 pairs; mooresbay_simple with bottom "
                      "reflection (n_reflections=1, up to 6 solutions, all three types); no attenuation"),
    "cfg5": dict(ice="southpole_2015", att="SP1", n_refl=0, n_freq=25, n_vertices=1_000_000, seed=5, dist="cyl", rmax=6000., zmin=-2700.,
                 antennas=np.array([[x, y, zz] for x in (np.arange(5) - 2) * 1500. for y in (np.arange(5) - 2) * 1500.
                                    for zz in (-145., -150., -155., -160.)]),
                 freqs=np.fft.rfftfreq(1022, 0.2), fmax=1.2, att_out="sparse",
                 text="cfg5 (BASELINE.json configs[4]): 1e6 vertices in r<6 km, z in [-2700,0] m x 100 channels (5x5 stations, 1.5 km pitch, "
                      "4 depths) = 1e8 pairs; southpole_2015; SP1; 512-bin grid 0-2.5 GHz, max_detector_freq 1.2 GHz, n_freq 25 -> 37 "
                      "integration frequencies; sparse attenuation output"),
}
CONFIGS["cfg4mb1"] = dict(CONFIGS["cfg4"], att="MB1", freqs=np.fft.rfftfreq(1022, 0.2), fmax=1.2, att_out="sparse",
                          text=CONFIGS["cfg4"]["text"].replace("no attenuation", "MB1 attenuation, 37 integration frequencies, sparse output")
                          .replace("cfg4 ", "cfg4mb1 "))


def vertices_of(cfg, n_vertices):
    """(3, Nv) SoA vertices of the configuration"""
    if cfg["dist"] == "T01":       # T01test_python_vs_cpp.py:12-23 (legacy numpy RNG)
        np.random.seed(cfg["seed"])
        r = np.random.triangular(50., 3000., 3000., n_vertices)
        phi = np.random.uniform(0, 2 * np.pi, n_vertices)
        z = np.random.uniform(0., cfg["zmin"], n_vertices)
    else:                          # uniform in a cylinder, EvtGen/generator.py:613-618
        rng = np.random.default_rng(cfg["seed"])
        r = np.sqrt(rng.uniform(0, cfg["rmax"] ** 2, n_vertices))
        phi = rng.uniform(0, 2 * np.pi, n_vertices)
        z = rng.uniform(cfg["zmin"], 0, n_vertices)
    return np.ascontiguousarray(np.array([r * np.cos(phi), r * np.sin(phi), z]))


def workload(n_vertices, config="cfg5"):
    """(V (3, Nv), A (3, Na), frequencies) -- kept for the scripts that import it"""
    cfg = CONFIGS[config]
    return vertices_of(cfg, n_vertices), np.ascontiguousarray(cfg["antennas"].T), cfg["freqs"]


# SURVEY.md 8(d) FLOP model of the REFERENCE algorithm (FP64 ops: add/mul 1, FMA 2, div/sqrt 4, transcendental 8)
def reference_model_flops(cfg, Fs):
    M = 1 + 2 * cfg["n_refl"]
    w_pair = 50 + M * 30 * 100 * (1 + (1 if cfg["n_refl"] else 0))       # G + M N_eval E
    c_f = 12 if cfg["att"] == "SP1" else 8
    w_att = (64 * (40 + Fs * c_f) + 8 * Fs) if cfg["att"] else 0          # Q (C_node + F C_f) + 8 F
    return w_pair, 600, w_att                                             # per pair, properties per solution, attenuation per solution


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax = float(r[2]); power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = [c for c in sm if smax and c > 0.3 * smax] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm)}


def config_dict(cfg, n_vertices, e2e_vertices, world):
    na = cfg["antennas"].shape[0]
    return {"workload": cfg["text"] + ("" if n_vertices == cfg["n_vertices"] else f" [run with {n_vertices} vertices]"),
            "pairs_per_step": n_vertices * na,
            "outputs": "n_sol,status,sol_offset per pair; per solution row (CSR): type,reflection,reflection_case,C0,C1,path_length,travel_time,"
                       "launch_vector,receive_vector,reflection_angle" + {None: "", "sparse": ",attenuation_sparse[Fs]",
                                                                           "dense": ",attenuation[F]"}[cfg["att_out"]],
            "l2": "inputs+outputs per step (>= GBs) far exceed the 126 MB L2; no explicit flush" if n_vertices * na >= 4_000_000
                  else "small workload: an 800 MB buffer is rewritten between timed steps to flush the 126 MB L2",
            "e2e_vertices": e2e_vertices,
            "parallelism": f"vertices sharded over {world} rank(s); no data-path collective in `value`; result gather timed separately (`gathered`)"}


# ------------------------------------------------------------------------------------------------------------------
# CPU arms
# ------------------------------------------------------------------------------------------------------------------
def pairs_of(V, A, n_pairs):
    na = A.shape[1]
    nv = max(1, -(-n_pairs // na))
    X1 = np.repeat(V[:, :nv].T, na, axis=0)[:n_pairs]
    X2 = np.tile(A.T, (nv, 1))[:n_pairs]
    return X1, X2


def cpu_port(cfg, V, A, n_pairs, threads):
    """the oracle port on `threads` host threads, reference quadrature tolerance (epsrel = 1e-2), same workload"""
    from oracle.oracle import Oracle
    X1, X2 = pairs_of(V, A, n_pairs)
    gl3 = None
    o = Oracle(cfg["ice"], attenuation_model=cfg["att"], n_reflections=cfg["n_refl"], n_freq=cfg["n_freq"] or 100, tight=False, gl3_table=gl3)
    t0 = time.perf_counter()
    out = o.trace(X1, X2, cfg["freqs"] if cfg["att"] else None, cfg["fmax"], n_threads=threads, dense=False)
    dt = time.perf_counter() - t0
    return {"value": len(X1) / dt, "unit": "pairs/s", "cores": threads, "kind": "port", "pairs": len(X1), "solutions": int(out["n_sol"].sum()),
            "seconds": dt, "what": "oracle/nrmc_oracle.c (C restatement of the reference's Python path, quad epsrel=1e-2), one thread per core"}


def cpu_python_reference(cfg, V, A, n_pairs, threads, numba):
    """the unmodified Python reference, one process per core (None when baseline/_ref did not travel)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pyref"))
    try:
        import ref_bench
        if not ref_bench.available() or (numba and cfg["n_refl"] > 0):
            return None
        X1, X2 = pairs_of(V, A, n_pairs)
        rcfg = dict(ice=cfg["ice"], att=cfg["att"], n_refl=cfg["n_refl"], n_freq=cfg["n_freq"], freqs=cfg["freqs"], fmax=cfg["fmax"])
        return ref_bench.run(rcfg, X1, X2, threads, numba=numba)
    except Exception as e:      # the reference arm must never take the bench line down
        return {"error": repr(e)[:300], "kind": "reference", "variant": "numba (warm)" if numba else "plain python"}


def cpu_baseline_all(cfg, V, A, threads, budget_s):
    """port + python reference (numba warm, plain) on bounded samples: ~budget_s seconds of CPU work each"""
    out = {}
    with_att = cfg["att"] is not None
    na = A.shape[1]
    cal = cpu_port(cfg, V, A, max(na, 100 * max(1, threads // 2)), threads)
    n_port = int(min(max(cal["value"] * budget_s, 1000), 2_000_000))
    out["port"] = cpu_port(cfg, V, A, n_port, threads)
    per_core = 4.0 if with_att else 150.0                 # pairs/s/core of the Python reference, to size the sample
    n_py = int(max(threads * 2, per_core * threads * budget_s * 0.5))
    for key, numba in (("python_numba", True), ("python_plain", False)):
        r = cpu_python_reference(cfg, V, A, n_py * (2 if numba else 1), threads, numba)
        if r is not None:
            out[key] = r
    return out


def headline_cpu(entries):
    """the reference's fastest own path that ran (numba, else plain), else the port"""
    for k in ("python_numba", "python_plain", "port"):
        if k in entries and "value" in entries[k]:
            return k, entries[k]
    return None, None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    nv_sample = min(cfg["n_vertices"], 20000)
    V, A, _ = workload(nv_sample, args.config)
    threads = len(os.sched_getaffinity(0))
    with_att = cfg["att"] is not None
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pyref"))
    have_py = False
    try:
        import ref_bench
        have_py = ref_bench.available()
    except Exception:
        pass
    extra = {}
    if have_py:
        # the reference's own fastest path: numba (its default when the C++ extension is missing, analyticraytracing.py:2024-2026)
        numba = cfg["n_refl"] == 0
        per_core = (8.0 if numba else 4.0) if with_att else (600.0 if numba else 100.0)
        sample_pairs = int(max(threads, per_core * threads * 3.0))          # ~3 s of CPU work per step
        kind, what = "reference", ("unmodified NuRadioMC Python reference (baseline/_ref), %s path, multiprocessing, one process per core"
                                   % ("numba (warm)" if numba else "plain"))

        def step(n):
            r = cpu_python_reference(cfg, V, A, n, threads, numba)
            if r is None or "value" not in r:
                raise RuntimeError(str(r))
            return r
    else:
        cal = cpu_port(cfg, V, A, 100 * max(1, threads), threads)
        sample_pairs = int(max(cal["value"] * 4.0, 1000))
        kind, what = "port", "oracle port (baseline/_ref absent)"

        def step(n):
            return cpu_port(cfg, V, A, n, threads)
    steps = max(1, min(args.steps, 8))                   # bounded: the whole run ends within a few minutes
    for _ in range(1 if have_py else min(args.warmup, 3)):
        step(max(threads, sample_pairs // 4))
    pairs, secs = 0, 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        r = step(sample_pairs)
        pairs += r["pairs"]
        secs += r.get("seconds_slowest_process", r.get("seconds", 0.0))
    wall = time.perf_counter() - t0
    value = pairs / secs              # compute time of the slowest process per step (pool start-up / jit warm-up excluded, as BASELINE.md 3 plans)
    extra["port"] = cpu_port(cfg, V, A, int(min(max(600.0 * threads * 3.0, 1000), 500_000)), threads) if have_py else None
    if have_py and cfg["n_refl"] == 0:
        extra["python_plain"] = cpu_python_reference(cfg, V, A, max(threads, sample_pairs // 2), threads, False)
    sample = (f"{sample_pairs} pairs/step x {steps} steps (first pairs, vertex-major, of the {args.config} workload); {what}; "
              f"rate = pairs / compute time of the slowest process (start-up and jit warm-up excluded; wall incl. them: {wall:.1f} s)")
    print(json.dumps({
        "impl": "reference", "metric": "ray-trace pairs/s (vertex x antenna pairs%s)" % (", with attenuation" if with_att else ""), "value": value,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": secs / steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(cfg, cfg["n_vertices"], None, 1),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": kind, "sample": sample,
                         **{k: v for k, v in extra.items() if v}},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg5", choices=sorted(CONFIGS))
    ap.add_argument("--vertices", type=int, default=0, help="0: the configuration's own number")
    ap.add_argument("--e2e-vertices", type=int, default=0, help="0: sized automatically from the host memory")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=8.0, help="seconds of CPU work per cpu_baseline entry")
    ap.add_argument("--gather", default="all", choices=["all", "p2p", "nccl", "none"])
    ap.add_argument("--gather-steps", type=int, default=0, help="0: min(steps, 5)")
    ap.add_argument("--gather-chunks", type=int, default=8, help="chunks per rank of the pipelined p2p-dma gather")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs CUDA devices (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from nuradiomc_b200 import distributed as nd
    numa = nd.bind_to_gpu_numa_node(local)        # before any pinned allocation: first touch puts the pages on the GPU's node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from nuradiomc_b200.SignalProp import propagation
    from nuradiomc_b200.SignalProp.analyticraytracing import measure_fp64_peak
    from nuradiomc_b200.utilities import medium

    cfg = CONFIGS[args.config]
    n_vertices = args.vertices or cfg["n_vertices"]
    V, A, ff = workload(n_vertices, args.config)
    na = A.shape[1]
    lo, hi = nd.shard_bounds(n_vertices, world, rank)
    Vr = np.ascontiguousarray(V[:, lo:hi])
    n_pairs_rank = (hi - lo) * na
    n_pairs_total = n_vertices * na
    with_att = cfg["att"] is not None
    rt = propagation.get_propagation_module("analytic")(medium.get_ice_model(cfg["ice"]), attenuation_model=cfg["att"] or "SP1",
                                                          n_frequencies_integration=cfg["n_freq"], n_reflections=cfg["n_refl"], device=local)
    S = rt.get_number_of_raytracing_solutions()
    dv, da = torch.tensor(Vr, device=dev), torch.tensor(A, device=dev)
    kw = dict(outer=True, compact=True)
    if with_att:
        kw.update(frequency=ff, max_detector_freq=cfg["fmax"], attenuation=cfg["att_out"])
    # rows the compact arrays hold: the workload's solutions per pair plus 15 % (checked through the synchronising stats call below)
    probe = rt.trace_batch_device(dv[:, :max(1, min(hi - lo, 2000))].contiguous(), da, outer=True, compact=True, sync_stats=True)
    sol_per_pair = probe.stats["n_solutions"] / max(probe.stats["n_pairs"], 1)
    rows_per_pair = min(float(S), 1.15 * sol_per_pair + 0.05)
    kw["row_capacity"] = int(n_pairs_rank * rows_per_pair) + 1024
    del probe
    flush = torch.empty(100_000_000, dtype=torch.float64, device=dev) if n_pairs_total < 4_000_000 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    out = None
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()          # sampled from the warm-up on (same load): short timed regions still get several samples under load
    for _ in range(args.warmup):
        out = rt.trace_batch_device(dv, da, out=out, **kw)
    barrier()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            out = rt.trace_batch_device(dv, da, out=out, **kw)
        e1.record()
        barrier()
        ms_local = e0.elapsed_time(e1)
    else:                       # small workloads: flush the L2 between steps, time every step on its own
        ms_local = 0.0
        barrier()
        for _ in range(args.steps):
            flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = rt.trace_batch_device(dv, da, out=out, **kw)
            e1.record()
            torch.cuda.synchronize(dev)
            ms_local += e0.elapsed_time(e1)
        barrier()
    ms = torch.tensor([ms_local], device=dev, dtype=torch.float64)
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())

    # per-kernel durations (CUDA events inside the library, on the launching stream) from one extra, untimed pass; the
    # synchronising call also reports a row_capacity overflow
    meas = rt.trace_batch_device(dv, da, out=out, sync_stats=True, **kw).stats
    launches_per_step = meas["n_launches"]
    n_sol_rank = torch.tensor([float(meas["n_solutions"])], device=dev, dtype=torch.float64)
    tt_local = out["travel_time"][:int(meas["n_solutions"])].sum().reshape(1) if "travel_time" in out else torch.zeros(1, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(n_sol_rank)
    n_solutions_total = int(n_sol_rank.item())

    # ---- gathered: the same step with the result rows landing on rank 0 inside the timed region ------------------
    gathered = None
    if args.gather != "none" and args.config in ("cfg5", "cfg3", "cfg4mb1", "cfg2", "cfg4", "cfg1"):
        gathered = run_gathered(args, rt, dv, da, kw, out, meas, n_pairs_rank, n_pairs_total, rows_per_pair, world, rank, dev, barrier, tt_local, cfg)

    # ---- e2e: host buffers through the public API, copies inside the timed region --------------------------------
    free_host = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    Fs = len(out.frequencies_sparse) if with_att else 0
    F = len(ff) if with_att else 0
    K1 = cfg["n_refl"] + 1
    att_bytes = {None: 0, "sparse": Fs * 8, "dense": F * 8}[cfg["att_out"]]
    bytes_per_row = 3 + 4 * 8 + 2 * 24 + 8 * K1 + att_bytes
    bytes_per_pair_out = 16 + S * bytes_per_row                       # pinned capacity: S rows per pair
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    e2e_v = args.e2e_vertices or int(min(hi - lo, max(100, 0.2 * free_host / local_world / (bytes_per_pair_out * na))))
    e2e_v = max(1, min(e2e_v, hi - lo))
    del out
    torch.cuda.empty_cache()
    Vh = np.ascontiguousarray(Vr[:, :e2e_v].T)
    Ah = np.ascontiguousarray(A.T)
    hres = None
    kwh = dict(outer=True, pinned=True, compact=True)
    if with_att:
        kwh.update(frequency=ff, max_detector_freq=cfg["fmax"], attenuation=cfg["att_out"])
    for _ in range(2):
        hres = rt.trace_batch(Vh, Ah, out=hres, **kwh)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        hres = rt.trace_batch(Vh, Ah, out=hres, **kwh)
    torch.cuda.synchronize(dev)
    dt_e2e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt_e2e, op=dist.ReduceOp.MAX)
    e2e_pairs_total = e2e_v * na * world
    e2e_value = e2e_pairs_total * e2e_steps / float(dt_e2e.item())
    h2d, d2h = int(hres.stats["h2d_bytes"]), int(hres.stats["d2h_bytes"])
    # the ceiling of the box: the same number of bytes as one plain D2H copy per rank into pinned memory, all ranks at once
    probe_bytes = max(1 << 20, min(d2h, 4 << 30))
    src = torch.empty(probe_bytes, dtype=torch.uint8, device=dev)
    dst = torch.empty(probe_bytes, dtype=torch.uint8, pin_memory=True)
    dst.copy_(src)
    barrier()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    dt_probe = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt_probe, op=dist.ReduceOp.MAX)
    d2h_probe_gbs = probe_bytes * reps * world / float(dt_probe.item()) / 1e9
    del src, dst
    numa_all = [numa]
    if world > 1:
        numa_all = [None] * world
        dist.all_gather_object(numa_all, numa)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms_total / args.steps
    value = n_pairs_total / (ms_per_step * 1e-3)
    fp64_peak, clk_nominal = measure_fp64_peak(local, 1.0)
    n_sol_launch = meas["n_solutions"]
    peaks, hw = {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    try:   # executed FP64 FLOPs / DRAM bytes per unit from the committed ncu captures (profiles/summarize_ncu.py)
        hw = json.load(open(os.path.join(ROOT, "profiles", "hw_counts.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    out_bytes = n_pairs_rank * 16 + n_sol_launch * bytes_per_row          # n_sol, status, sol_offset per pair + the rows
    mk = meas["ms_kernel"]
    w_pair, w_props, w_att = reference_model_flops(cfg, Fs)
    refl = cfg["n_refl"] > 0
    if not with_att:
        att_kernel = None
    elif cfg["att"] in ("MB1", "GL2") and (not refl or cfg["att_out"] == "sparse"):
        att_kernel = "K_att_sep"
    else:
        att_kernel = {"SP1": "K_att_sp1", "GL1": "K_att_gl1"}.get(cfg["att"], "K_att") if not refl else "K_att"
    names = {"classify": "K_classify_m" if refl else "K_classify", "hump": "K_hump_m" if refl else "K_hump",
             "roots": "K_roots_m" if refl else "K_roots", "attenuation_main": att_kernel}

    def kernel_entry(key, units, unit_name):
        """one kernel: live CUDA-event duration x the FP64 FLOPs / DRAM bytes per unit ncu counted for it (profiles/hw_counts.json)"""
        name, t = names[key], mk[key]
        e = {"kernel": name, "kernel_ms": t, "units_per_launch": units, "unit": unit_name, "share_of_step": t / max(meas["ms_total"], 1e-9)}
        h = hw.get(f"{args.config}:{name}") or hw.get(name)
        if h and t > 0:
            e["executed_fp64_flops_per_unit"] = h["fp64_flops_per_unit"]
            e["achieved"] = units * h["fp64_flops_per_unit"] / (t * 1e-3) / 1e12
            e["frac"] = e["achieved"] / fp64_peak
            e["traffic"] = units * h["dram_bytes_per_unit"]
            e["counts_from"] = h.get("report")
            # pipe utilisation of the same capture: FP64 instructions of any kind (a DADD or DMUL occupies the pipe like a DFMA but
            # counts one FLOP) / the pipe's issue rate, lanes active per warp instruction, resident warps
            for k_src, k_dst in (("fp64_pipe_active_pct", "ncu_fp64_pipe_active_pct"), ("threads_per_warp_instruction", "ncu_threads_per_warp_instruction"),
                                 ("achieved_occupancy_pct", "ncu_achieved_occupancy_pct"), ("registers", "registers")):
                if k_src in h:
                    e[k_dst] = h[k_src]
        return e
    kernels = {"classify": kernel_entry("classify", n_pairs_rank, "pairs"), "hump": kernel_entry("hump", n_pairs_rank, "pairs"),
               "roots": kernel_entry("roots", n_sol_launch, "solutions")}
    if with_att:
        kernels["attenuation_main"] = kernel_entry("attenuation_main", n_sol_launch, "solutions")
    kernels["other_ms"] = mk["attenuation_other"]
    top = max((k for k in kernels.values() if isinstance(k, dict)), key=lambda k: k["kernel_ms"])
    step_flops = sum(k["units_per_launch"] * k.get("executed_fp64_flops_per_unit", 0.0) for k in kernels.values() if isinstance(k, dict))
    model_flops = n_pairs_rank * w_pair + n_sol_launch * (w_props + w_att)
    roofline = {
        "bound": "fp64", "kernel": top["kernel"], "unit": "TFLOP/s", "peak": fp64_peak,
        "achieved": top.get("achieved"), "frac": top.get("frac"), "traffic": top.get("traffic"),
        "kernel_ms": top["kernel_ms"], "share_of_step": top["share_of_step"],
        "definition": "achieved = FP64 FLOPs the kernel EXECUTES per unit (ncu: DFMA x 2 + DMUL + DADD, predicated-on threads; "
                      "profiles/hw_counts.json, DESIGN.md section 5) x units per launch / live CUDA-event duration of the launch; "
                      "frac = achieved / peak <= 1.  The FLOP model of the REFERENCE algorithm (SURVEY.md 8(d)) is kept as algorithmic_speedup.",
        "peak_source": "measured live: independent DFMA chains on all SMs for 1 s (nrmc_rt_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
        "peak_clocks": {"measured_tflops": fp64_peak, "sm_mhz_nominal": clk_nominal, "sm_mhz_under_load": (clk or {}).get("sm_mhz"),
                        "sm_max_mhz": (clk or {}).get("sm_max_mhz"), "dfma_per_clk_per_sm": fp64_peak * 1e12 / 2 / 148 / max(clk_nominal * 1e6, 1.0)},
        "step": {"executed_tflops": step_flops / (meas["ms_total"] * 1e-3) / 1e12, "frac": step_flops / (meas["ms_total"] * 1e-3) / 1e12 / fp64_peak,
                 "executed_fp64_flops_per_pair": step_flops / n_pairs_rank},
        "algorithmic_speedup": {"reference_model_flops_per_pair": model_flops / n_pairs_rank,
                                "executed_flops_per_pair": step_flops / n_pairs_rank if step_flops else None,
                                "ratio": model_flops / step_flops if step_flops else None,
                                "reference_model_tflops": model_flops / (meas["ms_total"] * 1e-3) / 1e12,
                                "model": {"per_pair": w_pair, "properties_per_solution": w_props, "attenuation_per_solution": w_att}},
        "kernels": kernels,
        "hbm": {"achieved": out_bytes / (meas["ms_total"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": out_bytes / (meas["ms_total"] * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_pair": out_bytes / n_pairs_rank}}
    line = {
        "metric": "ray-trace pairs/s (vertex x antenna pairs%s)" % (", with attenuation" if with_att else ""), "value": value, "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(cfg, n_vertices, e2e_v * world, world),
        "solutions_per_s": n_solutions_total / (ms_per_step * 1e-3), "solutions_per_pair": n_solutions_total / n_pairs_total,
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                "pairs_per_step": e2e_pairs_total, "steps": e2e_steps, "timer": "host perf_counter around the blocking API call, max over ranks",
                "layout": "host numpy in, pinned numpy out, per-solution (CSR) rows: empty slots are not copied",
                "d2h_gbs": d2h * world * e2e_steps / float(dt_e2e.item()) / 1e9,
                "d2h_probe_gbs": d2h_probe_gbs,
                "d2h_probe": "one plain cudaMemcpy D2H of the same bytes per rank into pinned memory, all ranks at once: the box's ceiling for this step",
                "numa": numa_all},
        "gpu_launches": int(launches_per_step * args.steps * world),
        "roofline": roofline,
    }
    if gathered is not None:
        line["gathered"] = gathered
        best = gathered.get("headline")
        if best:
            line["value_gathered"] = gathered[best]["value"]
    if world == 1 and not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0))
        Vs, As, _ = workload(min(n_vertices, 20000), args.config)
        entries = cpu_baseline_all(cfg, Vs, As, threads, args.cpu_budget)
        key, head = headline_cpu(entries)
        line["cpu_baseline"] = {"value": head["value"], "unit": "pairs/s", "cores": threads, "kind": head["kind"],
                                "sample": f"{key}: first {head['pairs']} pairs (vertex-major) of the same workload, {head.get('solutions')} solutions; "
                                          "python_* = the unmodified NuRadioMC Python reference (one process per core; rate over the slowest process, "
                                          "start-up / jit warm-up excluded); port = oracle/nrmc_oracle.c with the reference's quadrature tolerance",
                                **entries}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_gathered(args, rt, dv, da, kw, out_local, meas, n_pairs_rank, n_pairs_total, rows_per_pair, world, rank, dev, barrier, tt_local, cfg):
    """value_gathered: K steps whose timed region contains the movement of the compact result rows to rank 0"""
    import torch
    import torch.distributed as dist
    from nuradiomc_b200 import distributed as nd
    steps = args.gather_steps or max(2, min(args.steps, 5))
    row_names = [k for k in out_local.keys() if k not in nd.PER_PAIR]
    Fs = len(out_local.frequencies_sparse) if out_local.frequencies_sparse is not None else 0
    F = len(cfg["freqs"]) if cfg["freqs"] is not None else 0
    res = {"steps": steps, "dst_rank": 0,
           "layout": "rank 0 holds n_sol/status/sol_offset of all pairs (global order) and every result row; p2p: rank r's rows in the segment "
                     "[row_base[r], row_base[r] + n_rows[r]); nccl: contiguous CSR"}
    kw_g = {k: v for k, v in kw.items() if k not in ("compact", "row_capacity")}

    def timed(fn, sync_each=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        t_host = time.perf_counter() - t0
        t = torch.tensor([e0.elapsed_time(e1), t_host * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t[0].item()) / steps, float(t[1].item()) / steps

    variants = []
    if args.gather in ("all", "p2p"):
        if world > 1:
            variants.append(("p2p-dma", "records+sparse", row_names))
        variants.append(("p2p", "records+sparse", row_names))
        if any(n.startswith("attenuation") for n in row_names):
            variants.append(("p2p", "records", [n for n in row_names if not n.startswith("attenuation")]))
    for kind, what, names in variants:
        key = f"{kind}:{what}"
        try:
            pg = nd.P2PGather(rt, n_pairs_rank, names=names, Fs=Fs, F=F, rows_per_pair=rows_per_pair + 1024.0 / max(n_pairs_rank, 1))
            run = (lambda: pg.trace_pushed(dv, da, n_chunks=args.gather_chunks, **kw_g)) if kind == "p2p-dma" else (lambda: pg.trace(dv, da, **kw_g))
            for _ in range(2):
                run()
            pg.finish()
            ms_dev, ms_host = timed(run)
            n_rows = int(meas["n_solutions"])
            nv = torch.tensor([float(pg.nvlink_bytes(n_rows))], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(nv)
            entry = {"value": n_pairs_total / (max(ms_dev, ms_host) * 1e-3), "unit": "pairs/s", "ms_per_step": max(ms_dev, ms_host),
                     "gather": what,
                     "variant": ("p2p-dma: chunk-pipelined -- while chunk c+1 computes, the copy engines push the rows of chunk c into rank 0's "
                                 "HBM (peer-to-peer cudaMemcpyAsync into the mapped block, %d chunks per rank)" % args.gather_chunks) if kind == "p2p-dma"
                     else "p2p: kernels store into rank 0's HBM through NVLink peer mappings (no collective, no staging)",
                     "nvlink_bytes_per_step": int(nv.item()), "rank0_block_bytes": pg.block_bytes}
            entry["nvlink_gbs_into_rank0"] = entry["nvlink_bytes_per_step"] / (entry["ms_per_step"] * 1e-3) / 1e9
            # integrity: a checksum of the gathered rows against the sum of the ranks' local checksums
            if "travel_time" in names:
                tt = tt_local.clone()
                if world > 1:
                    dist.all_reduce(tt)
                if rank == 0:
                    A = pg.arrays
                    n_sol = A["n_sol"].to(torch.int64)
                    total = 0.0
                    for r in range(world):
                        cnt = int(n_sol[pg.pair_base[r]:pg.pair_base[r + 1]].sum().item())
                        total += float(A["travel_time"][pg.row_base[r]:pg.row_base[r] + cnt].sum().item())
                    entry["checksum_ok"] = bool(abs(total - float(tt.item())) <= 1e-9 * abs(float(tt.item())))
                    entry["rows_on_rank0"] = int(n_sol.sum().item())
            pg.close()
            del pg
            torch.cuda.empty_cache()
            res[key] = entry
        except Exception as e:
            res[key] = {"error": repr(e)[:400]}
            if world > 1:
                try:
                    dist.barrier()
                except Exception:
                    pass
    if args.gather in ("all", "nccl") and world > 1:
        key = "nccl:records+sparse"
        try:
            state = {"res": None}

            def step():
                state["res"] = rt.trace_batch_device(dv, da, out=state["res"], **kw)
                nd.gather_compact_result(dict(state["res"]), n_pairs_rank)
            step()
            ms_dev, ms_host = timed(step)
            n_rows = int(meas["n_solutions"])
            per_row = sum(int(np.prod(out_local[n].shape[1:])) * out_local[n].element_size() for n in row_names)
            nb = torch.tensor([float(0 if rank == 0 else n_rows * per_row + n_pairs_rank * 16)], device=dev, dtype=torch.float64)
            dist.all_reduce(nb)
            res[key] = {"value": n_pairs_total / (max(ms_dev, ms_host) * 1e-3), "unit": "pairs/s", "ms_per_step": max(ms_dev, ms_host),
                        "gather": "records+sparse", "nccl_bytes_per_step": int(nb.item()),
                        "variant": "nccl: kernels, then the row counts (all_gather + host read), then one grouped ncclSend/ncclRecv per array"}
            res[key]["nccl_gbs_into_rank0"] = res[key]["nccl_bytes_per_step"] / (res[key]["ms_per_step"] * 1e-3) / 1e9
            state.clear()
            torch.cuda.empty_cache()
        except Exception as e:
            res[key] = {"error": repr(e)[:400]}
    ok = [k for k in ("p2p-dma:records+sparse", "p2p:records+sparse", "nccl:records+sparse", "p2p:records") if k in res and "value" in res[k]]
    full = [k for k in ok if res[k]["gather"] == "records+sparse"] or ok
    if full:
        res["headline"] = max(full, key=lambda k: res[k]["value"])
        h = res[res["headline"]]
        moved = h.get("nvlink_bytes_per_step", h.get("nccl_bytes_per_step", 0))
        res["limiter"] = ("single GPU: nothing crosses NVLink" if world == 1 else
                          "ingress of rank 0: %.1f GB per step land on one GPU (%.0f GB/s achieved; NVLink 5 gives ~900 GB/s per direction per GPU); "
                          "the kernels alone take %.1f ms" % (moved / 1e9, moved / (h["ms_per_step"] * 1e-3) / 1e9, meas["ms_total"]))
    return res


if __name__ == "__main__":
    main()
