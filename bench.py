#!/usr/bin/env python
"""
bench.py -- headline benchmark of the B200-native analytic ray tracer.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--vertices NV]

Metric (BASELINE.json): ray-trace pairs/s (vertex x antenna pairs fully processed: solutions, type, C0/C1, launch/receive
vectors, path length, travel time AND the attenuation factors), whole job over N GPUs of one node.

Workload (BASELINE.json configs[4], the configuration the metric is quoted on; SURVEY.md 8(d) "cfg5"): 1e6 vertices uniform in a
cylinder r < 6 km, z in [-2700, 0] m (seed 5) x 100 channels (5 x 5 stations on a 1.5 km grid, 4 channels per station at
-145/-150/-155/-160 m) = 1e8 pairs, southpole_2015 ice, SP1 attenuation on the 512-bin 0-2.5 GHz grid with
max_detector_freq = 1.2 GHz and n_freq = 25 -> 37 integration frequencies, sparse attenuation output.
A "step" is one pass of the hot path over all pairs.  N > 1: the vertices are sharded over the ranks (one process per
GPU, no data-path collective; strong scaling: the total stays 1e8 pairs).

  value  -- inputs resident in HBM when the timed region starts, outputs left in HBM (torch CUDA tensors).
  e2e    -- the same pass through the public API with HOST buffers: numpy in, pinned numpy out, H2D and D2H copies
            inside the timed region (on a slice of the workload sized to the host memory, see config.e2e_vertices).
--impl reference times the CPU oracle port (oracle/, a restatement of the reference's algorithm with the reference's
quadrature tolerance) on all host cores on a bounded sample of the same workload.
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
# stdout carries ONE JSON line.  NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION and honours NCCL_DEBUG_FILE only
# above that level: raise VERSION to WARN and send the log to stderr.
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
    os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

N_VERTICES = 1_000_000
ICE, ATT_MODEL, N_FREQ, FMAX = "southpole_2015", "SP1", 25, 1.2
# SURVEY.md 8(d) FLOP model (FP64 ops: add/mul 1, FMA 2, div/sqrt 4, transcendental 8)
W_SOLVE_PER_PAIR = 50 + 1 * 30 * 100            # G + M * N_eval * E            (M = 1 mode)
W_PROPS_PER_SOLUTION = 600                      # P
W_ATT_PER_SOLUTION = 64 * (40 + 37 * 12) + 8 * 37   # Q (C_node + F C_f) + 8 F = 31.3 kFLOP (SP1, F = 37)


def workload(n_vertices):
    rng = np.random.default_rng(5)
    r = np.sqrt(rng.uniform(0, 6000. ** 2, n_vertices))
    phi = rng.uniform(0, 2 * np.pi, n_vertices)
    z = rng.uniform(-2700., 0, n_vertices)
    V = np.array([r * np.cos(phi), r * np.sin(phi), z])                       # (3, Nv) SoA
    g = (np.arange(5) - 2) * 1500.
    A = np.array([[x, y, zz] for x in g for y in g for zz in (-145., -150., -155., -160.)]).T.copy()   # (3, 100)
    ff = np.fft.rfftfreq(1022, 0.2)                                            # 512 bins, 0 .. 2.5 GHz
    return np.ascontiguousarray(V), A, ff


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


def cpu_reference(V, A, ff, n_pairs, threads):
    """the oracle port on `threads` host threads, reference quadrature tolerance (epsrel = 1e-2), same workload"""
    from oracle.oracle import Oracle
    na = A.shape[1]
    nv = max(1, n_pairs // na)
    X1 = np.repeat(V[:, :nv].T, na, axis=0)
    X2 = np.tile(A.T, (nv, 1))
    o = Oracle(ICE, attenuation_model=ATT_MODEL, n_freq=N_FREQ, tight=False)
    t0 = time.perf_counter()
    out = o.trace(X1, X2, ff, FMAX, n_threads=threads, dense=False)
    dt = time.perf_counter() - t0
    return len(X1) / dt, len(X1), int(out["n_sol"].sum()), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    V, A, ff = workload(20000)
    threads = len(os.sched_getaffinity(0))
    sample_pairs = 100 * max(1, int(round(40 * threads)) // 1)      # ~ 4 s per step at ~600 pairs/s/thread
    for _ in range(args.warmup):
        cpu_reference(V, A, ff, max(100, sample_pairs // 10), threads)
    t0 = time.perf_counter()
    pairs = 0
    for _ in range(args.steps):
        _, n, _, _ = cpu_reference(V, A, ff, sample_pairs, threads)
        pairs += n
    dt = time.perf_counter() - t0
    value = pairs / dt
    sample = f"{sample_pairs} pairs/step ({sample_pairs // 100} vertices x 100 channels of the cfg5 workload), oracle port, quad epsrel=1e-2"
    print(json.dumps({
        "impl": "reference", "metric": "ray-trace pairs/s (vertex x antenna pairs, with attenuation)", "value": value,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(sample_pairs // 100, None),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def config_dict(n_vertices, e2e_vertices):
    return {"workload": "cfg5 (BASELINE.json configs[4]): %d vertices in r<6 km, z in [-2700,0] m x 100 channels (5x5 stations, "
                        "1.5 km pitch, 4 depths) = %d pairs; southpole_2015; SP1; 512-bin grid 0-2.5 GHz, max_detector_freq 1.2 GHz, "
                        "n_freq 25 -> 37 integration frequencies; sparse attenuation output" % (n_vertices, n_vertices * 100),
            "pairs_per_step": n_vertices * 100, "outputs": "n_sol,status,sol_offset per pair; per solution row (CSR): type,reflection,"
            "reflection_case,C0,C1,path_length,travel_time,launch_vector,receive_vector,reflection_angle,attenuation_sparse[37]",
            "l2": "inputs+outputs per step (>= GBs) far exceed the 126 MB L2; no explicit flush",
            "e2e_vertices": e2e_vertices, "parallelism": "vertices sharded over ranks, no data-path collective"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vertices", type=int, default=N_VERTICES)
    ap.add_argument("--e2e-vertices", type=int, default=0, help="0: sized automatically from the host memory")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from nuradiomc_b200.SignalProp import propagation
    from nuradiomc_b200.SignalProp.analyticraytracing import measure_fp64_peak
    from nuradiomc_b200.distributed import shard_bounds
    from nuradiomc_b200.utilities import medium

    V, A, ff = workload(args.vertices)
    lo, hi = shard_bounds(args.vertices, world, rank)
    Vr = np.ascontiguousarray(V[:, lo:hi])
    n_pairs_rank = (hi - lo) * A.shape[1]
    n_pairs_total = args.vertices * A.shape[1]
    rt = propagation.get_propagation_module("analytic")(medium.get_ice_model(ICE), attenuation_model=ATT_MODEL,
                                                          n_frequencies_integration=N_FREQ, device=local)
    dv, da = torch.tensor(Vr, device=dev), torch.tensor(A, device=dev)
    kw = dict(outer=True, frequency=ff, max_detector_freq=FMAX, attenuation="sparse", compact=True)

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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = rt.trace_batch_device(dv, da, out=out, **kw)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    n_sol_rank = out["n_sol"].sum().to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(n_sol_rank)
    n_solutions_total = int(n_sol_rank.item())

    # per-kernel durations (CUDA events inside the library, on the launching stream) from one extra, untimed pass
    meas = rt.trace_batch_device(dv, da, out=out, sync_stats=True, **kw).stats
    launches_per_step = meas["n_launches"]

    # ---- e2e: host buffers through the public API, copies inside the timed region --------------------------------
    free_host = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    bytes_per_pair_out = 4 + 4 + 3 * 2 + 4 * 16 + 2 * 48 + 16 + 2 * 37 * 8
    e2e_v = args.e2e_vertices or int(min(hi - lo, max(1000, 0.2 * free_host / world / (bytes_per_pair_out * A.shape[1]))))
    e2e_v = min(e2e_v, hi - lo)
    Vh = np.ascontiguousarray(Vr[:, :e2e_v].T)
    Ah = np.ascontiguousarray(A.T)
    hres = None
    kwh = dict(outer=True, frequency=ff, max_detector_freq=FMAX, attenuation="sparse", pinned=True, compact=True)
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
    e2e_pairs_total = e2e_v * A.shape[1] * world
    e2e_value = e2e_pairs_total * e2e_steps / float(dt_e2e.item())
    h2d, d2h = hres.stats["h2d_bytes"], hres.stats["d2h_bytes"]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms_total / args.steps
    value = n_pairs_total / (ms_per_step * 1e-3)
    fp64_peak, _ = measure_fp64_peak(local, 1.0)
    n_sol_launch = meas["n_solutions"] if meas["n_solutions"] else n_solutions_total / world
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
    bytes_per_row = 3 + 4 * 8 + 2 * 24 + 8 + 37 * 8                  # type/reflection/case, C0/C1/path/time, vectors, angle, 37 factors
    out_bytes = n_pairs_rank * (4 + 4 + 8) + n_sol_launch * bytes_per_row     # n_sol, status, sol_offset per pair + the rows
    mk = meas["ms_kernel"]

    def kernel_entry(name, ms, units, alg_flops_per_unit, unit_name):
        """roofline numbers of one kernel: algorithmic (SURVEY.md 8(d) model) and executed (ncu counters) FP64 rates"""
        e = {"kernel_ms": ms, "units_per_launch": units, "unit": unit_name, "share_of_step": ms / max(meas["ms_total"], 1e-9)}
        if alg_flops_per_unit is not None:
            e["algorithmic_flops_per_unit"] = alg_flops_per_unit
            e["achieved"] = units * alg_flops_per_unit / (ms * 1e-3) / 1e12
            e["frac"] = e["achieved"] / fp64_peak
        h = hw.get(name)
        if h:
            e["executed_fp64_flops_per_unit"] = h["fp64_flops_per_unit"]
            e["executed_tflops"] = units * h["fp64_flops_per_unit"] / (ms * 1e-3) / 1e12
            e["executed_frac"] = e["executed_tflops"] / fp64_peak
            e["traffic"] = units * h["dram_bytes_per_unit"]
        return e

    k_att = kernel_entry("K_att_sp1", mk["attenuation_main"], n_sol_launch, W_ATT_PER_SOLUTION, "solutions")
    k_cls = kernel_entry("K_classify", mk["classify"], n_pairs_rank, None, "pairs")
    k_hmp = kernel_entry("K_hump", mk["hump"], n_pairs_rank, None, "pairs")
    k_rts = kernel_entry("K_roots", mk["roots"], n_sol_launch, None, "solutions")
    solve_flops = n_pairs_rank * W_SOLVE_PER_PAIR + n_sol_launch * W_PROPS_PER_SOLUTION
    solver = {"kernel_ms": meas["ms_solve"], "achieved": solve_flops / (meas["ms_solve"] * 1e-3) / 1e12,
              "frac": solve_flops / (meas["ms_solve"] * 1e-3) / 1e12 / fp64_peak,
              "algorithmic_flops_per_pair": W_SOLVE_PER_PAIR, "algorithmic_flops_per_solution": W_PROPS_PER_SOLUTION,
              "share_of_step": meas["ms_solve"] / max(meas["ms_total"], 1e-9),
              "K_classify": k_cls, "K_hump": k_hmp, "K_roots": k_rts}
    if all("executed_tflops" in k for k in (k_cls, k_hmp, k_rts)):
        ex = sum(k["executed_tflops"] * k["kernel_ms"] for k in (k_cls, k_hmp, k_rts)) / max(meas["ms_solve"], 1e-9)
        solver["executed_tflops"], solver["executed_frac"] = ex, ex / fp64_peak
    roofline = {"bound": "fp64", "kernel": "K_att_sp1 (attenuation integral, SP1 moment form, thread per solution)",
                "achieved": k_att["achieved"], "peak": fp64_peak, "unit": "TFLOP/s", "frac": k_att["frac"],
                "peak_source": "measured live: independent DFMA chains on all SMs (nrmc_rt_measure_fp64_peak); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "algorithmic_flops_per_solution": W_ATT_PER_SOLUTION, "solutions_per_launch": n_sol_launch,
                "kernel_ms": k_att["kernel_ms"], "traffic": k_att.get("traffic"), "share_of_step": k_att["share_of_step"],
                "note": "achieved/frac use SURVEY.md 8(d)'s FLOP model of the REFERENCE algorithm (64 nodes x 37 frequencies x exp per "
                        "solution). The kernel integrates the same quantity with 12-24 nodes and frequency-independent moments, so it "
                        "executes ~6x fewer FLOPs: frac > 1 is an algorithmic gain, not a hardware rate. executed_* are the FP64 "
                        "FLOPs counted by ncu (DFMA x2 + DMUL + DADD) over the live duration: the true pipe utilisation.",
                "executed_fp64_flops_per_solution": k_att.get("executed_fp64_flops_per_unit"),
                "executed_tflops": k_att.get("executed_tflops"), "executed_frac": k_att.get("executed_frac"),
                "solver": solver,
                "hbm": {"achieved": out_bytes / (meas["ms_total"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": out_bytes / (meas["ms_total"] * 1e-3) / 1e9 / hbm_peak,
                        "algorithmic_bytes_per_pair": out_bytes / n_pairs_rank}}
    line = {
        "metric": "ray-trace pairs/s (vertex x antenna pairs, with attenuation)", "value": value, "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.vertices, e2e_v * world),
        "solutions_per_s": n_solutions_total / (ms_per_step * 1e-3), "solutions_per_pair": n_solutions_total / n_pairs_total,
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
                "pairs_per_step": e2e_pairs_total, "steps": e2e_steps, "timer": "host perf_counter around the blocking API call",
                "layout": "host numpy in, pinned numpy out, per-solution (CSR) rows: empty slots are not copied"},
        "gpu_launches": int(launches_per_step * args.steps * world),
        "roofline": roofline,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0))
        rate0, _, _, _ = cpu_reference(V, A, ff, 100 * max(1, threads // 2), threads)          # calibration
        sample_pairs = int(min(max(rate0 * 15, 1000), 2_000_000)) // 100 * 100                 # ~15 s of CPU work
        rate, n, nsol, dt = cpu_reference(V, A, ff, sample_pairs, threads)
        line["cpu_baseline"] = {"value": rate, "unit": "pairs/s", "cores": threads, "kind": "port",
                                "sample": f"first {n} pairs ({n // 100} vertices x 100 channels) of the same workload, {dt:.1f} s, "
                                          f"{nsol} solutions; oracle port with the reference's quadrature tolerance (epsrel=1e-2)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
