#!/usr/bin/env python
"""
Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into the small tracked files under profiles/:
  python profiles/summarize_ncu.py <tag> <units_per_launch> <report.ncu-rep> [...]   ->  profiles/<tag>_<kernel>.json / .txt
`units_per_launch` = pairs (solver kernels) or solutions (attenuation kernels) the profiled launch processed; the JSON
holds executed FP64 FLOPs and DRAM bytes per unit, which bench.py combines with its live CUDA-event durations.
Runs wherever the `ncu` CLI is installed (no GPU needed to read a report).
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def raw(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return [dict(zip(rows[0], r)) for r in rows[2:]], dict(zip(rows[0], rows[1]))


def f(d, k, default=None):
    try:
        return float(d[k].replace(",", ""))
    except (KeyError, ValueError):
        return default


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return value * scale.get(unit, 1)


def summarise(report, units):
    kernels, unit_row = raw(report)
    d = kernels[0]
    name = re.sub(r"^void ", "", d["Kernel Name"]).split("(")[0].split("<")[0]
    ms = f(d, "gpu__time_duration.sum")
    if unit_row.get("gpu__time_duration.sum") == "us":
        ms /= 1e3
    cycles = f(d, "smsp__cycles_elapsed.avg") or f(d, "sm__cycles_elapsed.avg")
    per_cycle = {op: f(d, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed", 0.0) for op in ("dfma", "dmul", "dadd")}
    flops = (2 * per_cycle["dfma"] + per_cycle["dmul"] + per_cycle["dadd"]) * cycles
    rd = to_bytes(f(d, "dram__bytes_read.sum", 0.0), unit_row.get("dram__bytes_read.sum", "byte"))
    wr = to_bytes(f(d, "dram__bytes_write.sum", 0.0), unit_row.get("dram__bytes_write.sum", "byte"))
    s = {
        "kernel": name, "report": os.path.basename(report), "units_per_launch": units, "duration_ms": ms,
        "grid": d.get("launch__grid_size"), "block": d.get("launch__block_size"), "registers": d.get("launch__registers_per_thread"),
        "fp64_flops_executed": flops, "fp64_flops_per_unit": flops / units, "fp64_tflops": flops / (ms * 1e-3) / 1e12,
        "dram_bytes": rd + wr, "dram_bytes_per_unit": (rd + wr) / units, "dram_gbs": (rd + wr) / (ms * 1e-3) / 1e9,
        "warp_instructions": f(d, "smsp__inst_executed.sum"),
        "threads_per_warp_instruction": f(d, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "fp64_pipe_active_pct": f(d, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": f(d, "smsp__issue_active.avg.pct"),
        "achieved_occupancy_pct": f(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "note": "ncu replays the launch with cold caches and serialised kernels: use the counts, not the duration",
    }
    return s


def main():
    tag, units, reports = sys.argv[1], float(sys.argv[2]), sys.argv[3:]
    for rep in reports:
        s = summarise(rep, units)
        base = os.path.join(HERE, f"{tag}_{s['kernel']}")
        json.dump(s, open(base + ".json", "w"), indent=1)
        det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
        open(base + "_ncu_details.txt", "w").write(det)
        print(json.dumps(s))


if __name__ == "__main__":
    main()
