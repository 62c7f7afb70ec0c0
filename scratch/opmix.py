"""SASS opcode mix of an ncu report weighted by executed warp instructions: python scratch/opmix.py report.ncu-rep [lo_line hi_line file]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
mix, smp = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr): continue
    sass = r[ix["Source"]].strip()
    parts = sass.split()
    if not parts: continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.split(".")[0]
    try:
        mix[op] += int(r[ix["Instructions Executed"]]); smp[op] += int(r[ix["# Samples"]])
    except ValueError:
        pass
T, S = sum(mix.values()), sum(smp.values())
fp64 = sum(v for k, v in mix.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"warp instructions {T}; FP64-pipe share {100 * fp64 / T:.1f}%")
for k, v in mix.most_common(28):
    print(f"{k:10s} {100 * v / T:5.1f}% inst  {100 * smp[k] / max(S, 1):5.1f}% samples")
