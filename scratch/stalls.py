"""Aggregate the per-instruction warp-stall samples of an ncu report's source page: python scratch/stalls.py report.ncu-rep"""
import csv, io, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in rows[2:]:
    if len(r) <= max(cols):
        continue
    for i in cols:
        try:
            tot[hdr[i]] += int(r[i])
        except ValueError:
            pass
s = sum(tot.values())
for k, v in tot.most_common():
    if v:
        print(f"{k:28s} {100 * v / s:5.1f} %")
