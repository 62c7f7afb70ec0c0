"""per-source-line profile of an ncu report: python scratch/srcprof.py report.ncu-rep [top]  (samples, warp instructions, avg threads,
excessive / local L2 sectors per CUDA line; stall mix of the whole kernel)"""
import csv, io, subprocess, sys, collections
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr = "", None
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0, ""])   # samples, inst, thread inst, excessive, global sectors, local sectors, text
stalls = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; ix = {h: i for i, h in enumerate(hdr)}
        # two columns are called "Source": first = CUDA text, second = SASS
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    def g(name):
        try: return int(r[ix[name]])
        except (ValueError, KeyError): return 0
    key = (fname, int(r[0]))
    a = agg[key]
    a[0] += g("# Samples"); a[1] += g("Instructions Executed"); a[2] += g("Thread Instructions Executed")
    a[3] += g("L2 Theoretical Sectors Global Excessive"); a[4] += g("L2 Theoretical Sectors Global"); a[5] += g("L2 Theoretical Sectors Local")
    a[6] = r[1].strip()[:90]
    for h, i in ix.items():
        if h.startswith("stall_") and "Not Issued" not in h:
            try: stalls[h] += int(r[i])
            except ValueError: pass
S = sum(a[0] for a in agg.values()); I = sum(a[1] for a in agg.values()); TI = sum(a[2] for a in agg.values())
E = sum(a[3] for a in agg.values()); G = sum(a[4] for a in agg.values()); L = sum(a[5] for a in agg.values())
print(f"samples {S}  warp-inst {I}  avg threads {TI / max(I, 1):.1f}  global sectors {G} (excessive {E})  local sectors {L}")
print("stalls: " + ", ".join(f"{k[6:]} {100 * v / max(sum(stalls.values()), 1):.0f}%" for k, v in stalls.most_common(7)))
print("--- top lines by samples")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * a[0] / max(S, 1):5.1f}% smp {100 * a[1] / max(I, 1):5.1f}% inst thr {a[2] / max(a[1], 1):4.1f} exc {a[3]:9d} loc {a[5]:8d}  {f}:{ln}  {a[6]}")
print("--- top lines by excessive + local sectors")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -(kv[1][3] + kv[1][5]))[:10]:
    if a[3] + a[5]:
        print(f"exc {a[3]:9d} of {a[4]:9d} loc {a[5]:8d}  {f}:{ln}  {a[6]}")
