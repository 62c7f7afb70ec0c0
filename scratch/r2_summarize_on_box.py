"""Run on the GPU box after the ncu captures of scratch/r2_final_profile.sh: summarise every gpurun_out/<tag>_K_*.ncu-rep with
profiles/summarize_ncu.py (units per launch from the bench line in the capture's log), copy the summaries to gpurun_out/ and drop
the reports that are not listed in KEEP (gpurun brings back at most 64 MiB)."""
import glob, json, os, re, shutil, subprocess, sys
tag = sys.argv[1]
KEEP = set(sys.argv[2:])
PAIR_KERNELS = {"K_classify", "K_hump", "K_classify_m", "K_hump_m"}
for rep in sorted(glob.glob(f"gpurun_out/{tag}_K_*.ncu-rep")):
    k = re.sub(rf"^gpurun_out/{tag}_", "", rep)[:-len(".ncu-rep")]
    line = None
    try:
        for l in open(f"gpurun_out/ncu_{tag}_{k}.log"):
            if l.startswith("{"):
                line = json.loads(l)
    except OSError:
        pass
    if line is None:
        print("no bench line for", k); continue
    pairs = line["config"]["pairs_per_step"]
    units = pairs if k in PAIR_KERNELS else pairs * line["solutions_per_pair"]
    subprocess.run([sys.executable, "profiles/summarize_ncu.py", tag, str(units), rep], stdout=subprocess.DEVNULL)
    for f in glob.glob(f"profiles/{tag}_{k}*"):
        shutil.copy(f, "gpurun_out/")
    if k not in KEEP:
        os.remove(rep)
