#!/bin/bash
# Round-2 judged artefacts in one GPU call (1 GPU): GPU test suite, ncu captures of the hot kernels of cfg5 / cfg3 / cfg4, the launch list of
# the bench command, bench lines of every configuration (never under ncu), the reference arm, parity report, compute-sanitizer.
# Outputs under gpurun_out/ with the tag $T; profiles/ is filled from them by scratch/r2_collect.sh on the build machine.
T=${1:-r2z}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
B5="python bench.py --vertices 100000 --steps 1 --warmup 1 --no-cpu-baseline --gather none --e2e-vertices 2000"
for k in K_att_sp1 K_roots K_classify K_hump; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 3 -c 1 -f -o gpurun_out/${T}_$k $B5 > gpurun_out/ncu_${T}_$k.log 2>&1
done
B3="python bench.py --config cfg3 --vertices 200000 --steps 1 --warmup 1 --no-cpu-baseline --gather none --e2e-vertices 500"
for k in K_att_gl1 K_gl1_item K_gl1_fine K_att_expand; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 3 -c 1 -f -o gpurun_out/${T}_$k $B3 > gpurun_out/ncu_${T}_$k.log 2>&1
done
B4="python bench.py --config cfg4 --vertices 100000 --steps 1 --warmup 1 --no-cpu-baseline --gather none --e2e-vertices 500"
for k in K_roots_m K_classify_m; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 3 -c 1 -f -o gpurun_out/${T}_$k $B4 > gpurun_out/ncu_${T}_$k.log 2>&1
done
python scratch/r2_summarize_on_box.py $T K_att_sp1 K_roots
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --gather none --e2e-vertices 2000 > /dev/null 2>&1
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 3 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err
for c in cfg1 cfg2 cfg3 cfg4 cfg4mb1; do
  python bench.py --config $c --steps 10 > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err
done
python scratch/scalar_latency.py > gpurun_out/${T}_scalar.json 2> gpurun_out/${T}_scalar.err
python profiles/parity_report.py > gpurun_out/${T}_parity.log 2>&1; tail -2 gpurun_out/${T}_parity.log
timeout 600 python scratch/stress_att.py 3000 > gpurun_out/${T}_stress_att.log 2>&1
SAN_N=150 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python scratch/sanitize_run.py > gpurun_out/${T}_memcheck.log 2>&1; echo memcheck rc=$?
SAN_N=60 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python scratch/sanitize_run.py > gpurun_out/${T}_racecheck.log 2>&1; echo racecheck rc=$?
SAN_N=60 timeout 600 compute-sanitizer --tool synccheck --error-exitcode 3 python scratch/sanitize_run.py > gpurun_out/${T}_synccheck.log 2>&1; echo synccheck rc=$?
for f in gpurun_out/${T}_bench*.json; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    print(f, '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], d.get('roofline',{}).get('frac'))
except Exception as e: print(f,'ERR',e)
PY
done
