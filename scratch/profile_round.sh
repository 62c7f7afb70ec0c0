#!/bin/bash
# One GPU call that refreshes the judged artefacts: GPU tests, ncu captures of the four bench kernels (1e7-pair launches), the
# launch list of the bench command, and the bench line itself (never under ncu).  Outputs under gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="python bench.py --vertices 100000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-vertices 2000"
for k in K_att_sp1 K_roots K_classify K_hump; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 3 -c 1 -f -o gpurun_out/r1_$k $B > gpurun_out/ncu_$k.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-vertices 2000 > /dev/null 2>&1
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
tail -c 600 gpurun_out/r1_bench.json
