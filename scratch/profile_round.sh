#!/bin/bash
# One GPU call that refreshes the judged artefacts: ncu captures of the bench kernels (1e7-pair launches), the launch list of the
# bench command, the bench line itself (never under ncu) and the reference arm.  Outputs under gpurun_out/.
# usage: bash scratch/profile_round.sh [kernels...]   (default: all four)
set -x
KERNELS=${@:-K_att_sp1 K_roots K_classify K_hump}
B="python bench.py --vertices 100000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-vertices 2000"
for k in $KERNELS; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 3 -c 1 -f -o gpurun_out/r1_$k $B > gpurun_out/ncu_$k.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-vertices 2000 > /dev/null 2>&1
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
python bench.py --impl reference > gpurun_out/r1_bench_reference_arm.json 2> gpurun_out/r1_bench_reference_arm.err
tail -c 300 gpurun_out/r1_bench.json; tail -c 400 gpurun_out/r1_bench_reference_arm.json
