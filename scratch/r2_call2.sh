#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2b_tests.log 2>&1; tail -12 gpurun_out/r2b_tests.log
bash scratch/ab.sh scratch/libs/r2a.so nuradiomc_b200/libnrmc_rt.so scratch/libs/kt10.so > gpurun_out/r2b_ab.log 2>&1; cat gpurun_out/r2b_ab.log
timeout 600 python scratch/stress_att.py 3000 > gpurun_out/r2b_stress_att.log 2>&1; grep -v "^$" gpurun_out/r2b_stress_att.log | tail -14
B="python bench.py --vertices 100000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-vertices 2000"
for k in K_roots K_classify K_att_sp1; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 3 -c 1 -f -o gpurun_out/r2b_$k $B > gpurun_out/ncu_r2b_$k.log 2>&1
done
ls -la gpurun_out/r2b*.ncu-rep | tail -4
