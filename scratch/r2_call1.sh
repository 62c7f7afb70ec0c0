#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; tail -5 gpurun_out/r2a_tests.log
bash scratch/ab.sh nuradiomc_b200/libnrmc_rt.so scratch/libs/r6.so scratch/libs/k88.so > gpurun_out/r2a_ab.log 2>&1; cat gpurun_out/r2a_ab.log
python scratch/scalar_latency.py > gpurun_out/r2a_scalar.json 2> gpurun_out/r2a_scalar.err; cat gpurun_out/r2a_scalar.json
B="python bench.py --vertices 100000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-vertices 2000"
for k in K_roots K_classify K_att_sp1; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 3 -c 1 -f -o gpurun_out/r2a_$k $B > gpurun_out/ncu_r2a_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -4
