#!/bin/bash
set -x
mkdir -p gpurun_out
bash scratch/ab_cfg.sh cfg3 scratch/libs/g4.so scratch/libs/g5.so scratch/libs/g6.so > gpurun_out/r2k_ab_cfg3.log 2>&1; cat gpurun_out/r2k_ab_cfg3.log
python -m pytest tests -m gpu -q -x > gpurun_out/r2k_tests.log 2>&1; tail -6 gpurun_out/r2k_tests.log
timeout 600 python scratch/stress_att.py 3000 > gpurun_out/r2k_stress_att.log 2>&1; grep "GL1" gpurun_out/r2k_stress_att.log
