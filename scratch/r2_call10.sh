#!/bin/bash
set -x
mkdir -p gpurun_out
bash scratch/ab_cfg.sh cfg3 nuradiomc_b200/libnrmc_rt.so > gpurun_out/r2l_ab_cfg3.log 2>&1; cat gpurun_out/r2l_ab_cfg3.log
python -m pytest tests -m gpu -q -x -k "gl1 or GL1 or greenland or attenuation or fixtures or cfg3" > gpurun_out/r2l_tests.log 2>&1; tail -6 gpurun_out/r2l_tests.log
timeout 600 python scratch/stress_att.py 6000 > gpurun_out/r2l_stress_att.log 2>&1; grep "GL1" gpurun_out/r2l_stress_att.log
