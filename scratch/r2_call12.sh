#!/bin/bash
set -x
mkdir -p gpurun_out
bash scratch/ab_cfg.sh cfg3 nuradiomc_b200/libnrmc_rt.so > gpurun_out/r2n_ab_cfg3.log 2>&1; cat gpurun_out/r2n_ab_cfg3.log
python -m pytest tests -m gpu -q -x > gpurun_out/r2n_tests.log 2>&1; tail -4 gpurun_out/r2n_tests.log
