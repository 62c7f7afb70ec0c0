#!/bin/bash
mkdir -p gpurun_out
( bash scratch/ab_cfg.sh cfg4mb1 scratch/libs/sp8stock.so scratch/libs/sp8.so scratch/libs/sp8stock.so scratch/libs/sp8.so ) > gpurun_out/r2y6_ab.log 2>&1
cat gpurun_out/r2y6_ab.log
NRMC_RT_LIB=$PWD/scratch/libs/sp8.so python -m pytest tests -m gpu -q -x -k "separable or attenuation or compact or effects or small_batch" 2>&1 | tail -3
NRMC_RT_LIB=$PWD/scratch/libs/sp8.so timeout 600 python scratch/stress_att.py 3000 2>&1 | grep "MB1\|GL2"
