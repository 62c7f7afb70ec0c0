#!/bin/bash
mkdir -p gpurun_out
( NRMC_SEP_GENERIC=1 bash scratch/ab_cfg.sh cfg4mb1 scratch/libs/sep5.so
bash scratch/ab_cfg.sh cfg4mb1 scratch/libs/sep4.so scratch/libs/sep5.so scratch/libs/sep6.so scratch/libs/sep8.so ) > gpurun_out/r2y5_ab.log 2>&1
cat gpurun_out/r2y5_ab.log
NRMC_RT_LIB=$PWD/scratch/libs/sep5.so python -m pytest tests -m gpu -q -x 2>&1 | tail -5
NRMC_RT_LIB=$PWD/scratch/libs/sep5.so timeout 600 python scratch/stress_att.py 3000 2>&1 | grep "MB1\|GL2"
