#!/bin/bash
mkdir -p gpurun_out
( NRMC_EXPAND_PLAIN=1 bash scratch/ab_cfg.sh cfg3 scratch/libs/x2.so
bash scratch/ab_cfg.sh cfg3 scratch/libs/x2.so scratch/libs/x3.so scratch/libs/x2w8.so scratch/libs/x4w2.so ) > gpurun_out/r2y10_ab.log 2>&1
cat gpurun_out/r2y10_ab.log
NRMC_RT_LIB=$PWD/scratch/libs/x2.so python -m pytest tests -m gpu -q -x 2>&1 | tail -3
