#!/bin/bash
mkdir -p gpurun_out
( bash scratch/ab.sh scratch/libs/c_late.so scratch/libs/c_first.so scratch/libs/c_late.so scratch/libs/c_first.so
for c in cfg2 cfg3; do bash scratch/ab_cfg.sh $c scratch/libs/c_late.so scratch/libs/c_first.so; done ) > gpurun_out/r2y3_ab.log 2>&1
cat gpurun_out/r2y3_ab.log
NRMC_RT_LIB=$PWD/scratch/libs/c_first.so python -m pytest tests -m gpu -q -x 2>&1 | tail -3
