#!/bin/bash
mkdir -p gpurun_out
( bash scratch/ab_cfg.sh cfg4 scratch/libs/mm.so scratch/libs/rc.so scratch/libs/mm.so scratch/libs/rc.so ) > gpurun_out/r2y8_ab.log 2>&1
cat gpurun_out/r2y8_ab.log
NRMC_RT_LIB=$PWD/scratch/libs/rc.so python -m pytest tests -m gpu -q -x 2>&1 | tail -3
