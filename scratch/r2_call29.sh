#!/bin/bash
mkdir -p gpurun_out
( bash scratch/ab_cfg.sh cfg4 scratch/libs/pm.so scratch/libs/mm.so scratch/libs/pm.so scratch/libs/mm.so
bash scratch/ab_cfg.sh cfg4mb1 scratch/libs/pm.so scratch/libs/mm.so ) > gpurun_out/r2y7_ab.log 2>&1
cat gpurun_out/r2y7_ab.log
NRMC_RT_LIB=$PWD/scratch/libs/mm.so python -m pytest tests -m gpu -q -x 2>&1 | tail -3
