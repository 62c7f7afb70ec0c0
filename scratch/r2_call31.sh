#!/bin/bash
mkdir -p gpurun_out
( bash scratch/ab.sh scratch/libs/c4.so scratch/libs/c5.so scratch/libs/c6.so scratch/libs/c8.so
bash scratch/ab_cfg.sh cfg4 scratch/libs/c4.so scratch/libs/c5.so scratch/libs/c6.so scratch/libs/c8.so ) > gpurun_out/r2y9_ab.log 2>&1
cat gpurun_out/r2y9_ab.log
