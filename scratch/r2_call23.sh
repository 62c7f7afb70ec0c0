#!/bin/bash
mkdir -p gpurun_out
( bash scratch/ab.sh scratch/libs/m3.so
bash scratch/ab_cfg.sh cfg4 scratch/libs/m3.so scratch/libs/m4.so scratch/libs/m5.so scratch/libs/m6.so scratch/libs/m8.so
bash scratch/ab_cfg.sh cfg4mb1 scratch/libs/a4.so scratch/libs/a5.so scratch/libs/a6.so scratch/libs/a8.so
bash scratch/ab_cfg.sh cfg3 scratch/libs/a4.so scratch/libs/g5.so scratch/libs/g6.so scratch/libs/g7.so ) > gpurun_out/r2y_ab.log 2>&1
cat gpurun_out/r2y_ab.log
