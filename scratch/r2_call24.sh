#!/bin/bash
mkdir -p gpurun_out
( bash scratch/ab.sh scratch/libs/k4.so scratch/libs/k6.so scratch/libs/k8.so scratch/libs/k10.so
bash scratch/ab_cfg.sh cfg4 scratch/libs/k4.so scratch/libs/k6.so scratch/libs/k8.so scratch/libs/k10.so ) > gpurun_out/r2y2_ab.log 2>&1
cat gpurun_out/r2y2_ab.log
