#!/bin/bash
mkdir -p gpurun_out
bash scratch/ab.sh scratch/libs/h_rb8_nov.so scratch/libs/h_rb7.so scratch/libs/h_rb8.so scratch/libs/h_rb9.so scratch/libs/h_rb10.so scratch/libs/h_rb8_nov.so > gpurun_out/r2x_ab.log 2>&1
cat gpurun_out/r2x_ab.log
for c in cfg2 cfg3 cfg4; do bash scratch/ab_cfg.sh $c scratch/libs/h_rb8_nov.so scratch/libs/h_rb8.so; done 2>&1 | tee gpurun_out/r2x_ab_cfgs.log
NRMC_RT_LIB=$PWD/scratch/libs/h_rb8.so python -m pytest tests -m gpu -q -x 2>&1 | tail -3
