#!/bin/bash
mkdir -p gpurun_out
bash scratch/ab.sh scratch/libs/v0.so scratch/libs/d1.so scratch/libs/d2.so scratch/libs/d4.so scratch/libs/d2b7.so scratch/libs/d4b7.so scratch/libs/d1rs.so scratch/libs/v0.so > gpurun_out/r2r_ab.log 2>&1
cat gpurun_out/r2r_ab.log
NRMC_RT_LIB=$PWD/scratch/libs/d1.so python -m pytest tests -m gpu -q -x -k "cfg5 or sp1 or SP1 or oracle or golden or fixture" 2>&1 | tail -3
