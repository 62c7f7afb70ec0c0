#!/bin/bash
mkdir -p gpurun_out
bash scratch/ab.sh scratch/libs/f_base.so scratch/libs/f_rb6.so scratch/libs/f_rb5.so scratch/libs/f_rb8.so scratch/libs/f_tol6.so scratch/libs/f_tol6_rb6.so scratch/libs/f_base.so > gpurun_out/r2w_ab.log 2>&1
cat gpurun_out/r2w_ab.log
