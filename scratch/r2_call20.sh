#!/bin/bash
mkdir -p gpurun_out
bash scratch/ab.sh scratch/libs/base_e037.so scratch/libs/e_tab2_ew2.so scratch/libs/e_tab2_ew3.so scratch/libs/e8_14_2.so scratch/libs/e8_14_1.so scratch/libs/e7_14_2.so scratch/libs/r_pk.so scratch/libs/base_e037.so > gpurun_out/r2v_ab.log 2>&1
cat gpurun_out/r2v_ab.log
NRMC_RT_LIB=$PWD/scratch/libs/r_pk.so python -m pytest tests -m gpu -q -x 2>&1 | tail -3
