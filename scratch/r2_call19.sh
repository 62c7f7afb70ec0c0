#!/bin/bash
mkdir -p gpurun_out
bash scratch/ab.sh scratch/libs/base_e037.so scratch/libs/e_tab.so scratch/libs/e_tab7.so scratch/libs/e_tab_ew2.so scratch/libs/e_nodetab.so scratch/libs/e_tab2.so scratch/libs/e_tab2_7.so scratch/libs/e_tab2_ew2.so scratch/libs/base_e037.so > gpurun_out/r2u_ab.log 2>&1
cat gpurun_out/r2u_ab.log
NRMC_RT_LIB=$PWD/scratch/libs/e_tab2.so python -m pytest tests -m gpu -q -x -k "cfg5 or sp1 or SP1 or oracle or golden or fixture" 2>&1 | tail -3
NRMC_RT_LIB=$PWD/scratch/libs/e_tab2.so python scratch/stress_att.py > gpurun_out/r2u_stress_att.log 2>&1; tail -12 gpurun_out/r2u_stress_att.log
