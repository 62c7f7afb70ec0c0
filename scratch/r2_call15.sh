#!/bin/bash
mkdir -p gpurun_out
bash scratch/ab.sh scratch/libs/v0.so scratch/libs/v_plan.so scratch/libs/v_store.so scratch/libs/v_ps.so scratch/libs/v_all.so scratch/libs/v_all72.so scratch/libs/v_all80.so scratch/libs/v0.so > gpurun_out/r2q_ab.log 2>&1
cat gpurun_out/r2q_ab.log
