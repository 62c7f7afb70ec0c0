#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --vertices 100000 --steps 1 --warmup 1 --no-cpu-baseline --gather none --e2e-vertices 2000"
for k in K_att_sp1 K_roots K_classify; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 3 -c 1 -f -o gpurun_out/r2t_$k $B > gpurun_out/ncu_r2t_$k.log 2>&1
done
ls -la gpurun_out/r2t_*
