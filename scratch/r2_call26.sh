#!/bin/bash
mkdir -p gpurun_out
( bash scratch/ab.sh scratch/libs/base_prev.so scratch/libs/pn.so scratch/libs/base_prev.so scratch/libs/pn.so
for c in cfg2 cfg4; do bash scratch/ab_cfg.sh $c scratch/libs/base_prev.so scratch/libs/pn.so; done ) > gpurun_out/r2y4_ab.log 2>&1
cat gpurun_out/r2y4_ab.log
NRMC_RT_LIB=$PWD/scratch/libs/pn.so python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
B="python bench.py --config cfg4mb1 --vertices 50000 --steps 1 --warmup 1 --no-cpu-baseline --gather none --e2e-vertices 500"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^K_att<" -s 3 -c 1 -f -o gpurun_out/r2z_K_att_generic $B > gpurun_out/ncu_r2z_K_att_generic.log 2>&1
ls -la gpurun_out/r2z_K_att_generic.ncu-rep
