#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --config cfg3 --steps 5 --no-cpu-baseline --gather none > gpurun_out/r2j_bench_cfg3.json 2> gpurun_out/r2j_bench_cfg3.err
python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_cfg3.json')); k=d['roofline']['kernels']
print('cfg3', 'value %.3e ms %.2f' % (d['value'], d['ms_per_step']), {n:(round(v['kernel_ms'],2) if isinstance(v,dict) else round(v,2)) for n,v in k.items()})" || tail -5 gpurun_out/r2j_bench_cfg3.err
B="python bench.py --config cfg3 --vertices 200000 --steps 1 --warmup 1 --no-cpu-baseline --gather none --e2e-vertices 500"
for k in K_att_gl1 K_gl1_item K_gl1_fine K_att_expand; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 3 -c 1 -f -o gpurun_out/r2j_$k $B > gpurun_out/ncu_r2j_$k.log 2>&1
done
B="python bench.py --config cfg4mb1 --vertices 100000 --steps 1 --warmup 1 --no-cpu-baseline --gather none --e2e-vertices 500"
for k in K_roots_m K_classify_m "K_att<"; do
  n=$(echo $k | tr -d '<')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^(void )?$k" -s 3 -c 1 -f -o gpurun_out/r2j_$n $B > gpurun_out/ncu_r2j_$n.log 2>&1
done
python scratch/run_effects.py 4000 > gpurun_out/r2j_effects.json 2> gpurun_out/r2j_effects.err; cat gpurun_out/r2j_effects.json
for k in K_apply_effects K_focusing; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 1 -c 1 -f -o gpurun_out/r2j_$k python scratch/run_effects.py 4000 > gpurun_out/ncu_r2j_$k.log 2>&1
done
ls -la gpurun_out/r2j*.ncu-rep
