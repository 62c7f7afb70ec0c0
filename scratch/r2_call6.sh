#!/bin/bash
set -x
mkdir -p gpurun_out
for c in cfg2 cfg3 cfg4 cfg4mb1; do
  timeout 600 python bench.py --config $c --steps 5 --no-cpu-baseline --gather none > gpurun_out/r2h_bench_$c.json 2> gpurun_out/r2h_bench_$c.err
  python -c "
import json; d=json.load(open('gpurun_out/r2h_bench_$c.json')); k=d['roofline']['kernels']
print('$c', 'value %.3e ms %.2f e2e %.3e sol/pair %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['solutions_per_pair']), {n:(round(v['kernel_ms'],2) if isinstance(v,dict) else round(v,2)) for n,v in k.items()})" || tail -5 gpurun_out/r2h_bench_$c.err
done
B="python bench.py --config cfg3 --vertices 100000 --steps 1 --warmup 1 --no-cpu-baseline --gather none --e2e-vertices 500"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^K_att_gl1" -s 3 -c 1 -f -o gpurun_out/r2h_K_att_gl1 $B > gpurun_out/ncu_r2h_gl1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^K_att<" -s 3 -c 1 -f -o gpurun_out/r2h_K_att_fallback $B > gpurun_out/ncu_r2h_fb.log 2>&1
ls -la gpurun_out/r2h*.ncu-rep
