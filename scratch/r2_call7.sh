#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "gl1 or GL1 or greenland or attenuation or small_batch or fixtures" > gpurun_out/r2i_tests.log 2>&1; tail -15 gpurun_out/r2i_tests.log
timeout 600 python scratch/stress_att.py 3000 > gpurun_out/r2i_stress_att.log 2>&1; grep -v "^$" gpurun_out/r2i_stress_att.log | tail -14
for c in cfg3; do
  timeout 600 python bench.py --config $c --steps 5 --no-cpu-baseline --gather none > gpurun_out/r2i_bench_$c.json 2> gpurun_out/r2i_bench_$c.err
  python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_$c.json')); k=d['roofline']['kernels']
print('$c', 'value %.3e ms %.2f e2e %.3e sol/pair %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['solutions_per_pair']), {n:(round(v['kernel_ms'],2) if isinstance(v,dict) else round(v,2)) for n,v in k.items()})" || tail -5 gpurun_out/r2i_bench_$c.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2i_launches_cfg3.csv python bench.py --config cfg3 --steps 1 --warmup 1 --no-cpu-baseline --gather none --e2e-vertices 500 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2i_launches_cfg3.csv')) if len(r)>5 and r[0].isdigit()]
# last pass: print kernel name + duration for the final 16 launches
for r in rows[-16:]:
    print(r[4][:40], r[-1], r[-2])
PY
