#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2s_tests.log 2>&1; tail -5 gpurun_out/r2s_tests.log
for c in cfg5 cfg3 cfg4mb1 cfg4 cfg2 cfg1; do
python bench.py --config $c --no-cpu-baseline --gather none --e2e-vertices 2000 --steps 5 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
k=d['roofline']['kernels']
print('$c', 'pairs/s %.3e' % d['value'], 'ms/step %.3f' % d['ms_per_step'], {n:(round(v['kernel_ms'],2) if isinstance(v,dict) else round(v,2)) for n,v in k.items()})"
done > gpurun_out/r2s_bench.log 2>&1
cat gpurun_out/r2s_bench.log
