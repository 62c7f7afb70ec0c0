#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2g_bench.json') if l.startswith('{')][-1])
print('%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], d['roofline']['frac'], {k:round(v['kernel_ms'],2) for k,v in d['roofline']['kernels'].items() if isinstance(v,dict)})
PY
