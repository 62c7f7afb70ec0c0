#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2e_tests.log 2>&1; tail -15 gpurun_out/r2e_tests.log
python scratch/scalar_latency.py > gpurun_out/r2e_scalar.json 2> gpurun_out/r2e_scalar.err; cat gpurun_out/r2e_scalar.json; tail -3 gpurun_out/r2e_scalar.err
python bench.py --config cfg1 --steps 10 --no-cpu-baseline > gpurun_out/r2e_bench_cfg1.json 2> gpurun_out/r2e_bench_cfg1.err; python -c "
import json; d=json.load(open('gpurun_out/r2e_bench_cfg1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; tail -3 gpurun_out/r2e_bench_cfg1.err
python -c "import __graft_entry__ as g; g.smoke()"
