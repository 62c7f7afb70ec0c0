#!/bin/bash
set -x
mkdir -p gpurun_out
bash scratch/ab.sh scratch/libs/r2c_nopf.so scratch/libs/r2c_pf.so scratch/libs/r2c_r6.so scratch/libs/r2c_r5.so > gpurun_out/r2c_ab.log 2>&1; cat gpurun_out/r2c_ab.log
python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 1500 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
python bench.py --config cfg1 --steps 10 > gpurun_out/r2c_bench_cfg1.json 2> gpurun_out/r2c_bench_cfg1.err; tail -c 600 gpurun_out/r2c_bench_cfg1.json; tail -5 gpurun_out/r2c_bench_cfg1.err
