#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu_nccl.py -m gpu -q -x > gpurun_out/r2g_tests.log 2>&1; tail -5 gpurun_out/r2g_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; tail -c 3500 gpurun_out/r2g_bench_n2.json; grep -v "^\[W\|^$\|^\*\*\|OMP_NUM" gpurun_out/r2g_bench_n2.err | tail -12
