#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2d_topo.txt 2>&1; head -12 gpurun_out/r2d_topo.txt
timeout 600 python -m pytest tests/test_multi_gpu_nccl.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2d_tests.log 2>&1; tail -15 gpurun_out/r2d_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; tail -c 3000 gpurun_out/r2d_bench_n2.json; tail -8 gpurun_out/r2d_bench_n2.err
