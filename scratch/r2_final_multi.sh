#!/bin/bash
# multi-GPU artefacts of round 2 (gpurun --gpus 8): bench at N = 8, 4, 2 (value, value_gathered with every gather variant, e2e) and the
# NCCL tests on 2 GPUs
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2z_topo8.txt 2>&1
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n bench.py --gpus $n --steps 10 > gpurun_out/r2z_bench_n$n.json 2> gpurun_out/r2z_bench_n$n.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2z_bench_n$n.json') if l.startswith('{')][-1])
    print('N=$n value %.3e ms %.2f e2e %.3e d2h %.1f probe %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['d2h_gbs'], d['e2e']['d2h_probe_gbs']))
    g=d.get('gathered',{})
    for k,v in g.items():
        if isinstance(v,dict): print('  ', k, '%.3e' % v.get('value',0), v.get('ms_per_step'), v.get('nvlink_gbs_into_rank0', v.get('nccl_gbs_into_rank0')), v.get('checksum_ok'), v.get('error'))
    print('  headline', g.get('headline'), d.get('value_gathered'))
except Exception as e:
    print('N=$n failed', e)
PY
done
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_multi_gpu_nccl.py -m gpu -q > gpurun_out/r2z_tests_nccl.log 2>&1; tail -3 gpurun_out/r2z_tests_nccl.log
