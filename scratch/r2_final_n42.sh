#!/bin/bash
mkdir -p gpurun_out
for n in 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2958$n bench.py --gpus $n --steps 10 > gpurun_out/r2g_bench_n$n.json 2> gpurun_out/r2g_bench_n$n.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2g_bench_n$n.json') if l.startswith('{')][-1])
print('N=$n value %.3e ms %.2f e2e %.3e d2h %.1f probe %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['d2h_gbs'], d['e2e']['d2h_probe_gbs']))
g=d.get('gathered',{})
for k,v in g.items():
    if isinstance(v,dict): print('  ', k, '%.3e' % v.get('value',0), v.get('ms_per_step'), v.get('nvlink_gbs_into_rank0', v.get('nccl_gbs_into_rank0')), v.get('checksum_ok'), v.get('error'))
PY
done
