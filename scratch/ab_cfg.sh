#!/bin/bash
# A/B runs of bench.py --config <cfg> against alternative builds: scratch/ab_cfg.sh cfg3 lib1.so lib2.so ...
cfg=$1; shift
for lib in "$@"; do
  NRMC_RT_LIB=$PWD/$lib python bench.py --config $cfg --no-cpu-baseline --gather none --e2e-vertices 500 --steps 5 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
k=d['roofline']['kernels']
print('$cfg $lib', 'pairs/s %.3e' % d['value'], 'ms/step %.2f' % d['ms_per_step'], {n:(round(v['kernel_ms'],2) if isinstance(v,dict) else round(v,2)) for n,v in k.items()})"
done
