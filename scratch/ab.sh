#!/bin/bash
# A/B runs of bench.py against alternative builds of the library: scratch/ab.sh lib1.so lib2.so ...
for lib in "$@"; do
  NRMC_RT_LIB=$PWD/$lib python bench.py --no-cpu-baseline --e2e-vertices 2000 --steps 5 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
r=d['roofline']; s=r['solver']
print('$lib', 'pairs/s %.3e' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'att %.2f' % r['kernel_ms'], 'classify %.2f hump %.2f roots %.2f' % (s['K_classify']['kernel_ms'], s['K_hump']['kernel_ms'], s['K_roots']['kernel_ms']))"
done
