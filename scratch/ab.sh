#!/bin/bash
# A/B runs of bench.py against alternative builds of the library: scratch/ab.sh lib1.so lib2.so ...
for lib in "$@"; do
  NRMC_RT_LIB=$PWD/$lib python bench.py --no-cpu-baseline --gather none --e2e-vertices 2000 --steps 5 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
k=d['roofline']['kernels']
print('$lib', 'pairs/s %.3e' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'att %.2f' % k['attenuation_main']['kernel_ms'], 'classify %.2f hump %.2f roots %.2f other %.2f' % (k['classify']['kernel_ms'], k['hump']['kernel_ms'], k['roots']['kernel_ms'], k['other_ms']))"
done
