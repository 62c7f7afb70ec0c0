#!/bin/bash
# A/B runs of bench.py against alternative builds of the library: scratch/ab.sh lib1.so lib2.so ...
for lib in "$@"; do
  NRMC_RT_LIB=$PWD/$lib python bench.py --no-cpu-baseline --e2e-vertices 2000 --steps 3 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
print('$lib', 'pairs/s %.3e' % d['value'], 'ms/step %.1f' % d['ms_per_step'], 'att_ms %.1f' % d['roofline']['kernel_ms'], 'solve_ms %.1f' % d['roofline']['K_solve']['kernel_ms'])"
done
