#!/bin/bash
# refresh of the judged artefacts on the FINAL build (1 GPU): tests, smoke, bench lines (cfg5 default, reference arm, cfg4, cfg4mb1, cfg2), ncu captures of the
# kernels that changed after r2z (K_att_sep new, K_roots_m / K_classify_m, K_roots, K_classify; generic K_att), solver stress run
T=r2f
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
B5="python bench.py --vertices 100000 --steps 1 --warmup 1 --no-cpu-baseline --gather none --e2e-vertices 2000"
for k in K_roots K_classify; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k\$" -s 3 -c 1 -f -o gpurun_out/${T}_$k $B5 > gpurun_out/ncu_${T}_$k.log 2>&1
done
B4="python bench.py --config cfg4mb1 --vertices 100000 --steps 1 --warmup 1 --no-cpu-baseline --gather none --e2e-vertices 500"
for k in K_roots_m K_classify_m K_att_sep; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k\$" -s 3 -c 1 -f -o gpurun_out/${T}_$k $B4 > gpurun_out/ncu_${T}_$k.log 2>&1
done
NRMC_SEP_GENERIC=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^K_att\$" -s 3 -c 1 -f -o gpurun_out/${T}_K_att $B4 > gpurun_out/ncu_${T}_K_att.log 2>&1
python scratch/r2_summarize_on_box.py $T
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 3 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err
for c in cfg2 cfg4 cfg4mb1; do
  python bench.py --config $c --steps 10 > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err
done
timeout 900 python scratch/stress_parity.py 200000 > gpurun_out/${T}_stress_parity.log 2>&1; grep -v "^   parity" gpurun_out/${T}_stress_parity.log | tail -8
for f in gpurun_out/${T}_bench*.json; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    print(f, '%.3e'%d['value'], '%.3f ms'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], d.get('roofline',{}).get('frac'))
except Exception as e: print(f,'ERR',e)
PY
done
