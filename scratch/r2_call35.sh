#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2i_tests.log 2>&1; tail -3 gpurun_out/r2i_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; tail -2 gpurun_out/r2i_smoke.log
python profiles/parity_report.py > gpurun_out/r2i_parity.log 2>&1; tail -2 gpurun_out/r2i_parity.log
cp profiles/r2_parity_report.json gpurun_out/r2i_parity_report.json
