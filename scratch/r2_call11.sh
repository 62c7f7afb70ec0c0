#!/bin/bash
set -x
mkdir -p gpurun_out
bash scratch/ab_cfg.sh cfg3 scratch/libs/e8.so scratch/libs/e16.so scratch/libs/e4.so scratch/libs/e8wt.so scratch/libs/e8cg.so > gpurun_out/r2m_ab_cfg3.log 2>&1; cat gpurun_out/r2m_ab_cfg3.log
python -m pytest tests -m gpu -q -x -k "gl1 or GL1 or greenland or cfg3" > gpurun_out/r2m_tests.log 2>&1; tail -4 gpurun_out/r2m_tests.log
