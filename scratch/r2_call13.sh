#!/bin/bash
set -x
mkdir -p gpurun_out
python profiles/parity_report.py > gpurun_out/r2o_parity.log 2>&1; tail -4 gpurun_out/r2o_parity.log
python scratch/sanitize_run.py > gpurun_out/r2o_sanitize_plain.log 2>&1; tail -3 gpurun_out/r2o_sanitize_plain.log
SAN_N=150 timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python scratch/sanitize_run.py > gpurun_out/r2o_memcheck.log 2>&1; echo memcheck rc=$?; grep -E "ERROR SUMMARY|done" gpurun_out/r2o_memcheck.log | tail -3
SAN_N=60 timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python scratch/sanitize_run.py > gpurun_out/r2o_racecheck.log 2>&1; echo racecheck rc=$?; grep -E "RACECHECK SUMMARY|done" gpurun_out/r2o_racecheck.log | tail -3
SAN_N=60 timeout 900 compute-sanitizer --tool synccheck --error-exitcode 3 python scratch/sanitize_run.py > gpurun_out/r2o_synccheck.log 2>&1; echo synccheck rc=$?; grep -E "ERROR SUMMARY|done" gpurun_out/r2o_synccheck.log | tail -3
