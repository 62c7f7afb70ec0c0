#!/bin/bash
mkdir -p gpurun_out
T=r2h
python scratch/sanitize_run.py > gpurun_out/${T}_sanitize_plain.log 2>&1; tail -3 gpurun_out/${T}_sanitize_plain.log
SAN_N=150 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python scratch/sanitize_run.py > gpurun_out/${T}_memcheck.log 2>&1; echo memcheck rc=$?; grep -E "ERROR SUMMARY" gpurun_out/${T}_memcheck.log | tail -1
SAN_N=60 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python scratch/sanitize_run.py > gpurun_out/${T}_racecheck.log 2>&1; echo racecheck rc=$?; grep -E "RACECHECK SUMMARY" gpurun_out/${T}_racecheck.log | tail -1
SAN_N=60 timeout 600 compute-sanitizer --tool synccheck --error-exitcode 3 python scratch/sanitize_run.py > gpurun_out/${T}_synccheck.log 2>&1; echo synccheck rc=$?; grep -E "ERROR SUMMARY" gpurun_out/${T}_synccheck.log | tail -1
