#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 240 python -m pytest tests -m gpu -q -k "stager or side_branches or pairwise_helpers" > gpurun_out/r2_newtests.log 2>&1
echo "newtests rc=$?" > gpurun_out/r2_status.txt
CAREL_SIDE_BRANCHES=3 timeout 420 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2_suite_sb3.log 2>&1
echo "suite sb3 rc=$?" >> gpurun_out/r2_status.txt
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/r2_suite_sb0.log 2>&1
echo "suite sb0 rc=$?" >> gpurun_out/r2_status.txt
cat gpurun_out/r2_status.txt; tail -4 gpurun_out/r2_newtests.log; tail -12 gpurun_out/r2_suite_sb3.log; tail -4 gpurun_out/r2_suite_sb0.log
