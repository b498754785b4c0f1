#!/bin/bash
# usage (on the GPU box): bash tools/step_profile.sh <tag>   -> gpurun_out/<tag>_{pytest,bench,launches}*
tag=$1
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${tag}_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench_err.log
tail -c 300 gpurun_out/${tag}_pytest.log
tail -c 300 gpurun_out/${tag}_bench_err.log
