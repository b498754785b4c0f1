#!/bin/bash
# round-2 experiment run: new tests, bench with side branches off / fwd / fwd+bwd, then the whole GPU suite with them on
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 240 python -m pytest tests -m gpu -q -k "stager or side_branches or pairwise_helpers" > gpurun_out/r1_newtests.log 2>&1
echo "newtests rc=$?" > gpurun_out/r1_status.txt
for m in 0 1 3; do
  CAREL_SIDE_BRANCHES=$m timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r1_bench_sb$m.json 2> gpurun_out/r1_bench_sb$m.err
  echo "bench sb$m rc=$?" >> gpurun_out/r1_status.txt
done
CAREL_SIDE_BRANCHES=3 timeout 420 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r1_suite_sb3.log 2>&1
echo "suite sb3 rc=$?" >> gpurun_out/r1_status.txt
tail -5 gpurun_out/r1_newtests.log; cat gpurun_out/r1_status.txt; tail -15 gpurun_out/r1_suite_sb3.log
for m in 0 1 3; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r1_bench_sb$m.json").read().strip().splitlines()[-1])
    print("sb$m", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("split"), d["e2e"].get("staging"), "ell", d["e2e_ell"]["value"], d["e2e_ell"]["resident_ms_per_step"], d["clocks"])
except Exception as e:
    print("sb$m failed", e)
PY
done
