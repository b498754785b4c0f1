#!/bin/bash
# usage (on the GPU box, from the repo root): bash tools/final_run.sh <tag>
# GPU suite, smoke, the contract bench (both arms), then the ncu launch list of the same bench command (short run).
tag=${1:-final}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/${tag}_status.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_status.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err; echo "bench rc=$?" >> gpurun_out/${tag}_status.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; echo "reference rc=$?" >> gpurun_out/${tag}_status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu rc=$?" >> gpurun_out/${tag}_status.txt
cat gpurun_out/${tag}_status.txt; tail -3 gpurun_out/${tag}_pytest.log; tail -2 gpurun_out/${tag}_smoke.log
python tools/ncu_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt 2>&1; head -12 gpurun_out/${tag}_launches_summary.txt
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "ell", d["e2e_ell"]["value"], "roof", d["roofline"]["tag"], d["roofline"]["frac"], d["roofline_step"]["frac"], "cpu", d["cpu_baseline"], d["clocks"])
PY
tail -c 600 gpurun_out/${tag}_bench_reference_arm.json
