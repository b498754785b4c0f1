#!/bin/bash
# usage (GPU box): bash tools/dp_overlap_ab.sh <n_gpus> [workload] [reserved list]  -> one bench line per reservation
n=$1; w=${2:-mmd}; shift; shift
rs=${@:-"0 16"}
port=29600
for r in $rs; do
  port=$((port+1))
  out=gpurun_out/dp_${w}_${n}gpu_r${r}.json
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --workload $w --steps 20 --warmup 5 --no-cpu --dp-reserved-sms $r > $out 2> gpurun_out/dp_${w}_${n}gpu_r${r}.err
  echo "== $w x$n reserved=$r rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$out").read().strip().splitlines()[-1])
    print("   ", d["value"], d.get("ms_per_step"), "graph", d.get("cuda_graph"), d.get("dp_overlap"), "identical", d.get("dp_params_identical_across_ranks"), "ell", d["e2e_ell"].get("resident_ms_per_step"))
except Exception as e:
    print("    parse error", e)
    print(open("gpurun_out/dp_${w}_${n}gpu_r${r}.err").read()[-1500:])
PY
done
