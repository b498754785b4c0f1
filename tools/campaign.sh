#!/bin/bash
# usage (on the GPU box): bash tools/campaign.sh <n_gpus> [workloads...]   -> gpurun_out/r02_<workload>_<n>gpu.json
# One bench.py line per workload (BASELINE.json configs): mmd (cfg2), hsic (cfg3), vi (cfg4), gan, gram_sweep (cfg5).
n=$1; shift
wl=${@:-"mmd vi gram_sweep"}
port=29540
for w in $wl; do
  port=$((port+1))
  out=gpurun_out/r02_${w}_${n}gpu.json
  extra="--steps 20 --warmup 5 --no-cpu"
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --workload $w $extra > $out 2> gpurun_out/r02_${w}_${n}gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --workload $w $extra > $out 2> gpurun_out/r02_${w}_${n}gpu.err
  fi
  echo "== $w x$n rc=$? $(tail -c 400 $out | head -c 0)"
  python - <<PY
import json
try:
    d = json.loads(open("$out").read().strip().splitlines()[-1])
    print("   ", d["metric"], d["value"], d.get("ms_per_step"), "graph", d.get("cuda_graph"), "identical", d.get("dp_params_identical_across_ranks"))
except Exception as e:
    print("    parse error", e)
    print(open("gpurun_out/r02_${w}_${n}gpu.err").read()[-600:])
PY
done
