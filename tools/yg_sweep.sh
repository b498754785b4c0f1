#!/bin/bash
# usage (GPU box): bash tools/yg_sweep.sh  -> decoder backward kernel times for Y/B ring depth splits (ygrad.cuh yg_stages)
# and for the load probes (CAREL_YG_PROBE: 1 = no B loads, 2 = no Y loads, 3 = neither; results garbage, timing only)
for cfg in "4,3 3,4 3" "4,3 3,4 7" "4,3 3,4 11" "4,3 3,4 19" "4,3 3,4 35" "4,3 3,4 59" "4,3 3,4 4" "4,3 3,4 8"; do
  set -- $cfg
  CAREL_YG_STAGES_DZ=$1 CAREL_YG_STAGES_DW=$2 CAREL_YG_PROBE=$3 python - <<PY
import sys, torch
sys.path.insert(0, ".")
from carel_vae_b200 import functional as F, _lib
from oracle import head_oracle as ho
dev = "cuda"
N, K, V = 8192, 256, 30000
torch.manual_seed(0)
Z = (torch.randn(N, K, device=dev) * 1.1).requires_grad_(True)
W = ((torch.rand(V, K, device=dev) * 2 - 1) / 16).requires_grad_(True)
b = ((torch.rand(V, device=dev) * 2 - 1) / 16).requires_grad_(True)
bow = ho.synthetic_batch(N, V, D=8, nnz=24, seed=0, device=dev)["bow"]
for i in range(6):
    if i == 3:
        _lib.timing_enable(True)
    loss = F.bow_reconstruction_loss(Z, W, b, bow, 0.1)
    loss.backward()
torch.cuda.synchronize()
rep = _lib.timing_report()
print("dz=$1 dw=$2 probe=$3", {k: round(1e3 * v[0] / v[1], 1) for k, v in rep.items()})
PY
done
