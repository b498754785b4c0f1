"""Run one op a few times (for ncu captures): python tools/prof_one.py mmd_fwd|mmd_bwd [n] [d]"""
import sys
import torch
sys.path.insert(0, ".")
from carel_vae_b200 import functional as F
what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
d = int(sys.argv[3]) if len(sys.argv) > 3 else 128
torch.manual_seed(0)
x = torch.tanh(torch.randn(n, d, device="cuda")).requires_grad_(True)
y = torch.tanh(torch.randn(n, d, device="cuda")).requires_grad_(True)
for _ in range(5):
    o = F.mmd(x, y, [0.1], full_gram=("full" in what))
    if "bwd" in what:
        o.backward()
torch.cuda.synchronize()
print("ok", float(o))
