"""Run one op a few times (for ncu captures): python tools/prof_one.py mmd_fwd|mmd_bwd|bow [n] [d]"""
import sys
import torch
sys.path.insert(0, ".")
from carel_vae_b200 import functional as F
what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
d = int(sys.argv[3]) if len(sys.argv) > 3 else 128
torch.manual_seed(0)
if what.startswith("bow"):
    V = 30000
    Z = (torch.randn(n, 2 * d, device="cuda") * 1.1).requires_grad_(True)
    W = ((torch.rand(V, 2 * d, device="cuda") * 2 - 1) / 16).requires_grad_(True)
    b = ((torch.rand(V, device="cuda") * 2 - 1) / 16).requires_grad_(True)
    idx = torch.randint(0, V, (n, 24), device="cuda")
    bow = torch.zeros(n, V, device="cuda")
    bow.scatter_add_(1, idx, torch.ones(n, 24, device="cuda"))
    bow = bow / bow.sum(1, keepdim=True).clamp_min(1)
    for _ in range(4):
        o = F.bow_reconstruction_loss(Z, W, b, bow)
        o.backward()
else:
    x = torch.tanh(torch.randn(n, d, device="cuda")).requires_grad_(True)
    y = torch.tanh(torch.randn(n, d, device="cuda")).requires_grad_(True)
    for _ in range(5):
        o = F.mmd(x, y, [0.1], full_gram=("full" in what))
        if "bwd" in what:
            o.backward()
torch.cuda.synchronize()
print("ok", float(o))
