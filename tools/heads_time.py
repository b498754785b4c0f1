"""Device time of the fused classifier heads (forward, backward) at N = 8192, L = 128."""
import sys, torch
import torch.nn as nn
sys.path.insert(0, ".")
from carel_vae_b200 import functional as F
dev = "cuda"
N, L = 8192, 128
torch.manual_seed(0)
z = torch.randn(N, 2 * L, device=dev, requires_grad=True)
ec, cc, pc = nn.Linear(L, 6).to(dev), nn.Linear(L, 1).to(dev), nn.Linear(2 * L, 1).to(dev)
emo = torch.randint(0, 6, (N, 1), device=dev)
cau = (torch.rand(N, 1, device=dev) < 0.5).float(); pair = (torch.rand(N, 1, device=dev) < 0.5).float()
masks = tuple(torch.nn.functional.dropout(torch.ones(N, k, device=dev), 0.5, True) for k in (L, L, 2 * L))
def fwd():
    return F.classifier_heads(z, masks, ec, cc, pc, emo, cau, pair, 0, 0.1, 1.0, 1.0, 1.0, True, (10.0, 10.0, 30.0))[0]
for _ in range(3):
    fwd().backward()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf = tb = 0.0
for _ in range(20):
    ev[0].record(); t = fwd(); ev[1].record(); t.backward(); ev[2].record(); torch.cuda.synchronize()
    tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
print(f"heads fwd {tf / 20 * 1e3:.1f} us, bwd {tb / 20 * 1e3:.1f} us (incl. launch gaps of the wrappers)")
