"""Timeline of the decoder chain kernels (debug build: CAREL_EXTRA_FLAGS=-DCAREL_CHAIN_TRACE python -m carel_vae_b200.build --force)
usage: python tools/chain_trace.py [dw|dz]"""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
from carel_vae_b200 import functional as F, _lib
which = sys.argv[1] if len(sys.argv) > 1 else "dz"
n, d, V = 8192, 128, 30000
torch.manual_seed(0)
Z = (torch.randn(n, 2 * d, device="cuda") * 1.1).requires_grad_(True)
W = ((torch.rand(V, 2 * d, device="cuda") * 2 - 1) / 16).requires_grad_(which == "dw")
b = ((torch.rand(V, device="cuda") * 2 - 1) / 16).requires_grad_(which == "dw")
idx = torch.randint(0, V, (n, 24), device="cuda")
bow = torch.zeros(n, V, device="cuda")
bow.scatter_add_(1, idx, torch.ones(n, 24, device="cuda"))
bow = bow / bow.sum(1, keepdim=True).clamp_min(1)
if which == "mmdfwd":
    x = torch.tanh(torch.randn(n, d, device="cuda"))
    y = torch.tanh(torch.randn(n, d, device="cuda"))
    for _ in range(2):
        F.mmd(x, y, [0.1])
elif which == "mmd":
    x = torch.tanh(torch.randn(n, d, device="cuda")).requires_grad_(True)
    y = torch.tanh(torch.randn(n, d, device="cuda")).requires_grad_(True)
    for _ in range(2):
        F.mmd(x, y, [0.1]).backward()
else:
    for _ in range(2):
        o = F.bow_reconstruction_loss(Z, W, b, bow)
        if which != "fwd":
            o.backward()
torch.cuda.synchronize()
lib = C.CDLL("carel_vae_b200/libcarel_head.so")
buf = (C.c_longlong * (8 * 256))()
assert (lib.carel_debug_chain_trace_mmd if which.startswith('mmd') else lib.carel_debug_chain_trace)(buf) == 0
ev = [[buf[s * 256 + i] for i in range(256)] for s in range(8)]
t0 = min(v for v in ev[0] if v > 0)
if which in ("fwd", "mmdfwd"):
    names = ["tile_load", "mma_start", "mma_issued", "epi_start", "epi_done", "-", "-", "-"]
else:
  names = (["y_slot_free", "b_loads_issued", "epi_y_ready", "epi_arrived", "mma_p_ready", "mma_b_ready", "mma_issued", "epi15_arrived"]
           if which in ("dz", "dw") else
           ["load_issue", "g1_issue", "g2_issue", "epi_acc1_ready", "epi_ld_done", "epi_p_arrive", "g1_issued", "g2_issued"])
print("tile " + " ".join(f"{n:>14s}" for n in names))
for i in list(range(0, 8)) + (list(range(8, 29)) if which in ('fwd', 'mmdfwd') else list(range(50, 70))):
    print(f"{i:4d} " + " ".join(f"{ev[s][i] - t0:14d}" for s in range(8)))
for s in range(8):
    d_ = [ev[s][i + 1] - ev[s][i] for i in (range(4, 26) if which in ('fwd', 'mmdfwd') else range(20, 100))]
    print(names[s], "mean period", sum(d_) / len(d_))
