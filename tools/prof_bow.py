"""ncu target: two decoder forward+backward passes at the bench size (N=8192, K=256, V=30000)."""
import sys
import torch
sys.path.insert(0, ".")
from carel_vae_b200 import functional as F  # noqa: E402
from oracle import head_oracle as ho  # noqa: E402
dev = "cuda"
N, K, V = 8192, 256, 30000
torch.manual_seed(0)
Z = (torch.randn(N, K, device=dev) * 1.1).requires_grad_(True)
W = ((torch.rand(V, K, device=dev) * 2 - 1) / 16).requires_grad_(True)
b = ((torch.rand(V, device=dev) * 2 - 1) / 16).requires_grad_(True)
bow = ho.synthetic_batch(N, V, D=8, nnz=24, seed=0, device=dev)["bow"]
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    loss = F.bow_reconstruction_loss(Z, W, b, bow, 0.1)
    loss.backward()
torch.cuda.synchronize()
print("done", loss.item())
