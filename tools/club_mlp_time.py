"""Developer script: CLUB approximation network forward + backward at N = 8192, L = 128: fused kernels vs the nn.Sequential
modules (cuBLAS fp32 + elementwise), CUDA-graph replay of each."""
import sys
import torch
import torch.nn as nn
sys.path.insert(0, ".")
from carel_vae_b200 import functional as F  # noqa: E402
dev = "cuda"
N, L = 8192, 128
torch.manual_seed(0)
ec_mu = nn.Sequential(nn.Linear(L, L), nn.ReLU(), nn.Linear(L, L)).to(dev)
ec_lv = nn.Sequential(nn.Linear(L, L), nn.ReLU(), nn.Linear(L, L), nn.Tanh()).to(dev)
z = torch.randn(N, 2 * L, device=dev, requires_grad=True)
gm, gl = torch.randn(N, L, device=dev), torch.randn(N, L, device=dev)
params = list(ec_mu.parameters()) + list(ec_lv.parameters())


def ours():
    mu, lv = F.club_mlp(z[:, L:], ec_mu, ec_lv)
    return torch.autograd.grad((mu * gm).sum() + (lv * gl).sum(), [z] + params)


def theirs():
    c = z[:, L:]
    mu, lv = ec_mu(c), ec_lv(c)
    return torch.autograd.grad((mu * gm).sum() + (lv * gl).sum(), [z] + params)


for name, fn in (("fused", ours), ("torch", theirs)):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        fn()
    for _ in range(5):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(name, "fwd+bwd us:", round(e0.elapsed_time(e1) / 50 * 1e3, 1))
