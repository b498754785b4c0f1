import sys, torch
sys.path.insert(0, ".")
from carel_vae_b200 import functional as F
from oracle import head_oracle as ho
DEV="cuda"
bf=lambda t: t.bfloat16().float()
torch.manual_seed(7)
N, K, V = 64, 48, 500
Z = bf(torch.randn(N, K, device=DEV) * 1.5).requires_grad_(True)
W = bf(torch.randn(V, K, device=DEV) * 0.8).requires_grad_(True)
b = bf(torch.randn(V, device=DEV)).requires_grad_(True)
bow = ho.synthetic_batch(N, V, D=8, nnz=6, seed=1, device=DEV)["bow"]
loss = F.bow_reconstruction_loss(Z, W, b, bow)
loss.backward()
Zd, Wd, bd = (t.detach().double().requires_grad_(True) for t in (Z, W, b))
ref = ho.bow_reconstruction(Zd, Wd, bd, bow.double())
ref.backward()
err = (Z.grad.double() - Zd.grad).norm(dim=1) / Zd.grad.norm(dim=1)
p = torch.softmax(Zd @ Wd.T + bd, 1)
for i in range(N):
    if err[i] > 1e-3:
        print(i, float(err[i]), "pmax", float(p[i].max()), "logpmin", float(p[i].log().min()), "gradnorm", float(Zd.grad[i].norm()))
print("W err", float((W.grad.double()-Wd.grad).norm()/Wd.grad.norm()), "b err", float((b.grad.double()-bd.grad).norm()/bd.grad.norm()))
