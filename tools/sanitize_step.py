"""compute-sanitizer target: one small train step (forward + backward) of every head variant in the bf16 tier and of
the MMD / HSIC variants in the fp32 tier, plus the wide-Gram paths.  Small shapes: the sanitizer slows kernels ~100x.
    compute-sanitizer --tool memcheck python tools/sanitize_step.py"""
import sys

import torch

sys.path.insert(0, ".")
import carel_vae_b200 as cv  # noqa: E402
from carel_vae_b200 import functional as F  # noqa: E402

dev = "cuda"
N, L, V, D = 200, 64, 1500, 768
g = torch.Generator().manual_seed(0)
X = torch.tanh(torch.randn(N, D, generator=g)).to(dev)
emo = torch.randint(0, 6, (N, 1), generator=g).to(dev)
pair = (torch.rand(N, 1, generator=g) < 0.5).float().to(dev)
idx = torch.randint(0, V, (N, 12), generator=g)
bow = torch.zeros(N, V)
bow.scatter_add_(1, idx, torch.ones(N, 12))
bow = (bow / bow.sum(1, keepdim=True).clamp_min(1.0)).to(dev)
for prec in ("bf16", "fp32"):
    cv.set_precision(prec)
    for variant in ("mmd", "hsic", "vi", "gan") if prec == "bf16" else ("mmd", "hsic"):
        opt = cv.default_opt(variant, ec_dim=L, pair_bow_dim=V, bert_dim=D, dropout=0.5)
        torch.manual_seed(1)
        model = cv.VARIANTS[variant](opt, encoder=cv.IdentityEncoder()).to(dev).train()
        e = emo if variant in ("mmd", "vi") else pair
        out = model(X, None, None, e, pair, pair, bow, 3)
        outs = out if isinstance(out, tuple) else (out,)
        total = outs[-1] if variant != "vi" else outs[2] + outs[3]
        if variant == "gan":
            total = outs[0] + outs[1] + outs[2]
        if variant == "vi":
            total = total + 0.1 * model.get_ec_upper_loss(outs[0], outs[1])
        total.backward()
        torch.cuda.synchronize()
        print(prec, variant, float(total.detach()), flush=True)
cv.set_precision("bf16")
# ELL target, wide Gram (CTA-pair forward + materialising backward), sharded row range
cols, vals, nnz = F.bow_pack_ell(bow, 16)
z = torch.randn(N, 2 * L, device=dev, requires_grad=True)
w = (torch.randn(V, 2 * L, device=dev) / 16).requires_grad_(True)
b = torch.zeros(V, device=dev, requires_grad=True)
F.bow_reconstruction_loss(z, w, b, None, 0.1, ell=(cols, vals, nnz)).backward()
x = torch.randn(300, 768, device=dev, requires_grad=True)
y = torch.randn(260, 768, device=dev, requires_grad=True)
cv.MMDStatistic(300, 260)(x, y, [0.01]).backward()
xs = torch.randn(300, 64, device=dev, requires_grad=True)
ys = torch.randn(300, 64, device=dev, requires_grad=True)
cv.HSIC(xs, ys, 50.0, 50.0).backward()
torch.cuda.synchronize()
print("sanitize target done")
