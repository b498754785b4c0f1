"""Developer script (GPU box): head-level value / gradient errors of every variant against the fp64 oracle, printed
per parameter (the numbers behind the tolerances of tests/test_gpu_head.py).  usage: python tools/head_err.py [fp32]"""
import sys

import torch

sys.path.insert(0, ".")
import carel_vae_b200 as cv  # noqa: E402
from oracle import head_oracle as ho  # noqa: E402

DEV = "cuda"
prec = sys.argv[1] if len(sys.argv) > 1 else None
if prec:
    cv.set_precision(prec)


def bf(x):
    return x.bfloat16().float()


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def total(variant, out):
    outs = out if isinstance(out, tuple) else (out,)
    if variant == "vi":
        return outs, outs[2] + outs[3]
    if variant == "gan":
        return outs, outs[0] + outs[1] + outs[2]
    return outs, outs[0]


for variant, N, L, V in [("mmd", 64, 64, 3000), ("mmd", 512, 128, 5000), ("mmd", 8192, 128, 30000), ("hsic", 512, 128, 5000),
                         ("vi", 512, 128, 5000), ("gan", 512, 128, 5000), ("hsic", 8192, 128, 30000)]:
    kw = dict(ec_dim=L, pair_bow_dim=V, bert_dim=768, dropout=0.0)
    opt = cv.default_opt(variant, **kw)
    torch.manual_seed(1234 + N)
    model = cv.VARIANTS[variant](opt, encoder=cv.IdentityEncoder()).to(DEV).train()
    rnd = (lambda t: t) if prec == "fp32" else bf
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(rnd(p))
    b = ho.synthetic_batch(N, V, D=768, seed=N, device=DEV, emotion_float=variant in ("hsic", "gan"))
    X = rnd(b["X"]).clone().requires_grad_(True)
    emo = b["emo"] if variant in ("mmd", "vi") else b["emo"].float()
    torch.manual_seed(99)
    out = model(X, None, None, emo, b["cau"], b["pair"], b["bow"], 3)
    outs, tot = total(variant, out)
    params = dict(model.named_parameters())
    grads = torch.autograd.grad(tot, [X] + list(params.values()), allow_unused=True)
    p64 = {k: v.detach().double().requires_grad_(True) for k, v in params.items()}
    X64 = X.detach().double().requires_grad_(True)
    torch.manual_seed(99)
    emo64 = emo if emo.dtype == torch.int64 else emo.double()
    oref = ho.HeadOracle(opt, variant).forward(p64, X64, emo64, b["cau"].double(), b["pair"].double(), b["bow"].double(), 3)
    orefs, rtot = total(variant, oref)
    rgrads = torch.autograd.grad(rtot, [X64] + list(p64.values()), allow_unused=True)
    print(f"== {variant} N={N} L={L} V={V} prec={prec}: outputs " +
          " ".join(f"{rel(o.detach(), r.detach()):.1e}" for o, r in zip(outs, orefs)))
    names = ["X"] + list(params.keys())
    print("   " + "  ".join(f"{k.replace('.weight', '.w').replace('.bias', '.b')} {rel(g, r):.1e}"
                            for k, g, r in zip(names, grads, rgrads) if r is not None and g is not None))
