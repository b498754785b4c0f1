"""TEST INFRASTRUCTURE — generates tests/golden/*.pt by running the UNMODIFIED reference code
(AST-extracted, see reference_extract.py) on seeded synthetic inputs in the build container.

    python oracle/make_golden.py

The fixtures hold the inputs (small) and the reference's outputs: loss values and gradients for
each loss primitive and for the full forward of the four trainer variants.  They travel to the GPU
box, where /root/reference does not exist.
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import head_oracle as ho  # noqa: E402
from oracle import reference_extract as rx  # noqa: E402

OUT = ROOT / "tests" / "golden"


def g_mmd():
    cases = []
    for (n1, n2, d, alphas, seed) in [(16, 16, 24, [0.1], 0), (40, 24, 64, [0.1], 1), (33, 33, 128, [0.1, 1.0, 0.01], 2),
                                      (64, 64, 64, [0.1], 3)]:
        g = torch.Generator().manual_seed(seed)
        x = (torch.randn(n1, d, generator=g) * 1.1).requires_grad_(True)
        y = (torch.randn(n2, d, generator=g) * 1.1 + 0.25).requires_grad_(True)
        ns = rx.load("mmd", ho.default_opt("mmd"))
        stat = ns["MMDStatistic"](n1, n2)
        val, K = stat(x, y, alphas, ret_matrix=True)
        gx, gy = torch.autograd.grad(val, (x, y))
        dist = ns["pdist"](x.detach(), y.detach())
        cases.append(dict(x=x.detach(), y=y.detach(), alphas=alphas, mmd=val.detach(), grad_x=gx, grad_y=gy,
                          kernels=K.detach(), pdist=dist, a00=stat.a00, a11=stat.a11, a01=stat.a01))
    torch.save(cases, OUT / "mmd.pt")


def g_hsic():
    cases = []
    with rx.cpu_cuda_shim():
        for (m, d, sx, sy, seed) in [(16, 24, 1.0, 1.0, 0), (48, 64, 1.0, 1.0, 1), (40, 24, 30.0, 50.0, 2),
                                     (64, 128, 300.0, 300.0, 3)]:
            g = torch.Generator().manual_seed(seed)
            x = torch.randn(m, d, generator=g).requires_grad_(True)
            y = (0.5 * x.detach() + torch.randn(m, d, generator=g)).requires_grad_(True)
            ns = rx.load("hsic", ho.default_opt("hsic"))
            val = ns["HSIC"](x, y, sx, sy)
            gx, gy = torch.autograd.grad(val, (x, y))
            cases.append(dict(x=x.detach(), y=y.detach(), s_x=sx, s_y=sy, hsic=val.detach(), grad_x=gx, grad_y=gy,
                              K=ns["GaussianKernelMatrix"](x.detach(), sx), d2=ns["pairwise_distances"](x.detach())))
    torch.save(cases, OUT / "hsic.pt")


def g_head():
    """Full forward (+ gradients w.r.t. the sentence embeddings and every head parameter) of the
    reference DrlClassifier for the four variants, dropout off and dropout on (seeded)."""
    cases = []
    for variant in ("mmd", "hsic", "vi", "gan"):
        for (N, L, V, D, dropout, seed) in [(16, 24, 150, 96, 0.0, 0), (20, 64, 211, 128, 0.0, 1), (24, 24, 150, 96, 0.5, 2)]:
            opt = ho.default_opt(variant, ec_dim=L, pair_bow_dim=V, bert_dim=D, dropout=dropout)
            ns = rx.load(variant, opt)
            torch.manual_seed(1234 + seed)
            model = ns["DrlClassifier"](opt)
            model.train()
            b = ho.synthetic_batch(N, V, D=D, seed=seed, emotion_float=variant in ("hsic", "gan"))
            X = b["X"].clone().requires_grad_(True)
            emo = b["emo"] if variant in ("mmd", "vi") else b["emo"].float()
            torch.manual_seed(77 + seed)
            with rx.cpu_cuda_shim():
                out = model(X, None, None, emo, b["cau"], b["pair"], b["bow"], 3)
            outs = out if isinstance(out, tuple) else (out,)
            if variant == "vi":
                total = outs[2] + outs[3]
            elif variant == "gan":
                total = outs[0] + outs[1] + outs[2]
            else:
                total = outs[0]
            params = {k: v for k, v in model.named_parameters() if not k.startswith("encoder")}
            grads = torch.autograd.grad(total, [X] + list(params.values()), allow_unused=True)
            rec = dict(variant=variant, N=N, L=L, V=V, D=D, dropout=dropout, seed=seed, iteration=3,
                       state={k: v.detach().clone() for k, v in model.state_dict().items()},
                       outputs=[o.detach().clone() if torch.is_tensor(o) else torch.tensor(float(o)) for o in outs],
                       grad_X=grads[0].detach(),
                       grads={k: (g.detach() if g is not None else None) for k, g in zip(params.keys(), grads[1:])},
                       params_in_get_params=[sum(p.numel() for p in grp) for grp in
                                             ([model.get_params()] if variant in ("mmd", "hsic") else
                                              [list(g) for g in model.get_params()])])
            # the stochastic inference path
            torch.manual_seed(5 + seed)
            rec["pair_preds"] = torch.tensor(model.get_pair_preds(X.detach(), None, None))
            cases.append(rec)
    torch.save(cases, OUT / "head.pt")


def g_misc():
    opt = ho.default_opt("mmd")
    ns = rx.load("mmd", opt)
    m = ns["DrlClassifier"].__new__(ns["DrlClassifier"])
    g = torch.Generator().manual_seed(9)
    mu, lv = torch.randn(20, 24, generator=g), 0.3 * torch.randn(20, 24, generator=g)
    V = 211
    logits = torch.randn(20, V, generator=g)
    bow = ho.synthetic_batch(20, V, seed=4)["bow"]
    torch.manual_seed(3)
    z = ns["DrlClassifier"].sample_prior(m, mu, lv)
    torch.manual_seed(3)
    eps = torch.randn(24)
    ns["opt"].pair_bow_dim = V
    rec = dict(mu=mu, lv=lv, z=z, eps=eps, kl=ns["DrlClassifier"].get_kl_loss(m, mu, lv),
               weights=[ns["DrlClassifier"].get_annealed_weight(m, it, 0.03) for it in (0, 3, 1000, 19999)],
               logits=logits, bow=bow,
               recon=ns["DrlClassifier"].get_reconstruct_loss(m, torch.softmax(logits, 1), bow))
    torch.save(rec, OUT / "misc.pt")


if __name__ == "__main__":
    if not rx.available():
        raise SystemExit("/root/reference is not mounted: golden vectors can only be generated in the build container")
    OUT.mkdir(parents=True, exist_ok=True)
    g_mmd(); g_hsic(); g_misc(); g_head()
    for f in sorted(OUT.glob("*.pt")):
        print(f.name, f.stat().st_size)
