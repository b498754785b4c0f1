"""GPU (-m gpu): the fp32 precision tier (north_star: "loss values and gradients within ... 1e-5 in fp32").

Protocol: arbitrary fp32 inputs and parameters (NOT rounded to bf16) go to the CUDA path in precision "fp32" and, as
fp64 copies, to the oracle restatement (oracle/head_oracle.py); values and gradients must agree to 1e-5 relative.
The same tier is then held against the outputs of the UNMODIFIED reference (tests/golden/*.pt, fp32 on the CPU):
values AND gradients at 1e-5, dropout cases included (the reference's CPU masks are replayed, see _ReplayDropout).
"""
import ctypes as C

import pytest
import torch

from oracle import head_oracle as ho

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import carel_vae_b200 as cv
    from carel_vae_b200 import functional as F
    from carel_vae_b200 import _lib
DEV = "cuda"
TOL = 1e-5


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def relv(a, b):
    return abs(float(a) - float(b)) / max(abs(float(b)), 1e-300)


@pytest.fixture
def fp32_tier():
    cv.set_precision("fp32")
    yield
    cv.set_precision("bf16")


# ------------------------------------------------------------------------------------------
# the fp32 GEMM itself: every layout switch, split-K, bias, ragged shapes
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ta,tb,M,N,K,mode,ksplit", [(0, 0, 300, 200, 77, 0, 1), (0, 1, 129, 257, 1000, 0, 1),
                                                     (1, 0, 64, 130, 513, 0, 1), (1, 1, 200, 96, 4096, 2, 0),
                                                     (1, 1, 30, 17, 5, 1, 1), (0, 0, 1, 1, 1, 0, 1)])
def test_f32_gemm(ta, tb, M, N, K, mode, ksplit):
    torch.manual_seed(M + N + K)
    a = torch.randn((K, M) if ta else (M, K), device=DEV)
    b = torch.randn((K, N) if tb else (N, K), device=DEV)
    bias = torch.randn(N, device=DEV)
    c0 = torch.randn(M, N, device=DEV) if mode == 1 else torch.zeros(M, N, device=DEV)
    c = c0.clone()
    rc = _lib.lib().carel_f32_gemm(C.c_void_p(a.data_ptr()), a.stride(0), ta, C.c_void_p(b.data_ptr()), b.stride(0), tb, M, N, K,
                                   C.c_void_p(c.data_ptr()), N, C.c_void_p(bias.data_ptr()), 0.5, mode, ksplit,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    A = (a.T if ta else a).double()
    B = (b if tb else b.T).double()
    ref = 0.5 * (A @ B) + bias.double() + (c0.double() if mode == 1 else 0.0)
    assert rel(c, ref) < 2e-6


# ------------------------------------------------------------------------------------------
# op level
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,D,L", [(64, 768, 64), (301, 100, 24), (8192, 768, 128)])
def test_latent_fp32(N, D, L):
    torch.manual_seed(N + L)
    x = torch.tanh(torch.randn(N, D, device=DEV)).requires_grad_(True)
    lins = [torch.nn.Linear(D, L).to(DEV) for _ in range(4)]
    eps_e, eps_c = torch.randn(L, device=DEV), torch.randn(L, device=DEV)
    wb = [t for lin in lins for t in (lin.weight, lin.bias)]
    z, kl, _ = F.latent(x, eps_e, eps_c, *wb, precision="fp32")
    gz = torch.randn_like(z)
    w_kl = torch.tensor([0.3, 0.7], device=DEV)
    grads = torch.autograd.grad((z * gz).sum() + (kl * w_kl).sum(), [x] + wb)
    xd = x.detach().double().requires_grad_(True)
    wbd = [t.detach().double().requires_grad_(True) for t in wb]
    mu_e, lv_e = xd @ wbd[0].T + wbd[1], xd @ wbd[2].T + wbd[3]
    mu_c, lv_c = xd @ wbd[4].T + wbd[5], xd @ wbd[6].T + wbd[7]
    zd = torch.cat((ho.sample_prior(mu_e, lv_e, eps_e.double()), ho.sample_prior(mu_c, lv_c, eps_c.double())), 1)
    kld = torch.stack((ho.kl_loss(mu_e, lv_e), ho.kl_loss(mu_c, lv_c)))
    rgrads = torch.autograd.grad((zd * gz.double()).sum() + (kld * w_kl.double()).sum(), [xd] + wbd)
    assert rel(z, zd) < TOL and rel(kl, kld) < TOL
    for g, r in zip(grads, rgrads):
        assert rel(g, r) < TOL


@pytest.mark.parametrize("n1,n2,d,alphas", [(2, 2, 8, [0.1]), (64, 64, 64, [0.1]), (300, 300, 128, [0.1, 0.01, 1.0]),
                                            (129, 257, 128, [0.1]), (500, 700, 768, [0.05]), (2048, 2048, 128, [0.1])])
def test_mmd_fp32(n1, n2, d, alphas):
    torch.manual_seed(n1 + d + 1)
    scale = 1.1 if d <= 128 else 0.3
    x = (torch.randn(n1, d, device=DEV) * scale).requires_grad_(True)
    y = (torch.randn(n2, d, device=DEV) * scale + 0.3 * scale).requires_grad_(True)
    xd, yd = x.detach().double().requires_grad_(True), y.detach().double().requires_grad_(True)
    ref = ho.mmd_statistic(xd, yd, alphas)
    rgx, rgy = torch.autograd.grad(ref, (xd, yd))
    out = F.mmd(x, y, alphas, precision="fp32")
    gx, gy = torch.autograd.grad(out, (x, y))
    assert relv(out, ref) < TOL
    assert rel(gx, rgx) < TOL and rel(gy, rgy) < TOL


def test_mmd_bf16_tier_is_differentiable_at_any_width():
    """d = 768 (the Gram sweep's width): the bf16 tier's forward runs on the tensor cores, its gradient on the
    materialising fp32 kernels (the fused backward holds <= 250 columns).  Reference MMDStatistic is differentiable
    at any width (mmd:547-596)."""
    torch.manual_seed(3)
    x = torch.tanh(torch.randn(700, 768, device=DEV))
    y = torch.tanh(torch.randn(900, 768, device=DEV))
    x = (x / x.norm(dim=1, keepdim=True)).bfloat16().float().requires_grad_(True)
    y = (y / y.norm(dim=1, keepdim=True)).bfloat16().float().requires_grad_(True)
    xd, yd = x.detach().double().requires_grad_(True), y.detach().double().requires_grad_(True)
    ref = ho.mmd_statistic(xd, yd, [0.1, 1.0])
    rgx, rgy = torch.autograd.grad(ref, (xd, yd))
    out = cv.MMDStatistic(700, 900)(x, y, [0.1, 1.0])
    gx, gy = torch.autograd.grad(out, (x, y))
    assert relv(out, ref) < 1e-4
    assert rel(gx, rgx) < TOL and rel(gy, rgy) < TOL


@pytest.mark.parametrize("m,dx,dy,s", [(16, 24, 24, 1.0), (300, 64, 48, 100.0), (1000, 128, 128, 300.0), (4096, 128, 128, 350.0)])
def test_hsic_fp32(m, dx, dy, s):
    torch.manual_seed(m + dx)
    x = (torch.randn(m, dx, device=DEV) * 1.1).requires_grad_(True)
    y = (0.5 * x.detach()[:, :dy] + torch.randn(m, dy, device=DEV)).requires_grad_(True)
    xd, yd = x.detach().double().requires_grad_(True), y.detach().double().requires_grad_(True)
    ref = ho.hsic_fast(xd, yd, s, s)
    rgx, rgy = torch.autograd.grad(ref, (xd, yd))
    out = F.hsic(x, y, s, s, precision="fp32")
    gx, gy = torch.autograd.grad(out, (x, y))
    assert relv(out, ref) < TOL
    # at a degenerate bandwidth (K ~ I) the gradient is ~1e-15 of the value and pure rounding in fp32 on
    # either side: it is compared against max(|grad|, |value|), i.e. absolutely on the value's scale
    for g, r in ((gx, rgx), (gy, rgy)):
        assert (g.double() - r).norm().item() < TOL * max(r.norm().item(), abs(float(ref)))


@pytest.mark.parametrize("N,K,V,nnz,ell", [(64, 48, 300, 5, False), (1, 48, 100, 3, False), (130, 128, 777, 40, True),
                                           (512, 256, 30000, 24, False), (300, 40, 1001, 5, True)])
def test_bow_fp32(N, K, V, nnz, ell):
    torch.manual_seed(N + V)
    Z = (torch.randn(N, K, device=DEV) * 1.1).requires_grad_(True)
    W = ((torch.rand(V, K, device=DEV) * 2 - 1) / K ** 0.5).requires_grad_(True)
    b = ((torch.rand(V, device=DEV) * 2 - 1) / K ** 0.5).requires_grad_(True)
    bow = ho.synthetic_batch(N, V, D=8, nnz=nnz, seed=N, device=DEV)["bow"]
    if ell:
        loss = F.bow_reconstruction_loss(Z, W, b, None, 0.1, ell=F.bow_pack_ell(bow, 64), precision="fp32")
    else:
        loss = F.bow_reconstruction_loss(Z, W, b, bow, 0.1, precision="fp32")
    grads = torch.autograd.grad(loss, (Z, W, b))
    Zd, Wd, bd = (t.detach().double().requires_grad_(True) for t in (Z, W, b))
    ref = ho.bow_reconstruction(Zd, Wd, bd, bow.double(), 0.1)
    rgrads = torch.autograd.grad(ref, (Zd, Wd, bd))
    assert relv(loss, ref) < TOL
    for g, r in zip(grads, rgrads):
        assert rel(g, r) < TOL


# ------------------------------------------------------------------------------------------
# whole head, every variant, incl. BASELINE configs[0] (N = 64, L = 64, V = 30000)
# ------------------------------------------------------------------------------------------
def _total(variant, out):
    outs = out if isinstance(out, tuple) else (out,)
    if variant == "vi":
        return outs, outs[2] + outs[3]
    if variant == "gan":
        return outs, outs[0] + outs[1] + outs[2]
    return outs, outs[0]


@pytest.mark.parametrize("variant", ["mmd", "hsic", "vi", "gan"])
@pytest.mark.parametrize("N,L,V,it", [(64, 64, 30000, 3), (96, 24, 1001, 3), (512, 128, 5000, 25000)])
def test_head_fp32_matches_oracle(fp32_tier, variant, N, L, V, it):
    """`it` = 25000 >= kl_ann_iterations exercises the un-annealed KL branch (mmd:242-250)."""
    opt = cv.default_opt(variant, ec_dim=L, pair_bow_dim=V, bert_dim=768, dropout=0.0)
    torch.manual_seed(1234 + N)
    model = cv.VARIANTS[variant](opt, encoder=cv.IdentityEncoder()).to(DEV).train()
    b = ho.synthetic_batch(N, V, D=768, seed=N, device=DEV, emotion_float=variant in ("hsic", "gan"))
    X = b["X"].clone().requires_grad_(True)
    emo = b["emo"] if variant in ("mmd", "vi") else b["emo"].float()
    torch.manual_seed(99)
    outs, total = _total(variant, model(X, None, None, emo, b["cau"], b["pair"], b["bow"], it))
    params = dict(model.named_parameters())
    grads = torch.autograd.grad(total, [X] + list(params.values()), allow_unused=True)
    p64 = {k: v.detach().double().requires_grad_(True) for k, v in params.items()}
    X64 = X.detach().double().requires_grad_(True)
    torch.manual_seed(99)
    emo64 = emo if emo.dtype == torch.int64 else emo.double()
    orefs, rtotal = _total(variant, ho.HeadOracle(opt, variant).forward(p64, X64, emo64, b["cau"].double(), b["pair"].double(),
                                                                        b["bow"].double(), it))
    rgrads = torch.autograd.grad(rtotal, [X64] + list(p64.values()), allow_unused=True)
    for o, r in zip(outs, orefs):
        assert rel(o.detach(), r.detach()) < TOL, (variant, "output")
    for k, g, r in zip(["X"] + list(params.keys()), grads, rgrads):
        if r is None:
            assert g is None or float(g.abs().max()) == 0.0
            continue
        assert rel(g, r) < TOL, (variant, k, rel(g, r))


class _ReplayDropout(torch.nn.Module):
    """Stands in for model.dropout: multiplies by the mask the REFERENCE drew on the CPU generator.  The reference's
    nn.Dropout on a CPU tensor draws `empty_like(x).bernoulli_(1 - p)`, which depends on the shape only; drawing
    `dropout(ones(shape))` on the CPU in the same call order therefore reproduces its masks exactly."""

    def __init__(self, p):
        super().__init__()
        self.p = p

    def forward(self, x):
        if not self.training or self.p == 0:
            return x
        return x * torch.nn.functional.dropout(torch.ones(x.shape), self.p, True).to(x.device)


@pytest.mark.parametrize("variant", ["mmd", "hsic", "vi", "gan"])
def test_golden_reference_head_fp32(fp32_tier, golden_dir, variant):
    """Outputs AND gradients of the unmodified reference (tests/golden/head.pt: fp32 on the CPU, seeded noise, dropout
    off and on) at 1e-5: the reference checkpoint loads unchanged, its CPU noise / dropout stream is replayed."""
    for c in [c for c in torch.load(golden_dir / "head.pt") if c["variant"] == variant]:
        opt = cv.default_opt(variant, ec_dim=c["L"], pair_bow_dim=c["V"], bert_dim=c["D"], dropout=c["dropout"])
        model = cv.VARIANTS[variant](opt, encoder=cv.IdentityEncoder())
        model.load_state_dict(c["state"])
        model = model.to(DEV).train()
        model.dropout = _ReplayDropout(c["dropout"])
        model.dropout.train()
        b = ho.synthetic_batch(c["N"], c["V"], D=c["D"], seed=c["seed"], device=DEV, emotion_float=variant in ("hsic", "gan"))
        emo = b["emo"] if variant in ("mmd", "vi") else b["emo"].float()
        X = b["X"].clone().requires_grad_(True)
        torch.manual_seed(77 + c["seed"])                      # the reference's CPU stream: randn(L), randn(L), masks ...
        orig = torch.randn
        try:
            torch.randn = lambda *a, **k: orig(*a).to(DEV)     # CPU draw (device kwarg dropped), moved to the GPU
            out = model(X, None, None, emo, b["cau"], b["pair"], b["bow"], c["iteration"])
        finally:
            torch.randn = orig
        outs, total = _total(variant, out)
        params = {k: v for k, v in model.named_parameters() if not k.startswith("encoder")}
        grads = torch.autograd.grad(total, [X] + list(params.values()), allow_unused=True)
        for o, r in zip(outs, c["outputs"]):
            assert rel(o.detach().cpu(), r) < TOL, (variant, c["dropout"], float(o.flatten()[0]), float(r.flatten()[0]))
        assert rel(grads[0].cpu(), c["grad_X"]) < 2 * TOL, (variant, c["dropout"], "grad X", rel(grads[0].cpu(), c["grad_X"]))
        for (k, g) in zip(params.keys(), grads[1:]):
            r = c["grads"][k]
            if r is None:
                assert g is None or float(g.abs().max()) == 0.0
                continue
            # the golden gradients are the reference's own fp32 results (1e-6-level rounding of their own)
            assert rel(g.cpu(), r) < 2 * TOL, (variant, c["dropout"], k, rel(g.cpu(), r))


def test_golden_reference_primitives_fp32(golden_dir):
    """tests/golden/{mmd,hsic}.pt: value and gradients of the reference's own MMDStatistic / HSIC at 1e-5."""
    for c in torch.load(golden_dir / "mmd.pt"):
        x, y = c["x"].to(DEV).requires_grad_(True), c["y"].to(DEV).requires_grad_(True)
        out = F.mmd(x, y, c["alphas"], precision="fp32")
        gx, gy = torch.autograd.grad(out, (x, y))
        # The statistic is a difference of kernel means (a00 sum K_xx + a11 sum K_yy + 2 a01 sum K_xy) that can cancel to
        # 1e-3 of their size; the golden value is the reference's own fp32 result and carries that cancellation noise.
        # Against it: 1e-5 of the kernel-mean scale.  Against the fp64 evaluation of the same inputs: 1e-5 of the value.
        scale = max(abs(float(c["mmd"])), float(c["kernels"].mean()))
        assert abs(float(out) - float(c["mmd"])) < TOL * scale
        assert relv(out, ho.mmd_statistic(c["x"].double(), c["y"].double(), c["alphas"])) < TOL
        assert rel(gx.cpu(), c["grad_x"]) < TOL and rel(gy.cpu(), c["grad_y"]) < TOL
    for c in torch.load(golden_dir / "hsic.pt"):
        x, y = c["x"].to(DEV).requires_grad_(True), c["y"].to(DEV).requires_grad_(True)
        out = F.hsic(x, y, c["s_x"], c["s_y"], precision="fp32")
        gx, gy = torch.autograd.grad(out, (x, y))
        assert relv(out, c["hsic"]) < TOL
        # degenerate default bandwidth (s = 1 at 24 dims, K ~ I): the true gradient is ~1e-12 and the reference's own
        # fp32 gradient is rounding noise (multiples of 2^-30); SURVEY §8(d): compare absolutely on the value's scale
        for g, r in ((gx, c["grad_x"]), (gy, c["grad_y"])):
            assert (g.cpu() - r).abs().max().item() < TOL * max(r.abs().max().item(), abs(float(c["hsic"])))
