"""GPU (-m gpu): the drop-in DrlClassifier variants against the oracle restatement of the reference
forward (oracle/head_oracle.py, itself pinned to the reference by tests/test_oracle_golden.py).

Protocol ("bf16 mode"): parameters and clause embeddings are rounded to bf16 once and handed to both
sides; the oracle then runs in fp64.  Same device RNG stream on both sides, so eps_e, eps_c and the
dropout masks are identical (RNG consumption order is part of the drop-in contract).
"""
import pytest
import torch

from oracle import head_oracle as ho

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import carel_vae_b200 as cv
DEV = "cuda"


def bf(x):
    return x.bfloat16().float()


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def _setup(variant, N, L, V, D, dropout, seed):
    opt = cv.default_opt(variant, ec_dim=L, pair_bow_dim=V, bert_dim=D, dropout=dropout)
    torch.manual_seed(1234 + seed)
    model = cv.VARIANTS[variant](opt, encoder=cv.IdentityEncoder()).to(DEV)
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(bf(p))
    b = ho.synthetic_batch(N, V, D=D, seed=seed, device=DEV, emotion_float=variant in ("hsic", "gan"))
    b["X"] = bf(b["X"])
    return opt, model, b


def _total(variant, out):
    outs = out if isinstance(out, tuple) else (out,)
    if variant == "vi":
        return outs, outs[2] + outs[3]
    if variant == "gan":
        return outs, outs[0] + outs[1] + outs[2]
    return outs, outs[0]


@pytest.mark.parametrize("variant", ["mmd", "hsic", "vi", "gan"])
@pytest.mark.parametrize("N,L,V,D,dropout", [(64, 64, 3000, 768, 0.0), (96, 24, 1001, 768, 0.5), (512, 128, 5000, 768, 0.0)])
def test_forward_backward_matches_oracle(variant, N, L, V, D, dropout):
    opt, model, b = _setup(variant, N, L, V, D, dropout, seed=N)
    model.train()
    X = b["X"].clone().requires_grad_(True)
    emo = b["emo"] if variant in ("mmd", "vi") else b["emo"].float()
    torch.manual_seed(99)
    out = model(X, None, None, emo, b["cau"], b["pair"], b["bow"], 3)
    outs, total = _total(variant, out)
    params = dict(model.named_parameters())
    grads = torch.autograd.grad(total, [X] + list(params.values()), allow_unused=True)

    # oracle on the same RNG stream.  fp64 when dropout is off; with dropout the masks only coincide if
    # the dropout inputs have the product's dtype (fp32), so the oracle then runs in fp32 like the reference.
    dt = torch.float64 if dropout == 0.0 else torch.float32
    p64 = {k: v.detach().to(dt).requires_grad_(True) for k, v in params.items()}
    X64 = b["X"].to(dt).requires_grad_(True)
    torch.manual_seed(99)
    emo64 = emo if emo.dtype == torch.int64 else emo.to(dt)
    oref = ho.HeadOracle(opt, variant).forward(p64, X64, emo64, b["cau"].to(dt), b["pair"].to(dt), b["bow"].to(dt), 3)
    orefs, rtotal = _total(variant, oref)
    rgrads = torch.autograd.grad(rtotal, [X64] + list(p64.values()), allow_unused=True)

    for o, r in zip(outs, orefs):
        assert rel(o.detach(), r.detach()) < (1e-4 if dropout == 0.0 else 1e-3), (variant, "output")
    # north_star tolerance for the bf16 tier: 1e-3 relative.  The latent sample z is an INTERMEDIATE, so its rounding
    # as a tensor-core operand is not shared with the oracle: the MMD and decoder kernels take it as FP16 (11 bits),
    # measured 1e-5..2e-4 (tools/head_err.py).  (With dropout the ORACLE runs in fp32 — see above — and carries its
    # own rounding noise on the Gram-loss gradient.)
    gtol = 1e-3
    assert rel(grads[0], rgrads[0]) < gtol, (variant, "grad X", rel(grads[0], rgrads[0]))
    for (k, g), r in zip(zip(params.keys(), grads[1:]), rgrads[1:]):
        if r is None:
            assert g is None or float(g.abs().max()) == 0.0
            continue
        assert rel(g, r) < gtol, (variant, k, rel(g, r))


@pytest.mark.parametrize("variant", ["mmd", "hsic", "vi", "gan"])
def test_golden_reference_head(golden_dir, variant):
    """bf16 tier against the reference's own outputs (fp32; inputs and parameters NOT bf16-representable, so the 2^-9
    rounding of the 768-wide projection operands is the whole error budget here: ~3e-3 on the latent codes).  The
    same golden vectors are met at 1e-5, gradients and dropout cases included, by the fp32 tier:
    tests/test_gpu_fp32.py::test_golden_reference_head_fp32."""
    for c in [c for c in torch.load(golden_dir / "head.pt") if c["variant"] == variant]:
        opt = cv.default_opt(variant, ec_dim=c["L"], pair_bow_dim=c["V"], bert_dim=c["D"], dropout=c["dropout"])
        model = cv.VARIANTS[variant](opt, encoder=cv.IdentityEncoder())
        model.load_state_dict(c["state"])          # reference checkpoint loads unchanged
        model = model.to(DEV).train()
        if c["dropout"] > 0:
            continue                               # CPU and CUDA generators differ: masks not comparable
        b = ho.synthetic_batch(c["N"], c["V"], D=c["D"], seed=c["seed"], device=DEV,
                               emotion_float=variant in ("hsic", "gan"))
        emo = b["emo"] if variant in ("mmd", "vi") else b["emo"].float()
        # inject the reference's CPU-generated noise: seed 77+seed, two randn(L) draws
        torch.manual_seed(77 + c["seed"])
        eps = [torch.randn(c["L"]), torch.randn(c["L"])]
        it = iter(eps)
        orig = torch.randn
        try:
            torch.randn = lambda *a, **k: next(it).to(DEV)
            out = model(b["X"], None, None, emo, b["cau"], b["pair"], b["bow"], c["iteration"])
        finally:
            torch.randn = orig
        outs = out if isinstance(out, tuple) else (out,)
        for o, r in zip(outs, c["outputs"]):
            assert rel(o.detach().cpu(), r) < 1e-2, (variant, float(o.flatten()[0]), float(r.flatten()[0]))


def test_vi_upper_loss_and_two_optimisers():
    """vi_final train loop (:766-781): aprx step, then get_ec_upper_loss with the updated net."""
    opt, model, b = _setup("vi", 128, 24, 1000, 768, 0.0, seed=3)
    aprx_params, core_params = model.get_params()
    o1 = torch.optim.Adam(aprx_params, lr=3e-3)
    o2 = torch.optim.Adam(core_params, lr=1e-5)
    z_e, z_c, aprx, loss = model(b["X"], None, None, b["emo"], b["cau"], b["pair"], b["bow"], 0)
    o1.zero_grad()
    aprx.backward(retain_graph=True)
    o1.step()
    torch.manual_seed(0)
    rj = model.get_ec_upper_loss(z_e, z_c)
    loss = loss + 0.5 * rj
    o2.zero_grad()
    loss.backward()
    o2.step()
    assert torch.isfinite(loss) and torch.isfinite(rj)


def test_get_pair_preds_and_inference_tensor():
    opt, model, b = _setup("mmd", 200, 24, 1000, 768, 0.5, seed=4)
    model.eval()
    torch.manual_seed(11)
    preds = model.get_pair_preds(b["X"], None, None)
    torch.manual_seed(11)
    t = model.get_pair_preds_tensor(b["X"], None, None)
    assert isinstance(preds, list) and len(preds) == 200 and len(preds[0]) == 1
    assert preds == t.detach().cpu().numpy().round().tolist()
    p = {k: v.detach() for k, v in model.named_parameters()}
    torch.manual_seed(11)
    ref = ho.HeadOracle(opt, "mmd").pair_preds(p, b["X"])
    assert (t.detach() - ref).abs().max().item() < 2e-3


def test_train_steps_reduce_loss():
    """A few Adam steps over get_params() (the reference's optimiser set) run under anomaly detection
    (mmd:837) and move the loss."""
    opt, model, b = _setup("mmd", 256, 64, 2000, 768, 0.0, seed=5)
    optim = torch.optim.Adam(model.get_params(), lr=1e-3)
    losses = []
    for it in range(5):
        torch.manual_seed(1)
        with torch.autograd.set_detect_anomaly(True):
            loss = model(b["X"], None, None, b["emo"], b["cau"], b["pair"], b["bow"], it)
            optim.zero_grad()
            loss.backward()
            optim.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]
    assert all(p.grad is not None for p in model.emotion_mu.parameters())   # projections receive gradients ...
    assert not any(p is q for p in model.get_params() for q in model.emotion_mu.parameters())  # ... but are not stepped


@pytest.mark.parametrize("variant", ["mmd", "hsic", "vi", "gan"])
def test_full_size_head_step_matches_oracle(variant):
    """BASELINE configs[1] (MMD), [2] (HSIC), [3] (VI) and the GAN trainer at their real size (N = 8192 clauses, d = 768,
    latent 128, vocab 30000, dropout off): every output and every gradient of one forward/backward against the fp64
    oracle, bf16 tier, north_star tolerance 1e-3."""
    N, L, V, D = 8192, 128, 30000, 768
    opt, model, b = _setup(variant, N, L, V, D, 0.0, seed=5)
    model.train()
    X = b["X"].clone().requires_grad_(True)
    emo = b["emo"] if variant in ("mmd", "vi") else b["emo"].float()
    torch.manual_seed(99)
    outs, total = _total(variant, model(X, None, None, emo, b["cau"], b["pair"], b["bow"], 3))
    params = dict(model.named_parameters())
    grads = torch.autograd.grad(total, [X] + list(params.values()), allow_unused=True)
    p64 = {k: v.detach().double().requires_grad_(True) for k, v in params.items()}
    X64 = b["X"].double().requires_grad_(True)
    torch.manual_seed(99)
    emo64 = emo if emo.dtype == torch.int64 else emo.double()
    orefs, rtotal = _total(variant, ho.HeadOracle(opt, variant).forward(p64, X64, emo64, b["cau"].double(), b["pair"].double(),
                                                                        b["bow"].double(), 3))
    rgrads = torch.autograd.grad(rtotal, [X64] + list(p64.values()), allow_unused=True)
    for o, r in zip(outs, orefs):
        assert rel(o.detach(), r.detach()) < 1e-4, (variant, "output", rel(o.detach(), r.detach()))
    assert rel(grads[0], rgrads[0]) < 1e-3, ("grad X", rel(grads[0], rgrads[0]))
    for (k, g), r in zip(zip(params.keys(), grads[1:]), rgrads[1:]):
        if r is None:
            continue
        assert rel(g, r) < 1e-3, (variant, k, rel(g, r))
    del p64, X64, orefs, rtotal, rgrads
    torch.cuda.empty_cache()


def test_weighted_sum_is_bit_identical_to_the_scalar_chain():
    """F.weighted_sum (loss assembly, mmd:256-261) against the torch expression it replaces: value bit-exact,
    gradients exact (w_i * g), 1-D terms (the two KL sums) get a gradient of their own shape."""
    from carel_vae_b200 import functional as F
    torch.manual_seed(5)
    mmd = (torch.rand((), device=DEV) * 1e-2).requires_grad_()
    heads = (torch.rand((), device=DEV) * 50).requires_grad_()
    kl = (torch.rand(2, device=DEV) * 30).requires_grad_()
    recon = (torch.rand((), device=DEV) * 1e-3).requires_grad_()
    w = 7.412e-6
    ours = F.weighted_sum([mmd, heads, kl, recon], [-30.0, 1.0, (w, w), 1.0])
    ref = 30 * (-mmd) + heads + w * kl[0] + w * kl[1] + recon
    assert ours.shape == () and torch.equal(ours.detach(), ref.detach())
    g_ours = torch.autograd.grad(ours * 3.0, [mmd, heads, kl, recon])
    g_ref = torch.autograd.grad(ref * 3.0, [mmd, heads, kl, recon])
    for a, b in zip(g_ours, g_ref):
        assert a.shape == b.shape and torch.equal(a, b)
    with pytest.raises(ValueError):
        F.weighted_sum([kl], [1.0])
    with pytest.raises(cv._lib.CarelError):
        F.weighted_sum([torch.ones(17, device=DEV)], [tuple([1.0] * 17)])


def test_vi_step_is_graph_capturable_and_draws_the_reference_permutation():
    """CLUB negatives: same CPU-generator permutation as `torch.randperm(N)` (vi_final:433), delivered by an async copy
    from pinned memory, so the VI train step can be replayed as a CUDA graph; the replay picks up the permutation
    drawn by redraw_negative_index."""
    opt, model, b = _setup("vi", 128, 24, 1000, 768, 0.0, seed=3)
    z_e, z_c = torch.randn(128, 24, device=DEV), torch.randn(128, 24, device=DEV)
    torch.manual_seed(11)
    ours = model.get_ec_upper_loss(z_e, z_c)
    torch.manual_seed(11)
    perm = torch.randperm(128).to(DEV)
    mu, lv = model.get_ec_emb(z_c)
    ref = ((-(mu - z_e) ** 2 / lv.exp()).sum(-1) - (-(mu - z_e[perm]) ** 2 / lv.exp()).sum(-1)).mean() / 2.0
    assert torch.allclose(ours, ref, rtol=1e-5, atol=1e-6)       # same permutation (a different one is off by O(1))
    out = torch.zeros((), device=DEV)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        out.copy_(model.get_ec_upper_loss(z_e, z_c).detach())
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(g):
        out.copy_(model.get_ec_upper_loss(z_e, z_c))
    for seed in (5, 6):
        torch.manual_seed(seed)
        model.redraw_negative_index(128)
        g.replay()
        torch.cuda.synchronize()
        torch.manual_seed(seed)
        perm = torch.randperm(128).to(DEV)
        ref = ((-(mu - z_e) ** 2 / lv.exp()).sum(-1) - (-(mu - z_e[perm]) ** 2 / lv.exp()).sum(-1)).mean() / 2.0
        assert torch.allclose(out, ref.detach(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("N,L", [(1, 24), (37, 5), (300, 64), (2048, 128)])
def test_club_kernels_match_the_reference_formulas(N, L):
    """carel_club_fwd/bwd against get_ec_aprx_loss / get_ec_upper_loss (vi_final:421-439) written out in fp64:
    value and all three gradients to 1e-5; e is a strided column view of z; the permutation has fixed points."""
    from carel_vae_b200 import functional as F
    torch.manual_seed(N + L)
    z = torch.randn(N, 2 * L, device=DEV)
    mu = torch.randn(N, L, device=DEV).requires_grad_()
    lv = torch.tanh(torch.randn(N, L, device=DEV)).requires_grad_()
    e = z[:, :L].clone().requires_grad_()
    perm = torch.randperm(N).tolist()
    for i in range(max(1, N // 7)):                 # a few fixed points pi(i) = i
        j = perm.index(i)
        perm[i], perm[j] = perm[j], perm[i]
    assert sorted(perm) == list(range(N)) and perm[0] == 0
    perm = torch.tensor(perm, device=DEV)
    mud, lvd, ed = (t.detach().double().requires_grad_() for t in (mu, lv, e))
    refs = {"aprx": -((-(mud - ed) ** 2 / lvd.exp() - lvd).sum(1).mean(0)),
            "upper": ((-(mud - ed) ** 2 / lvd.exp()).sum(-1) - (-(mud - ed[perm]) ** 2 / lvd.exp()).sum(-1)).mean() / 2.0}
    for name, ref in refs.items():
        ours = F.club_aprx_loss(mu, lv, e) if name == "aprx" else F.club_upper_bound(mu, lv, e, perm)
        assert ours.shape == () and abs(float(ours) - float(ref)) <= 1e-5 * max(abs(float(ref)), 1e-3), name
        g_ours = torch.autograd.grad(ours * 1.7, [mu, lv, e])
        g_ref = torch.autograd.grad(ref * 1.7, [mud, lvd, ed])
        floor = max(t.norm().item() for t in g_ref)      # N = 1, upper: e[perm] is e and d/de is exactly 0 in the reference
        for a, b in zip(g_ours, g_ref):
            assert (a.double() - b).norm().item() <= 1e-5 * max(b.norm().item(), floor) + 1e-12, name
        strided = F.club_aprx_loss(mu, lv, z[:, :L]) if name == "aprx" else F.club_upper_bound(mu, lv, z[:, :L], perm)
        assert torch.equal(strided.detach(), (F.club_aprx_loss(mu, lv, z[:, :L].contiguous()) if name == "aprx" else
                                              F.club_upper_bound(mu, lv, z[:, :L].contiguous(), perm)).detach())


def test_inference_entry_launches_and_precision():
    """get_pair_preds (mmd:265-282) through carel_pair_preds: at most 4 launches of the library (pack X, pack W4,
    projection GEMM, fused reparameterisation + pair logit + sigmoid), no torch kernels besides the two noise draws;
    1e-5 against the oracle in the fp32 tier."""
    opt, model, b = _setup("mmd", 700, 128, 1000, 768, 0.5, seed=6)
    model.eval()
    with torch.no_grad():
        torch.manual_seed(21)
        l0 = cv.launch_count()
        t = model.get_pair_preds_tensor(b["X"], None, None)
        assert cv.launch_count() - l0 <= 4
        assert t.shape == (700, 1) and not t.requires_grad
        p = {k: v.detach() for k, v in model.named_parameters()}
        torch.manual_seed(21)
        ref = ho.HeadOracle(opt, "mmd").pair_preds({k: v.double() for k, v in p.items()}, b["X"].double())
        assert (t.double() - ref).abs().max().item() < 1e-3
        cv.set_precision("fp32")
        try:
            torch.manual_seed(21)
            t32 = model.get_pair_preds_tensor(b["X"], None, None)
        finally:
            cv.set_precision("bf16")
        assert rel(t32, ref) < 1e-5


@pytest.mark.parametrize("N,L,drop", [(1, 24, False), (300, 128, True), (8192, 128, True)])
def test_gan_discriminator_kernels_match_the_reference_formulas(N, L, drop):
    """carel_disc_fwd/bwd against the reference's PyTorch expressions in fp64 (gan:430-477, forward :222-233): the three
    scalars and the parameter gradients for each of the trainer's three backward passes (gan:790-798), 1e-5."""
    import torch.nn as nn
    import torch.nn.functional as TF
    from carel_vae_b200 import functional as F
    torch.manual_seed(N + L)
    z = torch.randn(N, 2 * L, device=DEV, requires_grad=True)
    ec, ce = nn.Linear(L, 1).to(DEV), nn.Linear(L, 1).to(DEV)
    y_e = torch.ones(N, 1, device=DEV)
    y_c = (torch.rand(N, 1, device=DEV) < 0.5).float()
    m1 = TF.dropout(torch.ones(N, L, device=DEV), 0.5, True) if drop else None
    m2 = TF.dropout(torch.ones(N, L, device=DEV), 0.5, True) if drop else None
    outs = F.gan_discriminators(z, m1, m2, ec, ce, y_e, y_c, 0.1, 1, 1e-8)
    params = [ec.weight, ec.bias, ce.weight, ce.bias]
    p64 = [p.detach().double().requires_grad_(True) for p in params]
    zd = z.detach().double()
    xc = zd[:, L:] * (m1.double() if drop else 1.0)
    xe = zd[:, :L] * (m2.double() if drop else 1.0)
    p_ec = torch.sigmoid(xc @ p64[0].T + p64[1])
    p_ce = torch.sigmoid(xe @ p64[2].T + p64[3])
    refs = (TF.binary_cross_entropy(p_ec, y_e.double() * 0.9 + 0.1), TF.binary_cross_entropy(p_ce, y_c.double() * 0.9 + 0.1),
            (p_ec * torch.log(p_ec + 1e-8)).sum(1).mean() + (p_ce * torch.log(p_ce + 1e-8)).sum(1).mean())
    for o, r in zip(outs, refs):
        assert abs(float(o.detach()) - float(r)) <= 1e-5 * abs(float(r))
    for k in range(3):
        g = torch.autograd.grad(outs[k], params, retain_graph=True, allow_unused=True)
        rg = torch.autograd.grad(refs[k], p64, retain_graph=True, allow_unused=True)
        for a, b in zip(g, rg):
            if b is None:
                assert a is None or float(a.abs().max()) == 0.0
            else:
                assert (a.double() - b).norm().item() <= 1e-5 * b.norm().item() + 1e-12
    assert z.grad is None          # the latents are detached in this branch, as in the reference


@pytest.mark.parametrize("N,L", [(1, 24), (77, 64), (300, 128), (8192, 128)])
def test_club_mlp_kernels_match_torch(N, L):
    """carel_club_mlp_fwd/bwd against the nn.Sequential modules they replace (vi_final:155-162, :342-347) in fp64:
    outputs and the gradients of the input and of all eight parameters, 1e-5."""
    import copy
    import torch.nn as nn
    from carel_vae_b200 import functional as F
    torch.manual_seed(N + L)
    ec_mu = nn.Sequential(nn.Linear(L, L), nn.ReLU(), nn.Linear(L, L)).to(DEV)
    ec_lv = nn.Sequential(nn.Linear(L, L), nn.ReLU(), nn.Linear(L, L), nn.Tanh()).to(DEV)
    z = torch.randn(N, 2 * L, device=DEV)
    # ReLU's derivative is discontinuous: a hidden unit whose pre-activation lies within fp32 rounding of zero gets the
    # other side's mask in fp64 and moves the gradients by O(1e-3) (it is a different, equally valid subgradient).
    # Rows with such a unit (|pre| < 1e-4, a few per thousand) are left out of the comparison batch.
    with torch.no_grad():
        c64 = z[:, L:].double()
        pre = torch.cat((c64 @ ec_mu[0].weight.double().T + ec_mu[0].bias.double(),
                         c64 @ ec_lv[0].weight.double().T + ec_lv[0].bias.double()), 1)
        keep = pre.abs().min(dim=1).values > 1e-4
        if N > 1:
            z = z[keep].contiguous()
    z.requires_grad_(True)
    mu, lv = F.club_mlp(z[:, L:], ec_mu, ec_lv)            # strided column view, like z_c
    gm, gl = torch.randn_like(mu), torch.randn_like(lv)
    params = list(ec_mu.parameters()) + list(ec_lv.parameters())
    grads = torch.autograd.grad((mu * gm).sum() + (lv * gl).sum(), [z] + params)
    mu64, lv64 = copy.deepcopy(ec_mu).double(), copy.deepcopy(ec_lv).double()
    zd = z.detach().double().requires_grad_(True)
    rmu, rlv = mu64(zd[:, L:]), lv64(zd[:, L:])
    rgrads = torch.autograd.grad((rmu * gm.double()).sum() + (rlv * gl.double()).sum(),
                                 [zd] + list(mu64.parameters()) + list(lv64.parameters()))
    assert rel(mu, rmu) < 1e-5 and rel(lv, rlv) < 1e-5
    for g, r in zip(grads, rgrads):
        assert rel(g, r) < 1e-5


@pytest.mark.parametrize("variant", ["mmd", "hsic"])
def test_reserved_sms_do_not_change_the_result(variant):
    """carel_set_reserved_sms (the data-parallel overlap leaves SMs to NCCL): every kernel sizes its grid for fewer SMs
    and still computes the same step — values to 1e-6, gradients to 1e-5 of the full-grid run (different split-K /
    stream-K cuts change the fp32 summation order, nothing else) — at a size where every persistent kernel has work
    for all CTAs."""
    from carel_vae_b200 import _lib
    opt, model, b = _setup(variant, 2048, 128, 9000, 768, 0.0, seed=5)
    model.train()
    emo = b["emo"] if variant == "mmd" else b["emo"].float()
    res = []
    try:
        for reserved in (0, 24, 100):
            _lib.set_reserved_sms(reserved)
            assert _lib.reserved_sms() == reserved
            X = b["X"].clone().requires_grad_(True)
            torch.manual_seed(3)
            loss = model(X, None, None, emo, b["cau"], b["pair"], b["bow"], 3)
            params = list(model.parameters())
            grads = torch.autograd.grad(loss, [X] + params, allow_unused=True)
            res.append((loss.detach(), grads))
    finally:
        _lib.set_reserved_sms(0)
    (l0, g0) = res[0]
    for (l, g) in res[1:]:
        assert rel(l, l0) < 1e-6
        for a, r in zip(g, g0):
            if r is None:
                assert a is None
                continue
            assert rel(a, r) < 1e-5


def test_host_bow_stager_equals_the_gpu_scan():
    """data.HostBowStager: host rows scanned on the host, the rest copied dense and scanned on the GPU — the staged ELL
    triple holds exactly the non-zeros of the batch for any split (the host list is in ascending column order, the GPU
    scan interleaves the four columns of a lane inside a 128-column group: same set, same values)."""
    from carel_vae_b200.data import HostBowStager
    from carel_vae_b200 import functional as F
    N, V, cap = 300, 5000, 32
    bow = ho.synthetic_batch(N, V, D=8, nnz=24, seed=2, device="cpu")["bow"].pin_memory()
    ref = F.bow_pack_ell(bow.to(DEV), cap)
    for frac in (0.0, 0.37, 1.0):
        st = HostBowStager(N, V, DEV, cap=cap, host_fraction=frac, threads=3)
        for _ in range(2):                                  # second call reuses the pinned buffers
            cols, vals, nnz = st.stage(bow)
        torch.cuda.synchronize()
        assert torch.equal(nnz, ref[2])
        k = torch.arange(cap, device=DEV)[None, :] < nnz[:, None]
        back = torch.zeros(N, V, device=DEV)
        rows = torch.arange(N, device=DEV)[:, None].expand(N, cap)
        back[rows[k], cols[k].long()] = vals[k]
        assert torch.equal(back, bow.to(DEV))
    st.calibrate(bow)
    assert 0 <= st.host_rows <= N and st.bytes_per_stage() > 0


@pytest.mark.parametrize("with_out", [False, True])
def test_host_bow_stager_pipelined_keeps_batches_apart(with_out):
    """HostBowStager(pipelined=True): staging runs on the stager's own stream into two buffer sets while the current
    stream consumes the previous batch.  Six DIFFERENT batches in a row, each consumed by a (deliberately slow) kernel
    queue on the current stream: every staged triple and every side copy must be its own batch's, for a caller-owned
    `out` (graph inputs) and for the stager's own sets, with the split adapting on the way."""
    from carel_vae_b200.data import HostBowStager
    from carel_vae_b200 import functional as F
    N, V, cap = 264, 4000, 32
    batches = [ho.synthetic_batch(N, V, D=8, nnz=20, seed=10 + s, device="cpu")["bow"].pin_memory() for s in range(6)]
    extras = [torch.full((N, 8), float(s)).pin_memory() for s in range(6)]
    out = tuple(torch.zeros_like(t) for t in F.bow_pack_ell(batches[0].to(DEV), cap)) if with_out else None
    x_dev = torch.zeros(N, 8, device=DEV)
    st = HostBowStager(N, V, DEV, cap=cap, host_fraction=0.4, threads=2, out=out, pipelined=True, adapt=True)
    busy = torch.randn(2048, 2048, device=DEV)
    got = []
    for s, (bow, ex) in enumerate(zip(batches, extras)):
        cols, vals, nnz = st.stage(bow, copies=[(x_dev, ex)])
        for _ in range(4):                                   # keeps the current stream busy past the next stage() call
            busy = (busy @ busy).clamp_(-1, 1)
        k = torch.arange(cap, device=DEV)[None, :] < nnz[:, None]
        back = torch.zeros(N, V, device=DEV)
        rows = torch.arange(N, device=DEV)[:, None].expand(N, cap)
        back[rows[k], cols[k].long()] = vals[k]
        got.append((back, nnz.clone(), x_dev.clone()))
    torch.cuda.synchronize()
    for s, (back, nnz, x) in enumerate(got):
        assert torch.equal(back, batches[s].to(DEV)), f"batch {s}"
        assert torch.equal(nnz.cpu(), (batches[s] != 0).sum(1).int()), f"batch {s}"
        assert torch.equal(x.cpu(), extras[s]), f"batch {s}"
    assert 8 <= st.host_rows <= N - 8 and st.host_rows % 8 == 0
