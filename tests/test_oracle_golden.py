"""CPU: the oracle restatement (oracle/head_oracle.py) against the golden vectors produced by the
UNMODIFIED reference code (oracle/make_golden.py).  This is what pins the oracle."""
import torch
import pytest

from oracle import head_oracle as ho


def _close(a, b, rtol=2e-5, atol=1e-7):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    err = (a - b).abs().max().item() if a.numel() else 0.0
    scale = b.abs().max().item() if b.numel() else 0.0
    assert err <= atol + rtol * scale, f"max err {err:.3e} vs scale {scale:.3e}"


def test_mmd_primitives(golden_dir):
    for c in torch.load(golden_dir / "mmd.pt"):
        x, y = c["x"].clone().requires_grad_(True), c["y"].clone().requires_grad_(True)
        val, K = ho.mmd_statistic(x, y, c["alphas"], ret_matrix=True)
        gx, gy = torch.autograd.grad(val, (x, y))
        _close(val, c["mmd"]); _close(K, c["kernels"]); _close(gx, c["grad_x"]); _close(gy, c["grad_y"])
        _close(ho.pdist(c["x"], c["y"]), c["pdist"])
        a00, a11, a01 = ho.mmd_coefficients(x.shape[0], y.shape[0])
        assert (a00, a11, a01) == (c["a00"], c["a11"], c["a01"])
        # closed-form blocked fp64 evaluation (the large-N oracle) against the literal reference
        v64, g1, g2 = ho.mmd_blocked(c["x"].double(), c["y"].double(), c["alphas"], block=16)
        _close(v64, c["mmd"], rtol=5e-5); _close(g1, c["grad_x"], rtol=2e-4); _close(g2, c["grad_y"], rtol=2e-4)


def test_mmd_n1_raises():
    with pytest.raises(ZeroDivisionError):
        ho.mmd_coefficients(1, 1)


def test_hsic_primitives(golden_dir):
    for c in torch.load(golden_dir / "hsic.pt"):
        x, y = c["x"].clone().requires_grad_(True), c["y"].clone().requires_grad_(True)
        val = ho.hsic_trace(x, y, c["s_x"], c["s_y"])
        gx, gy = torch.autograd.grad(val, (x, y))
        _close(val, c["hsic"], rtol=1e-4); _close(gx, c["grad_x"], rtol=1e-3, atol=1e-9); _close(gy, c["grad_y"], rtol=1e-3, atol=1e-9)
        _close(ho.gaussian_kernel_matrix(c["x"], c["s_x"]), c["K"]); _close(ho.pairwise_distances(c["x"]), c["d2"])
        # O(m^2) identity in fp64 equals the O(m^3) trace form
        _close(ho.hsic_fast(c["x"].double(), c["y"].double(), c["s_x"], c["s_y"]), c["hsic"], rtol=2e-4)


def test_misc(golden_dir):
    c = torch.load(golden_dir / "misc.pt")
    _close(ho.sample_prior(c["mu"], c["lv"], c["eps"]), c["z"])
    _close(ho.kl_loss(c["mu"], c["lv"]), c["kl"])
    for it, w in zip((0, 3, 1000, 19999), c["weights"]):
        assert ho.annealed_weight(it, 0.03) == w
    _close(ho.reconstruct_loss(torch.softmax(c["logits"], 1), c["bow"]), c["recon"])


@pytest.mark.parametrize("variant", ["mmd", "hsic", "vi", "gan"])
def test_head_forward_backward(golden_dir, variant):
    cases = [c for c in torch.load(golden_dir / "head.pt") if c["variant"] == variant]
    assert len(cases) == 3
    for c in cases:
        opt = ho.default_opt(variant, ec_dim=c["L"], pair_bow_dim=c["V"], bert_dim=c["D"], dropout=c["dropout"])
        # default nn.Linear init in the reference's construction order reproduces its parameters
        st = ho.init_state(opt, variant, seed=1234 + c["seed"])
        ref_state = {k: v for k, v in c["state"].items() if not k.startswith("encoder")}
        assert set(st) == set(ref_state)
        for k in st:
            assert torch.equal(st[k], ref_state[k]), k
        p = {k: v.clone().requires_grad_(True) for k, v in st.items()}
        b = ho.synthetic_batch(c["N"], c["V"], D=c["D"], seed=c["seed"], emotion_float=variant in ("hsic", "gan"))
        X = b["X"].clone().requires_grad_(True)
        emo = b["emo"] if variant in ("mmd", "vi") else b["emo"].float()
        torch.manual_seed(77 + c["seed"])          # same RNG stream: eps_e, eps_c, dropout masks in order
        out = ho.HeadOracle(opt, variant).forward(p, X, emo, b["cau"], b["pair"], b["bow"], c["iteration"])
        outs = out if isinstance(out, tuple) else (out,)
        for o, r in zip(outs, c["outputs"]):
            _close(o, r, rtol=3e-5)
        total = outs[2] + outs[3] if variant == "vi" else (outs[0] + outs[1] + outs[2] if variant == "gan" else outs[0])
        names = list(p.keys())
        grads = torch.autograd.grad(total, [X] + [p[k] for k in names], allow_unused=True)
        _close(grads[0], c["grad_X"], rtol=2e-4, atol=1e-8)
        for k, g in zip(names, grads[1:]):
            r = c["grads"][k]
            if r is None:
                assert g is None or float(g.abs().max()) == 0.0
            else:
                _close(g, r, rtol=2e-4, atol=1e-8)
        torch.manual_seed(5 + c["seed"])
        preds = ho.HeadOracle(opt, variant).pair_preds(st, b["X"]).round()
        assert torch.equal(preds, c["pair_preds"].reshape(preds.shape))


@pytest.mark.parametrize("variant", ["mmd", "hsic", "vi", "gan"])
@pytest.mark.parametrize("N,L,V,D,dropout,seed", [(33, 32, 301, 64, 0.5, 11), (7, 8, 40, 16, 0.0, 12)])
def test_oracle_matches_the_live_reference_on_fresh_cases(variant, N, L, V, D, dropout, seed):
    """Build container only (skipped where /root/reference is absent): the reference's own DrlClassifier, executed now
    (oracle/reference_extract.py), against the restatement on shapes and seeds the golden files do not contain —
    ragged N, dropout on, all outputs, the gradient w.r.t. the embeddings and every parameter, and for the VI variant
    the CLUB upper bound with the CPU-generator permutation of vi_final:433."""
    from oracle import reference_extract as rx
    if not rx.available():
        pytest.skip("the reference is only present in the build container")
    opt = ho.default_opt(variant, ec_dim=L, pair_bow_dim=V, bert_dim=D, dropout=dropout)
    ns = rx.load(variant, opt)
    torch.manual_seed(1234 + seed)
    model = ns["DrlClassifier"](opt)
    model.train()
    b = ho.synthetic_batch(N, V, D=D, seed=seed, emotion_float=variant in ("hsic", "gan"))
    emo = b["emo"] if variant in ("mmd", "vi") else b["emo"].float()

    def total_of(outs):
        return outs[2] + outs[3] if variant == "vi" else (outs[0] + outs[1] + outs[2] if variant == "gan" else outs[0])

    X = b["X"].clone().requires_grad_(True)
    torch.manual_seed(77 + seed)
    with rx.cpu_cuda_shim():
        out = model(X, None, None, emo, b["cau"], b["pair"], b["bow"], 5)
    r_outs = out if isinstance(out, tuple) else (out,)
    names = [k for k, _ in model.named_parameters() if not k.startswith("encoder")]
    r_params = dict(model.named_parameters())
    r_grads = torch.autograd.grad(total_of(r_outs), [X] + [r_params[k] for k in names], allow_unused=True)

    st = ho.init_state(opt, variant, seed=1234 + seed)
    p = {k: v.clone().requires_grad_(True) for k, v in st.items()}
    assert set(p) == set(names)
    X2 = b["X"].clone().requires_grad_(True)
    torch.manual_seed(77 + seed)
    oracle = ho.HeadOracle(opt, variant)
    out2 = oracle.forward(p, X2, emo, b["cau"], b["pair"], b["bow"], 5)
    o_outs = out2 if isinstance(out2, tuple) else (out2,)
    for o, r in zip(o_outs, r_outs):
        _close(o, torch.as_tensor(r), rtol=3e-5)
    o_grads = torch.autograd.grad(total_of(o_outs), [X2] + [p[k] for k in names], allow_unused=True)
    for g, r in zip(o_grads, r_grads):
        if r is None:
            assert g is None or float(g.abs().max()) == 0.0
        else:
            _close(g, r, rtol=2e-4, atol=1e-8)
    if variant == "vi":
        z_e, z_c = torch.randn(N, L), torch.randn(N, L)
        torch.manual_seed(3)
        r_up = model.get_ec_upper_loss(z_e, z_c)
        torch.manual_seed(3)
        _close(oracle.ec_upper_loss(st, z_e, z_c), r_up, rtol=3e-5)
