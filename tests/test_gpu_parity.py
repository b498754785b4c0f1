"""GPU (-m gpu): the CUDA path (through the C ABI) against the oracle on identical seeded inputs.

Tolerances: the kernels compute in bf16 operands / fp32 accumulation ("bf16 mode" of BASELINE.json's
north_star, bound 1e-3 relative).  Measured and asserted much tighter: loss VALUES ~1e-7..1e-5, MMD / HSIC /
latent GRADIENTS ~1e-6..1e-5 (the in-kernel weight tile is a bf16 hi+lo pair), decoder gradients
6e-5..3e-4 (single bf16 dlogits tile), relative in the Frobenius norm.  Inputs are rounded to bf16 once and
fed to BOTH sides, as the north_star's "identical inputs" protocol prescribes (SURVEY.md §7 hard part 2).
"""
import pytest
import torch

from oracle import head_oracle as ho

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import carel_vae_b200 as cv
    from carel_vae_b200 import functional as F
DEV = "cuda"


def bf(x):
    return x.bfloat16().float()


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def relv(a, b):
    return abs(float(a) - float(b)) / max(abs(float(b)), 1e-300)


def test_extension_loaded_and_launches():
    before = cv.launch_count()
    x = torch.randn(64, 32, device=DEV)
    cv.MMDStatistic(64, 64)(x, x + 1, [0.1])
    torch.cuda.synchronize()
    assert cv.launch_count() > before        # our kernels ran, not a library fallback


@pytest.mark.parametrize("a_mn,b_mn,ks", [(0, 0, 1), (0, 1, 1), (1, 0, 2), (1, 1, 3)])
def test_gemm_core_layouts(a_mn, b_mn, ks):
    torch.manual_seed(0)
    M, N, K = 300, 520, 208
    a = torch.randn(M, K, device=DEV).bfloat16()
    b = torch.randn(N, K, device=DEV).bfloat16()
    ref = a.float() @ b.float().T
    aa = a.T.contiguous() if a_mn else a
    bb = b.T.contiguous() if b_mn else b
    if aa.stride(0) % 8 or bb.stride(0) % 8:
        pytest.skip("pitch not a multiple of 8")
    assert rel(F.gemm_bf16(aa, bb, bool(a_mn), bool(b_mn), ksplit=ks), ref) < 1e-5


@pytest.mark.parametrize("a_mn,b_mn,ks,M,N,K", [(0, 0, 1, 304, 520, 208), (0, 1, 2, 304, 520, 208), (1, 0, 1, 304, 520, 208),
                                                (1, 1, 3, 304, 520, 208), (1, 1, 5, 1000, 256, 2048), (0, 1, 4, 768, 256, 1000)])
def test_gemm_core_cta_pairs(a_mn, b_mn, ks, M, N, K):
    """The same products on CTA pairs (cta_group::2 UMMA of M = 256, split-K, K-major and MN-major operands)."""
    torch.manual_seed(1)
    a = torch.randn(M, K, device=DEV).bfloat16()
    b = torch.randn(N, K, device=DEV).bfloat16()
    ref = a.float() @ b.float().T
    aa = a.T.contiguous() if a_mn else a
    bb = b.T.contiguous() if b_mn else b
    assert rel(F.gemm_bf16(aa, bb, bool(a_mn), bool(b_mn), ksplit=ks, pairs=True), ref) < 1e-5


# ------------------------------------------------------------------------------------------
# MMD
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n1,n2,d,alphas", [(2, 2, 8, [0.1]), (64, 64, 64, [0.1]), (50, 70, 24, [0.1]),
                                            (300, 300, 128, [0.1, 0.01, 1.0]), (1000, 1500, 96, [0.05]),
                                            (129, 257, 128, [0.1]), (512, 512, 240, [0.02])])
def test_mmd_vs_oracle(n1, n2, d, alphas):
    torch.manual_seed(n1 + d)
    x = bf(torch.randn(n1, d, device=DEV) * 1.1).requires_grad_(True)
    y = bf(torch.randn(n2, d, device=DEV) * 1.1 + 0.3).requires_grad_(True)
    xd, yd = x.detach().double().requires_grad_(True), y.detach().double().requires_grad_(True)
    ref = ho.mmd_statistic(xd, yd, alphas)
    rgx, rgy = torch.autograd.grad(ref, (xd, yd))
    for full in (False, True):
        out = F.mmd(x, y, alphas, full_gram=full)
        gx, gy = torch.autograd.grad(out, (x, y))
        assert relv(out, ref) < 1e-4
        assert rel(gx, rgx) < 1e-4 and rel(gy, rgy) < 1e-4
    stat = cv.MMDStatistic(n1, n2)
    assert (stat.a00, stat.a11, stat.a01) == ho.mmd_coefficients(n1, n2)


def test_mmd_golden_reference_vectors(golden_dir):
    """Against the outputs of the unmodified reference (fp32 inputs, not bf16-representable: the
    bf16 rounding of the inputs is part of the error here — SURVEY.md measured 1.5e-4 / 4.5e-3)."""
    for c in torch.load(golden_dir / "mmd.pt"):
        x, y = c["x"].to(DEV).requires_grad_(True), c["y"].to(DEV).requires_grad_(True)
        out, K = cv.MMDStatistic(x.shape[0], y.shape[0])(x, y, c["alphas"], ret_matrix=True)
        gx, gy = torch.autograd.grad(out, (x, y))
        assert relv(out, c["mmd"]) < 5e-3
        assert rel(gx.cpu(), c["grad_x"]) < 1.5e-2 and rel(gy.cpu(), c["grad_y"]) < 1.5e-2
        assert rel(K.cpu(), c["kernels"]) < 1e-2
        assert rel(cv.pdist(x.detach(), y.detach()).cpu(), c["pdist"]) < 5e-3


def test_mmd_full_size_properties():
    """N = 8192, L = 128 (BASELINE config 2): blocked fp64 oracle, symmetry and triangular == full Gram."""
    torch.manual_seed(0)
    N, L = 8192, 128
    x = bf(torch.randn(N, L, device=DEV) * 1.15).requires_grad_(True)
    y = bf(torch.randn(N, L, device=DEV) * 1.15 + 0.2).requires_grad_(True)
    out = F.mmd(x, y, [0.1])
    gx, gy = torch.autograd.grad(out, (x, y))
    ref, rgx, rgy = ho.mmd_blocked(x.detach().double(), y.detach().double(), [0.1], block=2048)
    assert relv(out, ref) < 1e-4
    # north_star tolerance of the bf16 tier: 1e-3.  At M >= 4096 rows the backward carries its weight tile as ONE fp16
    # value per element (2^-12); in this regime (d2 ~ 340, K ~ e^-34) a gradient row is dominated by a handful of
    # nearest neighbours, so the rounding does not average out: measured 1.4e-4 (1e-5 with the hi + lo pair used
    # below 4096 rows, where accuracy is free).
    assert rel(gx, rgx) < 3e-4 and rel(gy, rgy) < 3e-4
    assert relv(F.mmd(x, y, [0.1], full_gram=True), out) < 1e-5          # triangular schedule == full Gram
    assert relv(F.mmd(y, x, [0.1]), out) < 1e-5                           # MMD(x, y) == MMD(y, x)
    same = F.mmd(x.detach(), x.detach(), [0.1])
    # identical samples: MMD^2 = (sum offdiag - sum all)/... = -(mean diagonal kernel - mean offdiag)/(n) < 0 and tiny
    assert float(same) <= 0.0 and abs(float(same)) < 1e-3


def test_mmd_gram_sweep_d768():
    """BASELINE config 5 shape (d = 768, unit-norm rows): value against the blocked fp64 oracle."""
    torch.manual_seed(1)
    N, D = 4096, 768
    x = torch.tanh(torch.randn(N, D, device=DEV))
    y = torch.tanh(torch.randn(N, D, device=DEV) + 0.05)
    x, y = bf(x / x.norm(dim=1, keepdim=True)), bf(y / y.norm(dim=1, keepdim=True))
    alphas = [0.1, 1.0, 10.0]
    out = F.mmd(x, y, alphas)
    ref, _, _ = ho.mmd_blocked(x.double(), y.double(), alphas, block=2048, need_grad=False)
    assert relv(out, ref) < 1e-3


# ------------------------------------------------------------------------------------------
# HSIC
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,d,sx,sy", [(2, 8, 4.0, 4.0), (64, 24, 30.0, 30.0), (200, 64, 100.0, 150.0),
                                       (1025, 128, 300.0, 300.0), (500, 128, 1.0, 1.0)])
def test_hsic_vs_oracle(m, d, sx, sy):
    torch.manual_seed(m)
    x = bf(torch.randn(m, d, device=DEV) * 1.1).requires_grad_(True)
    y = bf(0.5 * x.detach() + torch.randn(m, d, device=DEV)).requires_grad_(True)
    out = cv.HSIC(x, y, sx, sy)
    gx, gy = torch.autograd.grad(out, (x, y))
    xd, yd = x.detach().double().requires_grad_(True), y.detach().double().requires_grad_(True)
    ref = ho.hsic_trace(xd, yd, sx, sy) if m <= 1100 else ho.hsic_fast(xd, yd, sx, sy)
    rgx, rgy = torch.autograd.grad(ref, (xd, yd))
    assert relv(out, ref) < 1e-3
    if sx == 1.0:      # reference default bandwidth: degenerate at this width (K ~ I), gradient ~ 0: absolute check
        assert gx.abs().max().item() < 1e-8 and rgx.abs().max().item() < 1e-8
        assert abs(float(out) - 1.0 / (m - 1)) < 1e-3 / (m - 1)
    else:
        assert rel(gx, rgx) < 2e-4 and rel(gy, rgy) < 2e-4
    assert relv(cv.HSIC(y, x, sy, sx), out) < 1e-4       # HSIC(x, y) == HSIC(y, x)


def test_hsic_golden_reference_vectors(golden_dir):
    for c in torch.load(golden_dir / "hsic.pt"):
        x, y = c["x"].to(DEV).requires_grad_(True), c["y"].to(DEV).requires_grad_(True)
        out = cv.HSIC(x, y, c["s_x"], c["s_y"])
        assert relv(out, c["hsic"]) < 5e-3
        assert rel(cv.GaussianKernelMatrix(x.detach(), c["s_x"]).cpu(), c["K"]) < 1e-2
        d2 = cv.pairwise_distances(x.detach()).cpu()
        assert (d2 - c["d2"]).abs().max().item() < 2e-2 * c["d2"].abs().max().item()


def test_hsic_full_size():
    torch.manual_seed(2)
    m, L = 8192, 128
    x = bf(torch.randn(m, L, device=DEV) * 1.15).requires_grad_(True)
    y = bf(0.3 * x.detach() + torch.randn(m, L, device=DEV)).requires_grad_(True)
    s = 2.0 * L * 1.3
    out = cv.HSIC(x, y, s, s)
    gx, gy = torch.autograd.grad(out, (x, y))
    xd, yd = x.detach().double().requires_grad_(True), y.detach().double().requires_grad_(True)
    ref = ho.hsic_fast(xd, yd, s, s)
    rgx, rgy = torch.autograd.grad(ref, (xd, yd))
    assert relv(out, ref) < 1e-4
    assert rel(gx, rgx) < 2e-4 and rel(gy, rgy) < 2e-4


# ------------------------------------------------------------------------------------------
# latent layer and BoW reconstruction
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,D,L", [(1, 768, 24), (64, 768, 24), (300, 768, 64), (777, 96, 40), (8192, 768, 128)])
def test_latent_vs_oracle(N, D, L):
    torch.manual_seed(N)
    X = bf(torch.tanh(torch.randn(N, D, device=DEV))).requires_grad_(True)
    Ws = [bf(torch.randn(L, D, device=DEV) / D ** 0.5).requires_grad_(True) for _ in range(4)]
    bs = [bf(torch.randn(L, device=DEV) * 0.1).requires_grad_(True) for _ in range(4)]
    ee, ec = torch.randn(L, device=DEV), torch.randn(L, device=DEV)
    args = [X, ee, ec]
    for w_, b_ in zip(Ws, bs):
        args += [w_, b_]
    z, kl, pre = F.latent(*args)
    Xd = X.detach().double().requires_grad_(True)
    Wd = [w_.detach().double().requires_grad_(True) for w_ in Ws]
    bd = [b_.detach().double().requires_grad_(True) for b_ in bs]
    pr = [torch.nn.functional.linear(Xd, w_, b_) for w_, b_ in zip(Wd, bd)]
    zr = torch.cat([ho.sample_prior(pr[0], pr[1], ee.double()), ho.sample_prior(pr[2], pr[3], ec.double())], 1)
    klr = torch.stack([ho.kl_loss(pr[0], pr[1]), ho.kl_loss(pr[2], pr[3])])
    assert rel(z, zr) < 1e-5 and rel(kl, klr) < 1e-5 and rel(pre, torch.cat(pr, 1)) < 1e-5
    gz, gk = torch.randn_like(z), torch.tensor([0.7, -1.3], device=DEV)
    ((z * gz).sum() + (kl * gk).sum()).backward()
    ((zr * gz.double()).sum() + (klr * gk.double()).sum()).backward()
    assert rel(X.grad, Xd.grad) < 1e-4
    for a, b in zip(Ws + bs, Wd + bd):
        assert rel(a.grad, b.grad) < 1e-4


@pytest.mark.parametrize("N,K,V,nnz,cap", [(1, 48, 100, 3, 64), (64, 48, 300, 5, 64), (100, 128, 1000, 24, 64),
                                           (256, 256, 5000, 24, 64), (130, 128, 777, 40, 8), (512, 256, 30000, 24, 64)])
def test_bow_vs_oracle(N, K, V, nnz, cap):
    """cap=8 with 40 non-zeros per row exercises the dense re-scan of overflowing rows."""
    torch.manual_seed(V)
    Z = bf(torch.randn(N, K, device=DEV) * 1.1).requires_grad_(True)
    W = bf((torch.rand(V, K, device=DEV) * 2 - 1) / K ** 0.5).requires_grad_(True)
    b = bf((torch.rand(V, device=DEV) * 2 - 1) / K ** 0.5).requires_grad_(True)
    bow = ho.synthetic_batch(N, V, D=8, nnz=nnz, seed=N, device=DEV)["bow"]
    loss = F.bow_reconstruction_loss(Z, W, b, bow, 0.1, ell_cap=cap)
    loss.backward()
    Zd, Wd, bd = (t.detach().double().requires_grad_(True) for t in (Z, W, b))
    ref = ho.bow_reconstruction(Zd, Wd, bd, bow.double(), 0.1)
    ref.backward()
    assert relv(loss, ref) < 1e-5
    assert rel(Z.grad, Zd.grad) < 1e-3 and rel(W.grad, Wd.grad) < 1e-3 and rel(b.grad, bd.grad) < 1e-3


def test_bow_sparse_equals_dense_target():
    torch.manual_seed(5)
    N, K, V = 200, 128, 4000
    Z, W, b = torch.randn(N, K, device=DEV), torch.randn(V, K, device=DEV) / 12, torch.randn(V, device=DEV) / 12
    sb = ho.synthetic_batch(N, V, D=8, nnz=24, seed=3, device=DEV)
    dense = F.bow_reconstruction_loss(Z, W, b, sb["bow"])
    rows = []
    cap = 24
    cols = torch.zeros(N, cap, dtype=torch.int32, device=DEV)
    vals = torch.zeros(N, cap, device=DEV)
    nnz = torch.zeros(N, dtype=torch.int32, device=DEV)
    for i in range(N):
        nz = sb["bow"][i].nonzero().flatten()
        cols[i, :len(nz)] = nz.int()
        vals[i, :len(nz)] = sb["bow"][i, nz]
        nnz[i] = len(nz)
    sparse = F.bow_reconstruction_loss(Z, W, b, None, 0.1, ell=(cols, vals, nnz))
    assert relv(sparse, dense) < 1e-6


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_bow_ell_only_overflow_is_reported_not_read(precision):
    """An ELL triple WITHOUT the dense target whose nnz exceeds the list capacity (bow_pack_ell documents that it may)
    cannot be evaluated: the loss comes back NaN and the gradients finite or NaN — never a read through the missing
    dense pointer (round-1 advisor finding: illegal address in bow_sparse_kernel) — and the context stays usable."""
    torch.manual_seed(9)
    N, K, V, cap = 96, 64, 900, 8
    Z = torch.randn(N, K, device=DEV, requires_grad=True)
    W = (torch.randn(V, K, device=DEV) / 8).requires_grad_(True)
    b = torch.zeros(V, device=DEV, requires_grad=True)
    bow = ho.synthetic_batch(N, V, D=8, nnz=40, seed=1, device=DEV)["bow"]
    cols, vals, nnz = F.bow_pack_ell(bow, cap)
    assert int(nnz.max()) > cap
    loss = F.bow_reconstruction_loss(Z, W, b, None, 0.1, ell=(cols, vals, nnz), precision=precision)
    loss.backward()
    torch.cuda.synchronize()                       # a fault would surface here as a sticky CUDA error
    assert torch.isnan(loss)
    # the same call with a list that fits is fine afterwards
    cols, vals, nnz = F.bow_pack_ell(bow, 64)
    ok = F.bow_reconstruction_loss(Z, W, b, None, 0.1, ell=(cols, vals, nnz), precision=precision)
    ref = ho.bow_reconstruction(Z.detach().double(), W.detach().double(), b.detach().double(), bow.double(), 0.1)
    assert relv(ok, ref) < 1e-3


@pytest.mark.parametrize("N,V,nnz,cap", [(1, 7, 3, 4), (37, 1001, 5, 8), (300, 30000, 24, 64), (64, 2050, 40, 8)])
def test_bow_pack_ell_matches_nonzero(N, V, nnz, cap):
    """carel_bow_pack_ell (bit-exact index / byte work): every non-zero of a row exactly once (order within a 128-column
    group is by element-in-lane, not ascending), values copied verbatim, nnz counts the whole row even when it
    overflows the list (cap = 8 with 40 draws per row: the list then holds cap distinct non-zeros); ragged V too.
    The list handed back together with the dense rows gives the same loss as the dense target alone."""
    bow = ho.synthetic_batch(N, V, D=8, nnz=nnz, seed=N + V, device=DEV)["bow"]
    cols, vals, cnt = F.bow_pack_ell(bow, cap)
    assert cols.dtype == torch.int32 and vals.dtype == torch.float32 and cnt.dtype == torch.int32
    for i in range(N):
        nz = bow[i].nonzero().flatten()
        assert int(cnt[i]) == len(nz)
        k = min(len(nz), cap)
        got = cols[i, :k].long()
        assert len(set(got.tolist())) == k and set(got.tolist()) <= set(nz.tolist())
        if len(nz) <= cap:
            assert torch.equal(got.sort().values, nz)
        assert torch.equal(vals[i, :k], bow[i, got])
    torch.manual_seed(1)
    Z, W, b = torch.randn(N, 64, device=DEV), torch.randn(V, 64, device=DEV) / 8, torch.randn(V, device=DEV) / 8
    dense = F.bow_reconstruction_loss(Z, W, b, bow, 0.1, ell_cap=cap)
    both = F.bow_reconstruction_loss(Z, W, b, bow, 0.1, ell=(cols, vals, cnt))
    assert torch.equal(dense, both)


def test_bow_extreme_logits_take_the_exact_path():
    """Peaked softmax rows (p up to ~0.99, log p << -27 elsewhere) leave the series' range and take
    the exact per-element path.  (Not pushed to p -> 1 - 1e-8: there log(1 - p) is ill-conditioned in
    fp32 and the reference's own fp32 result differs from its fp64 value by tens of percent.)"""
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
    # fp64 oracle: log p never underflows, so the -100 clamp acts exactly where ours does (an fp32
    # softmax flushes p to 0 below ~-87/-104 and clamps earlier; that regime is ambiguous in the reference itself)
    assert relv(loss, ref) < 2e-3
    assert torch.isfinite(Z.grad).all() and torch.isfinite(W.grad).all() and torch.isfinite(b.grad).all()
    assert rel(Z.grad, Zd.grad) < 5e-3


def test_mmd_row_block_sharding_emulated_ranks():
    """each 'rank' owns a range of 128-row Gram tiles and sees every column.  Emulated on
    one GPU (ranks as consecutive launches): partial block sums add up to the full statistic and each
    rank's gradient rows are complete without any reduction."""
    import ctypes as C
    from carel_vae_b200 import parallel as par
    from carel_vae_b200._lib import lib, check, float_array
    torch.manual_seed(3)
    n1, n2, d, alphas, world = 1500, 1100, 64, [0.1, 0.02], 4
    x = bf(torch.randn(n1, d, device=DEV) * 1.1).requires_grad_(True)
    y = bf(torch.randn(n2, d, device=DEV) * 1.1 + 0.3).requires_grad_(True)
    full = F.mmd(x, y, alphas)
    gx, gy = torch.autograd.grad(full, (x, y))
    L = lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    total = torch.zeros(3, dtype=torch.float64, device=DEV)
    gxs, gys = torch.zeros_like(gx), torch.zeros_like(gy)
    one = torch.ones((), device=DEV)
    al = float_array(alphas)
    for (t0, t1) in par.balanced_tri_ranges(n1 + n2, world):
        ws = torch.empty(L.carel_mmd_workspace_bytes(n1, n2, d, 0), dtype=torch.uint8, device=DEV)
        part = torch.zeros(3, dtype=torch.float64, device=DEV)
        xd, yd = x.detach(), y.detach()
        check(L.carel_mmd_fwd(C.c_void_p(xd.data_ptr()), d, C.c_void_p(yd.data_ptr()), d, n1, n2, d, al, len(alphas),
                              1e-5, 0, t0, t1, C.c_void_p(part.data_ptr()), None, C.c_void_p(ws.data_ptr()), ws.numel(), 0, st), "fwd")
        total += part
        px, py = torch.empty_like(gx), torch.empty_like(gy)
        check(L.carel_mmd_bwd(C.c_void_p(one.data_ptr()), n1, n2, d, al, len(alphas), 1e-5, t0, t1,
                              C.c_void_p(px.data_ptr()), d, C.c_void_p(py.data_ptr()), d, C.c_void_p(ws.data_ptr()), ws.numel(), 0, st), "bwd")
        # rows outside the rank's block are untouched zeros; inside they are the complete gradient
        gxs += px
        gys += py
    val = par.finalize_mmd(total, n1, n2)
    assert relv(val, full) < 1e-6
    assert rel(gxs, gx) < 1e-5 and rel(gys, gy) < 1e-5


def test_hsic_row_block_sharding_emulated_ranks():
    """HSIC row-block sharding: 'ranks' run one after the other on one GPU; the all-reduce of the row sums
    and scalars is emulated by summing the ranks' workspace regions."""
    from carel_vae_b200 import parallel as par
    torch.manual_seed(4)
    m, d, s, world = 1300, 64, 150.0, 3
    x = bf(torch.randn(m, d, device=DEV) * 1.1)
    y = bf(0.5 * x + torch.randn(m, d, device=DEV))
    xr, yr = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    full = F.hsic(xr, yr, s, s)
    gx, gy = torch.autograd.grad(full, (xr, yr))
    # pass 1: every rank's partial regions (captured through reduce_fn), pass 2: feed the sums back
    captured = [[] for _ in range(world)]
    for r in range(world):
        t0, t1 = par.row_tile_range(m, r, world)
        F.hsic(x, y, s, s, tile_range=(t0, t1), reduce_fn=lambda t, r=r: captured[r].append(t.clone()))
    sums = [sum(captured[r][k] for r in range(world)) for k in range(3)]
    gxs, gys = torch.zeros_like(gx), torch.zeros_like(gy)
    for r in range(world):
        t0, t1 = par.row_tile_range(m, r, world)
        it = iter(sums)
        xs, ys = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
        val = F.hsic(xs, ys, s, s, tile_range=(t0, t1), reduce_fn=lambda t: t.copy_(next(it)))
        assert relv(val, full) < 1e-6
        px, py = torch.autograd.grad(val, (xs, ys))
        gxs += px
        gys += py
    assert rel(gxs, gx) < 1e-5 and rel(gys, gy) < 1e-5


# ------------------------------------------------------------------------------------------
# fp32 mode (north_star: 1e-5 relative): arbitrary fp32 inputs, NOT rounded to bf16 beforehand
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n1,n2,d,alphas", [(2, 3, 8, [0.1]), (64, 64, 64, [0.1]), (50, 70, 24, [0.1]),
                                            (300, 300, 64, [0.1, 0.01, 1.0]), (2048, 2048, 64, [0.1]),
                                            (700, 900, 80, [0.05])])
def test_mmd_fp32_mode(n1, n2, d, alphas):
    torch.manual_seed(n1 + d + 1)
    x = (torch.randn(n1, d, device=DEV) * 1.1).requires_grad_(True)
    y = (torch.randn(n2, d, device=DEV) * 1.1 + 0.3).requires_grad_(True)
    xd, yd = x.detach().double().requires_grad_(True), y.detach().double().requires_grad_(True)
    ref = ho.mmd_statistic(xd, yd, alphas)
    rgx, rgy = torch.autograd.grad(ref, (xd, yd))
    out = F.mmd(x, y, alphas, precision="fp32")
    gx, gy = torch.autograd.grad(out, (x, y))
    assert relv(out, ref) < 1e-5
    assert rel(gx, rgx) < 1e-5 and rel(gy, rgy) < 1e-5


# ------------------------------------------------------------------------------------------
# fused classifier heads (carel_heads_fwd/bwd) vs the reference's PyTorch formulas in fp64
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,L,Ce,mode,drop", [(1, 24, 6, 0, False), (64, 24, 6, 0, True), (300, 128, 6, 0, True),
                                              (257, 64, 1, 1, True), (8192, 128, 6, 0, True), (40, 100, 3, 0, False)])
def test_classifier_heads_vs_torch(N, L, Ce, mode, drop):
    import torch.nn as nn
    import torch.nn.functional as TF
    torch.manual_seed(N + L)
    z = torch.randn(N, 2 * L, device=DEV, requires_grad=True)
    ec, cc, pc = nn.Linear(L, Ce).to(DEV), nn.Linear(L, 1).to(DEV), nn.Linear(2 * L, 1).to(DEV)
    emo = torch.randint(0, Ce, (N, 1), device=DEV) if mode == 0 else (torch.rand(N, 1, device=DEV) < 0.4).float()
    cau = (torch.rand(N, 1, device=DEV) < 0.5).float()
    pair = (torch.rand(N, 1, device=DEV) < 0.3).float()
    if N > 1:
        pair[0], pair[1 % N] = 1.0, 0.0                     # both classes present: finite pos_weight
    else:
        pair[0] = 1.0
    masks = None
    if drop:
        masks = tuple(TF.dropout(torch.ones(N, k, device=DEV), 0.5, True) for k in (L, L, 2 * L))
    ls, n_e, n_c, n_p, w = 0.1, 1.0, 1.0, 1.0, (10.0, 7.0, 30.0)
    total, parts = F.classifier_heads(z, masks, ec, cc, pc, emo, cau, pair, mode, ls, n_e, n_c, n_p, True, w)
    total.backward()
    got = [z.grad.clone()] + [p_.grad.clone() for m in (ec, cc, pc) for p_ in (m.weight, m.bias)]
    # reference formulas (mmd:461-513 / hsic:459-505) in fp64
    zd = z.detach().double().requires_grad_(True)
    mods = [nn.Linear(L, Ce).to(DEV).double(), nn.Linear(L, 1).to(DEV).double(), nn.Linear(2 * L, 1).to(DEV).double()]
    for md, ms in zip(mods, (ec, cc, pc)):
        md.load_state_dict({k: v.double() for k, v in ms.state_dict().items()})
    me, mc, mp = (m.double() for m in masks) if masks is not None else (1.0, 1.0, 1.0)
    xe, xc, xp = mods[0](zd[:, :L] * me), mods[1](zd[:, L:] * mc), mods[2](zd * mp)
    if mode == 0:
        le = TF.cross_entropy(xe, emo.flatten())
    else:
        le = TF.binary_cross_entropy(torch.sigmoid(xe), emo.double() * (1 - ls) + ls / n_e)
    lc = TF.binary_cross_entropy(torch.sigmoid(xc), cau.double() * (1 - ls) + ls / n_c)
    pw = (len(pair) - pair.sum()) / pair.sum()
    lp = TF.binary_cross_entropy_with_logits(xp, pair.double() * (1 - ls) + ls / n_p, pos_weight=pw.double())
    ref = w[0] * le + w[1] * lc + w[2] * lp
    ref.backward()
    want = [zd.grad] + [p_.grad for m in mods for p_ in (m.weight, m.bias)]
    assert relv(total, ref) < 1e-5
    for a_, b_ in zip((le, lc, lp), parts[:3]):
        assert relv(b_, a_) < 1e-5
    for a_, b_ in zip(got, want):
        assert rel(a_, b_) < 2e-5


def test_classifier_heads_infinite_pair_loss_guard():
    """No positive pair label: pos_weight = N / 0 = inf -> the reference replaces the infinite loss by 0 (mmd:510-511)."""
    import torch.nn as nn
    torch.manual_seed(3)
    N, L = 32, 24
    z = torch.randn(N, 2 * L, device=DEV, requires_grad=True)
    ec, cc, pc = nn.Linear(L, 6).to(DEV), nn.Linear(L, 1).to(DEV), nn.Linear(2 * L, 1).to(DEV)
    emo = torch.randint(0, 6, (N, 1), device=DEV)
    cau = (torch.rand(N, 1, device=DEV) < 0.5).float()
    pair = torch.zeros(N, 1, device=DEV)
    total, parts = F.classifier_heads(z, None, ec, cc, pc, emo, cau, pair, 0, 0.1, 1.0, 1.0, 1.0, True, (1.0, 1.0, 1.0))
    total.backward()
    assert float(parts[2]) == 0.0 and torch.isfinite(total) and torch.isfinite(z.grad).all()
    assert pc.weight.grad.abs().max().item() == 0.0


# ------------------------------------------------------------------------------------------
# BASELINE.json's full decoder size (N = 8192 clauses, 2L = 256, V = 30000): every row tile, the stream-K unit
# ranges of both backward chains (each CTA spans several owner tiles) and the partial combine at their real shapes
# ------------------------------------------------------------------------------------------
def test_bow_full_size_vs_oracle():
    torch.manual_seed(11)
    N, K, V = 8192, 256, 30000
    Z = bf(torch.randn(N, K, device=DEV) * 1.1).requires_grad_(True)
    W = bf((torch.rand(V, K, device=DEV) * 2 - 1) / K ** 0.5).requires_grad_(True)
    b = bf((torch.rand(V, device=DEV) * 2 - 1) / K ** 0.5).requires_grad_(True)
    bow = ho.synthetic_batch(N, V, D=8, nnz=24, seed=2, device=DEV)["bow"]
    loss = F.bow_reconstruction_loss(Z, W, b, bow, 0.1)
    loss.backward()
    # fp64 oracle in row blocks (the [N, V] fp64 intermediates of one shot would be ~20 GB)
    Zd, Wd, bd = (t.detach().double().requires_grad_(True) for t in (Z, W, b))
    ref = torch.zeros((), dtype=torch.float64, device=DEV)
    for r0 in range(0, N, 1024):
        # ho.bow_reconstruction takes the mean over its block: re-weight to the mean over all N rows
        part = ho.bow_reconstruction(Zd[r0:r0 + 1024], Wd, bd, bow[r0:r0 + 1024].double(), 0.1) * (1024 / N)
        part.backward()
        ref = ref + part.detach()
    assert relv(loss, ref) < 1e-5
    assert rel(Z.grad, Zd.grad) < 1e-3 and rel(W.grad, Wd.grad) < 1e-3 and rel(b.grad, bd.grad) < 1e-3
    # size-independent property: the loss is invariant under a shift of all logits of a row (softmax), so the
    # gradient w.r.t. the bias sums to zero over the vocabulary
    # (db is the column sum of the fp16-rounded dlogit tiles the dW tensor-core product consumes: the zero sum holds
    # to their rounding noise, 2^-12 / sqrt(#terms) of sum |db| ~ 1e-5, far inside the 1e-3 gradient tolerance)
    assert abs(float(b.grad.double().sum())) < 1e-4 * float(b.grad.double().abs().sum())


@pytest.mark.parametrize("n1,n2,d", [(130, 257, 64), (512, 300, 100)])
def test_dense_pairwise_helpers_are_differentiable(n1, n2, d):
    """pdist / pairwise_distances / GaussianKernelMatrix are plain autograd code in the reference (mmd:580-596,
    hsic:529-537), so a loss built from them has gradients; here the backward is W = G f'(d2) followed by two fp32
    GEMMs of the library.  Against the fp64 oracle on the same bf16-rounded inputs; the forward matrix carries the
    bf16-tier's operand rounding, hence 2e-2 on the gradients (measured far below)."""
    torch.manual_seed(n1 + d)
    x = bf(torch.randn(n1, d, device=DEV)).requires_grad_(True)
    y = bf(0.5 * torch.randn(n2, d, device=DEV) + 0.3).requires_grad_(True)
    x64, y64 = x.detach().double().requires_grad_(True), y.detach().double().requires_grad_(True)
    # pdist(x, y): both gradients
    g = torch.randn(n1, n2, device=DEV)
    out = cv.pdist(x, y)
    ref = ho.pdist(x64, y64)
    assert rel(out, ref) < 5e-3
    gx, gy = torch.autograd.grad((out * g).sum(), [x, y])
    rx, ry = torch.autograd.grad((ref * g.double()).sum(), [x64, y64])
    assert rel(gx, rx) < 2e-2 and rel(gy, ry) < 2e-2
    # pairwise_distances(x) and GaussianKernelMatrix(x, sigma): one tensor on both sides
    g2 = torch.randn(n1, n1, device=DEV)
    for fn, rfn in ((cv.pairwise_distances, ho.pairwise_distances),
                    (lambda t: cv.GaussianKernelMatrix(t, 100.0), lambda t: ho.gaussian_kernel_matrix(t, 100.0))):
        out, ref = fn(x), rfn(x64)
        (gx,) = torch.autograd.grad((out * g2).sum(), [x])
        (rx,) = torch.autograd.grad((ref * g2.double()).sum(), [x64])
        assert rel(out, ref) < 1e-2 and rel(gx, rx) < 2e-2
    # no graph when nothing asks for one
    assert not cv.pdist(x.detach(), y.detach()).requires_grad
    with torch.no_grad():
        assert not cv.pairwise_distances(x).requires_grad
