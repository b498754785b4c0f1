"""Developer smoke script (GPU box): quick numerical + timing checks of the kernels while they
are being written.  The real parity tests live in tests/."""
import sys
import time

import torch

sys.path.insert(0, ".")
from carel_vae_b200 import functional as F  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def bf(x):
    return x.bfloat16().float()


def mmd_ref(x, y, alphas, eps=1e-5):
    n1 = x.shape[0]
    n2 = y.shape[0]
    s = torch.cat([x, y]).double()
    n = (s * s).sum(1, keepdim=True)
    d2 = n + n.T - 2 * s @ s.T
    dist2 = eps + d2.abs()
    K = sum(torch.exp(-a * dist2) for a in alphas)
    k1, k2, k12 = K[:n1, :n1], K[n1:, n1:], K[:n1, n1:]
    a00, a11, a01 = 1 / (n1 * (n1 - 1)), 1 / (n2 * (n2 - 1)), -1 / (n1 * n2)
    return 2 * a01 * k12.sum() + a00 * (k1.sum() - k1.trace()) + a11 * (k2.sum() - k2.trace())


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()


def timeit(fn, n=30, warm_ms=200.0, graph=True):
    """Median device time per call (ms).  Warms the clocks up first; replays a CUDA graph of
    `fn` so that host launch overhead is not what is measured."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    runner = fn
    if graph:
        try:
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                fn()
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=s):
                    fn()
            torch.cuda.current_stream().wait_stream(s)
            runner = g.replay
        except Exception as e:  # noqa: BLE001
            print("graph capture failed:", repr(e)[:200])
            runner = fn
    t0 = time.time()
    while (time.time() - t0) * 1e3 < warm_ms:
        for _ in range(10):
            runner()
        torch.cuda.synchronize()
    times = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        runner()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    return times[len(times) // 2]


def check_gemm():
    for (M, N, K) in [(128, 256, 64), (300, 500, 192), (1024, 768, 512), (8192, 512, 768)]:
        a = torch.randn(M, K, device=dev).bfloat16()
        b = torch.randn(N, K, device=dev).bfloat16()
        ref = a.float() @ b.float().T
        for a_mn in (False, True):
            for b_mn in (False, True):
                aa = a.T.contiguous() if a_mn else a
                bb = b.T.contiguous() if b_mn else b
                if (aa.stride(0) % 8) or (bb.stride(0) % 8):
                    continue
                for ks in (1, 3):
                    c = F.gemm_bf16(aa, bb, a_mn, b_mn, ksplit=ks)
                    print(f"gemm M{M} N{N} K{K} a_mn={int(a_mn)} b_mn={int(b_mn)} ksplit={ks}: rel {rel(c, ref):.2e}")


def check_mmd():
    for (n1, n2, d, alphas) in [(64, 64, 64, [0.1]), (50, 70, 24, [0.1]), (300, 300, 128, [0.1, 0.01]),
                                (2048, 2048, 128, [0.1]), (1000, 1500, 96, [0.05]), (512, 512, 240, [0.02])]:
        x = bf(torch.randn(n1, d, device=dev) * 1.1).requires_grad_(True)
        y = bf(torch.randn(n2, d, device=dev) * 1.1 + 0.3).requires_grad_(True)
        for full in (False, True):
            out = F.mmd(x, y, alphas, full_gram=full)
            ref = mmd_ref(x, y, alphas)
            gx, gy = torch.autograd.grad(out, (x, y))
            rgx, rgy = torch.autograd.grad(ref, (x, y))
            print(f"mmd n1={n1} n2={n2} d={d} alphas={alphas} full={int(full)}: val {out.item():.6e} ref {ref.item():.6e} "
                  f"rel {abs(out.item() - ref.item()) / abs(ref.item()):.2e} | gx {rel(gx, rgx):.2e} gy {rel(gy, rgy):.2e}")


def bench_mmd():
    for (n, d) in [(8192, 128), (8192, 768), (4096, 768)]:
        x = torch.tanh(torch.randn(n, d, device=dev))
        y = torch.tanh(torch.randn(n, d, device=dev))
        if d == 768:
            x = x / x.norm(dim=1, keepdim=True)
            y = y / y.norm(dim=1, keepdim=True)
        for full in (False, True):
            ms = timeit(lambda: F.mmd(x, y, [0.1], full_gram=full))
            fl = 2 * (2 * n) ** 2 * d
            print(f"mmd fwd n={n} d={d} full={int(full)}: {ms * 1e3:.1f} us  -> {fl / ms / 1e9:.1f} TFLOP/s (algorithmic, full Gram)")
        if d <= 256:
            xr = x.clone().requires_grad_(True)
            yr = y.clone().requires_grad_(True)

            def fb():
                o = F.mmd(xr, yr, [0.1])
                o.backward()
            ms = timeit(fb)
            print(f"mmd fwd+bwd n={n} d={d}: {ms * 1e3:.1f} us")





def check_latent():
    import torch.nn.functional as TF
    for (N, D, L) in [(64, 768, 24), (300, 768, 64), (8192, 768, 128)]:
        X = bf(torch.tanh(torch.randn(N, D, device=dev))).requires_grad_(True)
        Ws = [bf(torch.randn(L, D, device=dev) / D ** 0.5).requires_grad_(True) for _ in range(4)]
        bs = [bf(torch.randn(L, device=dev) * 0.1).requires_grad_(True) for _ in range(4)]
        ee, ec = torch.randn(L, device=dev), torch.randn(L, device=dev)
        args = [X, ee, ec]
        for w_, b_ in zip(Ws, bs):
            args += [w_, b_]
        z, kl, pre = F.latent(*args)
        # reference
        Xd = X.detach().double().requires_grad_(True)
        Wd = [w_.detach().double().requires_grad_(True) for w_ in Ws]
        bd = [b_.detach().double().requires_grad_(True) for b_ in bs]
        pr = [TF.linear(Xd, w_, b_) for w_, b_ in zip(Wd, bd)]
        ze = pr[0] + ee.double() * pr[1].exp()
        zc = pr[2] + ec.double() * pr[3].exp()
        zr = torch.cat([ze, zc], 1)
        klr = torch.stack([(-0.5 * (1 + pr[1] - pr[1].exp() - pr[0] ** 2).sum(1)).mean(),
                           (-0.5 * (1 + pr[3] - pr[3].exp() - pr[2] ** 2).sum(1)).mean()])
        gzs = torch.randn_like(z)
        gk = torch.tensor([0.7, -1.3], device=dev)
        (z * gzs).sum().add((kl * gk).sum()).backward()
        ((zr * gzs.double()).sum() + (klr * gk.double()).sum()).backward()
        print(f"latent N={N} L={L}: z {rel(z, zr):.2e} kl {rel(kl, klr):.2e} pre {rel(pre, torch.cat(pr, 1)):.2e} | "
              f"gX {rel(X.grad, Xd.grad):.2e} gW {max(rel(a.grad, b.grad) for a, b in zip(Ws, Wd)):.2e} "
              f"gb {max(rel(a.grad, b.grad) for a, b in zip(bs, bd)):.2e}")


def check_bow():
    import torch.nn.functional as TF
    for (N, K, V, nnz) in [(64, 48, 300, 5), (100, 128, 1000, 24), (256, 256, 5000, 24), (1024, 256, 30000, 24)]:
        Z = bf(torch.randn(N, K, device=dev) * 1.1).requires_grad_(True)
        W = bf((torch.rand(V, K, device=dev) * 2 - 1) / K ** 0.5).requires_grad_(True)
        b = bf((torch.rand(V, device=dev) * 2 - 1) / K ** 0.5).requires_grad_(True)
        idx = torch.randint(0, V, (N, nnz), device=dev)
        bow = torch.zeros(N, V, device=dev)
        bow.scatter_add_(1, idx, torch.ones(N, nnz, device=dev))
        bow = bow / bow.sum(1, keepdim=True).clamp_min(1)
        loss = F.bow_reconstruction_loss(Z, W, b, bow)
        loss.backward()
        Zd, Wd, bd = (t.detach().double().requires_grad_(True) for t in (Z, W, b))
        pr = torch.softmax(TF.linear(Zd, Wd, bd), 1)
        t = bow.double() * 0.9 + 0.1 / V
        lr = TF.binary_cross_entropy(pr, t)
        lr.backward()
        print(f"bow N={N} K={K} V={V}: loss {loss.item():.6e} ref {lr.item():.6e} rel {abs(loss.item() - lr.item()) / abs(lr.item()):.2e} | "
              f"gZ {rel(Z.grad, Zd.grad):.2e} gW {rel(W.grad, Wd.grad):.2e} gb {rel(b.grad, bd.grad):.2e}")


def check_hsic():
    for (m, d, sx, sy) in [(64, 24, 30.0, 30.0), (200, 64, 100.0, 150.0), (1024, 128, 300.0, 300.0), (500, 128, 1.0, 1.0),
                           (4096, 128, 300.0, 300.0)]:
        x = bf(torch.randn(m, d, device=dev) * 1.1).requires_grad_(True)
        y = bf(0.5 * x.detach() + torch.randn(m, d, device=dev)).requires_grad_(True)
        out = F.hsic(x, y, sx, sy)
        gx, gy = torch.autograd.grad(out, (x, y))
        xd, yd = x.detach().double().requires_grad_(True), y.detach().double().requires_grad_(True)

        def gram(t, s):
            n = (t * t).sum(1, keepdim=True)
            return torch.exp(-(n + n.T - 2 * t @ t.T) / s)
        K, L = gram(xd, sx), gram(yd, sy)
        ref = ((K * L).sum() - 2.0 / m * (K.sum(1) * L.sum(1)).sum() + K.sum() * L.sum() / m ** 2) / (m - 1) ** 2
        rgx, rgy = torch.autograd.grad(ref, (xd, yd))
        print(f"hsic m={m} d={d} s=({sx},{sy}): val {out.item():.6e} ref {ref.item():.6e} rel {abs(out.item() - ref.item()) / abs(ref.item()):.2e} | "
              f"gx {rel(gx, rgx):.2e} gy {rel(gy, rgy):.2e} (|g| {rgx.norm().item():.2e})")


def bench_head_parts():
    N, D, L, V = 8192, 768, 128, 30000
    X = torch.tanh(torch.randn(N, D, device=dev)).requires_grad_(True)
    Ws = [(torch.randn(L, D, device=dev) / D ** 0.5).requires_grad_(True) for _ in range(4)]
    bs = [(torch.randn(L, device=dev) * 0.1).requires_grad_(True) for _ in range(4)]
    ee, ec = torch.randn(L, device=dev), torch.randn(L, device=dev)
    args = [X, ee, ec]
    for w_, b_ in zip(Ws, bs):
        args += [w_, b_]

    def lat():
        z, kl, _ = F.latent(*args)
        (z.sum() + kl.sum()).backward()
    print(f"latent fwd+bwd: {timeit(lat) * 1e3:.1f} us")
    Z = (torch.randn(N, 2 * L, device=dev) * 1.1).requires_grad_(True)
    W = ((torch.rand(V, 2 * L, device=dev) * 2 - 1) / 16).requires_grad_(True)
    b = ((torch.rand(V, device=dev) * 2 - 1) / 16).requires_grad_(True)
    idx = torch.randint(0, V, (N, 24), device=dev)
    bow = torch.zeros(N, V, device=dev)
    bow.scatter_add_(1, idx, torch.ones(N, 24, device=dev))
    bow = bow / bow.sum(1, keepdim=True).clamp_min(1)
    print(f"bow fwd: {timeit(lambda: F.bow_reconstruction_loss(Z, W, b, bow)) * 1e3:.1f} us")

    def bw():
        F.bow_reconstruction_loss(Z, W, b, bow).backward()
    print(f"bow fwd+bwd: {timeit(bw) * 1e3:.1f} us")


if __name__ == "__main__":
    what = sys.argv[1:] or ["gemm", "mmd", "latent", "bow", "bench"]
    t = time.time()
    if "gemm" in what:
        check_gemm()
    if "mmd" in what:
        check_mmd()
    if "latent" in what:
        check_latent()
    if "bow" in what:
        check_bow()
    if "hsic" in what:
        check_hsic()
    if "bench" in what:
        bench_mmd()
    if "bench2" in what:
        bench_head_parts()
    print(f"done in {time.time() - t:.1f}s")
