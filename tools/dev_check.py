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


if __name__ == "__main__":
    what = sys.argv[1:] or ["gemm", "mmd", "bench"]
    t = time.time()
    if "gemm" in what:
        check_gemm()
    if "mmd" in what:
        check_mmd()
    if "bench" in what:
        bench_mmd()
    print(f"done in {time.time() - t:.1f}s")
