"""Developer script (GPU box): decoder (Y-staged) parity + timing, step by step, printing numbers instead of asserting."""
import sys
import time

import torch

sys.path.insert(0, ".")
from carel_vae_b200 import _lib, functional as F  # noqa: E402
from oracle import head_oracle as ho  # noqa: E402

dev = "cuda"


def bf(x):
    return x.bfloat16().float()


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()


def one(N, K, V, nnz, cap, wscale=None, zscale=1.1, seed=0, bscale=None, rounded=True):
    torch.manual_seed(seed + V)
    r = bf if rounded else (lambda t: t)
    Z = r(torch.randn(N, K, device=dev) * zscale).requires_grad_(True)
    ws = (1 / K ** 0.5) if wscale is None else wscale
    W = r((torch.rand(V, K, device=dev) * 2 - 1) * ws).requires_grad_(True)
    b = r((torch.rand(V, device=dev) * 2 - 1) * (ws if bscale is None else bscale)).requires_grad_(True)
    bow = ho.synthetic_batch(N, V, D=8, nnz=nnz, seed=N, device=dev)["bow"]
    t0 = time.time()
    loss = F.bow_reconstruction_loss(Z, W, b, bow, 0.1, ell_cap=cap)
    loss.backward()
    torch.cuda.synchronize()
    Zd, Wd, bd = (t.detach().double().requires_grad_(True) for t in (Z, W, b))
    ref = ho.bow_reconstruction(Zd, Wd, bd, bow.double(), 0.1)
    ref.backward()
    print(f"N={N} K={K} V={V} nnz={nnz} cap={cap} ws={ws:.3f}: loss {loss.item():.6e} ref {ref.item():.6e} "
          f"relv {abs(loss.item() - ref.item()) / abs(ref.item()):.1e}  dZ {rel(Z.grad, Zd.grad):.1e} dW {rel(W.grad, Wd.grad):.1e} "
          f"db {rel(b.grad, bd.grad):.1e}  ({time.time() - t0:.2f}s)", flush=True)


if __name__ == "__main__":
    for cfg in [(64, 48, 300, 5, 64), (1, 48, 100, 3, 64), (100, 128, 1000, 24, 64), (256, 256, 5000, 24, 64),
                (130, 128, 777, 40, 8), (512, 256, 30000, 24, 64), (96, 40, 1001, 5, 64)]:
        one(*cfg)
    one(64, 48, 500, 6, 64, wscale=0.8, zscale=1.5, bscale=1.0, seed=7)        # peaked rows: exact path
    one(300, 256, 3000, 24, 64, rounded=False)
    # timing at the bench size
    N, K, V = 8192, 256, 30000
    torch.manual_seed(0)
    Z = (torch.randn(N, K, device=dev) * 1.1).requires_grad_(True)
    W = ((torch.rand(V, K, device=dev) * 2 - 1) / 16).requires_grad_(True)
    b = ((torch.rand(V, device=dev) * 2 - 1) / 16).requires_grad_(True)
    bow = ho.synthetic_batch(N, V, D=8, nnz=24, seed=0, device=dev)["bow"]
    for _ in range(3):
        loss = F.bow_reconstruction_loss(Z, W, b, bow, 0.1)
        loss.backward()
    _lib.timing_enable(True)
    for _ in range(5):
        loss = F.bow_reconstruction_loss(Z, W, b, bow, 0.1)
        loss.backward()
    rep = _lib.timing_report()
    _lib.timing_enable(False)
    print({k: round(v[0] / v[1] * 1e3, 1) for k, v in rep.items()})
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        loss = F.bow_reconstruction_loss(Z, W, b, bow, 0.1)
        loss.backward()
    e1.record()
    torch.cuda.synchronize()
    print("decoder fwd+bwd eager ms:", e0.elapsed_time(e1) / 10)
    Zd, Wd, bd = (t.detach().double().requires_grad_(True) for t in (Z, W, b))
    Z.grad = W.grad = b.grad = None
    loss = F.bow_reconstruction_loss(Z, W, b, bow, 0.1)
    loss.backward()
    ref = ho.bow_reconstruction(Zd, Wd, bd, bow.double(), 0.1)
    ref.backward()
    print(f"full size: relv {abs(loss.item() - ref.item()) / abs(ref.item()):.1e} dZ {rel(Z.grad, Zd.grad):.1e} "
          f"dW {rel(W.grad, Wd.grad):.1e} db {rel(b.grad, bd.grad):.1e}")
