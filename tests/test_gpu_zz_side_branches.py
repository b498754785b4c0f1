"""GPU (-m gpu): side branches (carel_set_side_branches) against the serial launch order, eagerly and under CUDA-graph
capture.  Kept in its own file, collected last: a failed capture leaves torch's CUDA generator registered with the dead
graph and would take every later test in the process down with it."""
import pytest
import torch

from oracle import head_oracle as ho

pytestmark = pytest.mark.gpu

DEV = "cuda"


def bf(x):
    return x.bfloat16().float()


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


@pytest.mark.parametrize("target", ["dense", "ell"])
def test_side_branches_do_not_change_the_result(target):
    """carel_set_side_branches: the sparse-target branch of carel_bow_fwd (dense -> ELL scan / list copy, CSC view) on
    the library's second stream — same kernels on the same data, so loss and gradients are the serial ones (1e-6 /
    1e-5: only the order of the double atomics can differ) — run eagerly three times in a row (the branch of call k
    must not run into call k + 1) and as a captured CUDA graph (the fork / join events become graph dependencies)."""
    from carel_vae_b200 import _lib, functional as F
    N, K, V = 1024, 256, 9000
    torch.manual_seed(11)
    b = ho.synthetic_batch(N, V, D=8, seed=6, device=DEV)
    dense = b["bow"]
    ell = F.bow_pack_ell(dense, 64)
    # Everything below runs on the stream the graph is captured on: a leaf created on the legacy default stream would make
    # autograd synchronise with that stream during capture (bench.HeadStep warms up on its own stream for the same reason).
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    before = _lib.side_branches()
    runs = []
    try:
        with torch.cuda.stream(s):
            z = bf(torch.randn(N, K, device=DEV)).requires_grad_(True)
            W = bf(0.05 * torch.randn(V, K, device=DEV)).requires_grad_(True)
            bias = bf(0.1 * torch.randn(V, device=DEV)).requires_grad_(True)

            def step():
                if target == "dense":
                    loss = F.bow_reconstruction_loss(z, W, bias, dense, 0.1)
                else:
                    loss = F.bow_reconstruction_loss(z, W, bias, None, 0.1, ell=ell)
                return loss, torch.autograd.grad(loss, [z, W, bias])

            _lib.set_side_branches(False)
            assert _lib.side_branches() == 0
            l0, g0 = step()
            _lib.set_side_branches(True)
            assert _lib.side_branches() == 3
            runs = [step() for _ in range(3)]
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            lg, gg = step()
        for _ in range(2):
            g.replay()
            torch.cuda.synchronize()
            runs.append((lg.clone(), tuple(t.clone() for t in gg)))
    finally:
        _lib.set_side_branches(before)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    assert torch.isfinite(l0) and float(l0) > 0
    assert len(runs) == 5
    for (l, gr) in runs:
        assert rel(l, l0) < 1e-6
        for a, r in zip(gr, g0):
            assert rel(a, r) < 1e-5
