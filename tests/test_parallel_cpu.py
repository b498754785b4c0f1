"""CPU, world_size 2 over gloo: the host logic of the multi-GPU path (row-tile partition, partial-sum
all-reduce + finalize, flat gradient bucket).  The per-rank partial sums come from a torch stand-in of
the kernel's contract (block sums of a row-tile range) so no GPU is needed."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from carel_vae_b200 import parallel as par
from oracle import head_oracle as ho


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_world2(target, extra=()):
    """Two spawned gloo ranks; one retry on a fresh port (the free port can be taken between probing and rendezvous, and
    a loaded machine can miss the queue deadline)."""
    ctx = mp.get_context("spawn")
    last = None
    for _attempt in range(2):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=target, args=(r, 2, port, q) + tuple(extra)) for r in range(2)]
        for p in procs:
            p.start()
        try:
            res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
            for p in procs:
                p.join(timeout=60)
            if all(p.exitcode == 0 for p in procs):
                return res
            last = RuntimeError(f"exit codes {[p.exitcode for p in procs]}")
        except Exception as e:  # noqa: BLE001  (queue.Empty: a rank died or hung)
            last = e
        for p in procs:
            if p.is_alive():
                p.kill()
            p.join(timeout=10)
    raise AssertionError(f"world-2 gloo run failed twice: {last!r}")


def _partial_cpu(x, y, alphas, eps, t0, t1):
    """Contract of carel_mmd_fwd(row_tile_begin=t0, row_tile_end=t1) in its triangular mode: block sums over
    pairs (i, j) with i in the row tiles and j > i, each weighted 2."""
    s = torch.cat([x, y]).double()
    n1, M = x.shape[0], s.shape[0]
    n = (s * s).sum(1)
    r0, r1 = t0 * 128, min(t1 * 128, M)
    out = torch.zeros(3, dtype=torch.float64)
    if r1 <= r0:
        return out
    d2 = n[r0:r1, None] + n[None, :] - 2 * s[r0:r1] @ s.T
    K = sum(torch.exp(-a * (d2.abs() + eps)) for a in alphas)
    i = torch.arange(r0, r1)[:, None]
    j = torch.arange(M)[None, :]
    upper = (j > i).double() * 2.0
    rx, cx = i < n1, j < n1
    out[0] = (K * upper * (rx & cx)).sum()
    out[1] = (K * upper * (~rx & ~cx)).sum()
    out[2] = (K * upper * (rx != cx)).sum()
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(300, 16, generator=g), torch.randn(420, 16, generator=g) + 0.3
    val = par.sharded_mmd(x, y, [0.1, 1.0], partial_fn=_partial_cpu)
    ref = ho.mmd_statistic(x.double(), y.double(), [0.1, 1.0])
    # data-parallel gradient bucket
    torch.manual_seed(1)
    lin = torch.nn.Linear(8, 4)
    bucket = par.FlatGradBucket(lin.parameters())
    bucket.zero_()
    (lin(torch.ones(2, 8) * (rank + 1)).sum()).backward()
    local = bucket.flat.clone()
    bucket.allreduce_mean()
    q.put((rank, float(val), float(ref), local, bucket.flat.clone(), lin.weight.grad.data_ptr() == bucket.flat.data_ptr()))
    dist.destroy_process_group()


def test_row_tile_partition_covers_everything():
    for n_rows in (1, 128, 129, 16384, 262144, 1000):
        for world in (1, 2, 3, 8):
            tiles = (n_rows + 127) // 128
            cover = []
            for r in range(world):
                t0, t1 = par.row_tile_range(n_rows, r, world)
                cover += list(range(t0, t1))
            assert cover == list(range(tiles))
            rs = par.balanced_tri_ranges(n_rows, world)
            assert rs[0][0] == 0 and rs[-1][1] == tiles and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
    # triangular balance: the work per rank is within one row of tiles of the mean
    rs = par.balanced_tri_ranges(262144, 8)
    ncol = 262144 // 256
    work = [sum(ncol - t // 2 for t in range(a, b)) for a, b in rs]
    assert max(work) - min(work) <= 2 * ncol


def test_sharded_mmd_and_flat_bucket_world2():
    res = _run_world2(_worker)
    for rank, val, ref, local, reduced, is_view in res:
        assert abs(val - ref) <= 1e-6 * abs(ref)
        assert is_view
    assert torch.allclose(res[0][4], (res[0][3] + res[1][3]) / 2)
    assert torch.equal(res[0][4], res[1][4])


def _worker_reducer(rank, world, port, q, overlapped=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(3)                                   # same parameters on every rank
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 2))
    # the [16, 8] weight goes in place, the rest through the flat buffer
    red = (par.OverlappedGradReducer if overlapped else par.GradReducer)(net.parameters(), big_numel=100)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    same0 = par.identical_across_ranks(net.parameters())
    x = torch.ones(4, 8) * (rank + 1)
    opt.zero_grad()
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    red.reduce()
    opt.step()
    same1 = par.identical_across_ranks(net.parameters())
    differ = par.identical_across_ranks([torch.full((3,), float(rank))])
    # plain lists: a tensor in the queue is shared by file descriptor and this process may exit before the parent maps it
    q.put((rank, [t.tolist() for t in local], [p.grad.tolist() for p in net.parameters()], len(red.big), len(red.small),
           same0, same1, differ))
    dist.destroy_process_group()


@pytest.mark.parametrize("overlapped", [False, True])
def test_grad_reducer_world2(overlapped):
    """GradReducer (data-parallel head training): averaged gradients through both paths, replicas stay
    bit-identical after the optimiser step, and identical_across_ranks notices a difference.  OverlappedGradReducer
    keeps the same protocol; on CPU tensors (gloo) its hooks stay inert and reduce() is the plain reduction — the
    overlapped NCCL path is checked on GPUs inside bench.py (dp_overlap.max_rel_diff_vs_plain_allreduce)."""
    res = _run_world2(_worker_reducer, (overlapped,))
    (_, l0, g0, nb, ns, s00, s01, d0), (_, l1, g1, _, _, s10, s11, d1) = res
    assert nb == 1 and ns == 3
    for a, b, ga, gb in zip(l0, l1, g0, g1):
        a, b, ga, gb = (torch.tensor(t) for t in (a, b, ga, gb))
        assert torch.allclose(ga, (a + b) / 2) and torch.equal(ga, gb)
    assert s00 and s01 and s10 and s11 and not d0 and not d1
