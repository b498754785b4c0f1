"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL / NVLink).

Only what the hot path needs:

* `row_tile_range` / `sharded_mmd`: Gram row-block sharding.  Rank g owns a contiguous range of
  128-row tiles of the stacked sample and sees EVERY column, so its three block sums are partial and
  its gradient rows are complete: one all-reduce of 3 doubles in forward, no data-path collective in
  backward (the row blocks are all-gathered only if the caller wants the full gradient everywhere).
* `FlatGradBucket`: data-parallel head training — the gradients of the optimiser's parameters live
  in ONE flat fp32 buffer (param.grad are views), so the DP reduction is a single NCCL all-reduce
  of ~31 MB instead of one call per tensor.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def row_tile_range(n_rows: int, rank: int, world: int, tile: int = 128) -> Tuple[int, int]:
    """Contiguous, balanced split of ceil(n_rows / tile) row tiles over `world` ranks.

    Balanced by TILE COUNT for the full Gram; with the triangular schedule row tiles near the top own
    more column tiles, so `balanced_tri_ranges` should be preferred there."""
    tiles = (n_rows + tile - 1) // tile
    base, extra = divmod(tiles, world)
    t0 = rank * base + min(rank, extra)
    return t0, t0 + base + (1 if rank < extra else 0)


def balanced_tri_ranges(n_rows: int, world: int, tile: int = 128, col_tile: int = 256) -> List[Tuple[int, int]]:
    """Row-tile ranges that balance the number of (row tile, column tile) pairs on or above the
    diagonal — the work of the triangular MMD schedule (row tile t owns n_col_tiles - t // 2 tiles)."""
    tiles = (n_rows + tile - 1) // tile
    ncol = (n_rows + col_tile - 1) // col_tile
    work = [ncol - (t * tile) // col_tile for t in range(tiles)]
    total = sum(work)
    ranges, t0, acc = [], 0, 0
    for r in range(world):
        target = total * (r + 1) / world
        t1 = t0
        while t1 < tiles and (acc + work[t1] <= target or t1 == t0 and r == world - 1):
            acc += work[t1]
            t1 += 1
        if r == world - 1:
            t1 = tiles
        elif (t1 & 1) and t1 < tiles:
            # keep rank boundaries on 256-row (CTA-pair) tiles: the long-K Gram kernel pairs two row tiles
            acc += work[t1]
            t1 += 1
        ranges.append((t0, t1))
        t0 = t1
    return ranges


def finalize_mmd(partial: torch.Tensor, n_1: int, n_2: int) -> torch.Tensor:
    """mmd from the (all-reduced) block sums [sum_xx offdiag, sum_yy offdiag, sum_xy + sum_yx]."""
    a00 = 1.0 / (n_1 * (n_1 - 1))
    a11 = 1.0 / (n_2 * (n_2 - 1))
    a01 = -1.0 / (n_1 * n_2)
    return a00 * partial[0] + a11 * partial[1] + a01 * partial[2]


def sharded_mmd(x: torch.Tensor, y: torch.Tensor, alphas: Sequence[float], group=None, eps: float = 1e-5,
                partial_fn=None) -> torch.Tensor:
    """MMD^2 of replicated samples with the Gram row blocks sharded over the ranks of `group`.

    Every rank passes the SAME x, y (all-gather DP-sharded inputs first) and gets the same value.
    `partial_fn(x, y, alphas, eps, t0, t1) -> float64[3]` computes the block sums of row tiles
    [t0, t1); by default it launches the CUDA kernel (carel_mmd_fwd with a row-tile range)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ranges = balanced_tri_ranges(x.shape[0] + y.shape[0], world)
    t0, t1 = ranges[rank]
    if partial_fn is None:
        from . import functional as F
        partial = F._MMDFn.apply(x, y, tuple(alphas), eps, 0, (t0, t1 if t1 > t0 else t0), 0)
    else:
        partial = partial_fn(x, y, alphas, eps, t0, t1)
    partial = partial.clone()
    if world > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return finalize_mmd(partial, x.shape[0], y.shape[0]).to(torch.float32)


def sharded_hsic(x: torch.Tensor, y: torch.Tensor, s_x: float = 1.0, s_y: float = 1.0, group=None) -> torch.Tensor:
    """HSIC of replicated paired samples with the Gram row blocks sharded over the ranks of `group`
    (SURVEY.md §8(e)): forward partials = three scalars + the row sums of the rank's rows, combined by
    all-reduces of 2m + 4 doubles; the backward of a row block needs no collective (each rank gets the
    gradient rows of its block, zeros elsewhere — all-reduce / all-gather them if every rank needs all rows)."""
    from . import functional as F
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    t0, t1 = row_tile_range(x.shape[0], rank, world)

    def reduce_fn(t):
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return F.hsic(x, y, s_x, s_y, tile_range=(t0, t1), reduce_fn=reduce_fn)


class FlatGradBucket:
    """One flat fp32 gradient buffer for a parameter list; param.grad become views into it."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else "cpu"
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def allreduce_mean(self, group=None):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))
