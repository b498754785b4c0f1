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


class _ShardedMMDFn(torch.autograd.Function):
    """Row-block-sharded MMD^2 as ONE autograd node: forward = block sums of this rank's row tiles (carel_mmd_fwd with a
    row-tile range) -> all-reduce of 3 doubles -> carel_mmd_finalize; backward = carel_mmd_bwd over the same row tiles,
    which writes the COMPLETE gradient rows of the rank's block (it sees every column) and zeros elsewhere."""

    @staticmethod
    def forward(ctx, x, y, alphas, eps, t0, t1, group, all_rows):
        from . import functional as F
        F._require_cuda(x, y)
        xf, yf = F._rowmajor_f32(x), F._rowmajor_f32(y)
        n1, d = xf.shape
        n2 = yf.shape[0]
        L = F.lib()
        ws = F._workspace(L.carel_mmd_workspace_bytes(n1, n2, d, 0), xf.device)
        partial = torch.empty(3, dtype=torch.float64, device=xf.device)
        out = torch.empty((), dtype=torch.float32, device=xf.device)
        F.check(L.carel_mmd_fwd(F._p(xf), xf.stride(0), F._p(yf), yf.stride(0), n1, n2, d, F.float_array(alphas), len(alphas),
                                float(eps), 0, t0, t1, F._p(partial), F._p(None), F._p(ws), ws.numel(), 0, F._stream()),
                "carel_mmd_fwd")
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
        F.check(L.carel_mmd_finalize(F._p(partial), n1, n2, F._p(out), F._stream()), "carel_mmd_finalize")
        ctx.ws, ctx.group, ctx.all_rows = ws, group, all_rows
        ctx.meta = (n1, n2, d, tuple(float(a) for a in alphas), float(eps), t0, t1, x.dtype, y.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        from . import functional as F
        n1, n2, d, alphas, eps, t0, t1, dtx, dty = ctx.meta
        ws = ctx.ws
        g = g.reshape(()).float().contiguous()
        gxy = torch.empty((n1 + n2, d), dtype=torch.float32, device=ws.device)      # one buffer: one collective
        gx, gy = gxy[:n1], gxy[n1:]
        F.check(F.lib().carel_mmd_bwd(F._p(g), n1, n2, d, F.float_array(alphas), len(alphas), eps, t0, t1, F._p(gx), d,
                                      F._p(gy), d, F._p(ws), ws.numel(), 0, F._stream()), "carel_mmd_bwd")
        if ctx.all_rows and dist.is_initialized() and dist.get_world_size(ctx.group) > 1:
            dist.all_reduce(gxy, op=dist.ReduceOp.SUM, group=ctx.group)             # rows of other ranks are zero here
        return gx.to(dtx), gy.to(dty), None, None, None, None, None, None


def sharded_mmd(x: torch.Tensor, y: torch.Tensor, alphas: Sequence[float], group=None, eps: float = 1e-5,
                partial_fn=None, all_rows: bool = False) -> torch.Tensor:
    """MMD^2 of replicated samples with the Gram row blocks sharded over the ranks of `group`; differentiable.

    Every rank passes the SAME x, y (all-gather DP-sharded inputs first) and gets the same value.  In backward every
    rank receives the complete gradient rows of ITS row block and zeros elsewhere (no collective); with
    `all_rows=True` the blocks are summed over the ranks so that every rank holds the full gradient.
    `partial_fn(x, y, alphas, eps, t0, t1) -> float64[3]` replaces the CUDA launch by a host function (CPU tests of
    the sharding logic; not differentiable)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ranges = balanced_tri_ranges(x.shape[0] + y.shape[0], world)
    t0, t1 = ranges[rank]
    if partial_fn is None:
        return _ShardedMMDFn.apply(x, y, tuple(alphas), eps, t0, t1 if t1 > t0 else t0, group, all_rows)
    if x.requires_grad or y.requires_grad:
        raise RuntimeError("sharded_mmd(partial_fn=...) is a forward-only test hook; drop partial_fn for gradients")
    partial = partial_fn(x, y, alphas, eps, t0, t1).clone()
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


class GradReducer:
    """Data-parallel gradient averaging for one optimiser's parameter list (head DP): tensors of at
    least `big_numel` elements (the decoder weight: 99 % of the bytes) are all-reduced IN PLACE with ncclAvg — no
    staging copy, no divide —, all the small ones travel through one flat buffer (pack, one all-reduce, unpack: two
    multi-tensor launches).  Works on whatever `.grad` tensors the backward pass left, so the reference loops'
    `optimizer.zero_grad()` / `loss.backward()` stay as they are: call `reduce()` between backward and `step()`."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, big_numel: int = 1 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.big = [p for p in self.params if p.numel() >= big_numel]
        self.small = [p for p in self.params if p.numel() < big_numel]
        n = sum(p.numel() for p in self.small)
        dev = self.params[0].device if self.params else "cpu"
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.small:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def snapshot(self):
        """The gradient tensors as they are NOW (one per parameter, None where there is none).  CUDA-graph users: a
        replay writes into the tensors that existed at capture time, while `param.grad` keeps pointing at whatever
        the LAST captured segment assigned — take a snapshot right after capturing the segment whose gradients are
        to be reduced and hand it to `reduce()` on every replay."""
        return [p.grad for p in self.params]

    def reduce(self, grads=None) -> None:
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        if grads is None:
            grads = self.snapshot()
        gmap = {id(p): g for p, g in zip(self.params, grads)}
        avg = dist.ReduceOp.AVG if self.flat.is_cuda else dist.ReduceOp.SUM        # gloo (CPU tests) has no AVG
        scale = None if self.flat.is_cuda else 1.0 / dist.get_world_size(self.group)
        for p in self.big:
            g = gmap[id(p)]
            if g is not None:
                dist.all_reduce(g, op=avg, group=self.group)
                if scale is not None:
                    g.mul_(scale)
        live = [(gmap[id(p)], v) for p, v in zip(self.small, self.views) if gmap[id(p)] is not None]
        if not live:
            return
        if len(live) != len(self.small):
            self.flat.zero_()
        torch._foreach_copy_([v for _, v in live], [g for g, _ in live])
        dist.all_reduce(self.flat, op=avg, group=self.group)
        if scale is not None:
            self.flat.mul_(scale)
        torch._foreach_copy_([g for g, _ in live], [v for _, v in live])


class OverlappedGradReducer(GradReducer):
    """GradReducer that hides the big all-reduces under the rest of the backward pass (hard part 5).

    The decoder weight gradient (31 MB at V = 30000, L = 128: 99 % of the DP bytes) is complete as soon as the decoder's
    backward has run — the MMD / HSIC backward, the heads and the whole latent backward (~250 us at N = 8192) still
    follow.  The reducer forks a communication stream as early as the gradient exists and issues the in-place ncclAvg
    all-reduce on it: for the decoder weight that is INSIDE the decoder's backward, between the kernels that complete
    the weight gradient and the latent-gradient kernel (`functional.decoder_grad_hooks`, ~400 us of cover); any other
    big parameter is caught by a post-accumulate-grad hook.  `reduce()` joins and sends the small gradients through
    the flat bucket as before.  Two things make the overlap real:
      * the library's persistent kernels own every SM (one 225 KB CTA each), so a collective launched beside them would
        only queue behind them: between the hook and the join `carel_set_reserved_sms(reserved_sms)` makes every kernel
        of the library size its grid for that many fewer SMs, and NCCL (capped to as many CTAs: NCCL_MAX_CTAS) runs on
        the ones left free;
      * fork and join are plain stream dependencies, so the whole step — collective included — can be captured in ONE
        CUDA graph (call `reduce()` inside the captured region).
    The trainer-facing protocol is unchanged: `reducer.reduce()` between `loss.backward()` and `optimizer.step()`."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, big_numel: int = 1 << 20, reserved_sms: int = 16):
        super().__init__(params, group, big_numel)
        self.reserved_sms = int(reserved_sms)
        self.enabled = True
        self._comm = None
        self._done = {}          # id(param) -> the gradient tensor that was all-reduced from the hook in this pass
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_big_grad) for p in self.big]
        self._big_ids = {id(p): p for p in self.big}
        from . import functional as F
        F.decoder_grad_hooks.append(self._on_decoder_grad)

    def _active(self) -> bool:
        return self.enabled and self.flat.is_cuda and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _start(self, p: torch.nn.Parameter, g: torch.Tensor) -> None:
        from . import _lib
        main = torch.cuda.current_stream(g.device)
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=g.device)
        self._comm.wait_stream(main)                 # fork: the gradient is complete on `main` here
        with torch.cuda.stream(self._comm):
            dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
        self._done[id(p)] = g
        if self.reserved_sms > 0:
            _lib.set_reserved_sms(self.reserved_sms)  # everything launched until reduce() leaves these SMs to NCCL

    def _on_decoder_grad(self, dec_w, grad_w: torch.Tensor) -> bool:
        """Called by the decoder's backward with the freshly written weight gradient.  When the parameter has no
        gradient yet (`zero_grad(set_to_none=True)`, torch's default) the tensor BECOMES `.grad` here — what autograd's
        AccumulateGrad would do a moment later, except that autograd declines to adopt a tensor somebody else (the
        collective's work object) also references and would clone it underneath the running all-reduce — and its
        all-reduce starts; returns True so that the backward hands autograd None for it.  Otherwise (accumulation into
        an existing gradient) nothing happens here and the post-accumulate hook takes over."""
        p = self._big_ids.get(id(dec_w))
        if p is None or not self._active() or p.grad is not None or not grad_w.is_contiguous() or id(p) in self._done:
            return False
        if grad_w.dtype != p.dtype or grad_w.shape != p.shape:
            return False
        p.grad = grad_w
        self._start(p, grad_w)
        return True

    def _on_big_grad(self, p: torch.nn.Parameter) -> None:
        if not self._active() or p.grad is None or not p.grad.is_cuda or id(p) in self._done:
            return
        self._start(p, p.grad)

    def reduce(self, grads=None) -> None:
        if not self._active():
            return super().reduce(grads)
        from . import _lib
        if self.reserved_sms > 0:
            _lib.set_reserved_sms(0)
        if grads is None:
            grads = self.snapshot()
        gmap = {id(p): g for p, g in zip(self.params, grads)}
        if self._done:
            torch.cuda.current_stream().wait_stream(self._comm)      # join
        for p in self.big:
            g = gmap[id(p)]
            d = self._done.get(id(p))
            if g is not None and (d is None or d.data_ptr() != g.data_ptr()):     # not seen by a hook (or replaced since)
                dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
        self._done.clear()
        live = [(gmap[id(p)], v) for p, v in zip(self.small, self.views) if gmap[id(p)] is not None]
        if not live:
            return
        if len(live) != len(self.small):
            self.flat.zero_()
        torch._foreach_copy_([v for _, v in live], [g for g, _ in live])
        dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        torch._foreach_copy_([g for g, _ in live], [v for _, v in live])

    def close(self) -> None:
        from . import functional as F
        for h in self._hooks:
            h.remove()
        self._hooks = []
        if self._on_decoder_grad in F.decoder_grad_hooks:
            F.decoder_grad_hooks.remove(self._on_decoder_grad)


def identical_across_ranks(tensors: Iterable[torch.Tensor], group=None) -> bool:
    """True if every rank holds bit-identical tensors (data-parallel replicas must, after any number of steps): the
    int32 views are all-reduced with MIN and MAX and compared."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return True
    ok = True
    for t in tensors:
        bits = t.detach().contiguous().view(torch.int32).reshape(-1)
        lo, hi = bits.clone(), bits.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        ok = ok and bool(torch.equal(lo, hi))
    return ok
