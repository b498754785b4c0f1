"""torch.autograd wrappers over the C ABI (include/carel_head.h).

Every function here launches the hand-written sm_100a kernels through ctypes; PyTorch only owns
the memory and the stream.  Tensors must live on a CUDA device — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import torch

from . import _lib
from ._lib import check, float_array, lib


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _require_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if not t.is_cuda:
            raise _lib.CarelError(
                "carel_vae_b200 runs on CUDA tensors only (sm_100a kernels, no CPU fallback); "
                f"got a tensor on {t.device}")


def _rowmajor_f32(t: torch.Tensor) -> torch.Tensor:
    """fp32 view/copy with unit stride in the last dim (row pitch may be anything)."""
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D tensor, got shape {tuple(t.shape)}")
    if t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------------------
# MMD
# --------------------------------------------------------------------------------------------
def _mmd_wide(d: int, prec: int) -> bool:
    """Feature widths the fused tensor-core backward cannot hold (bf16: d + 6 columns > 256) take their gradient from
    the materialising fp32 kernels (carel_mmd_f32_bwd): MMDStatistic is differentiable at any width."""
    return prec == 0 and d > 250


def _mmd_forward(xf, ldx, yf, ldy, n1, n2, d, alphas, eps, flags, t0, t1, partial, out, prec):
    """Launch the forward of the chosen tier; returns the workspace the matching backward needs."""
    L = lib()
    al = float_array(alphas)
    if prec == 2:
        ws = _workspace(L.carel_mmd_f32_workspace_bytes(n1, n2, d), out.device if out is not None else partial.device)
        check(L.carel_mmd_f32_fwd(xf, ldx, yf, ldy, n1, n2, d, al, len(alphas), float(eps), t0, t1, _p(partial),
                                  _p(None if partial is not None else out), _p(ws), ws.numel(), _stream()), "carel_mmd_f32_fwd")
        return ws
    ws = _workspace(L.carel_mmd_workspace_bytes(n1, n2, d, prec), out.device if out is not None else partial.device)
    check(L.carel_mmd_fwd(xf, ldx, yf, ldy, n1, n2, d, al, len(alphas), float(eps), int(flags), t0, t1, _p(partial),
                          _p(None if partial is not None else out), _p(ws), ws.numel(), prec, _stream()), "carel_mmd_fwd")
    return ws


def _mmd_backward(g, xf, ldx, yf, ldy, n1, n2, d, alphas, eps, t0, t1, gx, ldgx, gy, ldgy, ws, prec, dev):
    L = lib()
    al = float_array(alphas)
    if prec == 2 or _mmd_wide(d, prec):
        if prec != 2:
            ws = _workspace(L.carel_mmd_f32_workspace_bytes(n1, n2, d), dev)
        check(L.carel_mmd_f32_bwd(_p(g), xf, ldx, yf, ldy, n1, n2, d, al, len(alphas), eps, t0, t1, gx, ldgx, gy, ldgy,
                                  _p(ws), ws.numel(), _stream()), "carel_mmd_f32_bwd")
    else:
        check(L.carel_mmd_bwd(_p(g), n1, n2, d, al, len(alphas), eps, t0, t1, gx, ldgx, gy, ldgy, _p(ws), ws.numel(), prec,
                              _stream()), "carel_mmd_bwd")


class _MMDFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, alphas, eps, flags, tile_range, prec=0):
        _require_cuda(x, y)
        xf, yf = _rowmajor_f32(x), _rowmajor_f32(y)
        n1, d = xf.shape
        n2, d2 = yf.shape
        if d != d2:
            raise ValueError("sample_1 and sample_2 must have the same feature width")
        out = torch.empty((), dtype=torch.float32, device=xf.device)
        t0, t1 = tile_range
        partial = None
        if t0 != 0 or t1 != -1:
            partial = torch.empty(3, dtype=torch.float64, device=xf.device)
        ctx.ws = _mmd_forward(_p(xf), xf.stride(0), _p(yf), yf.stride(0), n1, n2, d, alphas, eps, flags, t0, t1, partial, out,
                              prec)
        ctx.keep = (xf, yf) if (prec == 2 or _mmd_wide(d, prec)) else None      # the fp32 kernels re-read the samples
        ctx.meta = (n1, n2, d, tuple(float(a) for a in alphas), float(eps), t0, t1, x.dtype, y.dtype, prec)
        if partial is not None:
            ctx.mark_non_differentiable(partial)
            return partial
        return out

    @staticmethod
    def backward(ctx, g):
        n1, n2, d, alphas, eps, t0, t1, dtx, dty, prec = ctx.meta
        ws = ctx.ws
        g = g.reshape(()).float().contiguous()
        gx = torch.empty((n1, d), dtype=torch.float32, device=ws.device)
        gy = torch.empty((n2, d), dtype=torch.float32, device=ws.device)
        xf, yf = ctx.keep if ctx.keep is not None else (None, None)
        _mmd_backward(g, _p(xf), 0 if xf is None else xf.stride(0), _p(yf), 0 if yf is None else yf.stride(0), n1, n2, d,
                      alphas, eps, t0, t1, _p(gx), d, _p(gy), d, ws, prec, ws.device)
        return gx.to(dtx), gy.to(dty), None, None, None, None, None


class _MMDHalvesFn(torch.autograd.Function):
    """MMD^2 between the two column halves of ONE tensor z = [x | y] (the head's z_e / z_c): same kernels as _MMDFn
    on strided views, but the backward writes both gradient halves into one [N, 2L] tensor, so autograd needs no
    slice-backward zero-fill / copy / add for them."""

    @staticmethod
    def forward(ctx, z, L, alphas, eps, prec):
        _require_cuda(z)
        zf = _rowmajor_f32(z)
        N = zf.shape[0]
        if zf.shape[1] != 2 * L:
            raise ValueError("z must be [N, 2L]")
        out = torch.empty((), dtype=torch.float32, device=zf.device)
        x, y = zf[:, :L], zf[:, L:]
        ctx.ws = _mmd_forward(_p(x), zf.stride(0), _p(y), zf.stride(0), N, N, L, alphas, eps, 0, 0, -1, None, out, prec)
        ctx.keep = zf if (prec == 2 or _mmd_wide(L, prec)) else None
        ctx.meta = (N, L, tuple(float(a) for a in alphas), float(eps), z.dtype, prec)
        return out

    @staticmethod
    def backward(ctx, g):
        N, L, alphas, eps, dtz, prec = ctx.meta
        ws = ctx.ws
        g = g.reshape(()).float().contiguous()
        gz = torch.empty((N, 2 * L), dtype=torch.float32, device=ws.device)
        zf = ctx.keep
        xp = yp = C.c_void_p(0)
        ldz = 0
        if zf is not None:
            xp, yp, ldz = _p(zf[:, :L]), _p(zf[:, L:]), zf.stride(0)
        _mmd_backward(g, xp, ldz, yp, ldz, N, N, L, alphas, eps, 0, -1, _p(gz[:, :L]), 2 * L, _p(gz[:, L:]), 2 * L, ws, prec,
                      ws.device)
        return gz.to(dtz), None, None, None, None


def mmd_halves(z: torch.Tensor, L: int, alphas: Sequence[float], eps: float = 1e-5,
               precision: str | None = None) -> torch.Tensor:
    """mmd(z[:, :L], z[:, L:], alphas) with a single [N, 2L] gradient (see _MMDHalvesFn)."""
    return _MMDHalvesFn.apply(z, int(L), tuple(alphas), eps, _lib.precision_code(precision))


def mmd(sample_1: torch.Tensor, sample_2: torch.Tensor, alphas: Sequence[float], eps: float = 1e-5,
        full_gram: bool = False, precision: str | None = None) -> torch.Tensor:
    """Unbiased multi-bandwidth RBF MMD^2 (reference MMDStatistic.__call__,
    drl_classifier_ec_mmd_final_mul.py:547-569).  0-dim fp32 tensor, differentiable w.r.t. both samples."""
    flags = (1 if full_gram else 0) | (2 if os.environ.get("CAREL_MMD_NO_PAIRS") else 0)
    return _MMDFn.apply(sample_1, sample_2, tuple(alphas), eps, flags, (0, -1), _lib.precision_code(precision))


def _pairwise_raw(af: torch.Tensor, bf: torch.Tensor, mode: int, alphas, eps: float) -> torch.Tensor:
    n1, d = af.shape
    n2 = bf.shape[0]
    L = lib()
    ws = _workspace(L.carel_pairwise_workspace_bytes(n1, n2, d), af.device)
    out = torch.empty((n1, n2), dtype=torch.float32, device=af.device)
    al = float_array(list(alphas) if len(alphas) else [0.0])
    check(L.carel_pairwise(_p(af), af.stride(0), _p(bf), bf.stride(0), n1, n2, d, mode, al, len(alphas),
                           float(eps), _p(out), n2, _p(ws), ws.numel(), _stream()), "carel_pairwise")
    return out


def _f32_gemm(a, ta: bool, b, tb: bool, M: int, N: int, K: int, alpha: float = 1.0) -> torch.Tensor:
    """C[M, N] = alpha opA(A) opB(B) through carel_f32_gemm (ta: A is [K, M]; tb: B is [K, N], else [N, K])."""
    c = torch.empty((M, N), dtype=torch.float32, device=a.device)
    check(lib().carel_f32_gemm(_p(a), a.stride(0), int(ta), _p(b), b.stride(0), int(tb), M, N, K, _p(c), N, None,
                               float(alpha), 0, 1, _stream()), "carel_f32_gemm")
    return c


class _PairwiseFn(torch.autograd.Function):
    """Dense pairwise matrices as differentiable ops (the reference's pdist / pairwise_distances /
    GaussianKernelMatrix are plain autograd code, mmd:580-596, hsic:529-537).  With M_ij = f(d2_ij),
    d2_ij = |a_i|^2 + |b_j|^2 - 2 a_i.b_j and W = G * f'(d2):  grad_a = 2 (rowsum(W) a - W b),
    grad_b = 2 (colsum(W) b - W^T a); the two products run on the library's fp32 GEMM."""

    @staticmethod
    def forward(ctx, a, b, mode, alphas, eps):
        af, bf = _rowmajor_f32(a), _rowmajor_f32(b)
        out = _pairwise_raw(af, bf, mode, alphas, eps)
        ctx.save_for_backward(af, bf, out)
        ctx.meta = (mode, tuple(float(x) for x in alphas), float(eps), a.dtype, b.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        af, bf, out = ctx.saved_tensors
        mode, alphas, eps, dta, dtb = ctx.meta
        g = g.float().contiguous()
        if mode == 2:                      # d2 itself
            w = g
        elif mode == 3:                    # exp(-alpha d2)
            w = g * out * (-alphas[0])
        elif mode == 1:                    # sqrt(eps + |d2|): the sign of d2 is torch's abs() rule, 0 at 0
            d2 = _pairwise_raw(af, bf, 2, (), eps)
            w = g * torch.sign(d2) / (2.0 * out)
        else:
            raise NotImplementedError("the kernel-sum matrix (MMDStatistic(..., ret_matrix=True)) is returned detached")
        n1, d = af.shape
        n2 = bf.shape[0]
        ga = gb = None
        if ctx.needs_input_grad[0]:
            ga = 2.0 * (w.sum(1, keepdim=True) * af - _f32_gemm(w, False, bf, True, n1, d, n2))
            ga = ga.to(dta)
        if ctx.needs_input_grad[1]:
            gb = 2.0 * (w.sum(0)[:, None] * bf - _f32_gemm(w, True, af, True, n2, d, n1))
            gb = gb.to(dtb)
        return ga, gb, None, None, None


def pairwise(a: torch.Tensor, b: torch.Tensor, mode: int, alphas: Sequence[float] = (), eps: float = 1e-5):
    """Dense [n1, n2] matrix: mode 0 kernel sum, 1 pdist, 2 squared distances, 3 exp(-d2 * alphas[0]).
    Differentiable w.r.t. both inputs for modes 1-3 (`a is b` — pairwise_distances(x) — receives both halves)."""
    _require_cuda(a, b)
    if torch.is_grad_enabled() and (a.requires_grad or b.requires_grad):
        if mode == 0:
            raise NotImplementedError("pairwise mode 0 (kernel-sum matrix) is not differentiable; detach the inputs")
        return _PairwiseFn.apply(a, b, int(mode), tuple(alphas), float(eps))
    return _pairwise_raw(_rowmajor_f32(a), _rowmajor_f32(b), int(mode), alphas, eps)


def gemm_bf16(a: torch.Tensor, b: torch.Tensor, a_mn: bool = False, b_mn: bool = False, bias=None,
              ksplit: int = 1, pairs: bool = False) -> torch.Tensor:
    """Debug entry: C = op(A) op(B)^T on the tcgen05 core.  a: [M,K] (or [K,M] if a_mn), bf16."""
    _require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    a, b = a.contiguous(), b.contiguous()
    M, K = (a.shape[1], a.shape[0]) if a_mn else a.shape
    N = b.shape[1] if b_mn else b.shape[0]
    c = torch.zeros((M, N), dtype=torch.float32, device=a.device)
    fn = lib().carel_gemm_bf16_2sm if pairs else lib().carel_gemm_bf16
    check(fn(_p(a), a.stride(0), int(a_mn), _p(b), b.stride(0), int(b_mn), M, N, K,
             _p(c), N, _p(bias), ksplit, _stream()), "carel_gemm_bf16")
    return c


# --------------------------------------------------------------------------------------------
# latent layer: fused 4x Linear + reparameterisation + KL
# --------------------------------------------------------------------------------------------
def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = 0 if t is None else t.data_ptr()
    return arr


class _LatentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps_e, eps_c, prec, *wb):
        # wb = (W_emu, b_emu, W_elv, b_elv, W_cmu, b_cmu, W_clv, b_clv)
        _require_cuda(x, eps_e, eps_c, *wb)
        xf = _rowmajor_f32(x)
        N, D = xf.shape
        ws_ = [w.detach().float().contiguous() for w in wb[0::2]]
        bs_ = [b.detach().float().contiguous() for b in wb[1::2]]
        L = ws_[0].shape[0]
        ee, ec = eps_e.detach().float().contiguous(), eps_c.detach().float().contiguous()
        lb = lib()
        f32 = prec == 2
        ws = _workspace((lb.carel_latent_f32_workspace_bytes if f32 else lb.carel_latent_workspace_bytes)(N, D, L), xf.device)
        pre = torch.empty((N, 4 * L), dtype=torch.float32, device=xf.device)
        z = torch.empty((N, 2 * L), dtype=torch.float32, device=xf.device)
        kl = torch.empty(2, dtype=torch.float32, device=xf.device)
        check((lb.carel_latent_f32_fwd if f32 else lb.carel_latent_fwd)(
            _p(xf), xf.stride(0), N, D, L, _ptr_array(ws_), _ptr_array(bs_), _p(ee), _p(ec), _p(pre), _p(z), _p(kl), _p(ws),
            ws.numel(), _stream()), "carel_latent_fwd")
        ctx.ws, ctx.pre, ctx.eps = ws, pre, (ee, ec)
        ctx.f32_keep = (xf, ws_) if f32 else None       # the fp32 kernels read x and the weights again in backward
        ctx.dims = (N, D, L)
        ctx.x_dtype = x.dtype
        ctx.wb_meta = [(t.dtype, tuple(t.shape)) for t in wb]
        ctx.mark_non_differentiable(pre)
        return z, kl, pre

    @staticmethod
    def backward(ctx, gz, gkl, _gpre):
        N, D, L = ctx.dims
        ws, pre = ctx.ws, ctx.pre
        ee, ec = ctx.eps
        dev = ws.device
        gz = None if gz is None else _rowmajor_f32(gz)
        gkl = None if gkl is None else gkl.float().contiguous()
        need = ctx.needs_input_grad
        gx = torch.empty((N, D), dtype=torch.float32, device=dev) if need[0] else None
        gws, gbs = [], []
        for i in range(4):
            gws.append(torch.empty((L, D), dtype=torch.float32, device=dev) if need[4 + 2 * i] else None)
            gbs.append(torch.empty((L,), dtype=torch.float32, device=dev) if need[5 + 2 * i] else None)
        if ctx.f32_keep is not None:
            xf, ws_ = ctx.f32_keep
            check(lib().carel_latent_f32_bwd(_p(gz), 0 if gz is None else gz.stride(0), _p(gkl), _p(xf), xf.stride(0),
                                             _ptr_array(ws_), _p(pre), _p(ee), _p(ec), N, D, L, _p(gx), D, _ptr_array(gws),
                                             _ptr_array(gbs), _p(ws), ws.numel(), _stream()), "carel_latent_f32_bwd")
        else:
            check(lib().carel_latent_bwd(_p(gz), 0 if gz is None else gz.stride(0), _p(gkl), _p(pre), _p(ee), _p(ec),
                                         N, D, L, _p(gx), D, _ptr_array(gws), _ptr_array(gbs), _p(ws), ws.numel(),
                                         _stream()), "carel_latent_bwd")
        out = [None if gx is None else gx.to(ctx.x_dtype), None, None, None]
        for i in range(4):
            out.append(gws[i])
            out.append(gbs[i])
        return tuple(out)


def latent(x, eps_e, eps_c, w_emu, b_emu, w_elv, b_elv, w_cmu, b_cmu, w_clv, b_clv, precision: str | None = None):
    """Fused latent layer.  Returns (z [N, 2L] = [z_e | z_c], kl [2] = (kl_e, kl_c), pre [N, 4L] =
    [mu_e | lv_e | mu_c | lv_c] (detached))."""
    return _LatentFn.apply(x, eps_e, eps_c, _lib.precision_code(precision), w_emu, b_emu, w_elv, b_elv, w_cmu, b_cmu,
                           w_clv, b_clv)


def pair_preds(x, eps_e, eps_c, w_emu, b_emu, w_elv, b_elv, w_cmu, b_cmu, w_clv, b_clv, pair_w, pair_b,
               precision: str | None = None) -> torch.Tensor:
    """Inference path of the head (reference get_pair_preds, drl_classifier_ec_mmd_final_mul.py:265-282):
    sigmoid(pair_classifier([z_e | z_c])) as [N, 1], four launches, no KL / no z; not differentiable (the reference
    calls it under torch.no_grad())."""
    wb = (w_emu, b_emu, w_elv, b_elv, w_cmu, b_cmu, w_clv, b_clv)
    _require_cuda(x, eps_e, eps_c, pair_w, pair_b, *wb)
    if pair_w.shape[0] != 1:
        raise NotImplementedError("pair_preds implements pair_num_class = 1 (the reference's setting)")
    xf = _rowmajor_f32(x.detach())
    N, D = xf.shape
    ws_ = [w.detach().float().contiguous() for w in wb[0::2]]
    bs_ = [b.detach().float().contiguous() for b in wb[1::2]]
    L = ws_[0].shape[0]
    pw, pb = pair_w.detach().float().contiguous().reshape(-1), pair_b.detach().float().contiguous().reshape(-1)
    ee, ec = eps_e.detach().float().contiguous(), eps_c.detach().float().contiguous()
    lb = lib()
    ws = _workspace(lb.carel_pair_preds_workspace_bytes(N, D, L), xf.device)
    out = torch.empty((N, 1), dtype=torch.float32, device=xf.device)
    prec = _lib.precision_code(precision)
    check(lb.carel_pair_preds(_p(xf), xf.stride(0), N, D, L, _ptr_array(ws_), _ptr_array(bs_), _p(ee), _p(ec), _p(pw), _p(pb),
                              _p(out), _p(ws), ws.numel(), 2 if prec == 2 else 0, _stream()), "carel_pair_preds")
    return out


# --------------------------------------------------------------------------------------------
# classifier heads: three tiny Linear + loss heads in two kernels
# --------------------------------------------------------------------------------------------
class _HeadsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, m_e, m_c, m_p, w_e, b_e, w_c, b_c, w_p, b_p, emo_labels, cau_labels, pair_labels, cfg):
        emo_mode, ls, n_e, n_c, n_p, guard, wts = cfg
        _require_cuda(z, w_e, w_c, w_p)
        zf = _rowmajor_f32(z)
        N, twoL = zf.shape
        L = twoL // 2
        C_e = w_e.shape[0]
        masks = [None if m is None else m.detach().float().contiguous() for m in (m_e, m_c, m_p)]
        weights = [t.detach().float().contiguous() for t in (w_e, b_e, w_c, b_c, w_p, b_p)]
        emo = (emo_labels.reshape(-1).long() if emo_mode == 0 else emo_labels.reshape(-1).float()).contiguous()
        cau = cau_labels.reshape(-1).float().contiguous()
        pair = pair_labels.reshape(-1).float().contiguous()
        lb = lib()
        ws = _workspace(lb.carel_heads_workspace_bytes(N, C_e), zf.device)
        losses = torch.empty(4, dtype=torch.float32, device=zf.device)
        wv = float_array(wts)
        check(lb.carel_heads_fwd(_p(zf), zf.stride(0), N, L, C_e, int(emo_mode), _ptr_array(masks), _ptr_array(weights),
                                 _p(emo), _p(cau), _p(pair), float(ls), float(n_e), float(n_c), float(n_p), int(guard),
                                 wv, _p(losses), _p(ws), ws.numel(), _stream()), "carel_heads_fwd")
        ctx.keep = (zf, masks, weights, emo, cau, pair, ws)
        ctx.meta = (N, L, C_e, int(emo_mode), float(ls), float(n_e), float(n_c), float(n_p), int(guard), tuple(wts),
                    z.dtype, [t.dtype for t in (w_e, b_e, w_c, b_c, w_p, b_p)],
                    [tuple(t.shape) for t in (w_e, b_e, w_c, b_c, w_p, b_p)])
        ctx.mark_non_differentiable(losses)
        return losses[3].clone(), losses

    @staticmethod
    def backward(ctx, g, _g_losses):
        zf, masks, weights, emo, cau, pair, ws = ctx.keep
        N, L, C_e, emo_mode, ls, n_e, n_c, n_p, guard, wts, zdt, wdts, wshapes = ctx.meta
        g = g.reshape(()).float().contiguous()
        need = ctx.needs_input_grad
        gz = torch.empty((N, 2 * L), dtype=torch.float32, device=zf.device)
        # the six weight gradients are carved back to back out of one buffer: the library zeroes them with one memset
        counts = [int(torch.Size(shape).numel()) if need[4 + i] else 0 for i, shape in enumerate(wshapes)]
        flat = torch.empty(sum(counts), dtype=torch.float32, device=zf.device)
        gws, off = [], 0
        for n, shape in zip(counts, wshapes):
            gws.append(flat[off:off + n].view(shape) if n else None)
            off += n
        check(lib().carel_heads_bwd(_p(g), _p(zf), zf.stride(0), N, L, C_e, emo_mode, _ptr_array(masks),
                                    _ptr_array(weights), _p(emo), _p(cau), _p(pair), ls, n_e, n_c, n_p, guard,
                                    float_array(wts), _p(gz), 2 * L, _ptr_array(gws), _p(ws), ws.numel(), _stream()),
              "carel_heads_bwd")
        out = [gz.to(zdt) if need[0] else None, None, None, None]
        out += [None if t is None else t.to(dt) for t, dt in zip(gws, wdts)]
        out += [None, None, None, None]
        return tuple(out)


def classifier_heads(z, masks, emotion_classifier, cause_classifier, pair_classifier, emotion_labels, cause_labels,
                     pair_labels, emo_mode: int, label_smoothing: float, n_e: float, n_c: float, n_p: float,
                     guard_inf: bool, loss_weights: Sequence[float]):
    """The three classifier heads + their weighted sum (reference get_emotion_mul_loss / get_cause_mul_loss /
    get_pair_mul_loss, drl_classifier_ec_mmd_final_mul.py:461-513, and the head terms of the loss assembly :256-261).
    `masks` = (m_e [N,L], m_c [N,L], m_p [N,2L]) scaled dropout masks or None.  Returns (weighted sum — differentiable
    w.r.t. z and the six parameters —, losses[4] = emotion, cause, pair, weighted sum (detached))."""
    m_e, m_c, m_p = masks if masks is not None else (None, None, None)
    cfg = (emo_mode, label_smoothing, n_e, n_c, n_p, bool(guard_inf), tuple(float(w) for w in loss_weights))
    return _HeadsFn.apply(z, m_e, m_c, m_p, emotion_classifier.weight, emotion_classifier.bias,
                          cause_classifier.weight, cause_classifier.bias, pair_classifier.weight, pair_classifier.bias,
                          emotion_labels, cause_labels, pair_labels, cfg)


# --------------------------------------------------------------------------------------------
# GAN discriminators: both adversaries, their BCE losses and the entropy terms in one kernel
# --------------------------------------------------------------------------------------------
class _DiscFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, m_ec, m_ce, w_ec, b_ec, w_ce, b_ce, y_emo, y_cau, cfg):
        ls, n_class, eps = cfg
        _require_cuda(z, w_ec, w_ce)
        if w_ec.shape[0] != 1 or w_ce.shape[0] != 1:
            raise NotImplementedError("the discriminator kernels implement ec_num_class = 1 (the reference's setting)")
        zf = _rowmajor_f32(z.detach())
        N, twoL = zf.shape
        L = twoL // 2
        masks = [None if m is None else m.detach().float().contiguous() for m in (m_ec, m_ce)]
        params = [t.detach().float().contiguous() for t in (w_ec, b_ec, w_ce, b_ce)]
        ye, yc = y_emo.reshape(-1).float().contiguous(), y_cau.reshape(-1).float().contiguous()
        lb = lib()
        ws = _workspace(lb.carel_disc_workspace_bytes(N), zf.device)
        out = torch.empty(3, dtype=torch.float32, device=zf.device)
        check(lb.carel_disc_fwd(_p(zf), zf.stride(0), N, L, _ptr_array(masks), _ptr_array(params), _p(ye), _p(yc), float(ls),
                                float(n_class), float(eps), _p(out), _p(ws), ws.numel(), _stream()), "carel_disc_fwd")
        ctx.keep = (zf, masks, params, ye, yc, ws)
        ctx.meta = (N, L, float(ls), float(n_class), float(eps), [t.dtype for t in (w_ec, b_ec, w_ce, b_ce)],
                    [tuple(t.shape) for t in (w_ec, b_ec, w_ce, b_ce)])
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_ec, g_ce, g_ent):
        zf, masks, params, ye, yc, ws = ctx.keep
        N, L, ls, n_class, eps, dts, shapes = ctx.meta
        gs = [None if g is None else g.reshape(()).float().contiguous() for g in (g_ec, g_ce, g_ent)]
        flat = torch.empty(2 * L + 2, dtype=torch.float32, device=zf.device)
        grads = [flat[:L], flat[L:L + 1], flat[L + 1:2 * L + 1], flat[2 * L + 1:]]
        check(lib().carel_disc_bwd(_ptr_array(gs), _p(zf), zf.stride(0), N, L, _ptr_array(masks), _ptr_array(params), _p(ye),
                                   _p(yc), ls, n_class, eps, _ptr_array(grads), _p(ws), ws.numel(), _stream()), "carel_disc_bwd")
        need = ctx.needs_input_grad
        outs = [g.reshape(shape).to(dt) if need[3 + i] else None for i, (g, shape, dt) in enumerate(zip(grads, shapes, dts))]
        return (None, None, None, *outs, None, None, None)


def gan_discriminators(z, m_ec, m_ce, ec_disc, ce_disc, emotion_labels, cause_labels, label_smoothing: float,
                       n_class: float, eps: float):
    """The adversarial branch of the GAN trainer (drl_classifier_ec_gan.py:222-233, :430-477) on the DETACHED latents:
    returns (ec_disc_loss, ce_disc_loss, ec_entropy + ce_entropy), differentiable w.r.t. the two discriminators'
    parameters only (z carries no gradient here, as in the reference).  m_ec / m_ce: scaled dropout masks [N, L] or None."""
    return _DiscFn.apply(z, m_ec, m_ce, ec_disc.weight, ec_disc.bias, ce_disc.weight, ce_disc.bias, emotion_labels,
                         cause_labels, (label_smoothing, n_class, eps))


# --------------------------------------------------------------------------------------------
# loss assembly: one launch for the weighted sum of the loss terms
# --------------------------------------------------------------------------------------------
class _WeightedSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, *tensors):
        _require_cuda(*tensors)
        flat = [t.detach().float().contiguous() for t in tensors]
        ptrs, wts = [], []
        for t, w in zip(flat, weights):
            if t.numel() != len(w):
                raise ValueError(f"weighted_sum: {t.numel()} elements but {len(w)} weights")
            for j, wj in enumerate(w):
                ptrs.append(t.data_ptr() + 4 * j)
                wts.append(float(wj))
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        out = torch.empty((), dtype=torch.float32, device=flat[0].device)
        check(lib().carel_wsum_fwd(arr, float_array(wts), len(ptrs), _p(out), _stream()), "carel_wsum_fwd")
        ctx.wts = wts
        ctx.meta = [(tuple(t.shape), t.dtype) for t in tensors]
        ctx.keep = flat            # the kernel reads them on the stream: keep them alive until it has run
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.reshape(()).float().contiguous()
        grads = torch.empty(len(ctx.wts), dtype=torch.float32, device=g.device)
        check(lib().carel_wsum_bwd(_p(g), float_array(ctx.wts), len(ctx.wts), _p(grads), _stream()), "carel_wsum_bwd")
        out, off = [None], 0
        for i, (shape, dt) in enumerate(ctx.meta):
            n = 1
            for d in shape:
                n *= d
            out.append(grads[off:off + n].view(shape).to(dt) if ctx.needs_input_grad[1 + i] else None)
            off += n
        return tuple(out)


def weighted_sum(tensors, weights) -> torch.Tensor:
    """sum_i w_i t_i over small device tensors in ONE launch (and one in the backward): the loss assembly of the
    trainers (drl_classifier_ec_mmd_final_mul.py:256-261 and its siblings).  `tensors[i]` is a 0-dim or short 1-D fp32
    tensor, `weights[i]` a float (0-dim) or a sequence with one float per element; at most 16 elements in total.
    Products and sums are rounded separately, left to right: bit-identical to the torch expression."""
    ws = tuple(tuple(float(x) for x in w) if isinstance(w, (tuple, list)) else (float(w),) for w in weights)
    return _WeightedSumFn.apply(ws, *tensors)


# --------------------------------------------------------------------------------------------
# CLUB terms of the VI variant (elementwise tails after the two small MLPs)
# --------------------------------------------------------------------------------------------
class _ClubFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu_hat, lv_hat, e, perm):
        _require_cuda(mu_hat, lv_hat, e)
        mf, lf = mu_hat.detach().float().contiguous(), lv_hat.detach().float().contiguous()
        ef = _rowmajor_f32(e.detach())
        N, L = mf.shape
        if lf.shape != (N, L) or ef.shape != (N, L):
            raise ValueError(f"club: shapes {tuple(mf.shape)}, {tuple(lf.shape)}, {tuple(ef.shape)} do not match")
        pf = None
        if perm is not None:
            pf = perm.to(device=mf.device, dtype=torch.int64).contiguous()
            if pf.shape != (N,):
                raise ValueError(f"club: perm must have shape ({N},), got {tuple(pf.shape)}")
        out = torch.empty((), dtype=torch.float32, device=mf.device)
        acc = torch.empty(1, dtype=torch.float64, device=mf.device)
        check(lib().carel_club_fwd(_p(mf), _p(lf), _p(ef), ef.stride(0), _p(pf), N, L, _p(out), _p(acc), _stream()),
              "carel_club_fwd")
        ctx.keep = (mf, lf, ef, pf)
        ctx.dts = (mu_hat.dtype, lv_hat.dtype, e.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        mf, lf, ef, pf = ctx.keep
        N, L = mf.shape
        g = g.reshape(()).float().contiguous()
        need = ctx.needs_input_grad
        gs = [torch.empty((N, L), dtype=torch.float32, device=mf.device) if need[i] else None for i in range(3)]
        check(lib().carel_club_bwd(_p(g), _p(mf), _p(lf), _p(ef), ef.stride(0), _p(pf), N, L, _p(gs[0]), _p(gs[1]),
                                   _p(gs[2]), _stream()), "carel_club_bwd")
        return tuple(None if t is None else t.to(dt) for t, dt in zip(gs, ctx.dts)) + (None,)


class _ClubMlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, c, *w8):
        _require_cuda(c, *w8)
        cf = _rowmajor_f32(c)
        N, L = cf.shape
        ws8 = [t.detach().float().contiguous() for t in w8]
        lb = lib()
        ws = _workspace(lb.carel_club_mlp_workspace_bytes(N, L), cf.device)
        mu = torch.empty((N, L), dtype=torch.float32, device=cf.device)
        lv = torch.empty((N, L), dtype=torch.float32, device=cf.device)
        check(lb.carel_club_mlp_fwd(_p(cf), cf.stride(0), N, L, _ptr_array(ws8), _p(mu), _p(lv), _p(ws), ws.numel(), _stream()),
              "carel_club_mlp_fwd")
        ctx.keep = (cf, ws8, ws, lv)
        ctx.meta = (N, L, c.dtype, [t.dtype for t in w8], [tuple(t.shape) for t in w8])
        return mu, lv

    @staticmethod
    def backward(ctx, dmu, dlv):
        cf, ws8, ws, lv = ctx.keep
        N, L, cdt, wdts, wshapes = ctx.meta
        dmu = None if dmu is None else dmu.float().contiguous()
        dlv = None if dlv is None else dlv.float().contiguous()
        need = ctx.needs_input_grad
        dc = torch.empty((N, L), dtype=torch.float32, device=cf.device) if need[0] else None
        # the eight parameter gradients are carved out of ONE zeroed buffer (the kernels accumulate into them)
        sizes = [int(torch.Size(shape).numel()) if need[1 + i] else 0 for i, shape in enumerate(wshapes)]
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=cf.device)
        g8, off = [], 0
        for n, shape in zip(sizes, wshapes):
            g8.append(flat[off:off + n].view(shape) if n else None)
            off += n
        check(lib().carel_club_mlp_bwd(_p(dmu), _p(dlv), _p(cf), cf.stride(0), _p(lv), N, L, _ptr_array(ws8), _p(dc), L,
                                       _ptr_array(g8), _p(ws), ws.numel(), _stream()), "carel_club_mlp_bwd")
        return (None if dc is None else dc.to(cdt), *[None if g is None else g.to(dt) for g, dt in zip(g8, wdts)])


def club_mlp(c_embedding, ec_mu, ec_log_var):
    """(mu_hat, lv_hat) = (ec_mu(c), ec_log_var(c)) of the VI trainer's approximation network (reference get_ec_emb,
    drl_classifier_ec_vi_final.py:342-347; ec_mu = Sequential(Linear, ReLU, Linear), ec_log_var = Sequential(Linear, ReLU,
    Linear, Tanh), :155-162): both MLPs in one kernel forward and one backward."""
    return _ClubMlpFn.apply(c_embedding, ec_mu[0].weight, ec_mu[0].bias, ec_mu[2].weight, ec_mu[2].bias,
                            ec_log_var[0].weight, ec_log_var[0].bias, ec_log_var[2].weight, ec_log_var[2].bias)


def club_aprx_loss(mu_hat, lv_hat, e_embedding) -> torch.Tensor:
    """-mean_i sum_d [-(mu - e)^2 / exp(lv) - lv]  (get_ec_aprx_loss, drl_classifier_ec_vi_final.py:421-426, after the MLPs)."""
    return _ClubFn.apply(mu_hat, lv_hat, e_embedding, None)


def club_upper_bound(mu_hat, lv_hat, e_embedding, perm) -> torch.Tensor:
    """mean_i [sum_d -(mu - e)^2/exp(lv) - sum_d -(mu - e[perm])^2/exp(lv)] / 2  (get_ec_upper_loss, :428-439)."""
    return _ClubFn.apply(mu_hat, lv_hat, e_embedding, perm)


# --------------------------------------------------------------------------------------------
# BoW reconstruction: decoder GEMM fused with softmax + BCE
# --------------------------------------------------------------------------------------------
DEFAULT_ELL_CAP = 64

#: Callables `hook(dec_w, grad_w) -> bool` run INSIDE the decoder's backward, after the kernels that complete the weight
#: gradient have been launched and before the latent-gradient kernel is: a data-parallel reducer starts the all-reduce
#: of grad_w there, underneath everything that follows (parallel.OverlappedGradReducer registers itself here).  A hook
#: that returns True has taken ownership of grad_w (it has made it `dec_w.grad` itself); autograd then gets None.
decoder_grad_hooks: list = []


class _BowFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, dec_w, dec_b, bow_dense, ell, label_smoothing, ell_cap, prec=0):
        _require_cuda(z, dec_w, dec_b)
        zf = _rowmajor_f32(z)
        wf = dec_w.detach().float().contiguous()
        bf_ = dec_b.detach().float().contiguous()
        N, K = zf.shape
        V = wf.shape[0]
        bd = None
        if bow_dense is not None:
            bd = _rowmajor_f32(bow_dense)
            if bd.shape != (N, V):
                raise ValueError(f"content_bow must be [{N}, {V}], got {tuple(bd.shape)}")
        cols = vals = nnz = None
        if ell is not None:
            cols, vals, nnz = ell
            cols, vals, nnz = cols.int().contiguous(), vals.float().contiguous(), nnz.int().contiguous()
            ell_cap = cols.shape[1]
        lb = lib()
        loss = torch.empty((), dtype=torch.float32, device=zf.device)
        if K > 256:
            prec = 2         # the tensor-core decoder holds 2 ec_dim <= 256 input columns; wider ones take the fp32 kernels
        if prec == 2:        # fp32 tier: materialised logits, fp32 GEMMs on the CUDA cores
            ws = _workspace(lb.carel_bow_f32_workspace_bytes(N, K, V, 1 if bd is None else 0), zf.device)
            check(lb.carel_bow_f32_fwd(_p(zf), zf.stride(0), _p(wf), _p(bf_), _p(bd), 0 if bd is None else bd.stride(0),
                                       _p(cols), _p(vals), _p(nnz), N, K, V, ell_cap, float(label_smoothing), _p(loss),
                                       _p(ws), ws.numel(), _stream()), "carel_bow_f32_fwd")
            ctx.f32_keep = (zf, wf)
        else:
            ws = _workspace(lb.carel_bow_workspace_bytes(N, K, V, ell_cap), zf.device)
            check(lb.carel_bow_fwd(_p(zf), zf.stride(0), _p(wf), _p(bf_), _p(bd), 0 if bd is None else bd.stride(0),
                                   _p(cols), _p(vals), _p(nnz), N, K, V, ell_cap, float(label_smoothing), _p(loss),
                                   _p(ws), ws.numel(), _stream()), "carel_bow_fwd")
            ctx.f32_keep = None
        ctx.ws, ctx.bd = ws, bd
        ctx.dec_w = dec_w
        ctx.meta = (N, K, V, int(ell_cap), float(label_smoothing), z.dtype, dec_w.dtype, dec_b.dtype)
        return loss

    @staticmethod
    def backward(ctx, g):
        N, K, V, cap, ls, dtz, dtw, dtb = ctx.meta
        ws, bd = ctx.ws, ctx.bd
        dev = ws.device
        g = g.reshape(()).float().contiguous()
        need = ctx.needs_input_grad
        gz = torch.empty((N, K), dtype=torch.float32, device=dev)
        gw = torch.empty((V, K), dtype=torch.float32, device=dev) if need[1] else None
        gb = torch.empty((V,), dtype=torch.float32, device=dev) if need[2] else None
        if ctx.f32_keep is not None:
            zf, wf = ctx.f32_keep
            check(lib().carel_bow_f32_bwd(_p(g), _p(zf), zf.stride(0), _p(wf), _p(bd), 0 if bd is None else bd.stride(0), N, K,
                                          V, ls, _p(gz), K, _p(gw), _p(gb), _p(ws), ws.numel(), _stream()), "carel_bow_f32_bwd")
        elif decoder_grad_hooks and gw is not None and dtw == torch.float32:
            # weight gradient first, hooks (the DP all-reduce forks here), then the latent gradient underneath it
            args = (_p(g), _p(bd), 0 if bd is None else bd.stride(0), N, K, V, cap, ls, _p(gz), K, _p(gw), _p(gb), _p(ws),
                    ws.numel())
            check(lib().carel_bow_bwd_parts(*args, 1, _stream()), "carel_bow_bwd_parts")
            taken = False
            for hook in list(decoder_grad_hooks):
                taken = bool(hook(ctx.dec_w, gw)) or taken
            check(lib().carel_bow_bwd_parts(*args, 2, _stream()), "carel_bow_bwd_parts")
            if taken:
                gw = None
        else:
            check(lib().carel_bow_bwd(_p(g), _p(bd), 0 if bd is None else bd.stride(0), N, K, V, cap, ls, _p(gz), K,
                                      _p(gw), _p(gb), _p(ws), ws.numel(), _stream()), "carel_bow_bwd")
        return (gz.to(dtz) if need[0] else None, None if gw is None else gw.to(dtw),
                None if gb is None else gb.to(dtb), None, None, None, None, None)


def bow_pack_ell(content_bow: torch.Tensor, ell_cap: int = DEFAULT_ELL_CAP):
    """Dense [N, V] BoW target -> ELL triple (cols int32 [N, C], vals fp32 [N, C], nnz int32 [N]) on the current
    stream; `nnz` may exceed C (the row then needs the dense target next to the list)."""
    _require_cuda(content_bow)
    bd = _rowmajor_f32(content_bow)
    N, V = bd.shape
    cols = torch.empty((N, ell_cap), dtype=torch.int32, device=bd.device)
    vals = torch.empty((N, ell_cap), dtype=torch.float32, device=bd.device)
    nnz = torch.empty((N,), dtype=torch.int32, device=bd.device)
    check(lib().carel_bow_pack_ell(_p(bd), bd.stride(0), N, V, int(ell_cap), _p(cols), _p(vals), _p(nnz), _stream()),
          "carel_bow_pack_ell")
    return cols, vals, nnz


def bow_reconstruction_loss(z, dec_w, dec_b, content_bow=None, label_smoothing: float = 0.1, ell=None,
                            ell_cap: int = DEFAULT_ELL_CAP, precision: str | None = None):
    """BCELoss(softmax(z dec_w^T + dec_b), content_bow (1 - ls) + ls / V) without materialising [N, V]
    (reference :253-254 + get_reconstruct_loss :381-387).  `content_bow` is the reference's dense
    [N, V] target; `ell=(cols[N,C] int32, vals[N,C] fp32, nnz[N] int32)` is the compact alternative."""
    return _BowFn.apply(z, dec_w, dec_b, content_bow, ell, label_smoothing, ell_cap, _lib.precision_code(precision))


# --------------------------------------------------------------------------------------------
# HSIC
# --------------------------------------------------------------------------------------------
class _HSICFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, s_x, s_y, tile_range, reduce_fn, prec=0):
        _require_cuda(x, y)
        xf, yf = _rowmajor_f32(x), _rowmajor_f32(y)
        m, dx = xf.shape
        m2, dy = yf.shape
        if m != m2:
            raise ValueError("HSIC needs paired samples (same number of rows)")
        lb = lib()
        ctx.f32_keep = None
        if prec == 2:        # fp32 tier: both kernel matrices materialised (no row-block sharding in this tier)
            if reduce_fn is not None or tuple(tile_range) != (0, -1):
                raise ValueError("the fp32 tier of HSIC evaluates the whole Gram on one GPU")
            ws = _workspace(lb.carel_hsic_f32_workspace_bytes(m, dx, dy), xf.device)
            out = torch.empty((), dtype=torch.float32, device=xf.device)
            check(lb.carel_hsic_f32_fwd(_p(xf), xf.stride(0), _p(yf), yf.stride(0), m, dx, dy, float(s_x), float(s_y), _p(out),
                                        _p(ws), ws.numel(), _stream()), "carel_hsic_f32_fwd")
            ctx.ws, ctx.f32_keep = ws, (xf, yf)
            ctx.meta = (m, dx, dy, float(s_x), float(s_y), 0, -1, x.dtype, y.dtype)
            return out
        ws = _workspace(lb.carel_hsic_workspace_bytes(m, dx, dy), xf.device)
        out = torch.empty((), dtype=torch.float32, device=xf.device)
        t0, t1 = tile_range
        sharded = reduce_fn is not None
        check(lb.carel_hsic_fwd(_p(xf), xf.stride(0), _p(yf), yf.stride(0), m, dx, dy, float(s_x), float(s_y), t0, t1,
                                _p(None if sharded else out), _p(ws), ws.numel(), _stream()), "carel_hsic_fwd")
        if sharded:
            offs = (C.c_int64 * 3)()
            lb.carel_hsic_ws_layout(m, dx, dy, offs)
            for off, n in ((offs[0], m), (offs[1], m), (offs[2], 4)):
                reduce_fn(ws[off:off + 8 * n].view(torch.float64))     # sum over ranks, in place
            check(lb.carel_hsic_finalize(m, dx, dy, _p(out), _p(ws), ws.numel(), _stream()), "carel_hsic_finalize")
        ctx.ws = ws
        ctx.meta = (m, dx, dy, float(s_x), float(s_y), t0, t1, x.dtype, y.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        m, dx, dy, s_x, s_y, t0, t1, dtx, dty = ctx.meta
        ws = ctx.ws
        g = g.reshape(()).float().contiguous()
        need = ctx.needs_input_grad
        gx = torch.empty((m, dx), dtype=torch.float32, device=ws.device) if need[0] else None
        gy = torch.empty((m, dy), dtype=torch.float32, device=ws.device) if need[1] else None
        if ctx.f32_keep is not None:
            xf, yf = ctx.f32_keep
            check(lib().carel_hsic_f32_bwd(_p(g), _p(xf), xf.stride(0), _p(yf), yf.stride(0), m, dx, dy, s_x, s_y, _p(gx), dx,
                                           _p(gy), dy, _p(ws), ws.numel(), _stream()), "carel_hsic_f32_bwd")
        else:
            check(lib().carel_hsic_bwd(_p(g), m, dx, dy, s_x, s_y, t0, t1, _p(gx), dx, _p(gy), dy, _p(ws), ws.numel(),
                                       _stream()), "carel_hsic_bwd")
        return (None if gx is None else gx.to(dtx), None if gy is None else gy.to(dty), None, None, None, None, None)


def hsic(x: torch.Tensor, y: torch.Tensor, s_x: float = 1.0, s_y: float = 1.0, tile_range=(0, -1),
         reduce_fn=None, precision: str | None = None) -> torch.Tensor:
    """HSIC(x, y) of the reference (hsic.py:540-547) as a 0-dim fp32 tensor, differentiable.
    `tile_range` / `reduce_fn` implement the multi-GPU row-block sharding (see parallel.sharded_hsic):
    the rank computes Gram row tiles [t0, t1) only, `reduce_fn(t)` sums a float64 tensor over the ranks in
    place; the gradient rows of the rank's block are complete (other rows are zero)."""
    return _HSICFn.apply(x, y, s_x, s_y, tuple(tile_range), reduce_fn, _lib.precision_code(precision))
