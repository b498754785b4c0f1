"""One training iteration of each reference trainer, as the reference's `train()` loops run it, for the drop-in heads.

The reference trainers keep their own `train()` (drl_classifier_ec_mmd_final_mul.py:802-922, drl_classifier_ec_hsic.py,
drl_classifier_ec_vi_final.py:730-800, drl_classifier_ec_gan.py:753-830) and work unchanged with
`carel_vae_b200.DrlClassifier*`; these functions are the loop BODY of those trainers (same calls, same order, same
optimiser choreography) for callers that drive the head directly — bench.py, the data-parallel wrapper in
`parallel.py` and the drop-in test in tests/test_gpu_trainer.py.

`RunningLoss` is the sync-free replacement of the `training_loss += loss.item()` bookkeeping (mmd:845-852, SURVEY
§8(f).2): the sum stays on the device and is read once per report instead of once per iteration.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch


class RunningLoss:
    """Device-side running sum of the per-iteration losses: `add()` is one tiny elementwise launch and never
    synchronises the host; `mean()` / `pop_mean()` read the value back (one sync) when a report is due — every 10
    mini-batches and at the end of an epoch in the reference (mmd:849-856)."""

    def __init__(self, device):
        self.total = torch.zeros((), dtype=torch.float64, device=device)
        self.count = 0

    def add(self, *losses: torch.Tensor) -> None:
        for t in losses:
            self.total += t.detach().double()
        self.count += 1

    def mean(self) -> float:
        return float(self.total) / max(self.count, 1)

    def pop_mean(self) -> float:
        m = self.mean()
        self.total.zero_()
        self.count = 0
        return m


def batch_to_device(batch: dict, device, emo_dtype=torch.long) -> tuple:
    """The `.to(device, dtype=...)` block at the top of every reference loop body (mmd:824-833)."""
    ids = batch["input_ids"].to(device, non_blocking=True)
    att = batch.get("attention_masks")
    tok = batch.get("token_type_ids")
    att = None if att is None else att.to(device, dtype=torch.long, non_blocking=True)
    tok = None if tok is None else tok.to(device, dtype=torch.long, non_blocking=True)
    labels = batch["labels"].to(device, dtype=torch.float, non_blocking=True)
    emo = batch["emo_labels"].to(device, dtype=emo_dtype, non_blocking=True)
    cau = batch["cau_labels"].to(device, dtype=torch.float, non_blocking=True)
    bow = batch["bow_reps"]
    if isinstance(bow, (tuple, list)):          # ELL triple from carel_vae_b200.data.ell_collate_fn
        bow = tuple(t.to(device, non_blocking=True) for t in bow)
    else:
        bow = bow.to(device, dtype=torch.float, non_blocking=True)
    return ids, att, tok, emo, cau, labels, bow


def train_step_single(model, optimizer, inputs: Sequence, iteration: int,
                      reduce_grads: Optional[Callable[[], None]] = None) -> torch.Tensor:
    """MMD and HSIC trainers (mmd:835-843, hsic same): one loss, one optimiser.
    inputs = (ids, att_masks, token_type_ids, emo_labels, cau_labels, labels, bow_reps)."""
    ids, att, tok, emo, cau, labels, bow = inputs
    loss = model(ids, att, tok, emo, cau, labels, bow, iteration)
    optimizer.zero_grad()
    loss.backward()
    if reduce_grads is not None:
        reduce_grads()
    optimizer.step()
    return loss


def train_step_vi(model, optimizers: Sequence, inputs: Sequence, iteration: int, epoch: int = 1,
                  reduce_grads: Optional[Callable[[int], None]] = None):
    """VI trainer (vi_final:756-781): the CLUB approximation net steps first on its own loss (retain_graph), then
    the upper bound is evaluated with the UPDATED net and added with beta = min(0.1 (epoch - 1), 1).
    Returns (ec_aprx_loss, vae_and_cls_loss).  `reduce_grads(k)` (data parallel) runs before optimiser k steps."""
    ec_aprx_opt, vae_and_cls_opt = optimizers
    ids, att, tok, emo, cau, labels, bow = inputs
    e_embedding, c_embedding, ec_aprx_loss, vae_and_cls_loss = model(ids, att, tok, emo, cau, labels, bow, iteration)
    ec_aprx_opt.zero_grad()
    ec_aprx_loss.backward(retain_graph=True)
    if reduce_grads is not None:
        reduce_grads(0)
    ec_aprx_opt.step()
    rj_loss = model.get_ec_upper_loss(e_embedding, c_embedding)
    beta = min((epoch - 1) * 0.1, 1.0)
    vae_and_cls_loss = vae_and_cls_loss + beta * rj_loss
    vae_and_cls_opt.zero_grad()
    vae_and_cls_loss.backward()
    if reduce_grads is not None:
        reduce_grads(1)
    vae_and_cls_opt.step()
    return ec_aprx_loss, vae_and_cls_loss


def train_step_gan(model, optimizers: Sequence, inputs: Sequence, iteration: int,
                   reduce_grads: Optional[Callable[[], None]] = None):
    """GAN trainer (gan:784-802): three backward passes, then the three optimisers step.
    Returns (ec_disc_loss, ce_disc_loss, vae_and_cls_loss)."""
    ec_disc_opt, ce_disc_opt, vae_and_cls_opt = optimizers
    ids, att, tok, emo, cau, labels, bow = inputs
    ec_disc_loss, ce_disc_loss, vae_and_cls_loss = model(ids, att, tok, emo, cau, labels, bow, iteration)
    ec_disc_opt.zero_grad()
    ec_disc_loss.backward(retain_graph=True)
    ce_disc_opt.zero_grad()
    ce_disc_loss.backward(retain_graph=True)
    vae_and_cls_opt.zero_grad()
    vae_and_cls_loss.backward()
    if reduce_grads is not None:
        reduce_grads()
    ec_disc_opt.step()
    ce_disc_opt.step()
    vae_and_cls_opt.step()
    return ec_disc_loss, ce_disc_loss, vae_and_cls_loss
