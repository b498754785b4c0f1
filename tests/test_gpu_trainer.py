"""GPU (-m gpu): the drop-in heads inside the reference trainers' loops.

/root/reference is not present on the GPU box and its sources may not be copied, so the loops below RESTATE the body
of the reference `train()` functions (drl_classifier_ec_mmd_final_mul.py:802-856, drl_classifier_ec_vi_final.py:730-800,
drl_classifier_ec_gan.py:753-820) line for line — DataLoader batch dict -> `.to(device, dtype=...)` -> model(...) ->
`with torch.autograd.set_detect_anomaly(True)` -> zero_grad / backward / step -> `.item()` bookkeeping — around
`carel_vae_b200.DrlClassifier*` and `carel_vae_b200.data.ECPEDataset`, with an identity encoder standing in for the
sentence transformer (pre-computed clause embeddings).  They check that the unchanged loop runs (no anomaly, finite,
loss moves), that the library's own loop bodies (carel_vae_b200/train.py) are the same computation, and that the ELL
collate path is reachable from a DataLoader and gives the dense path's loss.
"""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import carel_vae_b200 as cv
    from carel_vae_b200 import train as T
DEV = "cuda"


def _frame(n, vocab_size, seed):
    import pandas as pd
    rng = np.random.default_rng(seed)
    pairs = [" ".join(f"w{int(j)}" for j in rng.integers(0, vocab_size + 50, size=rng.integers(3, 30))) for _ in range(n)]
    return pd.DataFrame({"pair": pairs, "label": rng.integers(0, 2, n), "emotion": rng.integers(0, 6, n)})


def _setup(variant, n=192, L=24, V=400, dense_bow=True, seed=0):
    opt = cv.default_opt(variant, ec_dim=L, pair_bow_dim=V, bert_dim=768, dropout=0.5)
    torch.manual_seed(11)
    model = cv.VARIANTS[variant](opt, encoder=cv.IdentityEncoder()).to(DEV)
    vocab = [f"w{i}" for i in range(V)]
    df = _frame(n, V, seed)
    if variant in ("hsic", "gan"):
        df["emotion"] = 1                         # hsic:78 / gan:83: those trainers feed all-ones emotion labels
    g = torch.Generator().manual_seed(seed)
    emb = torch.tanh(torch.randn(n, 768, generator=g))
    ds = cv.ECPEDataset(df, tokenizer=None, bow_features=vocab, embeddings=emb, dense_bow=dense_bow)
    return opt, model, ds


def test_reference_mmd_train_loop_runs_unchanged():
    """mmd:802-856 loop body, verbatim structure."""
    opt, model, ds = _setup("mmd")
    train_loader = DataLoader(ds, batch_size=64, shuffle=False)
    vae_and_cls_opt = torch.optim.Adam(model.get_params(), lr=1e-3)
    device = torch.device(DEV)
    epoch_losses = []
    for epoch in range(1, 4):
        training_loss = 0
        running_loss = 0
        model.train()
        for iteration, batch in enumerate(train_loader):
            ids = batch['input_ids'].to(device)                       # (embeddings: the identity encoder's input)
            att_masks = None
            token_type_ids = None
            labels = batch['labels'].to(device, dtype=torch.float)
            emo_labels = batch['emo_labels'].to(device, dtype=torch.long)
            cau_labels = batch['cau_labels'].to(device, dtype=torch.float)
            bow_reps = batch['bow_reps'].to(device, dtype=torch.float)
            with torch.autograd.set_detect_anomaly(True):
                vae_and_cls_loss = model(ids, att_masks, token_type_ids, emo_labels, cau_labels, labels, bow_reps, iteration)
                vae_and_cls_opt.zero_grad()
                vae_and_cls_loss.backward()
                vae_and_cls_opt.step()
            training_loss += vae_and_cls_loss.item()
            running_loss += vae_and_cls_loss.item()
        epoch_losses.append(training_loss / len(train_loader))
        model.eval()
        with torch.no_grad():
            preds = []
            for batch in train_loader:
                preds += model.get_pair_preds(batch['input_ids'].to(device), None, None)
        assert len(preds) == len(ds) and all(p in ([0.0], [1.0]) for p in preds)
    assert all(np.isfinite(epoch_losses)) and epoch_losses[-1] < epoch_losses[0]


def test_library_loop_body_equals_the_restated_loop():
    """carel_vae_b200.train.train_step_single is the same computation as the loop body above (same seeds -> same loss
    trajectory bit for bit), and RunningLoss reports the mean without per-iteration .item() syncs."""
    traj = []
    for use_lib in (False, True):
        opt, model, ds = _setup("mmd", seed=3)
        loader = DataLoader(ds, batch_size=64, shuffle=False)
        optim = torch.optim.Adam(model.get_params(), lr=1e-3)
        model.train()
        torch.manual_seed(5)
        run = T.RunningLoss(DEV)
        losses = []
        for iteration, batch in enumerate(loader):
            inputs = T.batch_to_device(batch, DEV)
            if use_lib:
                loss = T.train_step_single(model, optim, inputs, iteration)
            else:
                ids, att, tok, emo, cau, labels, bow = inputs
                loss = model(ids, att, tok, emo, cau, labels, bow, iteration)
                optim.zero_grad()
                loss.backward()
                optim.step()
            run.add(loss)
            losses.append(loss.item())
        assert abs(run.mean() - float(np.mean(losses))) < 1e-6 * abs(np.mean(losses))
        traj.append(losses)
    assert traj[0] == traj[1]


def test_reference_vi_and_gan_train_loops_run_unchanged():
    """vi_final:756-781 (two optimisers, aprx step first, upper bound with the updated net) and gan:784-802 (three
    backward passes, RMSprop x2 + Adam)."""
    device = torch.device(DEV)
    # ---- VI ----
    opt, model, ds = _setup("vi")
    loader = DataLoader(ds, batch_size=64, shuffle=False)
    aprx_params, core_params = model.get_params()
    ec_aprx_opt = torch.optim.Adam(aprx_params, lr=1e-3)
    vae_and_cls_opt = torch.optim.Adam(core_params, lr=1e-3)
    model.train()
    first = last = None
    for epoch in range(1, 4):
        for iteration, batch in enumerate(loader):
            ids = batch['input_ids'].to(device)
            labels = batch['labels'].to(device, dtype=torch.float)
            emo_labels = batch['emo_labels'].to(device, dtype=torch.long)
            cau_labels = batch['cau_labels'].to(device, dtype=torch.float)
            bow_reps = batch['bow_reps'].to(device, dtype=torch.float)
            e_embedding, c_embedding, ec_aprx_loss, vae_and_cls_loss = model(ids, None, None, emo_labels, cau_labels,
                                                                             labels, bow_reps, iteration)
            with torch.autograd.set_detect_anomaly(True):
                ec_aprx_opt.zero_grad()
                ec_aprx_loss.backward(retain_graph=True)
                ec_aprx_opt.step()
                Rj_loss = model.get_ec_upper_loss(e_embedding, c_embedding)
                beta = (epoch - 1) * 0.1
                if beta > 1:
                    beta = 1
                vae_and_cls_loss += beta * Rj_loss
                vae_and_cls_opt.zero_grad()
                vae_and_cls_loss.backward()
                vae_and_cls_opt.step()
            total = ec_aprx_loss.item() + vae_and_cls_loss.item()
            first = total if first is None else first
            last = total
    assert np.isfinite(first) and np.isfinite(last) and last < first
    # ---- GAN ----
    opt, model, ds = _setup("gan")
    loader = DataLoader(ds, batch_size=64, shuffle=False)
    p1, p2, p3 = model.get_params()
    ec_disc_opt = torch.optim.RMSprop(p1, lr=1e-3)
    ce_disc_opt = torch.optim.RMSprop(p2, lr=1e-3)
    vae_and_cls_opt = torch.optim.Adam(p3, lr=1e-3)
    model.train()
    first = last = None
    for epoch in range(1, 4):
        for iteration, batch in enumerate(loader):
            ids = batch['input_ids'].to(device)
            labels = batch['labels'].to(device, dtype=torch.float)
            emo_labels = batch['emo_labels'].to(device, dtype=torch.float)
            cau_labels = batch['cau_labels'].to(device, dtype=torch.float)
            bow_reps = batch['bow_reps'].to(device, dtype=torch.float)
            ec_disc_loss, ce_disc_loss, vae_and_cls_loss = model(ids, None, None, emo_labels, cau_labels, labels, bow_reps,
                                                                 iteration)
            with torch.autograd.set_detect_anomaly(True):
                ec_disc_opt.zero_grad()
                ec_disc_loss.backward(retain_graph=True)
                ce_disc_opt.zero_grad()
                ce_disc_loss.backward(retain_graph=True)
                vae_and_cls_opt.zero_grad()
                vae_and_cls_loss.backward()
                ec_disc_opt.step()
                ce_disc_opt.step()
                vae_and_cls_opt.step()
            total = ec_disc_loss.item() + ce_disc_loss.item() + vae_and_cls_loss.item()
            first = total if first is None else first
            last = total
    assert np.isfinite(first) and np.isfinite(last) and last < first


def test_ell_collate_path_gives_the_dense_loss():
    """DataLoader(ECPEDataset(dense_bow=False), collate_fn=ell_collate_fn()) -> forward(content_bow=(cols, vals, nnz)):
    same loss and gradients as the reference's dense [B, V] rows (the kernels see the same non-zeros either way)."""
    out = []
    for dense in (True, False):
        opt, model, ds = _setup("mmd", dense_bow=dense, seed=7)
        loader = DataLoader(ds, batch_size=96, shuffle=False, collate_fn=None if dense else cv.ell_collate_fn(cap=64))
        model.train()
        batch = next(iter(loader))
        inputs = T.batch_to_device(batch, DEV)
        torch.manual_seed(9)
        ids, att, tok, emo, cau, labels, bow = inputs
        loss = model(ids, att, tok, emo, cau, labels, bow, 0)
        loss.backward()
        out.append((loss.item(), model.decoder.weight.grad.clone(), model.emotion_mu.weight.grad.clone()))
    assert abs(out[0][0] - out[1][0]) <= 1e-6 * abs(out[0][0])
    for a, b in zip(out[0][1:], out[1][1:]):
        assert (a - b).norm() <= 1e-5 * a.norm()
