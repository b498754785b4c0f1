"""Drop-in `DrlClassifier` for the four disentanglement trainers.

Same constructor, method names, return types and state_dict keys as the reference classes
(drl_classifier_ec_mmd_final_mul.py:149-534, drl_classifier_ec_hsic.py, drl_classifier_ec_vi_final.py,
drl_classifier_ec_gan.py), so train() / get_pair_preds() / checkpoints are unchanged; everything
after `.pooler_output` runs on the sm_100a kernels of libcarel_head.so:

    latent projections + reparameterisation + KL   -> functional.latent   (carel_latent_fwd/bwd)
    MMD / HSIC independence losses                 -> losses.MMDStatistic / losses.HSIC
    BoW decoder + softmax + BCE                    -> functional.bow_reconstruction_loss

The three classifier heads + their weighted sum -> functional.classifier_heads
    (carel_heads_fwd/bwd; the per-head methods keep the reference's PyTorch form for external callers).  Hyper-parameters
are read from the `opt` passed to __init__ (the reference reads a module-global `opt`).

RNG order per forward matches the reference: randn(L) for emotion, randn(L) for
cause, then the dropout masks in call order.
"""
from __future__ import annotations

import contextlib
import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as TF

from . import functional as F
from .losses import HSIC, MMDStatistic


class IdentityEncoder(nn.Module):
    """Stand-in encoder for pre-computed clause embeddings: pooler_output = input_ids.
    (ECPE_dataset.py:137-154 produces exactly such [N, 768] fp32 embeddings.)"""

    def forward(self, input_ids, attention_mask=None, token_type_ids=None):
        return SimpleNamespace(pooler_output=input_ids)


def _build_encoder(opt):
    """The reference downloads hfl/chinese-roberta-wwm-ext or roberta-base (:158-163)."""
    from transformers import BertModel, RobertaModel  # imported lazily: only needed for real text
    if opt.language == "zh":
        return BertModel.from_pretrained("hfl/chinese-roberta-wwm-ext",
                                         cache_dir="model/hfl_chinese-roberta-wwm-ext", return_dict=True)
    return RobertaModel.from_pretrained("roberta-base", cache_dir="model/roberta-base", return_dict=True)


class _DrlClassifierBase(nn.Module):
    variant = "base"

    def __init__(self, opt, encoder: nn.Module | None = None):
        super().__init__()
        self.opt = opt
        self.encoder = encoder if encoder is not None else _build_encoder(opt)
        # emotion / cause latent embeddings (same member names => same state_dict keys)
        self.emotion_mu = nn.Linear(opt.bert_dim, opt.ec_dim)
        self.emotion_log_var = nn.Linear(opt.bert_dim, opt.ec_dim)
        self.cause_mu = nn.Linear(opt.bert_dim, opt.ec_dim)
        self.cause_log_var = nn.Linear(opt.bert_dim, opt.ec_dim)
        self._build_variant_modules(opt)
        self.emotion_classifier = nn.Linear(opt.ec_dim, self._emotion_classes(opt))
        self.cause_classifier = nn.Linear(opt.ec_dim, self._cause_classes(opt))
        self.pair_classifier = nn.Linear(opt.ec_dim * 2, opt.pair_num_class)
        self.decoder = nn.Linear(opt.ec_dim * 2, opt.pair_bow_dim)
        self.dropout = nn.Dropout(opt.dropout)
        #: capacity of the per-row non-zero list the dense BoW target is packed into on the GPU
        self.bow_ell_cap = getattr(opt, "bow_ell_cap", F.DEFAULT_ELL_CAP)

    # ---- variant hooks -------------------------------------------------------------------
    def _build_variant_modules(self, opt):
        pass

    # ---- stream fork/join for the tiny classifier heads --------------------------------------
    def _side(self, device):
        """The classifier heads are ~100 launch-latency-bound PyTorch kernels; they run on a side stream
        (forked after the latent layer, joined before the loss is assembled) so that they overlap the
        tensor-core kernels instead of serialising with them.  Host program order — and therefore the
        RNG consumption order of the dropout masks — is unchanged.  Autograd replays each backward op on
        the stream of its forward op, so the backward pass overlaps the same way; the fork/join is
        captured by CUDA graphs as parallel branches."""
        if not getattr(self.opt, "overlap_heads", True):
            return None
        streams = self.__dict__.setdefault("_side_streams", {})
        key = torch.device(device).index
        if key not in streams:
            streams[key] = torch.cuda.Stream(device=device)
        return streams[key]

    def _emotion_classes(self, opt):
        return opt.ec_num_class

    def _cause_classes(self, opt):
        return opt.ec_num_class

    def _label_classes(self):
        return self.opt.ec_num_class

    # ---- latent layer --------------------------------------------------------------------
    def _sentence_emb(self, input_ids, att_masks, token_type_ids):
        return self.encoder(input_ids, attention_mask=att_masks, token_type_ids=token_type_ids).pooler_output

    def _latents_and_heads(self, emb, emotion_labels, cause_labels, pair_labels):
        """Latent layer on the current stream and the classifier heads on the forked side stream (see _side).
        Returns (z, kl, heads, join): call join() before the heads' loss is consumed on the current stream.
        The three dropout masks depend on nothing but the generator: they are drawn right after the two noise draws
        (reference order: randn, randn, emotion / cause / pair masks) on the side stream, so they run underneath
        the latent layer's small kernels instead of in front of the heads."""
        L = self.opt.ec_dim
        side, main = self._side(emb.device), torch.cuda.current_stream()
        masks = []

        def draw_masks():
            if side is None:
                return
            n = emb.shape[0]
            side.wait_stream(main)
            with torch.cuda.stream(side):
                masks.extend((self._mask(n, L, emb.device), self._mask(n, L, emb.device), self._mask(n, 2 * L, emb.device)))

        z, kl, _ = self._latents(emb, after_noise=draw_masks)
        if side is not None:
            side.wait_stream(main)
        with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
            heads = self._heads(z, emotion_labels, cause_labels, pair_labels, masks=tuple(masks) if masks else None)

        def join():
            if side is not None:
                main.wait_stream(side)
        return z, kl, heads, join

    def _latents(self, sentence_emb, after_noise=None):
        """Fused path.  Returns (z [N, 2L], kl [2], pre [N, 4L]).  `after_noise()` runs between the two noise draws
        and the latent kernels: host-side RNG order is still randn, randn, <whatever it draws>, as in the reference."""
        L = self.opt.ec_dim
        eps_e = torch.randn(L, device=sentence_emb.device)      # same two draws, same order as
        eps_c = torch.randn(L, device=sentence_emb.device)      # sample_prior(emotion), sample_prior(cause)
        if after_noise is not None:
            after_noise()
        return F.latent(sentence_emb, eps_e, eps_c,
                        self.emotion_mu.weight, self.emotion_mu.bias,
                        self.emotion_log_var.weight, self.emotion_log_var.bias,
                        self.cause_mu.weight, self.cause_mu.bias,
                        self.cause_log_var.weight, self.cause_log_var.bias)

    def get_emotion_emb(self, sentence_emb):
        return self.emotion_mu(sentence_emb), self.emotion_log_var(sentence_emb)

    def get_cause_emb(self, sentence_emb):
        return self.cause_mu(sentence_emb), self.cause_log_var(sentence_emb)

    def sample_prior(self, mu, log_var):
        """One epsilon vector of shape [L] shared by every row, scaled by exp(log_var) (:345-351)."""
        epsilon = torch.randn(mu.size(1), device=mu.device)
        return mu + epsilon * torch.exp(log_var)

    def get_kl_loss(self, mu, log_var):
        return torch.mean(-0.5 * torch.sum(1 + log_var - log_var.exp() - mu.pow(2), dim=1))

    def get_annealed_weight(self, iteration, lambda_weight):
        T = self.opt.kl_ann_iterations
        return (math.tanh((iteration - T * 1.5) / (T / 3)) + 1) * lambda_weight

    def get_reconstruct_loss(self, reconstruct_pred, content_bow):
        """Reference signature (probabilities in): BCELoss against the smoothed target (:381-387).
        forward() does not come through here — it uses the fused decoder+loss kernel."""
        o = self.opt
        target = content_bow * (1 - o.label_smoothing) + o.label_smoothing / o.pair_bow_dim
        return TF.binary_cross_entropy(reconstruct_pred, target)

    def _assemble(self, leading, kl, iteration, reconstruction_loss):
        """leading terms + w KL_e + w KL_c + reconstruction in ONE launch (F.weighted_sum): `leading` = [(tensor,
        weight), ...] in the reference's order of addition; the value is bit-identical to the chain of scalar ops."""
        w = 1.0
        if iteration < self.opt.kl_ann_iterations:
            w = self.get_annealed_weight(iteration, self.opt.ec_kl_lambda)
        return F.weighted_sum([t for t, _ in leading] + [kl, reconstruction_loss],
                              [wt for _, wt in leading] + [(w, w), 1.0])

    def _reconstruction(self, z, content_bow):
        """`content_bow` is the reference's dense [N, V] fp32 target, or — SURVEY §8f.1 — the compact
        ELL triple (cols int32 [N, C], vals fp32 [N, C], nnz int32 [N]) of its non-zeros."""
        if isinstance(content_bow, (tuple, list)):
            return F.bow_reconstruction_loss(z, self.decoder.weight, self.decoder.bias, None,
                                             self.opt.label_smoothing, ell=tuple(content_bow))
        return F.bow_reconstruction_loss(z, self.decoder.weight, self.decoder.bias, content_bow,
                                         self.opt.label_smoothing, ell_cap=self.bow_ell_cap)

    # ---- small classifier heads (PyTorch) -------------------------------------------------
    def _drop(self, x):
        """nn.Dropout on a CONTIGUOUS tensor: the mask torch draws depends on the memory layout, and the
        reference's dropout inputs are contiguous (z_e / z_c are column views of one buffer here)."""
        if self.training and self.dropout.p > 0:
            return self.dropout(x.contiguous())
        return x

    def get_emotion_mul_loss(self, emotion_emb, emotion_labels):
        preds = torch.sigmoid(self.emotion_classifier(self._drop(emotion_emb)))
        o = self.opt
        t = emotion_labels * (1 - o.label_smoothing) + o.label_smoothing / self._label_classes()
        return TF.binary_cross_entropy(preds, t)

    def get_cause_mul_loss(self, cause_emb, cause_labels):
        preds = torch.sigmoid(self.cause_classifier(self._drop(cause_emb)))
        o = self.opt
        t = cause_labels * (1 - o.label_smoothing) + o.label_smoothing / self._cause_label_classes()
        return TF.binary_cross_entropy(preds, t)

    def _cause_label_classes(self):
        return self.opt.ec_num_class

    def _pair_label_classes(self):
        return self.opt.ec_num_class

    def get_pair_mul_loss(self, pair_emb, pair_labels):
        preds = self.pair_classifier(self._drop(pair_emb))
        o = self.opt
        t = pair_labels * (1 - o.label_smoothing) + o.label_smoothing / self._pair_label_classes()
        pos_weight = (len(pair_labels) - torch.sum(pair_labels)) / torch.sum(pair_labels)
        return TF.binary_cross_entropy_with_logits(preds, t, pos_weight=pos_weight)

    # ---- the three heads fused into two CUDA kernels ------------------------------
    _emo_mode = 1          # 0: cross-entropy over the emotion classes, 1: sigmoid + BCE (base / hsic / gan)
    _guard_inf = False     # reference mmd:510-511 only

    def _fused_heads_ok(self):
        """The head kernels cover every configuration the reference trainers use (ec_dim <= 128, one cause / pair
        logit, <= 8 emotion classes).  Anything else raises: there is no eager PyTorch fallback on the forward path
        (the per-head get_*_mul_loss methods stay available to outside callers under the reference's names)."""
        o = self.opt
        ok = (o.ec_dim <= 128 and self.cause_classifier.out_features == 1 and self.pair_classifier.out_features == 1 and
              (self.emotion_classifier.out_features == 1 if self._emo_mode == 1
               else self.emotion_classifier.out_features <= 8))
        if not ok:
            raise NotImplementedError(
                "carel_vae_b200 classifier-head kernels need ec_dim <= 128, one cause / pair logit and at most 8 emotion "
                f"classes (got ec_dim={o.ec_dim}, emotion classes={self.emotion_classifier.out_features}, cause "
                f"logits={self.cause_classifier.out_features}, pair logits={self.pair_classifier.out_features})")
        return True

    def _mask(self, n, k, device):
        """Scaled dropout mask with the RNG consumption of `self.dropout(x)` on a contiguous [n, k] tensor (the
        mask torch draws depends on shape and layout only), or None in eval mode / p = 0."""
        if self.training and self.dropout.p > 0:
            # the all-ones input is a constant: keep one per shape instead of filling [n, k] three times a step
            cache = self.__dict__.setdefault("_ones_cache", {})
            key = (n, k, str(device))
            if key not in cache:
                cache[key] = torch.ones((n, k), device=device)
            return self.dropout(cache[key])
        return None

    def _head_weights(self):
        o = self.opt
        return (o.ec_mul_loss_weight, o.ec_mul_loss_weight, o.pair_mul_loss_weight)

    def _heads(self, z, emotion_labels, cause_labels, pair_labels, masks=None):
        """w_e emo_mul_loss + w_c cau_mul_loss + w_p pair_mul_loss through carel_heads_fwd/bwd.  `masks` = pre-drawn (m_e, m_c, m_p) if the caller had to interleave other
        dropout draws (GAN); otherwise they are drawn here in the reference order emotion, cause, pair."""
        w = self._head_weights()
        L = self.opt.ec_dim
        self._fused_heads_ok()
        if masks is None:
            n = z.shape[0]
            masks = (self._mask(n, L, z.device), self._mask(n, L, z.device), self._mask(n, 2 * L, z.device))
        o = self.opt
        total, parts = F.classifier_heads(z, masks, self.emotion_classifier, self.cause_classifier, self.pair_classifier,
                                          emotion_labels, cause_labels, pair_labels, self._emo_mode, o.label_smoothing,
                                          self._label_classes() if self._emo_mode == 1 else 1,
                                          self._cause_label_classes(), self._pair_label_classes(),
                                          self._guard_inf, w)
        self.last_head_losses = parts          # (emotion, cause, pair, weighted sum), detached, on the device
        return total

    # ---- inference -----------------------------------------------------------------------
    def get_pair_preds_tensor(self, input_ids, att_masks, token_type_ids):
        """Device-side result of get_pair_preds (no host round trip): sigmoid(pair logits) [N, 1], through the
        dedicated inference entry (carel_pair_preds: projection GEMM + one fused reparameterisation / dot product /
        sigmoid pass; the same two noise draws as the training path)."""
        emb = self._sentence_emb(input_ids, att_masks, token_type_ids)
        L = self.opt.ec_dim
        eps_e = torch.randn(L, device=emb.device)
        eps_c = torch.randn(L, device=emb.device)
        return F.pair_preds(emb, eps_e, eps_c, self.emotion_mu.weight, self.emotion_mu.bias,
                            self.emotion_log_var.weight, self.emotion_log_var.bias, self.cause_mu.weight, self.cause_mu.bias,
                            self.cause_log_var.weight, self.cause_log_var.bias, self.pair_classifier.weight,
                            self.pair_classifier.bias)

    def get_pair_preds(self, input_ids, att_masks, token_type_ids):
        """Reference return type: rounded python list (:265-282).  Still stochastic at eval."""
        return self.get_pair_preds_tensor(input_ids, att_masks, token_type_ids).cpu().detach().numpy().round().tolist()

    def _core_params(self):
        return (list(self.encoder.parameters()) + list(self.decoder.parameters()) +
                list(self.emotion_classifier.parameters()) + list(self.cause_classifier.parameters()) +
                list(self.pair_classifier.parameters()))


class DrlClassifierMMD(_DrlClassifierBase):
    """drl_classifier_ec_mmd_final_mul.py (and its newsplit / emnlp siblings with --adapter false)."""
    variant = "mmd"
    _emo_mode = 0
    _guard_inf = True

    def _head_weights(self):
        o = self.opt
        return (o.emo_mul_loss_weight, o.cau_mul_loss_weight, o.pair_mul_loss_weight)

    def _emotion_classes(self, opt):
        return opt.e_num_class

    def _cause_classes(self, opt):
        return opt.c_num_class

    def _cause_label_classes(self):
        return self.opt.c_num_class

    def _pair_label_classes(self):
        return self.opt.pair_num_class

    def get_emotion_mul_loss(self, emotion_emb, emotion_labels):
        preds = self.emotion_classifier(self._drop(emotion_emb))          # 6-class CE (:461-476)
        return TF.cross_entropy(preds, emotion_labels.flatten())

    def get_pair_mul_loss(self, pair_emb, pair_labels):
        loss = super().get_pair_mul_loss(pair_emb, pair_labels)
        # reference :510-511 replaces an infinite loss by 0 after a host sync; same value, no sync
        return torch.where(torch.isinf(loss), torch.zeros_like(loss), loss)

    def forward(self, input_ids, att_masks, token_type_ids, emotion_labels, cause_labels, pair_labels, content_bow,
                iteration):
        o = self.opt
        emb = self._sentence_emb(input_ids, att_masks, token_type_ids)
        z, kl, heads, join = self._latents_and_heads(emb, emotion_labels, cause_labels, pair_labels)
        mmd = MMDStatistic(input_ids.shape[0], input_ids.shape[0]).halves(z, [0.1])
        reconstruction_loss = self._reconstruction(z, content_bow)
        join()
        # mmd_loss_weight * (-mmd) + heads + kl_e + kl_c + reconstruction_loss   (:256-261)
        return self._assemble([(mmd, -float(o.mmd_loss_weight)), (heads, 1.0)], kl, iteration, reconstruction_loss)

    def get_params(self):
        """The four latent projections are NOT handed to the optimiser, as in the reference (:284-297)."""
        return self._core_params()


class DrlClassifierHSIC(_DrlClassifierBase):
    """drl_classifier_ec_hsic.py."""
    variant = "hsic"

    def forward(self, input_ids, att_masks, token_type_ids, emotion_labels, cause_labels, pair_labels, content_bow,
                iteration):
        o, L = self.opt, self.opt.ec_dim
        emb = self._sentence_emb(input_ids, att_masks, token_type_ids)
        z, kl, heads, join = self._latents_and_heads(emb, emotion_labels, cause_labels, pair_labels)
        z_e, z_c = z[:, :L], z[:, L:]
        hsic_loss = HSIC(z_e, z_c)
        reconstruction_loss = self._reconstruction(z, content_bow)
        join()
        return self._assemble([(hsic_loss, 1.0), (heads, 1.0)], kl, iteration, reconstruction_loss)

    def get_params(self):
        return self._core_params()


class DrlClassifierVI(_DrlClassifierBase):
    """drl_classifier_ec_vi_final.py / drl_classifier_ec_vi.py: KL + CLUB upper bound."""
    variant = "vi"
    _emo_mode = 0

    def _build_variant_modules(self, opt):
        L = opt.ec_dim
        self.ec_mu = nn.Sequential(nn.Linear(L, L), nn.ReLU(), nn.Linear(L, L))
        self.ec_log_var = nn.Sequential(nn.Linear(L, L), nn.ReLU(), nn.Linear(L, L), nn.Tanh())

    def _emotion_classes(self, opt):
        return 6

    def get_emotion_mul_loss(self, emotion_emb, emotion_labels):
        preds = self.emotion_classifier(self._drop(emotion_emb))
        return TF.cross_entropy(preds, emotion_labels.flatten())

    def get_ec_emb(self, cause_emb):
        """(ec_mu(c), ec_log_var(c)) — vi_final:342-347 — through the fused MLP kernels (carel_club_mlp_fwd/bwd) when the
        latent width fits them (ec_dim <= 128, multiple of 4: every configuration of the reference trainers); the two
        nn.Sequential modules stay the owners of the parameters (state_dict keys unchanged)."""
        L = self.opt.ec_dim
        if L <= 128 and L % 4 == 0:
            return F.club_mlp(cause_emb, self.ec_mu, self.ec_log_var)
        raise NotImplementedError(f"the CLUB approximation-network kernels need ec_dim <= 128 and a multiple of 4 (got {L})")

    def get_ec_aprx_loss(self, e_embedding, c_embedding):
        ec_mu, ec_log_var = self.get_ec_emb(c_embedding.detach())
        return F.club_aprx_loss(ec_mu, ec_log_var, e_embedding)          # raises on non-CUDA tensors: no CPU path

    # ---- CLUB negative sampling ---------------------------------------------------------------
    def _negative_slots(self, n):
        """Two pinned host buffers per batch size: the permutation is drawn on the CPU generator exactly like the
        reference's `torch.randperm(N)` (vi_final:433) and reaches the GPU by an ASYNC copy from pinned memory — no
        host sync, and under CUDA-graph capture the copy becomes a graph node that re-reads slot 0 on every replay."""
        slots = self.__dict__.setdefault("_perm_slots", {})
        if n not in slots:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("run one eager step (or redraw_negative_index) before capturing the VI step")
            slots[n] = {"host": [torch.empty(n, dtype=torch.long).pin_memory() for _ in range(2)],
                        "busy": [None, None], "next": 0}
        return slots[n]

    def redraw_negative_index(self, n):
        """For CUDA-graph users: draw the next permutation into the slot the captured copy reads (call between replays,
        once the previous replay's copy has run)."""
        torch.randperm(n, out=self._negative_slots(n)["host"][0])

    def _negative_index(self, n, device):
        s = self._negative_slots(n)
        if torch.cuda.is_current_stream_capturing():
            torch.randperm(n, out=s["host"][0])
            return s["host"][0].to(device, non_blocking=True)
        k = s["next"]
        s["next"] ^= 1
        if s["busy"][k] is not None:
            s["busy"][k].synchronize()              # the copy issued two calls ago has long finished
        torch.randperm(n, out=s["host"][k])
        idx = s["host"][k].to(device, non_blocking=True)
        s["busy"][k] = torch.cuda.Event()
        s["busy"][k].record()
        return idx

    def get_ec_upper_loss(self, e_embedding, c_embedding):
        ec_mu, ec_log_var = self.get_ec_emb(c_embedding)
        F._require_cuda(e_embedding)                                      # no CPU path
        random_index = self._negative_index(e_embedding.shape[0], e_embedding.device)   # CPU generator, as the reference
        return F.club_upper_bound(ec_mu, ec_log_var, e_embedding, random_index)

    def forward(self, input_ids, att_masks, token_type_ids, emotion_labels, cause_labels, pair_labels, content_bow,
                iteration):
        o, L = self.opt, self.opt.ec_dim
        emb = self._sentence_emb(input_ids, att_masks, token_type_ids)
        z, kl, heads, join = self._latents_and_heads(emb, emotion_labels, cause_labels, pair_labels)
        z_e, z_c = z[:, :L], z[:, L:]
        ec_aprx_loss = self.get_ec_aprx_loss(z_e, z_c)
        reconstruction_loss = self._reconstruction(z, content_bow)
        join()
        loss = self._assemble([(heads, 1.0)], kl, iteration, reconstruction_loss)
        return z_e, z_c, ec_aprx_loss, loss

    def get_params(self):
        ec_aprx_params = list(self.ec_mu.parameters()) + list(self.ec_log_var.parameters())
        return ec_aprx_params, self._core_params()


class DrlClassifierGAN(_DrlClassifierBase):
    """drl_classifier_ec_gan.py: adversarial discriminators on the detached latents + entropy terms (forward() runs them
    through carel_disc_fwd/bwd; the get_*_disc_* methods keep the reference's PyTorch form for outside callers)."""
    variant = "gan"

    def _build_variant_modules(self, opt):
        self.ec_disc = nn.Linear(opt.ec_dim, opt.ec_num_class)
        self.ce_disc = nn.Linear(opt.ec_dim, opt.ec_num_class)

    def get_ec_disc_preds(self, cause_emb):
        return torch.sigmoid(self.ec_disc(self._drop(cause_emb.detach())))

    def get_ce_disc_preds(self, emotion_emb):
        return torch.sigmoid(self.ce_disc(self._drop(emotion_emb.detach())))

    def get_ec_disc_loss(self, ec_disc_preds, ec_labels):
        o = self.opt
        t = ec_labels * (1 - o.label_smoothing) + o.label_smoothing / o.ec_num_class
        return TF.binary_cross_entropy(ec_disc_preds, t)

    def get_entropy_loss(self, preds):
        return torch.mean(torch.sum(preds * torch.log(preds + self.opt.epsilon), dim=1))

    def forward(self, input_ids, att_masks, token_type_ids, emotion_labels, cause_labels, pair_labels, content_bow,
                iteration):
        o, L = self.opt, self.opt.ec_dim
        z, kl, _ = self._latents(self._sentence_emb(input_ids, att_masks, token_type_ids))
        z_e, z_c = z[:, :L], z[:, L:]
        n = z.shape[0]
        self._fused_heads_ok()
        # dropout draws in the reference's order: ec_disc, emotion head, ce_disc, cause head, pair head
        m_ec = self._mask(n, L, z.device)
        m_e = self._mask(n, L, z.device)
        m_ce = self._mask(n, L, z.device)
        m_c = self._mask(n, L, z.device)
        m_p = self._mask(n, 2 * L, z.device)
        # both discriminators, their BCE losses and the entropy terms: one kernel (carel_disc_fwd)
        ec_disc_loss, ce_disc_loss, entropy = F.gan_discriminators(z, m_ec, m_ce, self.ec_disc, self.ce_disc, emotion_labels,
                                                                   cause_labels, o.label_smoothing, o.ec_num_class, o.epsilon)
        heads = self._heads(z, emotion_labels, cause_labels, pair_labels, masks=(m_e, m_c, m_p))
        reconstruction_loss = self._reconstruction(z, content_bow)
        loss = self._assemble([(entropy, float(o.ecce_adv_loss_weight)), (heads, 1.0)], kl, iteration, reconstruction_loss)
        return ec_disc_loss, ce_disc_loss, loss

    def get_params(self):
        return self.ec_disc.parameters(), self.ce_disc.parameters(), self._core_params()


VARIANTS = {"mmd": DrlClassifierMMD, "hsic": DrlClassifierHSIC, "vi": DrlClassifierVI, "gan": DrlClassifierGAN}

#: the canonical trainer's class under the reference's own name
DrlClassifier = DrlClassifierMMD


def default_opt(variant: str = "mmd", **kw) -> SimpleNamespace:
    """argparse defaults of the reference trainers (mmd:30-58, hsic:30-57, vi_final:30-56, gan:30-57)."""
    o = SimpleNamespace(bert_dim=768, ec_dim=24, pair_bow_dim=23771, kl_ann_iterations=20000, ec_kl_lambda=0.03,
                        label_smoothing=0.1, dropout=0.5, epsilon=1e-8, pair_num_class=1, language="en")
    if variant == "mmd":
        o.e_num_class, o.c_num_class = 6, 1
        o.mmd_loss_weight, o.emo_mul_loss_weight, o.cau_mul_loss_weight, o.pair_mul_loss_weight = 30, 10, 10, 30
    else:
        o.ec_num_class, o.ec_mul_loss_weight = 1, 10
        o.pair_mul_loss_weight = 30 if variant == "vi" else 25
        o.ecce_adv_loss_weight = 1 if variant == "gan" else 3
    for k, v in kw.items():
        setattr(o, k, v)
    return o
