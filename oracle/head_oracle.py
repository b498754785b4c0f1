"""TEST INFRASTRUCTURE — not product code.

Plain-PyTorch restatement of the CAREL-VAE disentanglement head (the reference's algorithm for
the hot path), used ONLY as the checker by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under carel_vae_b200/ imports this module.

Every function cites the reference lines it restates (paths relative to /root/reference):
  mmd    = drl_classifier_ec_mmd_final_mul.py
  hsic   = drl_classifier_ec_hsic.py
  vi     = drl_classifier_ec_vi_final.py
  gan    = drl_classifier_ec_gan.py

Parity pin: the reference ships no tests / golden vectors, so this restatement is
pinned against the reference ITSELF: oracle/make_golden.py AST-extracts the reference classes in
the build container, runs them on seeded inputs and commits the results under tests/golden/;
tests/test_oracle_golden.py checks this file against those vectors.

Works in any dtype/device torch supports (fp32 like the reference, fp64 for tolerance accounting).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# L0 loss primitives
# --------------------------------------------------------------------------------------------
def pdist(sample_1: torch.Tensor, sample_2: torch.Tensor, norm: float = 2, eps: float = 1e-5):
    """mmd:580-596.  norm==2: sqrt(eps + | |a|^2 + |b|^2 - 2 a.b |); else generic p-norm."""
    if float(norm) == 2.0:
        n1 = (sample_1 * sample_1).sum(dim=1, keepdim=True)
        n2 = (sample_2 * sample_2).sum(dim=1, keepdim=True)
        d2 = n1 + n2.t() - 2.0 * sample_1 @ sample_2.t()
        return torch.sqrt(eps + d2.abs())
    diff = (sample_1[:, None, :] - sample_2[None, :, :]).abs() ** float(norm)
    return (eps + diff.sum(dim=2)) ** (1.0 / float(norm))


def mmd_coefficients(n_1: int, n_2: int):
    """mmd:538-545 (raises ZeroDivisionError for n == 1 exactly like the reference)."""
    return 1.0 / (n_1 * (n_1 - 1)), 1.0 / (n_2 * (n_2 - 1)), -1.0 / (n_1 * n_2)


def mmd_statistic(sample_1, sample_2, alphas: Sequence[float], ret_matrix: bool = False):
    """mmd:547-569.  Unbiased MMD^2 with the kernel sum_a exp(-a * (|d2| + 1e-5))."""
    n_1, n_2 = sample_1.shape[0], sample_2.shape[0]
    a00, a11, a01 = mmd_coefficients(n_1, n_2)
    s = torch.cat((sample_1, sample_2), 0)
    dist = pdist(s, s, norm=2)
    kernels = None
    for alpha in alphas:
        k = torch.exp(-alpha * dist ** 2)
        kernels = k if kernels is None else kernels + k
    k_1, k_2, k_12 = kernels[:n_1, :n_1], kernels[n_1:, n_1:], kernels[:n_1, n_1:]
    mmd = 2 * a01 * k_12.sum() + a00 * (k_1.sum() - torch.trace(k_1)) + a11 * (k_2.sum() - torch.trace(k_2))
    return (mmd, kernels) if ret_matrix else mmd


def mmd_blocked(sample_1, sample_2, alphas: Sequence[float], eps: float = 1e-5, block: int = 2048,
                need_grad: bool = True):
    """Row-block streaming evaluation of the closed forms of SURVEY.md §8(a) a7 (value AND gradient)
    that never forms the [M, M] matrix: the oracle for sizes the literal form cannot hold.
    Returns (mmd, grad_1, grad_2) in the dtype of the inputs (use fp64)."""
    n_1, n_2 = sample_1.shape[0], sample_2.shape[0]
    a00, a11, a01 = mmd_coefficients(n_1, n_2)
    s = torch.cat((sample_1, sample_2), 0)
    M = s.shape[0]
    nrm = (s * s).sum(1)
    is_x = torch.arange(M, device=s.device) < n_1
    total = s.new_zeros(())
    grad = torch.zeros_like(s)
    for r0 in range(0, M, block):
        r1 = min(M, r0 + block)
        d2 = nrm[r0:r1, None] + nrm[None, :] - 2.0 * s[r0:r1] @ s.t()
        a = d2.abs() + eps
        rx, cx = is_x[r0:r1, None], is_x[None, :]
        C = torch.where(rx & cx, a00, 0.0) + torch.where(~rx & ~cx, a11, 0.0) + torch.where(rx != cx, a01, 0.0)
        C = C.to(s.dtype)
        idx = torch.arange(r0, r1, device=s.device)
        C[idx - r0, idx] = 0.0
        K = sum(torch.exp(-al * a) for al in alphas)
        total = total + (C * K).sum()
        if need_grad:
            Kp = sum(-al * torch.exp(-al * a) for al in alphas) * torch.sign(d2)
            W = C * Kp
            grad[r0:r1] = 4.0 * (W.sum(1, keepdim=True) * s[r0:r1] - W @ s)
    return total, grad[:n_1], grad[n_1:]


def pairwise_distances(x):
    """hsic:529-532 (no abs, no eps)."""
    n = (x * x).sum(-1).reshape(-1, 1)
    return -2.0 * x @ x.t() + n + n.t()


def gaussian_kernel_matrix(x, sigma: float = 1.0):
    """hsic:535-537."""
    return torch.exp(-pairwise_distances(x) / sigma)


def hsic_trace(x, y, s_x: float = 1.0, s_y: float = 1.0):
    """hsic:540-547, literally: trace(L H K H) / (m-1)^2 with three dense m x m products
    (without the hard-coded .cuda() / .double().float() round trip, which is an identity on values)."""
    m = x.shape[0]
    K = gaussian_kernel_matrix(x, s_x)
    L = gaussian_kernel_matrix(y, s_y)
    H = torch.eye(m, dtype=x.dtype, device=x.device) - 1.0 / m * torch.ones((m, m), dtype=x.dtype, device=x.device)
    return torch.trace(L @ (H @ (K @ H))) / ((m - 1) ** 2)


def hsic_fast(x, y, s_x: float = 1.0, s_y: float = 1.0):
    """O(m^2) identity of SURVEY.md §8(a) a9:
    [sum K.L - (2/m) sum_i rK_i rL_i + (sum K)(sum L)/m^2] / (m-1)^2."""
    m = x.shape[0]
    K = gaussian_kernel_matrix(x, s_x)
    L = gaussian_kernel_matrix(y, s_y)
    rK, rL = K.sum(1), L.sum(1)
    return ((K * L).sum() - (2.0 / m) * (rK * rL).sum() + K.sum() * L.sum() / (m * m)) / ((m - 1) ** 2)


def sample_prior(mu, log_var, epsilon: Optional[torch.Tensor] = None):
    """mmd:345-351.  ONE epsilon vector of shape [L] shared by every row, scaled by exp(log_var)
    (not exp(0.5 log_var)).  epsilon=None draws it exactly like the reference."""
    if epsilon is None:
        epsilon = torch.randn(mu.size(1), device=mu.device)
    return mu + epsilon.to(mu.dtype) * torch.exp(log_var)


def kl_loss(mu, log_var):
    """mmd:525-534."""
    return torch.mean(-0.5 * torch.sum(1 + log_var - log_var.exp() - mu.pow(2), dim=1))


def annealed_weight(iteration: int, lambda_weight: float, kl_ann_iterations: int = 20000):
    """mmd:515-523."""
    return (math.tanh((iteration - kl_ann_iterations * 1.5) / (kl_ann_iterations / 3)) + 1) * lambda_weight


def reconstruct_loss(reconstruct_pred, content_bow, label_smoothing: float = 0.1):
    """mmd:381-387.  BCELoss(softmax probabilities, smoothed BoW), mean over N*V, logs clamped at -100."""
    V = content_bow.shape[1]
    target = content_bow * (1 - label_smoothing) + label_smoothing / V
    return F.binary_cross_entropy(reconstruct_pred, target)


def bow_reconstruction(z_cat, dec_w, dec_b, content_bow, label_smoothing: float = 0.1):
    """mmd:253-254: softmax(decoder(z)) then reconstruct_loss."""
    return reconstruct_loss(torch.softmax(F.linear(z_cat, dec_w, dec_b), dim=1), content_bow, label_smoothing)


# --------------------------------------------------------------------------------------------
# L1 head: functional forward of the four trainer variants over a reference-layout state_dict
# --------------------------------------------------------------------------------------------
def default_opt(variant: str = "mmd", **kw) -> SimpleNamespace:
    """Head hyper-parameters with the reference defaults (mmd:30-58, hsic:30-57, vi:30-56, gan:30-57)."""
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


def init_state(opt, variant: str = "mmd", seed: int = 1234, device="cpu") -> Dict[str, torch.Tensor]:
    """Head parameters with nn.Linear's default init, in the reference's module order and
    state_dict key names (mmd:166-179; vi:155-179; gan:160-176)."""
    import torch.nn as nn
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    L, D, V = opt.ec_dim, opt.bert_dim, opt.pair_bow_dim
    mods = [("emotion_mu", nn.Linear(D, L)), ("emotion_log_var", nn.Linear(D, L)),
            ("cause_mu", nn.Linear(D, L)), ("cause_log_var", nn.Linear(D, L))]
    if variant == "vi":
        mods += [("ec_mu.0", nn.Linear(L, L)), ("ec_mu.2", nn.Linear(L, L)),
                 ("ec_log_var.0", nn.Linear(L, L)), ("ec_log_var.2", nn.Linear(L, L))]
    if variant == "gan":
        mods += [("ec_disc", nn.Linear(L, opt.ec_num_class)), ("ce_disc", nn.Linear(L, opt.ec_num_class))]
    if variant == "mmd":
        e_cls, c_cls = opt.e_num_class, opt.c_num_class
    elif variant == "vi":
        e_cls, c_cls = 6, opt.ec_num_class
    else:
        e_cls, c_cls = opt.ec_num_class, opt.ec_num_class
    mods += [("emotion_classifier", nn.Linear(L, e_cls)), ("cause_classifier", nn.Linear(L, c_cls)),
             ("pair_classifier", nn.Linear(2 * L, opt.pair_num_class)), ("decoder", nn.Linear(2 * L, V))]
    sd = {}
    for name, m in mods:
        sd[name + ".weight"] = m.weight.detach().clone().to(device)
        sd[name + ".bias"] = m.bias.detach().clone().to(device)
    torch.random.set_rng_state(g)
    return sd


class HeadOracle:
    """Functional restatement of DrlClassifier.forward after `.pooler_output` for the four variants.

    `params` is a dict with the reference's state_dict keys (tensors, requires_grad as the caller
    likes).  RNG: with eps_e/eps_c None and training=True the draws happen in the reference's order
    (randn(L) emotion, randn(L) cause, then the dropout masks in call order), so seeding the same
    device generator reproduces the reference bit for bit.
    """

    def __init__(self, opt, variant: str = "mmd"):
        assert variant in ("mmd", "hsic", "vi", "gan")
        self.opt, self.variant = opt, variant

    # -- helpers --------------------------------------------------------------------------
    def _lin(self, p, name, x):
        return F.linear(x, p[name + ".weight"], p[name + ".bias"])

    def _drop(self, x, training):
        return F.dropout(x, self.opt.dropout, training)

    def latents(self, p, sentence_emb, eps_e=None, eps_c=None):
        """mmd:209-216: four projections then the two reparameterisations (emotion first)."""
        mu_e, lv_e = self._lin(p, "emotion_mu", sentence_emb), self._lin(p, "emotion_log_var", sentence_emb)
        mu_c, lv_c = self._lin(p, "cause_mu", sentence_emb), self._lin(p, "cause_log_var", sentence_emb)
        z_e = sample_prior(mu_e, lv_e, eps_e)
        z_c = sample_prior(mu_c, lv_c, eps_c)
        return mu_e, lv_e, mu_c, lv_c, z_e, z_c

    def emotion_mul_loss(self, p, z_e, labels, training):
        o = self.opt
        if self.variant in ("mmd", "vi"):      # mmd:461-476, vi:465-476 — 6-class CE
            return F.cross_entropy(self._lin(p, "emotion_classifier", self._drop(z_e, training)), labels.flatten())
        preds = torch.sigmoid(self._lin(p, "emotion_classifier", self._drop(z_e, training)))  # hsic:455-469
        t = labels * (1 - o.label_smoothing) + o.label_smoothing / o.ec_num_class
        return F.binary_cross_entropy(preds, t)

    def cause_mul_loss(self, p, z_c, labels, training):
        o = self.opt                             # mmd:478-492 / hsic:471-485
        preds = torch.sigmoid(self._lin(p, "cause_classifier", self._drop(z_c, training)))
        ncls = o.c_num_class if self.variant == "mmd" else o.ec_num_class
        t = labels * (1 - o.label_smoothing) + o.label_smoothing / ncls
        return F.binary_cross_entropy(preds, t)

    def pair_mul_loss(self, p, pair_emb, labels, training):
        o = self.opt                             # mmd:494-513 / hsic:487-504
        logits = self._lin(p, "pair_classifier", self._drop(pair_emb, training))
        ncls = o.pair_num_class if self.variant == "mmd" else o.ec_num_class
        t = labels * (1 - o.label_smoothing) + o.label_smoothing / ncls
        pos_weight = (len(labels) - torch.sum(labels)) / torch.sum(labels)
        loss = F.binary_cross_entropy_with_logits(logits, t, pos_weight=pos_weight)
        if self.variant == "mmd" and bool(loss.isinf().any()):   # mmd:510-511
            return 0
        return loss

    def ec_aprx_net(self, p, c_emb):
        """vi:342-347 with the Sequential members of vi:155-162."""
        mu = self._lin(p, "ec_mu.2", torch.relu(self._lin(p, "ec_mu.0", c_emb)))
        lv = torch.tanh(self._lin(p, "ec_log_var.2", torch.relu(self._lin(p, "ec_log_var.0", c_emb))))
        return mu, lv

    def ec_aprx_loss(self, p, z_e, z_c):
        mu, lv = self.ec_aprx_net(p, z_c.detach())                # vi:421-426
        return -((-(mu - z_e) ** 2 / lv.exp() - lv).sum(dim=1).mean(dim=0))

    def ec_upper_loss(self, p, z_e, z_c, perm=None):
        mu, lv = self.ec_aprx_net(p, z_c)                         # vi:428-439
        if perm is None:
            perm = torch.randperm(z_e.shape[0]).long()            # CPU generator, like the reference
        pos = -(mu - z_e) ** 2 / lv.exp()
        neg = -(mu - z_e[perm]) ** 2 / lv.exp()
        return (pos.sum(dim=-1) - neg.sum(dim=-1)).mean() / 2.0

    def entropy_loss(self, preds):
        return torch.mean(torch.sum(preds * torch.log(preds + self.opt.epsilon), dim=1))   # gan:472-477

    # -- forward ---------------------------------------------------------------------------
    def forward(self, p, sentence_emb, emotion_labels, cause_labels, pair_labels, content_bow, iteration,
                eps_e=None, eps_c=None, training=True, return_parts=False):
        o, v = self.opt, self.variant
        mu_e, lv_e, mu_c, lv_c, z_e, z_c = self.latents(p, sentence_emb, eps_e, eps_c)
        gen = torch.cat((z_e, z_c), dim=1)
        parts = {}
        if v == "gan":                                            # gan:222-233 (call order = RNG order)
            ec_preds = torch.sigmoid(self._lin(p, "ec_disc", self._drop(z_c.detach(), training)))
            t_e = emotion_labels * (1 - o.label_smoothing) + o.label_smoothing / o.ec_num_class
            ec_disc_loss = F.binary_cross_entropy(ec_preds, t_e)
            ec_ent = self.entropy_loss(ec_preds)
            emo = self.emotion_mul_loss(p, z_e, emotion_labels, training)
            ce_preds = torch.sigmoid(self._lin(p, "ce_disc", self._drop(z_e.detach(), training)))
            t_c = cause_labels * (1 - o.label_smoothing) + o.label_smoothing / o.ec_num_class
            ce_disc_loss = F.binary_cross_entropy(ce_preds, t_c)
            ce_ent = self.entropy_loss(ce_preds)
            cau = self.cause_mul_loss(p, z_c, cause_labels, training)
        else:
            emo = self.emotion_mul_loss(p, z_e, emotion_labels, training)
            cau = self.cause_mul_loss(p, z_c, cause_labels, training)
        if v == "mmd":
            indep = -mmd_statistic(z_e, z_c, [0.1])               # mmd:231-233
        elif v == "hsic":
            indep = hsic_trace(z_e, z_c)                          # hsic:214
        elif v == "vi":
            aprx = self.ec_aprx_loss(p, z_e, z_c)                 # vi:226
        pair = self.pair_mul_loss(p, gen, pair_labels, training)
        kl_e, kl_c = kl_loss(mu_e, lv_e), kl_loss(mu_c, lv_c)
        if iteration < o.kl_ann_iterations:                       # mmd:242-250
            w = annealed_weight(iteration, o.ec_kl_lambda, o.kl_ann_iterations)
            kl_e, kl_c = w * kl_e, w * kl_c
        recon = bow_reconstruction(gen, p["decoder.weight"], p["decoder.bias"], content_bow, o.label_smoothing)
        if v == "mmd":                                            # mmd:256-261
            loss = (o.mmd_loss_weight * indep + o.emo_mul_loss_weight * emo + o.cau_mul_loss_weight * cau +
                    o.pair_mul_loss_weight * pair + kl_e + kl_c + recon)
            out = loss
        elif v == "hsic":                                         # hsic:253-258
            loss = indep + o.ec_mul_loss_weight * (emo + cau) + o.pair_mul_loss_weight * pair + kl_e + kl_c + recon
            out = loss
        elif v == "vi":                                           # vi:261-266
            loss = o.ec_mul_loss_weight * (emo + cau) + o.pair_mul_loss_weight * pair + kl_e + kl_c + recon
            out = (z_e, z_c, aprx, loss)
        else:                                                     # gan:275-281
            loss = (o.ecce_adv_loss_weight * (ec_ent + ce_ent) + o.ec_mul_loss_weight * (emo + cau) +
                    o.pair_mul_loss_weight * pair + kl_e + kl_c + recon)
            out = (ec_disc_loss, ce_disc_loss, loss)
        if return_parts:
            parts.update(mu_e=mu_e, lv_e=lv_e, mu_c=mu_c, lv_c=lv_c, z_e=z_e, z_c=z_c, emo=emo, cau=cau, pair=pair,
                         kl_e=kl_e, kl_c=kl_c, recon=recon)
            if v in ("mmd", "hsic"):
                parts["indep"] = indep
            return out, parts
        return out

    def pair_preds(self, p, sentence_emb, eps_e=None, eps_c=None):
        """mmd:265-282 (still stochastic at eval: sample_prior is called)."""
        *_, z_e, z_c = self.latents(p, sentence_emb, eps_e, eps_c)
        return torch.sigmoid(self._lin(p, "pair_classifier", torch.cat((z_e, z_c), dim=1)))


# --------------------------------------------------------------------------------------------
# Synthetic inputs (SURVEY.md §8(d) canonical definition)
# --------------------------------------------------------------------------------------------
def synthetic_batch(N: int, V: int, D: int = 768, nnz: int = 24, seed: int = 0, device="cpu",
                    emotion_float: bool = False):
    g = torch.Generator().manual_seed(seed)
    X = torch.tanh(torch.randn(N, D, generator=g))
    emo = torch.randint(0, 6, (N, 1), generator=g)
    pair = (torch.rand(N, 1, generator=g) < 0.5).float()
    idx = torch.randint(0, V, (N, nnz), generator=g)
    bow = torch.zeros(N, V)
    bow.scatter_add_(1, idx, torch.ones(N, nnz))
    bow = bow / bow.sum(1, keepdim=True).clamp_min(1.0)
    if emotion_float:
        emo = torch.ones(N, 1)          # hsic:78 / gan:83: emotion labels are all ones in those trainers
    return dict(X=X.to(device), emo=emo.to(device), cau=pair.clone().to(device), pair=pair.to(device),
                bow=bow.to(device), bow_idx=idx.to(device))
