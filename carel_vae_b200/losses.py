"""Drop-in loss primitives with the reference's names and call signatures.

    from carel_vae_b200.losses import MMDStatistic, pdist, HSIC, GaussianKernelMatrix, pairwise_distances

replaces the copies every drl_classifier_ec_*.py script defines
(drl_classifier_ec_mmd_final_mul.py:537-600, drl_classifier_ec_hsic.py:529-547).
All of them launch the sm_100a kernels of libcarel_head.so; inputs must be CUDA tensors.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import functional as F


class MMDStatistic:
    """Unbiased MMD^2 statistic (reference :537-577; same attributes n_1, n_2, a00, a11, a01)."""

    def __init__(self, n_1: int, n_2: int):
        self.n_1 = n_1
        self.n_2 = n_2
        # The three constants used in the test (ZeroDivisionError for a sample of size 1, as in
        # the reference, whose DataLoader can produce a final batch of one row).
        self.a00 = 1.0 / (n_1 * (n_1 - 1))
        self.a11 = 1.0 / (n_2 * (n_2 - 1))
        self.a01 = -1.0 / (n_1 * n_2)

    def __call__(self, sample_1: torch.Tensor, sample_2: torch.Tensor, alphas: Sequence[float],
                 ret_matrix: bool = False):
        if sample_1.shape[0] != self.n_1 or sample_2.shape[0] != self.n_2:
            raise ValueError("sample sizes do not match the (n_1, n_2) this statistic was built for")
        mmd = F.mmd(sample_1, sample_2, list(alphas))
        if ret_matrix:
            # materialising path (no caller in the reference uses it; kept for API completeness)
            s = torch.cat((sample_1.detach(), sample_2.detach()), 0)
            return mmd, F.pairwise(s, s, 0, list(alphas), 1e-5)
        return mmd

    def halves(self, z: torch.Tensor, alphas: Sequence[float]):
        """self(z[:, :L], z[:, L:], alphas) for z = [sample_1 | sample_2] stored side by side (the head's latent
        tensor): identical value, one [N, 2L] gradient instead of two sliced ones."""
        if z.shape[0] != self.n_1 or self.n_1 != self.n_2 or z.shape[1] % 2:
            raise ValueError("halves() needs z of shape [n, 2L] with n == n_1 == n_2")
        return F.mmd_halves(z, z.shape[1] // 2, list(alphas))

    def pval(self, distances, n_permutations=1000):
        """Dead in the reference too: permutation_test_mat is a stub returning None (:599-600)."""
        return permutation_test_mat(distances, self.n_1, self.n_2, n_permutations,
                                    a00=self.a00, a11=self.a11, a01=self.a01)


def pdist(sample_1: torch.Tensor, sample_2: torch.Tensor, norm: float = 2, eps: float = 1e-5):
    """Dense [n_1, n_2] pairwise distances (reference :580-596).  Only norm == 2 is ever used by the
    reference and only that is implemented on the GPU path."""
    if float(norm) != 2.0:
        raise NotImplementedError("carel_vae_b200.pdist implements norm=2 (the only value the reference calls)")
    return F.pairwise(sample_1, sample_2, 1, (), eps)


def permutation_test_mat(*args, **kwargs):  # real signature unknown in the reference as well
    pass


def pairwise_distances(x: torch.Tensor):
    """-2 x x^T + |x_i|^2 + |x_j|^2 (reference hsic.py:529-532)."""
    return F.pairwise(x, x, 2)


def GaussianKernelMatrix(x: torch.Tensor, sigma: float = 1):
    """exp(-pairwise_distances(x) / sigma) (reference hsic.py:535-537)."""
    return F.pairwise(x, x, 3, (1.0 / float(sigma),))


def HSIC(x: torch.Tensor, y: torch.Tensor, s_x: float = 1, s_y: float = 1):
    """trace(L H K H) / (m-1)^2 (reference hsic.py:540-547) evaluated in O(m^2) without the centring
    matmuls; differentiable w.r.t. x and y."""
    return F.hsic(x, y, float(s_x), float(s_y))
