"""carel_vae_b200 — B200-native (sm_100a) implementation of the CAREL-VAE disentanglement head.

Drop-in symbols (same names / signatures / state_dict keys as the reference trainers):

    from carel_vae_b200 import DrlClassifier, MMDStatistic, pdist, HSIC, GaussianKernelMatrix, pairwise_distances

Host code is PyTorch (memory, streams, autograd plumbing); the arithmetic runs in the hand-written
CUDA kernels of libcarel_head.so, reached through the C ABI in include/carel_head.h.  There is no
CPU / PyTorch fallback: without the built extension every op raises.
"""
from ._lib import CarelError, get_precision, launch_count, lib, set_precision  # noqa: F401
from . import data, functional, parallel, train  # noqa: F401
from .losses import (GaussianKernelMatrix, HSIC, MMDStatistic, pairwise_distances, pdist,  # noqa: F401
                     permutation_test_mat)
from .data import ECPEDataset, ell_collate_fn  # noqa: F401
from .head import (DrlClassifier, DrlClassifierGAN, DrlClassifierHSIC, DrlClassifierMMD, DrlClassifierVI,  # noqa: F401
                   IdentityEncoder, VARIANTS, default_opt)

__version__ = "0.1.0"
