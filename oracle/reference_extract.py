"""TEST INFRASTRUCTURE — not product code.

Executes the reference's OWN hot-path code (read-only at /root/reference, present only in the build
container) without copying it: the trainer scripts cannot be imported (module-level argparse, jieba,
HF downloads, stdout redirection — SURVEY.md §8(c)), so the top-level class / function definitions
of the head are AST-extracted and exec'd in a namespace that provides `torch, nn, math, Variable,
opt` and an identity stand-in for BertModel / RobertaModel whose `pooler_output` is the tensor passed
as `input_ids`.  Feeding [N, 768] float embeddings as `input_ids` then runs the reference forward,
losses and autograd verbatim.
"""
from __future__ import annotations

import ast
import math
from pathlib import Path
from types import SimpleNamespace

import torch
import torch.nn as nn

REFERENCE_ROOT = Path("/root/reference")
FILES = {
    "mmd": "drl_classifier_ec_mmd_final_mul.py",
    "hsic": "drl_classifier_ec_hsic.py",
    "vi": "drl_classifier_ec_vi_final.py",
    "gan": "drl_classifier_ec_gan.py",
}
WANTED = {"DrlClassifier", "MMDStatistic", "pdist", "permutation_test_mat", "HSIC", "GaussianKernelMatrix",
          "pairwise_distances"}


class _IdentityEncoder(nn.Module):
    """Stand-in for the HF encoder: pooler_output = input_ids (the synthetic clause embeddings)."""

    @classmethod
    def from_pretrained(cls, *a, **k):
        return cls()

    def forward(self, input_ids, attention_mask=None, token_type_ids=None):
        return SimpleNamespace(pooler_output=input_ids)


def available() -> bool:
    return (REFERENCE_ROOT / FILES["mmd"]).exists()


def load(variant: str, opt) -> dict:
    """Return the namespace holding the reference's DrlClassifier / loss callables for `variant`,
    bound to the module-global `opt` the reference methods read."""
    import warnings
    src = (REFERENCE_ROOT / FILES[variant]).read_text()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)   # the reference's un-raw regex strings
        tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in WANTED]
    mod = ast.Module(body=keep, type_ignores=[])
    from torch.autograd import Variable
    ns = {"torch": torch, "nn": nn, "math": math, "Variable": Variable, "opt": opt,
          "BertModel": _IdentityEncoder, "RobertaModel": _IdentityEncoder}
    exec(compile(mod, FILES[variant], "exec"), ns)
    return ns


class cpu_cuda_shim:
    """hsic.py:545 hard-codes `.cuda()`; on a CPU-only box make it the identity while HSIC runs."""

    def __enter__(self):
        self._orig = torch.Tensor.cuda
        if not torch.cuda.is_available():
            torch.Tensor.cuda = lambda self, *a, **k: self
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda = self._orig
        return False
