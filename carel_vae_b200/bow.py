"""Host-side production of the BoW target.

The reference builds one dense fp32 [V] count vector per sample with an O(V) `list.index` lookup per token
(`ECPEDataset._get_bow_representations`, drl_classifier_ec_mmd_final_mul.py:100-119; en variant with
`bow_tokenize`, drl_classifier_ec_mmd_final_mul_newsplit_emnlp.py:132-154, bow_util.py:42-48) and ships
dense [B, V] rows to the GPU every step (1 GB at B = 8192, V = 30000).  `BowEncoder` produces the same
numbers through a dict lookup, either as the reference's dense rows or as the ELL triple
(cols, vals, nnz) that `DrlClassifier.forward(..., content_bow=(cols, vals, nnz), ...)` accepts.
Tokenisation itself is the caller's: pass word lists (jieba / `bow_tokenize` output).
"""
from __future__ import annotations

import re
from typing import Iterable, List, Sequence, Tuple

import numpy as np
import torch


def bow_tokenize(sentence: str) -> List[str]:
    """English tokenisation of the reference (bow_util.py:42-48): lower-case, strip punctuation, split on
    single spaces, drop RoBERTa space markers and empty tokens."""
    sentence = re.sub(r"[^\w\s]", "", sentence.lower())
    return [t for t in (tok.replace("Ġ", "") for tok in sentence.split(" ")) if t != ""]


class BowEncoder:
    def __init__(self, bow_features: Sequence[str]):
        self.V = len(bow_features)
        # first occurrence wins, like list.index
        self.index = {}
        for i, w in enumerate(bow_features):
            self.index.setdefault(w, i)

    def counts(self, words: Iterable[str]):
        c = {}
        for w in words:
            j = self.index.get(w)
            if j is not None:
                c[j] = c.get(j, 0) + 1
        return c

    def dense(self, words: Iterable[str]) -> np.ndarray:
        """Reference layout: fp32 [V], counts / max(sum(counts), 1)."""
        out = np.zeros(self.V, dtype=np.float32)
        for j, n in self.counts(words).items():
            out[j] += n
        out /= np.max([np.sum(out), 1])
        return out

    def ell_batch(self, batch: Sequence[Iterable[str]], cap: int = 64) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(cols int32 [B, cap], vals fp32 [B, cap], nnz int32 [B]) of the same targets (columns ascending).
        Raises if a sample has more than `cap` distinct in-vocabulary words."""
        B = len(batch)
        cols = np.zeros((B, cap), dtype=np.int32)
        vals = np.zeros((B, cap), dtype=np.float32)
        nnz = np.zeros(B, dtype=np.int32)
        for b, words in enumerate(batch):
            c = self.counts(words)
            if len(c) > cap:
                raise ValueError(f"sample {b} has {len(c)} distinct vocabulary words > cap={cap}")
            total = np.float32(max(sum(c.values()), 1))
            for q, j in enumerate(sorted(c)):
                cols[b, q] = j
                vals[b, q] = np.float32(c[j]) / total
            nnz[b] = len(c)
        return torch.from_numpy(cols), torch.from_numpy(vals), torch.from_numpy(nnz)
