"""CPU: host-side BoW target production against a literal restatement of the reference loop
(ECPEDataset._get_bow_representations, drl_classifier_ec_mmd_final_mul.py:107-119)."""
import numpy as np
import torch

from carel_vae_b200.bow import BowEncoder, bow_tokenize


def _reference_rows(bow_features, words):
    rep = np.zeros(shape=len(bow_features), dtype=np.float32)
    for word in words:
        if word in bow_features:
            rep[bow_features.index(word)] += 1
    rep /= np.max([np.sum(rep), 1])
    return rep


def test_dense_and_ell_match_reference_loop():
    rng = np.random.default_rng(0)
    vocab = [f"w{i}" for i in range(500)] + ["w3"]          # a duplicate entry: list.index takes the first
    enc = BowEncoder(vocab)
    batch = [[f"w{int(j)}" for j in rng.integers(0, 700, size=rng.integers(0, 40))] for _ in range(50)]
    batch.append([])                                          # empty sample: all-zero target
    dense = np.stack([enc.dense(w) for w in batch])
    ref = np.stack([_reference_rows(vocab, w) for w in batch])
    assert np.array_equal(dense, ref)
    cols, vals, nnz = enc.ell_batch(batch, cap=48)
    rebuilt = torch.zeros(len(batch), len(vocab))
    for b in range(len(batch)):
        n = int(nnz[b])
        rebuilt[b, cols[b, :n].long()] = vals[b, :n]
    assert torch.equal(rebuilt, torch.from_numpy(ref))


def test_bow_tokenize_matches_reference_rule():
    # NB: like the reference, the text is lower-cased BEFORE the 'Ġ' marker is stripped, so a marker that
    # survives as 'ġ' stays in the token (bow_util.py:43-47) — reproduced, not fixed.
    assert bow_tokenize("Hello, World!  it's Ġfine.") == ["hello", "world", "its", "ġfine"]
    assert bow_tokenize("a  b") == ["a", "b"]


def test_bow_tokenize_matches_the_live_reference_function():
    """Build container only: bow_util.py's own bow_tokenize (AST-extracted: the module imports jieba at the top) on
    random punctuation / case / marker / whitespace mixes."""
    import ast
    import re
    from pathlib import Path
    import pytest
    src = Path("/root/reference/bow_util.py")
    if not src.exists():
        pytest.skip("the reference is only present in the build container")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)          # the reference's un-raw regex strings
        tree = ast.parse(src.read_text())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "bow_tokenize"]
    ns = {"re": re}
    exec(compile(ast.Module(body=fn, type_ignores=[]), "bow_util.py", "exec"), ns)
    rng = np.random.default_rng(1)
    alphabet = list("abcXYZ09 _-.,!?'\"Ġġé中  \t")
    for _ in range(300):
        s = "".join(rng.choice(alphabet, size=rng.integers(0, 60)))
        assert bow_tokenize(s) == ns["bow_tokenize"](s), repr(s)
