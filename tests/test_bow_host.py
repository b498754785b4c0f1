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


def test_drop_in_dataset_dense_rows_and_ell_collate():
    """carel_vae_b200.data.ECPEDataset (reference :86-146): same item dict (keys, dtypes, dense fp32 [V] bow_reps equal
    to the reference loop); with dense_bow=False + ell_collate_fn the batch carries the same targets as an ELL triple."""
    import pandas as pd
    from torch.utils.data import DataLoader
    from carel_vae_b200.data import ECPEDataset, ell_collate_fn

    rng = np.random.default_rng(1)
    vocab = [f"w{i}" for i in range(300)]
    pairs = [" ".join(f"w{int(j)}" for j in rng.integers(0, 400, size=rng.integers(0, 30))) for _ in range(37)]
    df = pd.DataFrame({"pair": pairs, "label": rng.integers(0, 2, 37), "emotion": rng.integers(0, 6, 37)})
    emb = torch.randn(37, 16)
    ds = ECPEDataset(df, tokenizer=None, bow_features=vocab, embeddings=emb)
    item = ds[5]
    assert set(item) == {"input_ids", "labels", "emo_labels", "cau_labels", "bow_reps"}
    assert item["labels"].dtype == torch.float32 and item["emo_labels"].dtype == torch.int64
    assert item["bow_reps"].dtype == torch.float32 and item["bow_reps"].shape == (300,)
    for i in range(37):
        assert np.array_equal(ds[i]["bow_reps"].numpy(), _reference_rows(vocab, bow_tokenize(pairs[i])))
    dense_batches = list(DataLoader(ds, batch_size=16))
    ds2 = ECPEDataset(df, tokenizer=None, bow_features=vocab, embeddings=emb, dense_bow=False)
    ell_batches = list(DataLoader(ds2, batch_size=16, collate_fn=ell_collate_fn(cap=32)))
    assert len(dense_batches) == len(ell_batches) == 3
    for bd, be in zip(dense_batches, ell_batches):
        cols, vals, nnz = be["bow_reps"]
        assert cols.dtype == torch.int32 and vals.dtype == torch.float32 and nnz.dtype == torch.int32
        rebuilt = torch.zeros_like(bd["bow_reps"])
        for b in range(cols.shape[0]):
            n = int(nnz[b])
            rebuilt[b, cols[b, :n].long()] = vals[b, :n]
        assert torch.equal(rebuilt, bd["bow_reps"])
        for k in ("input_ids", "labels", "emo_labels", "cau_labels"):
            assert torch.equal(bd[k], be[k])


def test_host_ell_packer_matches_a_plain_scan():
    """carel_bow_pack_ell_host (data.pack_ell_host): ascending columns, v != 0 (so -0.0 is skipped), nnz counts every
    non-zero while only the first `cap` are stored, ragged widths / unaligned rows, any thread count.  Pure host code:
    runs without a GPU.  (The GPU scan carel_bow_pack_ell is held to the same outputs in tests/test_gpu_parity.py.)"""
    import numpy as np
    import torch
    from carel_vae_b200 import _lib
    from carel_vae_b200.data import pack_ell_host
    if not _lib.LIB_PATH.exists():
        pytest.skip("libcarel_head.so not built")
    rng = np.random.default_rng(0)
    for (B, V, per_row, cap, threads) in [(1, 7, 3, 4, 1), (37, 1001, 5, 8, 3), (64, 2050, 40, 8, 0), (300, 30000, 24, 64, 4)]:
        dense = np.zeros((B, V + 3), dtype=np.float32)[:, 1:V + 1]          # row pitch != V, rows not 8-byte aligned
        for r in range(B):
            idx = rng.choice(V, size=min(per_row, V), replace=False)
            dense[r, idx] = rng.random(len(idx)).astype(np.float32) + 0.01
        dense[0, 0] = -0.0
        t = torch.from_numpy(dense)
        cols, vals, nnz = pack_ell_host(t, cap, threads=threads)
        for r in range(B):
            nz = np.flatnonzero(dense[r] != 0)
            assert int(nnz[r]) == len(nz)
            k = min(len(nz), cap)
            assert np.array_equal(cols[r, :k].numpy(), nz[:k].astype(np.int32))
            assert np.array_equal(vals[r, :k].numpy(), dense[r, nz[:k]])


def _pack_check(seed, B, V, cap, threads):
    import numpy as np
    import torch
    from carel_vae_b200.data import pack_ell_host
    rng = np.random.default_rng(seed)
    dense = np.zeros((B, V), dtype=np.float32)
    for r in range(B):
        idx = rng.choice(V, size=int(rng.integers(0, min(V, cap) + 1)), replace=False)
        dense[r, idx] = rng.random(len(idx)).astype(np.float32) + 0.01
    cols, vals, nnz = pack_ell_host(torch.from_numpy(dense), cap, threads=threads)
    back = np.zeros_like(dense)
    for r in range(B):
        n = int(nnz[r])
        back[r, cols[r, :n].numpy()] = vals[r, :n].numpy()
    return bool(np.array_equal(back, dense) and np.array_equal(nnz.numpy(), (dense != 0).sum(1).astype(np.int32)))


def _pack_in_child(q):
    q.put(_pack_check(5, 200, 3000, 16, 4))


def test_host_ell_packer_worker_pool():
    """The scan's persistent worker pool: many calls in a row with changing thread counts and batch shapes (the pool
    grows, workers beyond the requested count sit a job out, chunks come from a shared counter), calls from two Python
    threads at once (they serialise inside the library) and a call in a forked child after the parent has used the
    pool (threads do not survive fork: the child builds its own)."""
    import multiprocessing as mp
    import threading
    from carel_vae_b200 import _lib
    if not _lib.LIB_PATH.exists():
        pytest.skip("libcarel_head.so not built")
    for i, (B, V, cap, threads) in enumerate([(64, 500, 8, 2), (513, 4000, 32, 8), (40, 100, 4, 16), (1000, 2000, 16, 3),
                                              (16, 64, 4, 8), (257, 9000, 64, 5)] * 3):
        assert _pack_check(i, B, V, cap, threads), (B, V, cap, threads)
    ok = []
    ts = [threading.Thread(target=lambda s=s: ok.append(_pack_check(100 + s, 300, 5000, 32, 4))) for s in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert ok == [True] * 4
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    p = ctx.Process(target=_pack_in_child, args=(q,))
    p.start()
    assert q.get(timeout=60) is True
    p.join(timeout=30)
    assert p.exitcode == 0


def test_host_stager_split_converges_to_equal_times():
    """HostBowStager.balanced_rows (the adaptive host / copy-engine split, pure arithmetic): iterating it against a
    model with constant costs on both sides reaches the split where both routes take the same time; degenerate
    measurements leave the split alone; the result stays a multiple of 8 with rows on both sides."""
    from carel_vae_b200.data import HostBowStager as S
    B = 8192
    host = lambda h: 0.3 + h * (5.2 / B)              # ms: pool wake-up + scan
    dev = lambda h: 0.9 + (B - h) * (19.5 / B)        # ms: X copy + list copy + dense rows over PCIe + GPU scan
    h = 4096
    for _ in range(40):
        h = S.balanced_rows(B, h, h, host(h), dev(h))
        assert h % 8 == 0 and 8 <= h <= B - 8
    assert abs(host(h) - dev(h)) < 0.02 * host(h)
    exact = (0.9 - 0.3 + 19.5) / (5.2 + 19.5) * B
    assert abs(h - exact) <= 16
    # a much slower copy route pushes everything but the minimum to the host, and back
    h2 = 4096
    for _ in range(60):
        h2 = S.balanced_rows(B, h2, h2, 1e-3 * h2, 10.0 * (B - h2))
    assert h2 == B - 8
    for _ in range(60):
        h2 = S.balanced_rows(B, h2, h2, 10.0 * h2, 1e-3 * (B - h2))
    assert h2 == 8
    assert S.balanced_rows(B, 4096, 0, 1.0, 1.0) == 4096 and S.balanced_rows(B, 4096, B, 1.0, 1.0) == 4096
    assert S.balanced_rows(B, 4096, 100, 0.0, 1.0) == 4096 and S.balanced_rows(16, 8, 8, 1.0, 1.0) == 8
