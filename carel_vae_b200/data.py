"""Drop-in `ECPEDataset` + an ELL `collate_fn`: the trainer-facing end of the compact BoW target.

The reference's `ECPEDataset` (drl_classifier_ec_mmd_final_mul.py:86-146) reads the module globals `tokenizer`, `opt`
and `bow`; here they are constructor arguments.  `__getitem__` returns the same dict with the same keys and dtypes, so
the reference's `DataLoader(...)` / `train()` work unchanged.  Two additions:

* the BoW counts are produced by a dict lookup (`bow.BowEncoder`) instead of one O(V) `list.index` per token, and are
  kept SPARSE (column, value pairs) in the dataset: `bow_reps` is expanded to the reference's dense fp32 [V] row only
  when an item is fetched (`dense_bow=True`, the default: bit-identical to the reference's row);
* with `dense_bow=False` an item carries its non-zero list instead, and `ell_collate_fn` stacks a batch into the ELL
  triple `(cols int32 [B, cap], vals fp32 [B, cap], nnz int32 [B])` that `DrlClassifier.forward(content_bow=...)`
  accepts: 27 MB instead of 1 GB per 8192-clause batch over PCIe, and no dense -> ELL pass on the GPU.

Tokenisation of the BoW words is the caller's (`word_fn`): `jieba.lcut` after the non-Chinese filter for zh (mmd:100-106),
`bow.bow_tokenize` for en (bow_util.py:42-48).  `tokenizer=None` skips the encoder inputs (pre-computed clause
embeddings: `input_ids` then holds the embedding row passed in `embeddings`).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np
import torch

from .bow import BowEncoder, bow_tokenize


class ECPEDataset(torch.utils.data.Dataset):
    def __init__(self, df, tokenizer=None, bow_features: Sequence[str] = (), max_len: int = 128,
                 word_fn: Callable[[str], Sequence[str]] = bow_tokenize, dense_bow: bool = True,
                 embeddings: Optional[torch.Tensor] = None):
        self.tokenizer = tokenizer
        self.pairs = df["pair"]
        self.labels = df["label"].values
        self.emo_labels = df["emotion"].values
        self.cau_labels = df["label"].values          # the reference feeds the pair label to the cause head (:92)
        self.max_len = max_len
        self.bow_features = bow_features
        self.dense_bow = dense_bow
        self.embeddings = embeddings
        self._encoder = BowEncoder(bow_features)
        self._counts = [self._encoder.counts(word_fn(str(p))) for p in self.pairs]

    def __len__(self):
        return len(self.pairs)

    def _get_bow_representations(self, index: int) -> np.ndarray:
        """The reference's dense fp32 [V] row (mmd:100-119) of sample `index`."""
        out = np.zeros(self._encoder.V, dtype=np.float32)
        for j, n in self._counts[index].items():
            out[j] += n
        out /= np.max([np.sum(out), 1])
        return out

    def bow_nonzeros(self, index: int):
        """(cols int32 [n], vals fp32 [n]) of the same row, columns ascending."""
        c = self._counts[index]
        total = np.float32(max(sum(c.values()), 1))
        cols = np.fromiter(sorted(c), dtype=np.int32, count=len(c))
        vals = np.array([np.float32(c[j]) / total for j in cols], dtype=np.float32)
        return cols, vals

    def __getitem__(self, index):
        instance = {
            "labels": torch.FloatTensor([self.labels[index]]),
            "emo_labels": torch.LongTensor([self.emo_labels[index]]),
            "cau_labels": torch.FloatTensor([self.cau_labels[index]]),
        }
        if self.tokenizer is not None:
            inputs = self.tokenizer.encode_plus(str(self.pairs[index]), None, add_special_tokens=True, max_length=self.max_len,
                                                padding="max_length", return_token_type_ids=True, truncation=True,
                                                return_attention_mask=True, return_tensors="pt")
            instance.update(input_ids=inputs["input_ids"].flatten(), attention_masks=inputs["attention_mask"].flatten(),
                            token_type_ids=inputs["token_type_ids"].flatten())
        elif self.embeddings is not None:
            instance["input_ids"] = self.embeddings[index]
        if self.dense_bow:
            instance["bow_reps"] = torch.from_numpy(self._get_bow_representations(index))
        else:
            cols, vals = self.bow_nonzeros(index)
            instance["bow_cols"], instance["bow_vals"] = torch.from_numpy(cols), torch.from_numpy(vals)
        return instance


def ell_collate_fn(cap: int = 64):
    """collate_fn for `DataLoader(ECPEDataset(..., dense_bow=False))`: default stacking for every key except the BoW
    non-zeros, which become `batch["bow_reps"] = (cols [B, cap], vals [B, cap], nnz [B])` (zero padded).  A sample with
    more than `cap` distinct vocabulary words raises (the kernels would otherwise need the dense row)."""
    from torch.utils.data import default_collate

    def collate(items):
        B = len(items)
        cols = torch.zeros((B, cap), dtype=torch.int32)
        vals = torch.zeros((B, cap), dtype=torch.float32)
        nnz = torch.zeros((B,), dtype=torch.int32)
        rest = []
        for b, it in enumerate(items):
            n = it["bow_cols"].numel()
            if n > cap:
                raise ValueError(f"sample {b} has {n} distinct vocabulary words > cap={cap}")
            cols[b, :n] = it["bow_cols"]
            vals[b, :n] = it["bow_vals"]
            nnz[b] = n
            rest.append({k: v for k, v in it.items() if k not in ("bow_cols", "bow_vals")})
        out = default_collate(rest)
        out["bow_reps"] = (cols, vals, nnz)
        return out
    return collate


def pack_ell_host(content_bow: torch.Tensor, cap: int = 64, out=None, threads: int = 0):
    """Dense fp32 [B, V] BoW rows in HOST memory -> the ELL triple (cols int32 [B, cap], vals fp32 [B, cap], nnz int32 [B])
    in host memory, by the library's multi-threaded scan (carel_bow_pack_ell_host): the same non-zeros as
    `functional.bow_pack_ell`, in ascending column order.  `out` = a previously returned triple to reuse (pinned buffers stay pinned)."""
    import ctypes as C
    from ._lib import check, lib
    if content_bow.is_cuda or content_bow.dtype != torch.float32 or content_bow.dim() != 2:
        raise ValueError("pack_ell_host expects a 2-D fp32 CPU tensor (use functional.bow_pack_ell on the GPU)")
    if content_bow.stride(1) != 1:
        content_bow = content_bow.contiguous()
    B, V = content_bow.shape
    if out is None:
        out = (torch.zeros((B, cap), dtype=torch.int32), torch.zeros((B, cap), dtype=torch.float32),
               torch.zeros((B,), dtype=torch.int32))
    cols, vals, nnz = out
    if tuple(cols.shape) != (B, cap) or tuple(vals.shape) != (B, cap) or tuple(nnz.shape) != (B,):
        raise ValueError("pack_ell_host: `out` does not match the batch")
    check(lib().carel_bow_pack_ell_host(C.c_void_p(content_bow.data_ptr()), content_bow.stride(0), B, V, cap,
                                        C.c_void_p(cols.data_ptr()), C.c_void_p(vals.data_ptr()),
                                        C.c_void_p(nnz.data_ptr()), int(threads)), "carel_bow_pack_ell_host")
    return out


class HostBowStager:
    """Moves a HOST-resident dense BoW target [B, V] (the reference's layout, mmd:100-119 / :830) to the GPU as the ELL
    triple `DrlClassifier.forward(content_bow=...)` accepts, using the host cores and the PCIe link TOGETHER: the first
    `host_rows` rows are scanned by the host threads (pack_ell_host; 27 MB of list instead of 983 MB of rows per 8192
    clauses) while the copy engine ships the remaining rows dense and the GPU scans those (functional.bow_pack_ell).
    `calibrate()` times both routes once and splits the rows so that they finish together; `adapt=True` keeps moving
    the split towards the point where the two routes measured in the previous stages take the same time.

    `pipelined=False`: everything is issued on the current stream; the returned tensors are the stager's own device
    buffers (or `out`), valid until the next `stage()`.

    `pipelined=True`: the copies and the GPU scan run on the stager's own stream into one of TWO buffer sets, so batch
    k + 1 is staged (host scan and PCIe copy) while the current stream still trains on batch k; the current stream only
    waits for the set to be complete and, when `out` is given (e.g. the static inputs of a captured CUDA graph), copies
    it there device-to-device.  Without `out` the returned set stays valid until the next-but-one `stage()`, and
    whatever reads it must be enqueued on the current stream before the next `stage()` call.  `stage(bow, copies=...)`
    takes the batch's other host tensors along the same way (`(device_dst, host_src)` pairs).

    No call synchronises the host with the GPU except on the buffer set used two stages (pipelined) / one stage ago.
    One stager per (batch size, vocabulary, device)."""

    def __init__(self, batch: int, vocab: int, device, cap: int = 64, host_fraction: float = 0.5, threads: int = 0,
                 out=None, pipelined: bool = False, adapt: bool = False):
        self.B, self.V, self.cap, self.device, self.threads = batch, vocab, cap, torch.device(device), threads
        self.host_rows = max(0, min(batch, int(round(batch * host_fraction))))
        self.pipelined, self.adapt = bool(pipelined), bool(adapt)
        nset = 2 if self.pipelined else 1

        def dev_set():
            return (torch.zeros((batch, cap), dtype=torch.int32, device=self.device),
                    torch.zeros((batch, cap), dtype=torch.float32, device=self.device),
                    torch.zeros((batch,), dtype=torch.int32, device=self.device))
        self._hosts = [(torch.zeros((batch, cap), dtype=torch.int32).pin_memory(),
                        torch.zeros((batch, cap), dtype=torch.float32).pin_memory(),
                        torch.zeros((batch,), dtype=torch.int32).pin_memory()) for _ in range(nset)]
        self._host = self._hosts[0]
        self._out = tuple(out) if out is not None else None   # caller-owned device buffers
        if self.pipelined:
            self._sets = [dev_set() for _ in range(nset)]
        else:
            self._sets = [self._out if self._out is not None else dev_set()]
        self.cols, self.vals, self.nnz = self._out if self._out is not None else self._sets[0]
        self._dense = torch.empty((batch, vocab), dtype=torch.float32, device=self.device)   # landing zone of the dense part
        self._side = torch.cuda.Stream(self.device) if self.pipelined else None
        self._ready = [None] * nset       # copy-stream event: set i is complete (its pinned lists have been read)
        self._consumed = [None] * nset    # current-stream event: everything that read set i has been enqueued before it
        self._meas = [None] * nset        # (host rows, host scan ms, timing events) of the stage that filled set i
        self._staging = [dict() for _ in range(nset)]
        self._k = 0
        self._last = None
        self.calibration = None

    def bytes_per_stage(self) -> int:
        h = self.host_rows
        return h * (self.cap * 8 + 4) + (self.B - h) * self.V * 4

    @staticmethod
    def balanced_rows(batch: int, current: int, measured_rows: int, t_host: float, t_dev: float) -> int:
        """New `host_rows`: half-way from `current` to the split at which the two routes finish together, given that
        the host scanned `measured_rows` rows in `t_host` ms while the copy engine + GPU took `t_dev` ms for the other
        `batch - measured_rows`.  Rates are taken as proportional to the rows, so the fixed point of repeated calls is
        t_host == t_dev whatever constant costs either side carries.  Multiples of 8, at least 8 rows on each side;
        unusable measurements (a route without rows, a zero time) leave the split alone."""
        if batch < 24 or not (0 < measured_rows < batch) or t_host <= 0.0 or t_dev <= 0.0:
            return current
        r_host, r_dev = measured_rows / t_host, (batch - measured_rows) / t_dev
        target = batch * r_host / (r_host + r_dev)
        new = current + 0.5 * (target - current)
        return max(8, min(batch - 8, int(round(new / 8)) * 8))

    def _rebalance(self, i: int) -> None:
        m, self._meas[i] = self._meas[i], None
        if m is None or not self.adapt:
            return
        h, t_host, e0, e1, e2, e3 = m
        t_dev = e0.elapsed_time(e1) + (e2.elapsed_time(e3) if e2 is not None else 0.0)   # dense route + list copy
        self.host_rows = self.balanced_rows(self.B, self.host_rows, h, t_host, t_dev)

    def stage(self, content_bow: torch.Tensor, copies=()):
        import time
        from . import functional as F
        if content_bow.is_cuda or tuple(content_bow.shape) != (self.B, self.V):
            raise ValueError(f"HostBowStager.stage expects a CPU tensor of shape ({self.B}, {self.V})")
        main = torch.cuda.current_stream(self.device)
        i = (self._k & 1) if self.pipelined else 0
        self._k += 1
        if self.pipelined and self._last is not None and self._out is None:
            ev = torch.cuda.Event()           # the caller has enqueued its reads of the previous set by now
            ev.record(main)
            self._consumed[self._last] = ev
        if self._ready[i] is not None:
            self._ready[i].synchronize()      # the copies that read pinned set i (one / two stages ago) have finished
            self._rebalance(i)
        h = self.host_rows
        side = self._side if self.pipelined else main
        cols, vals, nnz = self._sets[i]
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        staged = []
        with torch.cuda.stream(side):
            if self.pipelined and self._consumed[i] is not None:
                side.wait_event(self._consumed[i])          # set i is free to be overwritten
            e0.record(side)
            for j, (dst, src) in enumerate(copies):
                if self.pipelined:
                    stg = self._staging[i].get(j)
                    if stg is None or stg.shape != dst.shape or stg.dtype != dst.dtype:
                        stg = self._staging[i][j] = torch.empty_like(dst)
                    stg.copy_(src, non_blocking=True)
                    staged.append((dst, stg))
                else:
                    dst.copy_(src, non_blocking=True)
            # (1) dense rows [h, B): async copy + GPU scan — enqueued first, so the DMA runs while the host scans
            if h < self.B:
                self._dense[h:].copy_(content_bow[h:], non_blocking=True)
                c, v, n = F.bow_pack_ell(self._dense[h:], self.cap)
                cols[h:].copy_(c, non_blocking=True)
                vals[h:].copy_(v, non_blocking=True)
                nnz[h:].copy_(n, non_blocking=True)
            e1.record(side)
        # (2) rows [0, h): host scan into pinned buffers, then the (small) list copy
        t_host = 0.0
        e2 = e3 = None
        if h > 0:
            hc, hv, hn = (t[:h] for t in self._hosts[i])
            t0 = time.perf_counter()
            pack_ell_host(content_bow[:h], self.cap, out=(hc, hv, hn), threads=self.threads)
            t_host = 1e3 * (time.perf_counter() - t0)
            e2 = torch.cuda.Event(enable_timing=True)
            e3 = torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(side):
                e2.record(side)
                cols[:h].copy_(hc, non_blocking=True)
                vals[:h].copy_(hv, non_blocking=True)
                nnz[:h].copy_(hn, non_blocking=True)
                e3.record(side)
        ready = torch.cuda.Event()
        ready.record(side)
        self._ready[i] = ready
        self._meas[i] = (h, t_host, e0, e1, e2, e3)
        self._last = i
        if not self.pipelined:
            return self._sets[0]
        main.wait_event(ready)
        for dst, stg in staged:
            dst.copy_(stg, non_blocking=True)
        if self._out is None:
            self.cols, self.vals, self.nnz = cols, vals, nnz
            return cols, vals, nnz
        for dst, src in zip(self._out, (cols, vals, nnz)):
            dst.copy_(src, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(main)
        self._consumed[i] = ev
        return self._out

    def calibrate(self, content_bow: torch.Tensor, rounds: int = 2) -> float:
        """Times the host scan and the dense copy + GPU scan (each on the full batch, alone) and sets `host_rows` to the
        split at which both finish together; returns the host fraction."""
        import time
        from . import functional as F
        t_host = 1e30
        for _ in range(rounds):
            t0 = time.perf_counter()
            pack_ell_host(content_bow, self.cap, out=self._hosts[0], threads=self.threads)
            t_host = min(t_host, time.perf_counter() - t0)
        t_dev = 1e30
        for _ in range(rounds):
            torch.cuda.synchronize(self.device)
            t0 = time.perf_counter()
            self._dense.copy_(content_bow, non_blocking=True)
            F.bow_pack_ell(self._dense, self.cap)
            torch.cuda.synchronize(self.device)
            t_dev = min(t_dev, time.perf_counter() - t0)
        frac = t_dev / (t_dev + t_host)
        self.host_rows = max(0, min(self.B, int(round(self.B * frac / 8)) * 8))
        self.calibration = {"host_scan_ms": round(1e3 * t_host, 3), "copy_plus_gpu_scan_ms": round(1e3 * t_dev, 3),
                            "host_fraction": round(self.host_rows / self.B, 3)}
        return self.host_rows / self.B
