"""Host-side batch builders for the IRN hot path (reference: data_provider.py:568-575 train collate, :591-617
DataLoaderEvalIRS._collate_fn, :769-783 DataLoaderEvalNN1._collate_fn).

The reference builds every batch row by row in Python (np.zeros + slice assignment per sample, then
``torch.LongTensor(np.array(list_of_rows))``) from pickled object arrays of ragged id lists.  Here the ragged
histories live in one flat int32 CSR pair (values, offsets) and a batch is assembled with a handful of vectorised numpy
operations straight into PINNED int64 host tensors, so the ``.to(device)`` of pipeline.py is an asynchronous DMA.  The
produced tensors are element-for-element what the reference's collate functions return."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch


def to_csr(seqs: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
    """Ragged id lists (data_pack.npy['data'], the first column of the irs_*.npy object arrays) -> (values int32
    [total], offsets int64 [n+1]).  numpy-2 safe: no object arrays, no pickle."""
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    values = np.empty(int(offsets[-1]), dtype=np.int32)
    for i, s in enumerate(seqs):
        values[offsets[i]:offsets[i + 1]] = np.asarray(s, dtype=np.int64)
    return values, offsets


def _pinned(shape, dtype=torch.int64, pin=True):
    t = torch.empty(shape, dtype=dtype)
    if pin and torch.cuda.is_available():
        t = t.pin_memory()
    return t


def collate_eval_irs(values: np.ndarray, offsets: np.ndarray, rows: np.ndarray, users, targets, labels, seq_len: int,
                     gap_len: int = 0, pin: bool = True):
    """DataLoaderEvalIRS._collate_fn (data_provider.py:591-617) for the samples ``rows`` of a CSR history store:
    window = [0.., last (seq_len - gap_len - 1) history items, gap_len zeros, target]  (pre-padded).
    Returns (raw list of LongTensor, seqs [B,seq_len], users [B], targets [B], labels [B])."""
    rows = np.asarray(rows, dtype=np.int64)
    B = rows.shape[0]
    l_hist = seq_len - gap_len - 1
    start, end = offsets[rows], offsets[rows + 1]
    n = np.minimum(end - start, l_hist)                              # items of each history that fit
    seqs = _pinned((B, seq_len), pin=pin)
    out = seqs.numpy()
    out[:] = 0
    # window column c in [0, l_hist) holds history item (end - l_hist + c) when that index is >= end - n
    col = np.arange(l_hist, dtype=np.int64)[None, :]
    src = (end - l_hist)[:, None] + col
    valid = col >= (l_hist - n)[:, None]
    out[:, :l_hist][valid] = values[src[valid]]
    out[:, -1] = np.asarray(targets, dtype=np.int64)
    raw = [torch.from_numpy(values[s:e].astype(np.int64)) for s, e in zip(start, end)]
    cast = lambda a: torch.from_numpy(np.asarray(a, dtype=np.int64))
    return raw, seqs, cast(users), cast(targets), torch.as_tensor(np.asarray(labels))


def collate_train(windows: np.ndarray, users, rows=None, pin: bool = True):
    """Train / validation collate (data_provider.py:568-575): fixed-length pre-padded windows [n, L] (any integer or
    float dtype, as stored in irs_train_seq / irs_valid_seq) -> (seqs int64 [B,L] pinned, users int64 [B])."""
    w = windows if rows is None else windows[np.asarray(rows, dtype=np.int64)]
    u = np.asarray(users) if rows is None else np.asarray(users)[np.asarray(rows, dtype=np.int64)]
    seqs = _pinned(w.shape, pin=pin)
    seqs.numpy()[:] = w.astype(np.int64, copy=False)
    return seqs, torch.from_numpy(u.astype(np.int64))


def collate_eval_nn1(histories: np.ndarray, new_seqs: np.ndarray, targets, start_pos, l_path, pin: bool = True):
    """DataLoaderEvalNN1._collate_fn (data_provider.py:769-783): stacked evaluator inputs as pinned tensors."""
    h = _pinned(histories.shape, pin=pin)
    h.numpy()[:] = histories.astype(np.int64, copy=False)
    s = _pinned(new_seqs.shape, pin=pin)
    s.numpy()[:] = new_seqs.astype(np.int64, copy=False)
    return h, s, torch.as_tensor(np.asarray(targets)), torch.as_tensor(np.asarray(start_pos)), torch.as_tensor(np.asarray(l_path))
