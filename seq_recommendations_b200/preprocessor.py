"""Batch format of the reference's preprocessor.py, same class / method names and return shapes, vectorised numpy.

Format contract (preprocessor.py:16-20, :30-60, :67-94):
  * a sequence s of length L gives inputs s[0..L-2] and targets s[1..L-1]
  * sequences are LEFT padded (padding='pre') with `pad_value` and LEFT truncated (truncating='pre': the tail is
    kept) to `seq_length` (default: the longest sequence); `self.seq_length` is updated like the reference does
  * dense mode: x, y one-hot (N,T,V) float64; c = history features xs[:-1] (N,T,V) float64
  * sparse mode: x, y are ids (N,T,1) float64 -- an id TRANSPORT format for the B200 backend.  The reference pads ids
    with 0, which collides with item 0 (preprocessor.py:80 vs model.py:335), so here pads are -1 unless a negative
    `pad_value` is given; `IdBatch` (below) is the zero-copy int32 form the engine consumes directly.
"""
import numpy as np


def pad_pre(rows, maxlen, feat_shape, dtype, value):
    """keras.preprocessing.sequence.pad_sequences(padding='pre', truncating='pre') for a list of (L_i, *feat) arrays."""
    n = len(rows)
    out = np.full((n, maxlen) + tuple(feat_shape), value, dtype=dtype)
    for i, r in enumerate(rows):
        L = len(r)
        if L == 0:
            continue
        t = np.asarray(r[-maxlen:], dtype=dtype)
        out[i, maxlen - len(t):] = t.reshape((len(t),) + tuple(feat_shape))
    return out


def one_hot(ids, n_classes):
    ids = np.asarray(ids, dtype=np.int64)
    out = np.zeros((len(ids), n_classes), dtype=np.float64)
    out[np.arange(len(ids)), ids] = 1.0
    return out


class Preprocessor(object):
    def __init__(self, vocab, pad_value=0., seq_length=None, sparse=False):
        self.vocab = vocab
        self.seq_length = seq_length
        self.pad_value = pad_value
        self.sparse = sparse

    def _maxlen(self, rows):
        return self.seq_length if self.seq_length is not None else max((len(r) for r in rows), default=0)

    def _pad_sequences(self, rows, feat_shape, dtype=np.float64, value=None):
        maxlen = self._maxlen(rows)
        padded = pad_pre(rows, maxlen, feat_shape, dtype, self.pad_value if value is None else value)
        self.seq_length = padded.shape[1]
        return padded

    def transform_data(self, sequences, xs=None, pad=True):
        pass


class BaselinePreprocessor(Preprocessor):
    """x = [onehot(item) || xs] features, y = onehot(next item); sequences shorter than 2 are dropped
    (preprocessor.py:30-60)."""

    def __init__(self, vocab, pad_value=0., seq_length=None):
        Preprocessor.__init__(self, vocab, pad_value, seq_length)

    def transform_data(self, sequences, xs=None, pad=True):
        V = len(self.vocab)
        x_data, y_data = [], []
        for index, seq in enumerate(sequences):
            if len(seq) < 2:
                continue
            feats = one_hot(seq[:-1], V)
            if xs is not None:
                feats = np.concatenate([feats, np.asarray(xs[index][:len(seq) - 1], dtype=np.float64)], axis=1)
            x_data.append(feats)
            y_data.append(one_hot(seq[1:], V))
        if not pad:
            return [r.tolist() for r in x_data], [r.tolist() for r in y_data]
        features_dim = V + (V if xs is not None else 0)
        return (self._pad_sequences(x_data, (features_dim,)), self._pad_sequences(y_data, (V,)))


class FullModelPreprocessor(Preprocessor):
    """x = onehot(s[:-1]), y = onehot(s[1:]), c = xs[:-1]  (preprocessor.py:67-94).  Length-1 sequences stay as
    all-pad rows."""

    def __init__(self, vocab, pad_value=0., seq_length=None, sparse=False):
        Preprocessor.__init__(self, vocab, pad_value, seq_length, sparse=sparse)

    def transform_data(self, sequences, xs, pad=True):
        V = len(self.vocab)
        c_rows = [np.asarray(x[:-1], dtype=np.float64).reshape(-1, V) for x in xs]
        if not self.sparse:
            x_rows = [one_hot(s[:-1], V) for s in sequences]
            y_rows = [one_hot(s[1:], V) for s in sequences]
            feat, value = (V,), None
        else:
            x_rows = [np.asarray(s[:-1], dtype=np.float64).reshape(-1, 1) for s in sequences]
            y_rows = [np.asarray(s[1:], dtype=np.float64).reshape(-1, 1) for s in sequences]
            feat, value = (1,), (self.pad_value if self.pad_value < 0 else -1.0)
        x = self._pad_sequences(x_rows, feat, value=value)
        c = self._pad_sequences(c_rows, (V,))
        y = self._pad_sequences(y_rows, feat, value=value)
        return x, y, c

    def transform_ids(self, sequences):
        """New, additive: the same batch as int32 ids (N,T) with -1 pads -- what the engine stages to HBM.  Avoids the
        (N,T,V) one-hot entirely (cfg2 dense: 1 GB per batch; cfg3: 41 GB)."""
        x_rows = [np.asarray(s[:-1], dtype=np.int32).reshape(-1, 1) for s in sequences]
        y_rows = [np.asarray(s[1:], dtype=np.int32).reshape(-1, 1) for s in sequences]
        x = self._pad_sequences(x_rows, (1,), dtype=np.int32, value=-1)
        y = self._pad_sequences(y_rows, (1,), dtype=np.int32, value=-1)
        return x[:, :, 0], y[:, :, 0]


def ragged(sequences):
    """List of item-id sequences -> (flat int32 items, int64 offsets of length n+1)."""
    lens = np.fromiter((len(s) for s in sequences), dtype=np.int64, count=len(sequences))
    offs = np.zeros(len(sequences) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    flat = np.fromiter((int(v) for s in sequences for v in s), dtype=np.int32, count=int(offs[-1]))
    return flat, offs


def transform_ids_device(sequences, seq_length=None, device=None):
    """`FullModelPreprocessor.transform_ids` on the GPU: the ragged corpus goes to HBM once as
    (items, offsets) and the left-padded / left-truncated (N,T) int32 id and target batches are built by
    `seqrec_pad_sequences` -- bit-identical to the host path, no Python loop over sequences, no one-hot.
    `sequences` is a list of id sequences or an (items, offsets) pair.  Returns device tensors (ids, targets)."""
    import ctypes
    import torch
    from ._lib import call, ptr
    flat, offs = sequences if isinstance(sequences, tuple) else ragged(sequences)
    n = len(offs) - 1
    lens = np.diff(offs)
    T = int(seq_length) if seq_length is not None else int(max(int(lens.max(initial=1)) - 1, 0))
    if n == 0 or T <= 0:
        raise ValueError("need at least one sequence and a positive sequence length")
    dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    d_items = torch.from_numpy(np.ascontiguousarray(flat, dtype=np.int32)).to(dev)
    if d_items.numel() == 0:
        d_items = torch.zeros(1, dtype=torch.int32, device=dev)
    d_offs = torch.from_numpy(np.ascontiguousarray(offs, dtype=np.int64)).to(dev)
    ids = torch.empty((n, T), dtype=torch.int32, device=dev)
    tgt = torch.empty((n, T), dtype=torch.int32, device=dev)
    call("seqrec_pad_sequences", ptr(d_items), ptr(d_offs), ptr(ids), ptr(tgt), n, T,
         ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    return ids, tgt


def to_id_batch(a, n_classes=None):
    """Any accepted batch encoding -> (ids int32 (N,T) with pad=-1).

    (N,T,V) with V>1: dense one-hot rows (all-zero row = pad, model.py:335 Masking);
    (N,T,1) or (N,T): ids, negative = pad."""
    a = np.asarray(a)
    if a.ndim == 3 and a.shape[2] > 1:
        valid = (a != 0).any(axis=2)
        ids = a.argmax(axis=2).astype(np.int32)
        ids[~valid] = -1
        return ids
    if a.ndim == 3:
        a = a[:, :, 0]
    if a.ndim != 2:
        raise ValueError("expected (N,T,V) one-hot or (N,T[,1]) ids, got shape %s" % (a.shape,))
    ids = np.where(a < 0, -1, a).astype(np.int32)
    return ids


def is_one_hot(a):
    """True when every (n,t) row of a (N,T,F) array is all-zero or a single 1.0 (so the gather path applies)."""
    a = np.asarray(a)
    if a.ndim != 3:
        return False
    nz = (a != 0)
    cnt = nz.sum(axis=2)
    if cnt.max(initial=0) > 1:
        return False
    return bool(np.all(a[nz] == 1.0))
