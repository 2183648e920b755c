"""The reference's scoring metrics (utils.py:145-178) as device reductions (csrc/likelihood.cu).

`predictions` is what the reference passes: a list of per-sequence probability arrays (ragged), an (n, T) array of
per-step probabilities of left-padded batches, or -- additive -- the (n, T) device tensor the scoring kernels produce
(no host round trip).  The results are host floats, like the reference's."""
import ctypes

import numpy as np
import torch

from ._lib import call, ptr


def _pack(predictions, orig_lengths=None):
    """-> (P (n,T) float32 device tensor, right-aligned; lengths int32 device tensor or None)."""
    dev = torch.device("cuda:%d" % torch.cuda.current_device())
    if isinstance(predictions, torch.Tensor):
        P = predictions.to(dev, dtype=torch.float32).contiguous()
        lengths = None
    elif isinstance(predictions, np.ndarray) and predictions.ndim == 2:
        P = torch.from_numpy(np.ascontiguousarray(predictions, dtype=np.float32)).to(dev)
        lengths = None
    else:
        rows = [np.asarray(p, dtype=np.float32).reshape(-1) for p in predictions]
        T = max(1, max((len(r) for r in rows), default=1))
        host = np.full((len(rows), T), 1.0, dtype=np.float32)
        for i, r in enumerate(rows):
            if len(r):
                host[i, T - len(r):] = r
        P = torch.from_numpy(host).to(dev)
        lengths = np.asarray([len(r) for r in rows], dtype=np.int64)
    if orig_lengths is not None:
        lengths = np.asarray(orig_lengths).astype(np.int64)
    L = None
    if lengths is not None:
        if len(lengths) != P.shape[0]:
            raise ValueError("%d lengths for %d sequences" % (len(lengths), P.shape[0]))
        L = torch.from_numpy(lengths.astype(np.int32)).to(dev)
    return P, L, lengths


def compute_likelihood_cut(predictions, train_percent, orig_lengths=None, count_first_prob=False):
    """utils.py:145-163: (train NLL, val NLL) -- per sequence the mean -log p over its first ceil(train_percent*L) and its
    last floor((1-train_percent)*L) steps, averaged over the sequences that have such a part.  `orig_lengths` selects the
    last L steps of each (left-padded) row and, as in the reference, overrides count_first_prob."""
    assert train_percent <= 1.0, "ERROR: train_percent should be <= 1.0"
    P, L, host_len = _pack(predictions, orig_lengths)
    if host_len is not None and orig_lengths is None and (host_len == 0).any():
        # ragged input with an empty sequence: its window is empty (the packed row is all filler)
        keep = torch.from_numpy(np.flatnonzero(host_len > 0)).to(P.device)
        P, L = P.index_select(0, keep), L.index_select(0, keep)
    skip = 0 if (orig_lengths is not None or count_first_prob) else 1
    out = torch.zeros(4, dtype=torch.float64, device=P.device)
    st = ctypes.c_void_p(torch.cuda.current_stream(P.device).cuda_stream)
    call("seqrec_likelihood_cut", ptr(P), ptr(L), P.shape[0], P.shape[1], skip, float(train_percent), ptr(out), st)
    s_tr, n_tr, s_va, n_va = out.cpu().tolist()
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.float64(s_tr) / np.float64(n_tr), np.float64(s_va) / np.float64(n_va)


def compute_likelihood(predictions, count_first_prob=False, lengths=None):
    """utils.py:166-178: mean over sequences of the mean -log(clip(p, 1e-7, 1-1e-7)) over the sequence's steps (the
    first step dropped unless count_first_prob)."""
    P, L, host_len = _pack(predictions, lengths)
    if host_len is not None and (host_len == 0).any():
        keep = torch.from_numpy(np.flatnonzero(host_len > 0)).to(P.device)
        P, L = P.index_select(0, keep), L.index_select(0, keep)
    out = torch.zeros(2, dtype=torch.float64, device=P.device)
    st = ctypes.c_void_p(torch.cuda.current_stream(P.device).cuda_stream)
    call("seqrec_likelihood", ptr(P), ptr(L), P.shape[0], P.shape[1], 0 if count_first_prob else 1, ptr(out), st)
    s, n = out.cpu().tolist()
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.float64(s) / np.float64(n)
