"""History features of the reference's datasets.py (`build_xs`, datasets.py:97-113) -- the one function of that module
the model inputs depend on (its loaders and splitters are file parsing, out of scope).

`build_xs` keeps the reference's name, arguments and nesting (one (L, V) block of rows per sequence), so the drivers'
recipe -- `xs = build_xs(seqs, vocab, freq=True)`, `np.log(x + 1)` on every entry (experiments_server.py:33-36),
`FullModelPreprocessor.transform_data(seqs, xs=xs)` (experiments_methods.py:12-16) -- runs unchanged.

`history_features_device` is the same pipeline as ONE kernel on the GPU: the ragged corpus goes to HBM once and
`seqrec_history_features` writes the left-padded (N, T, V) float32 block `c` that transform_data returns as its third
array -- no O(N*L*V) Python lists, no float64 staging copy.  The result can be passed as the `xs` input of
RNNFullModel / NoRecurrenceModel (a device tensor is accepted wherever the numpy array is).
"""
import numpy as np

from .preprocessor import ragged


def build_xs(sequences, vocab, freq=False):
    """xs[i][j][v] = 1 if item v occurred among sequences[i][0..j] (freq=False) or how often it did (freq=True).
    Returns a list of (L_i, V) int64 arrays (the reference returns the same numbers as nested lists)."""
    V = len(vocab)
    xs = []
    for seq in sequences:
        s = np.asarray(seq, dtype=np.int64).reshape(-1)
        if s.size and (s.max() >= V or s.min() < -V):
            raise IndexError("list index out of range")            # what `xi[s]` raises in the reference
        onehot = np.zeros((len(s), V), dtype=np.int64)
        onehot[np.arange(len(s)), s] = 1                           # (a negative id wraps, as a Python list index does)
        counts = np.cumsum(onehot, axis=0)
        xs.append(counts if freq else (counts > 0).astype(np.int64))
    return xs


def history_features_device(sequences, n_items, seq_length=None, freq=False, log1p=False, device=None):
    """c = FullModelPreprocessor(seq_length).transform_data(seqs, xs=f(build_xs(seqs, vocab, freq)))[2] as a float32
    device tensor (N, T, V), with f = identity or np.log(x + 1) (log1p=True; the reference's drivers apply it to the
    counts).  `sequences`: list of id sequences or an (items, offsets) pair (see preprocessor.ragged).  T defaults to the
    longest sequence minus one, like the preprocessor."""
    import ctypes
    import torch
    from ._lib import SeqrecError, call, ptr
    flat, offs = sequences if isinstance(sequences, tuple) else ragged(sequences)
    n = len(offs) - 1
    lens = np.diff(offs)
    T = int(seq_length) if seq_length is not None else int(max(int(lens.max(initial=1)) - 1, 0))
    V = int(n_items)
    if n == 0 or T <= 0 or V <= 0:
        raise ValueError("need at least one sequence, a positive sequence length and a positive item count")
    dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    d_items = torch.from_numpy(np.ascontiguousarray(flat, dtype=np.int32)).to(dev)
    if d_items.numel() == 0:
        d_items = torch.zeros(1, dtype=torch.int32, device=dev)
    d_offs = torch.from_numpy(np.ascontiguousarray(offs, dtype=np.int64)).to(dev)
    table = None
    n_table = 0
    if log1p:
        # float64 log rounded once to float32: the value the reference's float64 features take at the Theano boundary
        n_table = int(lens.max(initial=1)) + 1 if freq else 2
        table = torch.from_numpy(np.log(np.arange(n_table, dtype=np.float64) + 1.0).astype(np.float32)).to(dev)
    out = torch.empty((n, T, V), dtype=torch.float32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    call("seqrec_history_features", ptr(d_items), ptr(d_offs), ptr(out), n, T, V, 1 if freq else 0,
         ptr(table) if table is not None else None, n_table, ptr(err),
         ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if int(err.item()) != 0:
        raise SeqrecError("history_features_device: an item id lies outside [0, %d)" % V)
    return out
