"""Pure-Python restatement of the reference's history features: `build_xs` (datasets.py:97-113), the drivers'
`np.log(x + 1)` transform (experiments_server.py:35-36) and the `c` array of FullModelPreprocessor.transform_data
(preprocessor.py:71: c = xs[:-1]; :89, :92: Keras pad_sequences(padding='pre', truncating='pre'), reshape to (N, T, V)).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): pinned against the reference's own datasets.build_xs /
preprocessor.py through tests/golden/history_features.npz, and the checker of `seqrec_history_features`."""
import numpy as np


def build_xs(sequences, n_items, freq=False):
    """datasets.py:97-113: the running presence / count vector after every item of every sequence."""
    xs = []
    for seq in sequences:
        cur, rows = [0] * n_items, []
        for s in seq:
            cur[s] = cur[s] + 1 if freq else 1
            rows.append(cur[:])
        xs.append(rows)
    return xs


def history_block(sequences, n_items, seq_length=None, freq=False, log1p=False):
    """(N, T, V) float64: xs[:-1] of every sequence, optionally log(x + 1), left-padded with zeros and left-truncated
    (the tail is kept) to T = seq_length or the longest row count."""
    xs = build_xs(sequences, n_items, freq)
    if log1p:
        xs = [[[np.log(x + 1) for x in row] for row in rows] for rows in xs]
    rows = [x[:-1] for x in xs]
    T = seq_length if seq_length is not None else max(len(r) for r in rows)
    out = np.zeros((len(rows), T, n_items), dtype=np.float64)
    for i, r in enumerate(rows):
        r = r[-T:] if T > 0 else []
        if len(r):
            out[i, T - len(r):] = np.asarray(r, dtype=np.float64)
    return out
