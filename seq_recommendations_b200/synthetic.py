"""Seeded synthetic workloads of the BASELINE.json configs: id-format batches (B,T) int32, pad = -1,
LEFT padded like preprocessor.py:16-20, inputs s[:-1] and targets s[1:]."""
import numpy as np

CONFIGS = {
    # name: cell, act, V, H, T, B
    "cfg1_msnbc_lstm100": dict(cell="LSTM", act="relu", V=17, H=100, T=50, B=100),
    "cfg1_msnbc_gru100": dict(cell="GRU", act="tanh", V=17, H=100, T=50, B=100),
    "cfg2_reddit_gru128": dict(cell="GRU", act="tanh", V=10000, H=128, T=50, B=256),
    "cfg3_lstm256_50k": dict(cell="LSTM", act="tanh", V=50000, H=256, T=100, B=1024),
    "cfg4_gru256_1m": dict(cell="GRU", act="tanh", V=1000000, H=256, T=50, B=1024),
    "cfg5_score_gru256_100k": dict(cell="GRU", act="tanh", V=100000, H=256, T=200, B=4096),
}


def zipf_items(rng, V, size, s=1.1):
    """Zipf(s) over a finite catalog: p(i) ~ 1/(i+1)^s, via inverse CDF."""
    p = 1.0 / np.power(np.arange(1, V + 1, dtype=np.float64), s)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return np.searchsorted(cdf, rng.random(size), side="left").astype(np.int32)


def make_batch(V, T, B, seed=0, min_len=None, max_len=None, zipf_s=1.1, all_valid=False):
    """Returns ids (B,T) int32 and tgt (B,T) int32; a sequence of L+1 items fills the last L steps."""
    rng = np.random.default_rng(seed)
    if min_len is None:
        min_len = max(1, T // 2)
    if max_len is None:
        max_len = T
    L = np.full(B, T) if all_valid else rng.integers(min_len, max_len + 1, size=B)
    items = zipf_items(rng, V, (B, T + 1), zipf_s)
    ids = np.full((B, T), -1, dtype=np.int32)
    tgt = np.full((B, T), -1, dtype=np.int32)
    for b in range(B):
        l = int(L[b])
        ids[b, T - l:] = items[b, :l]
        tgt[b, T - l:] = items[b, 1:l + 1]
    return ids, tgt


def glorot_uniform(rng, shape):
    lim = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def make_weights(cell, V, H, seed=0, out_bias=False, F=None):
    """glorot-uniform W_in / W_out, orthogonal U per gate block, zero bias (+1 forget) -- SURVEY §8(d)."""
    rng = np.random.default_rng(seed + 1000)
    G = {"simpleRNN": 1, "LSTM": 4, "GRU": 3}[cell]
    F = V if F is None else F
    W_in = glorot_uniform(rng, (F, G * H))
    blocks = []
    for _ in range(G):
        q, r = np.linalg.qr(rng.standard_normal((H, H)))
        blocks.append((q * np.sign(np.diag(r))).astype(np.float32))
    U = np.concatenate(blocks, axis=1)
    b = np.zeros(G * H, dtype=np.float32)
    if cell == "LSTM":
        b[H:2 * H] = 1.0
    ws = [W_in, U, b, glorot_uniform(rng, (H, V))]
    if out_bias:
        ws.append(np.zeros(V, dtype=np.float32))
    return ws
