"""The five BASELINE.json configurations through the CUDA path.  cfg1 and cfg2-shaped problems are compared with the
oracle at sizes it finishes in seconds; the large ones (H=256, V up to 1M, T up to 200) run at reduced batch with the
size-independent properties the domain offers (the optimisation step lowers the loss on its own batch, row-sparse update
leaves untouched rows untouched, probabilities are normalised, top-k is sorted and consistent with target scores)."""
import numpy as np
import pytest
import torch

from oracle import keras_semantics as ks
from seq_recommendations_b200 import synthetic
from seq_recommendations_b200.engine import HotPath

from gpu_util import as_t, rel_err

pytestmark = pytest.mark.gpu


def test_cfg1_msnbc_lstm100_full_size_matches_oracle(golden_dir):
    """configs[0]: MSNBC-like, V=17, LSTM-100/relu (the reference's real cell, experiments_server.py:40-42), B=100,
    T=50, sequences from the reference's own MCSampler (tests/golden/mc_sequences.npz)."""
    import os
    from seq_recommendations_b200.preprocessor import FullModelPreprocessor
    g = np.load(os.path.join(golden_dir, "mc_sequences.npz"))
    seqs = [g["flat"][g["offs"][i]:g["offs"][i + 1]].tolist() for i in range(100)]
    ids, tgt = FullModelPreprocessor(vocab=dict(zip(range(17), range(17))), seq_length=50).transform_ids(seqs)
    for cell, act in (("LSTM", "relu"), ("GRU", "tanh")):
        ws = synthetic.make_weights(cell, 17, 100, seed=1)
        hot = HotPath(cell, act, 17, 100, 17, weights=ws)
        hot.set_optimizer("adagrad", lr=0.01, epsilon=1e-8, clipnorm=1.0)
        ora = ks.Model(cell, act, ws, dtype=torch.float64)
        for _ in range(2):
            loss = float(hot.train_batch(ids, tgt).item())
            rl, _, _ = ora.train_step(as_t(ids), as_t(tgt), as_t(ids) >= 0, lr=0.01, epsilon=1e-8, clipnorm=1.0)
            assert abs(loss - float(rl)) <= 1e-4 * abs(float(rl))
        for w, r in zip(hot.get_weights(), ora.numpy_weights()):
            assert rel_err(w, r) <= 1e-4


def _step_properties(cfg, B, steps=2, V=None, tc="x3"):
    V = V or cfg["V"]
    H, T = cfg["H"], cfg["T"]
    ws = synthetic.make_weights(cfg["cell"], V, H, seed=0)
    ids, tgt = synthetic.make_batch(V, T, B, seed=0)
    hot = HotPath(cfg["cell"], cfg["act"], V, H, V, weights=ws, tc=tc)
    hot.set_optimizer("adagrad", lr=0.05, epsilon=1e-8, clipnorm=1.0)
    losses = [float(hot.train_batch(ids, tgt).item()) for _ in range(steps)]
    assert np.all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert abs(losses[0] - np.log(V)) < 0.5                       # random init: close to the uniform-guess loss
    used = np.unique(ids[ids >= 0])
    W_in = hot.W_in.cpu().numpy()
    untouched = np.setdiff1d(np.arange(min(V, 5000)), used)[:200]
    assert np.array_equal(W_in[untouched], ws[0][untouched])      # row-sparse update
    assert float(hot.dW_in.abs().max().item()) == 0.0 and int(hot.touched.sum().item()) == 0
    return hot, ids, tgt, losses


def test_cfg3_lstm256_50k_reduced_batch():
    """configs[2]: V=50k, LSTM-256, T=100 (B reduced 1024 -> 32): tensor-core forward and backward (Hk = 256)."""
    cfg = synthetic.CONFIGS["cfg3_lstm256_50k"]
    hot, ids, tgt, _ = _step_properties(cfg, B=32)
    w = hot.work(32, cfg["T"])
    assert w.tc["fwd"] and w.tc["bwd"]
    # x3 forward statistics agree with the exact-fp32 SIMT kernels on the same weights
    ref = HotPath(cfg["cell"], cfg["act"], cfg["V"], cfg["H"], cfg["V"], weights=hot.get_weights(), tc="off")
    a, _ = hot.loss_batch(ids, tgt)
    b, _ = ref.loss_batch(ids, tgt)
    assert abs(float(a.item()) / float(b.item()) - 1) <= 1e-5


def test_cfg4_gru256_1m_items_reduced_batch():
    """configs[3]: 1M-item catalog, GRU-256 (B reduced to 16, T to 10; tensor-core logits forward and backward): 3 GB input table, row-sparse everything."""
    cfg = dict(synthetic.CONFIGS["cfg4_gru256_1m"], T=10)
    _step_properties(cfg, B=16, steps=2, tc="x3")


def test_cfg5_scoring_topk_and_target_prob_reduced_batch():
    """configs[4]: T=200, V=100k, GRU-256 inference (B reduced 4096 -> 48): top-20 at the last step, p(true item) for
    every step."""
    cfg = synthetic.CONFIGS["cfg5_score_gru256_100k"]
    V, H, T, B, k = cfg["V"], cfg["H"], cfg["T"], 48, 20
    ws = synthetic.make_weights(cfg["cell"], V, H, seed=0)
    ws[3] = ws[3] * 20.0                                          # non-trivial ranking
    ids, tgt = synthetic.make_batch(V, T, B, seed=1)
    hot = HotPath(cfg["cell"], cfg["act"], V, H, V, weights=ws)
    top_i, top_p = hot.topk_batch(ids, k, last_step_only=True)
    top_i, top_p = top_i.cpu().numpy(), top_p.cpu().numpy()
    assert top_i.shape == (B, k) and top_i.min() >= 0 and top_i.max() < V
    assert np.all(np.diff(top_p, axis=1) <= 0) and np.all(top_p > 0) and np.all(top_p.sum(1) <= 1 + 1e-5)
    assert all(len(set(r)) == k for r in top_i.tolist())
    # exact check of one row against a float64 recomputation of its logits
    hid = hot.hidden_batch(ids)[:, -1].double()
    z = hid[0] @ hot.W_out.double()
    want = torch.topk(z, k).indices.cpu().numpy()
    assert np.array_equal(np.sort(want), np.sort(top_i[0]))
    p = torch.softmax(z, 0)[torch.as_tensor(top_i[0].astype(np.int64), device=z.device)].cpu().numpy()
    assert np.abs(p - top_p[0]).max() <= 1e-4 * p.max() + 1e-7
    # p(true next item): consistent with the top-k probabilities wherever the target is in the top-k
    py = hot.target_prob_batch(ids, tgt).cpu().numpy()
    assert py.shape == (B, T) and np.all(py >= np.float32(ks.EPS32)) and np.all(py <= 1)
    last_t = tgt[:, -1]
    for b in range(B):
        hit = np.nonzero(top_i[b] == last_t[b])[0]
        if len(hit):
            assert abs(py[b, -1] - top_p[b, hit[0]]) <= 1e-4 * top_p[b, hit[0]] + 1e-7
