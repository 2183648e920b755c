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


# ---------------------------------------------------------------------------------------------------------------------
# The named BASELINE shapes against the float64 ORACLE (not against the CUDA path itself): loss and all four gradients,
# norm-wise (1e-4, north_star) AND element-wise (|a - b| <= 1e-4 * max|b|: a localised error cannot hide in a norm).
# cfg2 runs at its full size; cfg3 / cfg4 / cfg5 keep their catalog width, hidden size and sequence length and reduce
# only the batch (factor stated per test) so that the CPU oracle finishes in seconds.
def _elementwise_ok(a, b, tol=1e-4):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max()) <= tol * float(np.abs(b).max()) + 1e-12


def _oracle_grads_compact(cell, act, ws, ids, tgt):
    """float64 oracle loss / gradients with the INPUT table restricted to the rows the batch touches (the dense
    (V, G*H) float64 gradient of a 1M-row table would be 6 GB of zeros); returns the used row ids with it."""
    used = np.unique(ids[ids >= 0])
    remap = np.full(int(ids.max()) + 2, -1, dtype=np.int64)
    remap[used] = np.arange(len(used))
    cid = np.where(ids >= 0, remap[np.maximum(ids, 0)], -1)
    ora = ks.Model(cell, act, [ws[0][used]] + list(ws[1:]), dtype=torch.float64)
    loss, gs = ora.grads(as_t(cid), as_t(tgt), as_t(ids) >= 0)
    return float(loss), [g.numpy() for g in gs], used


def _check_grads_vs_oracle(cfg, B, T=None, seed=3):
    V, H = cfg["V"], cfg["H"]
    T = T or cfg["T"]
    ws = synthetic.make_weights(cfg["cell"], V, H, seed=seed)
    ids, tgt = synthetic.make_batch(V, T, B, seed=seed + 1)
    hot = HotPath(cfg["cell"], cfg["act"], V, H, V, weights=ws, tc="x3")
    w = hot.work(B, T)
    assert w.tc["fwd"] and w.tc["bwd"], "the named shapes must run on the tensor-core logits kernels"
    loss, grads, extra = hot.grad_batch(ids, tgt)
    rl, rg, used = _oracle_grads_compact(cfg["cell"], cfg["act"], ws, ids, tgt)
    assert abs(loss - rl) <= 1e-4 * abs(rl), (loss, rl)
    assert np.array_equal(np.sort(extra["rows"]), used)              # exactly the touched rows carry a gradient
    names = ["dW_in", "dU", "db", "dW_out"]
    mine = [grads[0][used]] + grads[1:]
    for n, g, r in zip(names, mine, rg):
        assert rel_err(g, r) <= 1e-4, (n, rel_err(g, r))
        assert _elementwise_ok(g, r), (n, float(np.abs(g - r).max()), float(np.abs(r).max()))
    untouched = np.setdiff1d(np.arange(min(V, 20000)), used)[:500]
    assert not grads[0][untouched].any()


def test_cfg2_full_size_gradients_match_float64_oracle():
    """configs[1] at FULL size: V=10k, GRU-128, T=50, B=256 (12 800 tokens x 10 000 items)."""
    _check_grads_vs_oracle(synthetic.CONFIGS["cfg2_reddit_gru128"], B=256)


def test_cfg3_shape_gradients_match_float64_oracle():
    """configs[2]: V=50k, LSTM-256, T=100; batch reduced 1024 -> 8 (factor 128)."""
    _check_grads_vs_oracle(synthetic.CONFIGS["cfg3_lstm256_50k"], B=8)


def test_cfg4_shape_gradients_match_float64_oracle():
    """configs[3]: V=1M, GRU-256, T=50; batch reduced 1024 -> 4 (factor 256)."""
    _check_grads_vs_oracle(synthetic.CONFIGS["cfg4_gru256_1m"], B=4)


def test_cfg5_shape_scoring_matches_float64_oracle():
    """configs[4]: V=100k, GRU-256, T=200; batch reduced 4096 -> 4 (factor 1024): p(true item) at every step within
    1e-4 relative, top-20 of the last step bit-exact wherever the oracle's ranking is decided by more than fp32 noise."""
    cfg = synthetic.CONFIGS["cfg5_score_gru256_100k"]
    V, H, T, B, k = cfg["V"], cfg["H"], cfg["T"], 4, 20
    ws = synthetic.make_weights(cfg["cell"], V, H, seed=0)
    ws[3] = ws[3] * 20.0                                          # a ranking that is not flat
    ids, tgt = synthetic.make_batch(V, T, B, seed=1)
    hot = HotPath(cfg["cell"], cfg["act"], V, H, V, weights=ws)
    ora = ks.Model(cfg["cell"], cfg["act"], ws, dtype=torch.float64)
    ref = ora.predict_proba(ids=as_t(ids), mask=as_t(ids) >= 0)
    py = hot.target_prob_batch(ids, tgt).cpu().numpy()
    rpy = ks.target_prob(ref, as_t(tgt), as_t(ids) >= 0).numpy()
    assert np.abs(py / rpy - 1).max() <= 1e-4
    ti, tp = hot.topk_batch(ids, k, last_step_only=True)
    want = ks.topk_items(ref[:, -1], k)
    refp = np.take_along_axis(ref[:, -1].numpy(), want.astype(np.int64), axis=1)
    got = ti.cpu().numpy()
    for b in range(B):
        gaps = np.abs(np.diff(refp[b])) / refp[b, :-1]
        if gaps.min() > 1e-5:
            assert np.array_equal(got[b], want[b])
        else:
            assert np.array_equal(np.sort(got[b]), np.sort(want[b]))
    assert np.abs(tp.cpu().numpy() / np.take_along_axis(ref[:, -1].numpy(), got.astype(np.int64), axis=1) - 1).max() <= 1e-4
