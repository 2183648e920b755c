"""Device-side batch formatting (seqrec_pad_sequences) against the host preprocessor, which tests/test_format.py pins to
the REFERENCE's own preprocessor.py through tests/golden/batch_format.npz: bit-exact ids and targets."""
import os

import numpy as np
import pytest

from seq_recommendations_b200 import preprocessor as pp
from seq_recommendations_b200 import synthetic
from seq_recommendations_b200.engine import HotPath

pytestmark = pytest.mark.gpu


def _host(seqs, V, L):
    p = pp.FullModelPreprocessor(vocab=dict(zip(range(V), range(V))), seq_length=L)
    return p.transform_ids(seqs)


def test_device_formatter_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "batch_format.npz"))
    seqs = [g["flat"][g["offs"][i]:g["offs"][i + 1]].tolist() for i in range(len(g["offs"]) - 1)]
    V = int(g["V"])
    for tag, L in (("full", None), ("trunc", 4)):
        hi, ht = _host(seqs, V, L)
        di, dt = pp.transform_ids_device(seqs, seq_length=L)
        assert hi.shape[1] == int(g["T_" + tag])
        assert np.array_equal(di.cpu().numpy(), hi) and np.array_equal(dt.cpu().numpy(), ht)
        # and the one-hot batch of the reference is exactly this id batch
        x_ref = g["x_" + tag] if ("x_" + tag) in g.files else None
        if x_ref is not None:
            assert np.array_equal(pp.to_id_batch(x_ref), hi)


def test_device_formatter_ragged_edge_cases_and_training_equivalence():
    rng = np.random.default_rng(0)
    V, T = 500, 12
    seqs = [rng.integers(0, V, size=int(L)).tolist() for L in rng.integers(0, 40, size=300)]
    seqs[0], seqs[1], seqs[2] = [], [7], list(range(T + 1))          # empty, length 1 (all-pad rows), exactly T pairs
    hi, ht = _host(seqs, V, T)
    di, dt = pp.transform_ids_device(pp.ragged(seqs), seq_length=T)
    assert np.array_equal(di.cpu().numpy(), hi) and np.array_equal(dt.cpu().numpy(), ht)
    assert np.all(hi[0] == -1) and np.all(hi[1] == -1) and np.all(hi[2] >= 0)
    # the device batch feeds the engine directly (no host copy) and gives the same loss as the host batch
    ws = synthetic.make_weights("GRU", V, 32, seed=3)
    hot = HotPath("GRU", "tanh", V, 32, V, weights=ws)
    a, _ = hot.loss_batch(hi, ht)
    b, _ = hot.loss_batch(di, dt)
    assert float(a.item()) == float(b.item())


# ---- history features on the device (seqrec_history_features) ---------------------------------------------------------
def test_device_history_features_match_reference_fixture(golden_dir):
    """Bit-exact against the reference's own build_xs + FullModelPreprocessor output (float64 -> float32 once, the cast
    the reference's arrays take at the Theano boundary): binary / counts, raw / log(x + 1), full / left-truncated."""
    from seq_recommendations_b200 import datasets
    g = np.load(os.path.join(golden_dir, "history_features.npz"))
    seqs = [g["flat"][g["offs"][i]:g["offs"][i + 1]].tolist() for i in range(len(g["offs"]) - 1)]
    V = int(g["V"])
    for freq in (False, True):
        for tag, L in (("full", None), ("trunc", int(g["T_trunc"]))):
            for name in ("raw", "log"):
                ref = g["c_%s_%s_%s" % ("freq" if freq else "bin", name, tag)]
                c = datasets.history_features_device(seqs, V, seq_length=L, freq=freq, log1p=(name == "log"))
                assert tuple(c.shape) == ref.shape
                assert np.array_equal(c.cpu().numpy(), ref.astype(np.float32)), (freq, tag, name)


def test_device_history_features_ragged_corpus_against_oracle():
    """A corpus with empty / length-1 / heavily truncated sequences and a catalog wider than a warp, against the
    pure-Python oracle; an out-of-range item raises."""
    from oracle import history
    from seq_recommendations_b200 import datasets
    from seq_recommendations_b200._lib import SeqrecError
    rng = np.random.default_rng(4)
    V, T = 333, 9
    seqs = [rng.integers(0, V, size=int(L)).tolist() for L in rng.integers(0, 40, size=120)]
    seqs[0], seqs[1], seqs[2] = [], [5], [3] * 30                   # empty, one item (all-pad row), one item 30 times
    for freq, log1p in ((True, True), (True, False), (False, False)):
        ref = history.history_block(seqs, V, seq_length=T, freq=freq, log1p=log1p)
        c = datasets.history_features_device(pp.ragged(seqs), V, seq_length=T, freq=freq, log1p=log1p)
        assert np.array_equal(c.cpu().numpy(), ref.astype(np.float32))
    assert float(c[0].abs().sum()) == 0.0 and float(c[1].abs().sum()) == 0.0
    assert float(c[2, -1, 3]) == 1.0 and float(c[2].sum()) == T
    with pytest.raises(SeqrecError):
        datasets.history_features_device([[1, 2, V, 3]], V, seq_length=4)


@pytest.mark.parametrize("V,T", [(2, 37), (4, 37), (7, 37), (16, 37), (17, 37), (32, 37), (33, 37), (40, 37),
                                 (17, 130), (5, 300)])
def test_device_history_features_every_kernel_variant(V, T):
    """Small catalogs (a group of 4 / 8 / 16 / 32 lanes per sequence, several sequences per warp with different
    lengths; presence, counts and table modes), the scalar kernel (V = 33) and the float4 kernel (V = 40), with
    truncation that drops counted rows."""
    from oracle import history
    from seq_recommendations_b200 import datasets
    rng = np.random.default_rng(100 + V)
    seqs = [rng.integers(0, V, size=int(L)).tolist() for L in rng.integers(0, 90, size=203)]
    seqs[5], seqs[6], seqs[7] = [], [V - 1], [0] * 80
    seqs[8] = rng.integers(0, V, size=T + 45).tolist()
    for freq, log1p in ((True, True), (True, False), (False, False), (False, True)):
        ref = history.history_block(seqs, V, seq_length=T, freq=freq, log1p=log1p)
        c = datasets.history_features_device(seqs, V, seq_length=T, freq=freq, log1p=log1p)
        assert np.array_equal(c.cpu().numpy(), ref.astype(np.float32)), (V, freq, log1p)
    with pytest.raises(Exception):
        datasets.history_features_device([[0, 1], [1, V, 0, 1]], V, seq_length=3)


def test_device_history_features_feed_the_history_model_like_the_host_arrays():
    """x_to_y / x_to_z model (experiments_server.py:116-191 variants): the same losses whether the history features come
    from the host recipe (numpy, float64) or from the device kernel (a CUDA tensor passed as the `xs` input)."""
    from seq_recommendations_b200 import datasets
    from seq_recommendations_b200.model import RNNFullModel
    from seq_recommendations_b200.optimizers import Adagrad
    rng = np.random.default_rng(8)
    V, T = 11, 7
    seqs = [rng.integers(0, V, size=int(L)).tolist() for L in rng.integers(2, 12, size=48)]
    vocab = dict(zip(range(V), range(V)))
    xs = [np.log(x + 1.0) for x in datasets.build_xs(seqs, vocab, freq=True)]
    x, y, c = pp.FullModelPreprocessor(vocab=vocab, seq_length=T).transform_data(seqs, xs=xs)
    c_dev = datasets.history_features_device(seqs, V, seq_length=T, freq=True, log1p=True)
    assert np.array_equal(c_dev.cpu().numpy(), c.astype(np.float32))
    losses = []
    for feats in (c, c_dev):
        m = RNNFullModel(T, V, V, z_dim=8, rnn_type="LSTM", y_to_z=True, y_to_y=False, x_to_y=True, x_to_z=True, seed=1)
        m.compile_model(optimizer=Adagrad(lr=0.05, epsilon=1e-8, clipnorm=1.0))
        np.random.seed(3)
        h = m.fit_model([x, feats], y, validation_data=([x, feats], y), n_epochs=2, batch_size=16, verbose=0)
        losses.append(h.history["loss"] + h.history["val_loss"] + [m.evaluate([x, feats], y)[1][0]])
    assert np.allclose(losses[0], losses[1], rtol=1e-5), losses
