"""End-to-end parity through the reference-facing surface (model.py mirror): fit_model / evaluate / predict / weights,
driven exactly the way experiments_methods.py:19-50 drives the reference.  Run with -m gpu on a B200."""
import os

import numpy as np
import pytest
import torch

from oracle import keras_semantics as ks
from seq_recommendations_b200 import callbacks as cb
from seq_recommendations_b200 import model as M
from seq_recommendations_b200 import preprocessor as pp
from seq_recommendations_b200.optimizers import Adagrad

from gpu_util import as_t, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_ids(V, T, n, seed):
    from seq_recommendations_b200 import synthetic
    return synthetic.make_batch(V, T, n, seed=seed, min_len=1)


def mc_data(n=120, T=20):
    g = np.load(os.path.join(ROOT, "tests", "golden", "mc_sequences.npz"))
    seqs = [g["flat"][g["offs"][i]:g["offs"][i + 1]].tolist() for i in range(n)]
    V = 17
    vocab = dict(zip(range(V), range(V)))
    xs = [[[0] * V for _ in s] for s in seqs]
    p = pp.FullModelPreprocessor(vocab=vocab, pad_value=0., seq_length=T)
    x, y, _ = p.transform_data(seqs, xs=xs)
    return x, y, V, seqs


def oracle_fit(ora, x_ids, y_ids, epochs, batch_size, lr, seed, val=None):
    """Keras fit loop restated on the oracle."""
    np.random.seed(seed)
    n = len(x_ids)
    index = np.arange(n)
    hist = {"loss": [], "val_loss": []}
    for _ in range(epochs):
        np.random.shuffle(index)
        tot = 0.0
        for lo in range(0, n, batch_size):
            sel = index[lo:lo + batch_size]
            i, t = as_t(x_ids[sel]), as_t(y_ids[sel])
            l, _, _ = ora.train_step(i, t, i >= 0, lr=lr, epsilon=1e-8, clipnorm=1.0)
            tot += float(l) * len(sel)
        hist["loss"].append(tot / n)
        if val is not None:
            hist["val_loss"].append(oracle_eval(ora, val[0], val[1], batch_size))
    return hist


def oracle_eval(ora, x_ids, y_ids, batch_size):
    tot = 0.0
    for lo in range(0, len(x_ids), batch_size):
        i, t = as_t(x_ids[lo:lo + batch_size]), as_t(y_ids[lo:lo + batch_size])
        with torch.no_grad():
            l, _, _ = ora.loss(i, t, i >= 0)
        tot += float(l) * len(i)
    return tot / len(x_ids)


@pytest.mark.parametrize("rnn_type,act", [("LSTM", "relu"), ("GRU", "tanh"), ("simpleRNN", "relu")])
def test_fit_evaluate_predict_like_run_model(rnn_type, act):
    x, y, V, _ = mc_data()
    xv, yv = x[100:], y[100:]
    x, y = x[:100], y[:100]
    model = M.RNNFullModel(timesteps=x.shape[1], x_dim=V, y_dim=V, z_dim=24, model_name="ytoz", rnn_type=rnn_type,
                           z_to_z_activation=act, y_to_z=True, y_to_y=False, x_to_y=False, x_to_z=False, seed=3)
    assert model.n_classes == V and model.rnn_type == rnn_type
    ws = model.model.get_weights()
    G = {"LSTM": 4, "GRU": 3, "simpleRNN": 1}[rnn_type]
    assert [w.shape for w in ws] == [(V, G * 24), (24, G * 24), (G * 24,), (24, V)]
    ora = ks.Model(rnn_type, act, ws, dtype=torch.float64)
    model.compile_model(loss="categorical_crossentropy", metrics=[], optimizer=Adagrad(lr=0.05, epsilon=1e-08,
                                                                                        decay=0.0, clipnorm=1.))
    np.random.seed(11)
    hist = model.fit_model([x], y, validation_data=([xv], yv), n_epochs=3, batch_size=32, verbose=0)
    ref = oracle_fit(ora, pp.to_id_batch(x), pp.to_id_batch(y), 3, 32, 0.05, 11,
                     val=(pp.to_id_batch(xv), pp.to_id_batch(yv)))
    assert np.allclose(hist.history["loss"], ref["loss"], rtol=TOL)
    assert np.allclose(hist.history["val_loss"], ref["val_loss"], rtol=TOL)
    assert hist.history["categorical_crossentropy"] == hist.history["loss"]
    for w, r in zip(model.model.get_weights(), ora.numpy_weights()):
        assert rel_err(w, r) <= TOL
    names, scores = model.evaluate([xv], yv, batch_size=7)
    assert names == ["loss", "categorical_crossentropy"] and len(scores) == 2 and scores[0] == scores[1]
    assert abs(scores[0] - oracle_eval(ora, pp.to_id_batch(xv), pp.to_id_batch(yv), 7)) <= TOL * scores[0]
    probs = model.predict([xv], batch_size=10, verbose=0)
    i = as_t(pp.to_id_batch(xv))
    rp = ora.predict_proba(ids=i, mask=i >= 0).numpy()
    assert probs.shape == rp.shape and probs.dtype == np.float32 and np.abs(probs - rp).max() <= TOL
    # id-format batches give the same numbers as the dense one-hot batches
    probs_ids = model.predict(pp.to_id_batch(xv), batch_size=10, verbose=0)
    assert np.array_equal(probs, probs_ids)
    # the reference's scoring consumer (model.py:106-112) on the fused target-prob path
    py = model.predict_target_prob([xv], yv)
    want = np.clip(np.max(rp * yv, axis=2), ks.EPS32, ks.ONE_MINUS_EPS32)
    assert np.abs(py - want).max() <= TOL


def test_callbacks_early_stopping_checkpoint_and_val_cut(tmp_path):
    x, y, V, seqs = mc_data(60, 12)
    model = M.RNNFullModel(timesteps=12, x_dim=V, y_dim=V, z_dim=8, model_name="m", rnn_type="LSTM", y_to_z=True,
                           y_to_y=False, x_to_y=False, x_to_z=False, seed=5)
    model.compile_model(optimizer=Adagrad(lr=0.0, epsilon=1e-08, decay=0.0, clipnorm=1.))   # lr 0: loss cannot improve
    lengths = np.minimum([len(s) - 1 for s in seqs], 12)
    keep = lengths > 0
    val_hist = M.ValLossHistoryCut(([x[keep]], y[keep]), lengths[keep])
    ckpt = cb.ModelCheckpoint(str(tmp_path / "w.{epoch:02d}-{val_loss:.2f}.npz"), monitor="val_loss",
                              save_weights_only=True, save_best_only=True)
    stop = cb.EarlyStopping(monitor="val_loss", min_delta=0, patience=2, verbose=0, mode="auto")
    hist = model.fit_model([x], y, validation_data=([x], y), n_epochs=50, batch_size=16, verbose=0,
                           callbacks=[val_hist, ckpt, stop])
    assert len(hist.history["loss"]) == 4                     # epoch 0 best, then patience 2 (+1 Keras off-by-one)
    assert len(val_hist.val_lossses) == 4 and np.isfinite(val_hist.val_lossses).all()
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 1 and files[0].startswith("w.00-")
    before = model.model.get_weights()
    model.load_model_weights(str(tmp_path / files[0]))
    for a, b in zip(before, model.model.get_weights()):
        assert np.array_equal(a, b)


def test_weight_surface_layers_and_freezing(tmp_path):
    V = 9
    model = M.RNNFullModel(timesteps=5, x_dim=V, y_dim=V, z_dim=6, model_name="ws", rnn_type="LSTM", y_to_z=True,
                           y_to_y=False, x_to_y=False, x_to_z=False, toy_bias=False, seed=1)
    rnn = model.get_layer_weights("z_to_z_output")
    assert [w.shape for w in rnn] == [(V, 24), (6, 24), (24,)] and np.all(rnn[2][6:12] == 1.0)   # unit_forget_bias
    assert [w.shape for w in model.get_layer_weights("to_y_output")] == [(6, V)]
    assert [w.shape for w in model.get_layer_weights(3)] == [(V, 24), (6, 24), (24,)]
    new = [np.full_like(w, 0.5) for w in rnn]
    model.set_layer_weights("z_to_z_output", new)
    assert np.all(model.get_layer_weights("z_to_z_output")[0] == 0.5)
    model.set_layer_weights_trainable("to_y_output", trainable=False)
    tr, ntr = model.get_model_weights()
    assert ntr == ["W_out"] and "W_out" not in tr
    model.save_model_weights(str(tmp_path) + "/")
    other = M.RNNFullModel(timesteps=5, x_dim=V, y_dim=V, z_dim=6, model_name="o", rnn_type="LSTM", y_to_z=True,
                           y_to_y=False, x_to_y=False, x_to_z=False, seed=2)
    other.load_model_weights(str(tmp_path) + "/ws.npz")
    for a, b in zip(model.model.get_weights(), other.model.get_weights()):
        assert np.array_equal(a, b)
    acts = model.get_activations("z_to_z_output", [np.eye(V)[None, [1, 2, 3, 4, 0]]], ["y_input"])
    assert acts.shape == (1, 5, 6)


def test_baseline_model_with_history_features():
    """RNNBaseline (model.py:241-258): [onehot || xs] features through the K2 projection, logits with bias."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "batch_format.npz"))
    xb, yb = g["xb_xs"], g["yb_xs"]
    model = M.RNNBaseline(timesteps=xb.shape[1], features=xb.shape[2], n_classes=4, rnn_type="LSTM", z_dim=5, seed=7)
    ws = model.model.get_weights()
    assert [w.shape for w in ws] == [(8, 20), (5, 20), (20,), (5, 4), (4,)]
    ora = ks.Model("LSTM", "relu", ws, dtype=torch.float64)
    model.compile_model(optimizer=Adagrad(lr=0.1, epsilon=1e-08, decay=0.0, clipnorm=1.))
    hist = model.fit_model(xb, yb, n_epochs=2, batch_size=5, verbose=0)
    t = as_t(pp.to_id_batch(yb))
    xd = torch.tensor(xb, dtype=torch.float64)
    want = []
    for _ in range(2):
        l, _, _ = ora.train_step(None, t, ks.derive_mask(xd), lr=0.1, epsilon=1e-8, clipnorm=1.0, x_dense=xd)
        want.append(float(l))
    assert np.allclose(hist.history["loss"], want, rtol=TOL)
    for w, r in zip(model.model.get_weights(), ora.numpy_weights()):
        assert rel_err(w, r) <= TOL
    # plain one-hot features take the gather path and agree with the dense path's oracle too
    xp_, yp_ = g["xb_plain"], g["yb_plain"]
    m2 = M.RNNBaseline(timesteps=xp_.shape[1], features=4, n_classes=4, rnn_type="simpleRNN", z_dim=3, seed=8)
    o2 = ks.Model("simpleRNN", "relu", m2.model.get_weights(), dtype=torch.float64)
    p = m2.predict(xp_, batch_size=2, verbose=0)
    xd = torch.tensor(xp_, dtype=torch.float64)
    assert np.abs(p - o2.predict_proba(x_dense=xd, mask=ks.derive_mask(xd)).numpy()).max() <= TOL


def test_unsupported_branches_raise():
    M.RNNFullModel(timesteps=5, x_dim=4, y_dim=4)                 # reference defaults (y_to_y / x_to_y branches) build
    M.NoRecurrenceModel(5, 4, 4)
    with pytest.raises(NotImplementedError):                      # what no driver of the reference ever passes
        M.RNNFullModel(timesteps=5, x_dim=4, y_dim=4, toy_regularizer=object())
    with pytest.raises(NotImplementedError):
        M.NoRecurrenceModel(5, 4, 4, embed_y=True)
    with pytest.raises(ValueError):
        M.RNNFullModel(timesteps=5, x_dim=4, y_dim=4, y_to_z=False, x_to_z=False)
    m = M.RNNFullModel(timesteps=5, x_dim=4, y_dim=4, y_to_z=True, y_to_y=False, x_to_y=False, x_to_z=False)
    m.compile_model()                                              # 'adam' string is accepted at compile time ...
    with pytest.raises(NotImplementedError):
        m.fit_model(np.zeros((2, 5, 4)), np.zeros((2, 5, 4)), verbose=0)   # ... but the reference never trains with it


def test_device_likelihood_metrics_match_golden_and_oracle(golden_dir):
    """compute_likelihood / compute_likelihood_cut (utils.py:145-178) as device reductions: against the fixture produced
    by the REFERENCE's own utils.py (tests/golden/likelihood.npz) and against the numpy restatement on random ragged and
    padded inputs; probabilities travel as float32 (what the scoring kernels produce), hence 1e-6."""
    import os
    from oracle import likelihood as ol
    g = np.load(os.path.join(golden_dir, "likelihood.npz"))
    preds = [g["preds"][g["offs"][i]:g["offs"][i + 1]] for i in range(len(g["offs"]) - 1)]
    assert np.isclose(M.compute_likelihood(preds, count_first_prob=False), g["ll"], rtol=1e-6)
    assert np.isclose(M.compute_likelihood(preds, count_first_prob=True), g["ll_first"], rtol=1e-6)
    tr, va = M.compute_likelihood_cut(preds, 0.7, count_first_prob=False)
    assert np.isclose(tr, g["cut_tr"], rtol=1e-6) and np.isclose(va, g["cut_va"], rtol=1e-6)
    tr, va = M.compute_likelihood_cut(g["padded"], 0.7, orig_lengths=g["lengths"])
    assert np.isclose(tr, g["cut_tr_l"], rtol=1e-6) and np.isclose(va, g["cut_va_l"], rtol=1e-6)
    rng = np.random.default_rng(2)
    n, T = 300, 37
    lengths = rng.integers(1, T + 1, size=n)
    padded = np.full((n, T), 1e-7, dtype=np.float32)
    for i, L in enumerate(lengths):
        padded[i, T - L:] = rng.uniform(1e-4, 0.999, size=L)
    for tp in (0.7, 0.5, 1.0):
        a = M.compute_likelihood_cut(padded, tp, orig_lengths=lengths)
        b = ol.compute_likelihood_cut(padded.astype(np.float64), tp, orig_lengths=lengths)
        assert np.allclose(a, b, rtol=1e-6, equal_nan=True), (tp, a, b)
        a = M.compute_likelihood_cut(torch.tensor(padded).cuda(), tp, orig_lengths=lengths)      # device tensor in
        assert np.allclose(a, b, rtol=1e-6, equal_nan=True)
    ragged = [padded[i, T - L:] for i, L in enumerate(lengths)]
    for first in (False, True):
        assert np.isclose(M.compute_likelihood(ragged, count_first_prob=first),
                          ol.compute_likelihood([r.astype(np.float64) for r in ragged], count_first_prob=first), rtol=1e-6)
        a = M.compute_likelihood_cut(ragged, 0.7, count_first_prob=first)
        b = ol.compute_likelihood_cut([r.astype(np.float64) for r in ragged], 0.7, count_first_prob=first)
        assert np.allclose(a, b, rtol=1e-6)


def test_state_checkpoint_resumes_with_adagrad_accumulators(tmp_path):
    """save_state / load_state (weights + Adagrad accumulators + epoch): a resumed model continues EXACTLY like the
    uninterrupted one (to summation order); a weights-only checkpoint (the reference's, model.py:201-218) restarts the accumulators and does
    not.  A path that claims to be HDF5 gets '.npz' appended; a real HDF5 file is refused with a clear message."""
    V, H, T, B = 400, 32, 8, 64
    ids, tgt = make_ids(V, T, 4 * B, seed=3)

    def build():
        m = M.RNNFullModel(T, V, V, z_dim=H, rnn_type="GRU", z_to_z_activation="tanh", y_to_y=False, x_to_y=False, seed=5)
        m.compile_model(optimizer=Adagrad(lr=0.05, epsilon=1e-8, clipnorm=1.0))
        return m

    def steps(m, lo, hi):
        return [m.model.train_on_batch(ids[b * B:(b + 1) * B], tgt[b * B:(b + 1) * B]) for b in range(lo, hi)]

    a = build()
    steps(a, 0, 2)
    ck = str(tmp_path / "state.hdf5")
    a.model.save_state(ck, epoch=7)
    assert os.path.exists(ck + ".npz") and not os.path.exists(ck)
    wk = str(tmp_path / "weights_only")
    a.model.save_weights(wk)
    rest = steps(a, 2, 4)
    b = build()
    assert b.model.load_state(ck) == 7
    # same continuation, up to the summation order of the logits kernels' red.global accumulations (not fixed from run
    # to run: a few ulps in the gradients).  Adagrad's first update of a coordinate is lr * g / (|g| + eps): where a
    # gradient is a cancellation residue those ulps are a finite relative error, so a handful of coordinates may move
    # visibly -- the bulk must agree closely, and the losses must.
    assert np.allclose(steps(b, 2, 4), rest, rtol=1e-5)
    for x, y in zip(a.model.get_weights(), b.model.get_weights()):
        assert np.mean(~np.isclose(x, y, rtol=1e-4, atol=1e-6)) < 0.01
        assert np.abs(x - y).max() <= 2.1 * 0.05                     # never more than the two steps' lr
    c = build()
    c.model.load_weights(wk)
    assert steps(c, 2, 4)[1] != rest[1]                             # accumulators restarted: a different trajectory
    fake = tmp_path / "keras.h5"
    fake.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(IOError, match="HDF5"):
        c.model.load_weights(str(fake))


def test_fit_epoch_from_hbm_equals_fit_from_host_batches(monkeypatch):
    """fit keeps the id arrays in HBM and gathers each shuffled batch on the device; the losses must equal the host-slicing
    path batch for batch (same shuffle)."""
    V, H, T, B = 300, 24, 7, 32
    ids, tgt = make_ids(V, T, 5 * B + 7, seed=9)
    hist = []
    for resident in (True, False):
        m = M.RNNFullModel(T, V, V, z_dim=H, rnn_type="LSTM", y_to_y=False, x_to_y=False, seed=2)
        m.compile_model(optimizer=Adagrad(lr=0.05, epsilon=1e-8, clipnorm=1.0))
        if not resident:
            monkeypatch.setattr(type(m.model), "_resident", lambda self, a: None)
        np.random.seed(11)
        h = m.fit_model(ids, tgt, validation_data=(ids, tgt), n_epochs=2, batch_size=B, verbose=0)
        hist.append((h.history["loss"], h.history["val_loss"], m.model.get_weights()))
    # (equal up to the summation order of the logits kernels' red.global accumulations, which is not fixed from run to
    #  run; see test_state_checkpoint_resumes_with_adagrad_accumulators for why a few weights may move visibly)
    assert np.allclose(hist[0][0], hist[1][0], rtol=1e-5) and np.allclose(hist[0][1], hist[1][1], rtol=1e-5)
    for x, y in zip(hist[0][2], hist[1][2]):
        assert np.mean(~np.isclose(x, y, rtol=1e-4, atol=1e-6)) < 0.01
