"""Host-side batch format, likelihood metrics and config generators against fixtures produced by the REFERENCE's own
preprocessor.py / utils.py / sampler.py (tests/golden/make_golden.py)."""
import os

import numpy as np

from seq_recommendations_b200 import model as m
from seq_recommendations_b200 import preprocessor as pp
from seq_recommendations_b200 import synthetic


def unragged(flat, offs):
    return [flat[offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]


def build_xs(seqs, V, freq=False):
    xs = []
    for s in seqs:
        cur, rows = [0] * V, []
        for it in s:
            cur[it] = cur[it] + 1 if freq else 1
            rows.append(cur[:])
        xs.append(rows)
    return xs


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_full_model_preprocessor_matches_reference(golden_dir):
    g = load(golden_dir, "batch_format.npz")
    seqs = unragged(g["flat"], g["offs"])
    V = int(g["V"])
    vocab = dict(zip(range(V), range(V)))
    xs = build_xs(seqs, V)
    for tag, L in (("full", None), ("trunc", 4)):
        p = pp.FullModelPreprocessor(vocab=vocab, pad_value=0., seq_length=L)
        x, y, c = p.transform_data(seqs, xs=xs)
        assert p.seq_length == int(g["T_" + tag])
        for mine, ref in ((x, g["x_" + tag]), (y, g["y_" + tag]), (c, g["c_" + tag])):
            assert mine.dtype == ref.dtype == np.float64 and mine.shape == ref.shape
            assert np.array_equal(mine, ref)
    p = pp.FullModelPreprocessor(vocab=vocab, pad_value=0., seq_length=None)
    _, _, c = p.transform_data(seqs, xs=build_xs(seqs, V, freq=True))
    assert np.array_equal(c, g["c_freq"])


def test_sparse_id_format_against_reference(golden_dir):
    """Same ids at the same (left-padded) positions as the reference's sparse=True output; pads are -1 here because the
    reference's 0 pad collides with item 0 (documented deviation)."""
    g = load(golden_dir, "batch_format.npz")
    seqs = unragged(g["flat"], g["offs"])
    V = int(g["V"])
    vocab = dict(zip(range(V), range(V)))
    p = pp.FullModelPreprocessor(vocab=vocab, pad_value=0., seq_length=None, sparse=True)
    x, y, c = p.transform_data(seqs, xs=build_xs(seqs, V))
    dense_mask = (g["x_full"] != 0).any(axis=2)
    assert x.shape == g["x_sparse"].shape and x.dtype == np.float64
    assert np.array_equal(x[:, :, 0] >= 0, dense_mask)
    assert np.array_equal(x[:, :, 0][dense_mask], g["x_sparse"][:, :, 0][dense_mask])
    assert np.array_equal(y[:, :, 0][dense_mask], g["y_sparse"][:, :, 0][dense_mask])
    assert np.array_equal(c, g["c_sparse"])
    ids, tgt = pp.FullModelPreprocessor(vocab=vocab).transform_ids(seqs)
    assert ids.dtype == np.int32 and np.array_equal(ids, pp.to_id_batch(g["x_full"]))
    assert np.array_equal(tgt, pp.to_id_batch(g["y_full"]))
    assert np.array_equal(pp.to_id_batch(x), ids)


def test_baseline_preprocessor_matches_reference(golden_dir):
    g = load(golden_dir, "batch_format.npz")
    seqs = unragged(g["flat"], g["offs"])
    V = int(g["V"])
    vocab = dict(zip(range(V), range(V)))
    xb, yb = pp.BaselinePreprocessor(vocab=vocab).transform_data(seqs, xs=build_xs(seqs, V))
    assert np.array_equal(xb, g["xb_xs"]) and np.array_equal(yb, g["yb_xs"])
    xb, yb = pp.BaselinePreprocessor(vocab=vocab).transform_data(seqs, xs=None)
    assert np.array_equal(xb, g["xb_plain"]) and np.array_equal(yb, g["yb_plain"])
    assert pp.is_one_hot(g["xb_plain"]) and not pp.is_one_hot(g["xb_xs"])


def test_to_id_batch_edge_cases():
    assert pp.to_id_batch(np.zeros((2, 3, 4))).tolist() == [[-1] * 3] * 2          # all-pad rows
    a = np.array([[[0.0], [2.0], [-1.0]]])
    assert pp.to_id_batch(a).tolist() == [[0, 2, -1]]
    assert pp.to_id_batch(np.array([[5, -7]])).tolist() == [[5, -1]]


def test_likelihood_metrics_match_reference(golden_dir):
    """The numpy restatement (oracle/likelihood.py: the checker of the device reductions) against the reference's own
    utils.py."""
    from oracle import likelihood as m
    g = load(golden_dir, "likelihood.npz")
    preds = unragged(g["preds"], g["offs"])
    assert np.isclose(m.compute_likelihood(preds, count_first_prob=False), g["ll"], rtol=1e-12)
    assert np.isclose(m.compute_likelihood(preds, count_first_prob=True), g["ll_first"], rtol=1e-12)
    tr, va = m.compute_likelihood_cut(preds, 0.7, count_first_prob=False)
    assert np.isclose(tr, g["cut_tr"], rtol=1e-12) and np.isclose(va, g["cut_va"], rtol=1e-12)
    tr, va = m.compute_likelihood_cut(g["padded"], 0.7, orig_lengths=g["lengths"])
    assert np.isclose(tr, g["cut_tr_l"], rtol=1e-12) and np.isclose(va, g["cut_va_l"], rtol=1e-12)


def test_mc_fixture_is_a_valid_cfg1_workload(golden_dir):
    g = load(golden_dir, "mc_sequences.npz")
    seqs = unragged(g["flat"], g["offs"])
    assert len(seqs) == 400 and min(len(s) for s in seqs) >= 2 and max(max(s) for s in seqs) <= 16
    ids, tgt = pp.FullModelPreprocessor(vocab=dict(zip(range(17), range(17))), seq_length=50).transform_ids(seqs)
    assert ids.shape == (400, 50) and ((ids >= 0) == (tgt >= 0)).all()
    # left padding: once a row turns valid it stays valid
    valid = ids >= 0
    assert (np.diff(valid.astype(int), axis=1) >= 0).all()


def test_synthetic_batches():
    ids, tgt = synthetic.make_batch(1000, 12, 16, seed=3)
    valid = ids >= 0
    assert ids.dtype == np.int32 and ids.shape == (16, 12) and ids.max() < 1000
    assert (np.diff(valid.astype(int), axis=1) >= 0).all() and valid[:, -1].all()
    assert np.array_equal(ids[:, 1:][valid[:, :-1]], tgt[:, :-1][valid[:, :-1]])       # targets are the shifted inputs
    a, _ = synthetic.make_batch(1000, 12, 16, seed=3)
    assert np.array_equal(a, ids)
    ws = synthetic.make_weights("LSTM", 50, 8)
    assert [w.shape for w in ws] == [(50, 32), (8, 32), (32,), (8, 50)] and ws[2][8:16].min() == 1.0
    assert np.allclose(ws[1][:, :8].T @ ws[1][:, :8], np.eye(8), atol=1e-5)


# ---- history features (datasets.py:97-113 build_xs -> c of FullModelPreprocessor) ------------------------------------
def _history_cases(g):
    for freq in (False, True):
        for tag, L in (("full", None), ("trunc", int(g["T_trunc"]))):
            for name in ("raw", "log"):
                yield freq, L, name == "log", g["c_%s_%s_%s" % ("freq" if freq else "bin", name, tag)]


def test_history_oracle_matches_reference_fixture(golden_dir):
    """oracle/history.py against the reference's own build_xs + FullModelPreprocessor (+ the drivers' log(x + 1))."""
    from oracle import history
    g = load(golden_dir, "history_features.npz")
    seqs = unragged(g["flat"], g["offs"])
    V = int(g["V"])
    for freq, L, log1p, ref in _history_cases(g):
        mine = history.history_block(seqs, V, seq_length=L, freq=freq, log1p=log1p)
        assert mine.shape == ref.shape and np.array_equal(mine, ref)
    assert int(g["T_full"]) == max(len(s) for s in seqs) - 1


def test_host_build_xs_matches_reference_fixture(golden_dir):
    """The package's build_xs (vectorised) through the drivers' recipe and the package's preprocessor."""
    from seq_recommendations_b200 import datasets
    g = load(golden_dir, "history_features.npz")
    seqs = unragged(g["flat"], g["offs"])
    V = int(g["V"])
    vocab = dict(zip(range(V), range(V)))
    for freq, L, log1p, ref in _history_cases(g):
        xs = datasets.build_xs(seqs, vocab, freq=freq)
        assert [np.asarray(x).tolist() for x in xs] == build_xs(seqs, V, freq=freq)
        if log1p:
            xs = [[[np.log(x + 1) for x in row] for row in rows] for rows in xs]      # experiments_server.py:35
        _, _, c = pp.FullModelPreprocessor(vocab=vocab, pad_value=0., seq_length=L).transform_data(seqs, xs=xs)
        assert c.dtype == np.float64 and np.array_equal(c, ref)
    try:
        datasets.build_xs([[0, V]], vocab)
        assert False, "an id >= V must raise like the reference's list index"
    except IndexError:
        pass


def test_host_build_xs_and_preprocessor_equal_the_history_oracle_on_random_corpora():
    """Property test (hypothesis): for any ragged corpus, catalog width and truncation length, the package's host recipe
    (datasets.build_xs -> optional log(x + 1) -> FullModelPreprocessor) equals oracle/history.py, which the fixture above
    pins to the reference's own code."""
    from hypothesis import given, settings, strategies as st
    from oracle import history
    from seq_recommendations_b200 import datasets

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 9).flatmap(lambda V: st.tuples(
        st.just(V), st.lists(st.lists(st.integers(0, V - 1), min_size=0, max_size=14), min_size=1, max_size=7),
        st.one_of(st.none(), st.integers(1, 12)), st.booleans(), st.booleans())))
    def check(case):
        V, seqs, L, freq, log1p = case
        if L is None and max(len(s) for s in seqs) < 2:
            return                                             # (T = 0: the reference's pad_sequences has nothing to pad to)
        vocab = dict(zip(range(V), range(V)))
        xs = datasets.build_xs(seqs, vocab, freq=freq)
        assert [np.asarray(x).reshape(-1, V).tolist() for x in xs] == [r for r in history.build_xs(seqs, V, freq)]
        if log1p:
            xs = [np.log(np.asarray(x, dtype=np.float64) + 1) for x in xs]
        _, _, c = pp.FullModelPreprocessor(vocab=vocab, pad_value=0., seq_length=L).transform_data(seqs, xs=xs)
        ref = history.history_block(seqs, V, seq_length=L, freq=freq, log1p=log1p)
        assert c.shape == ref.shape and np.array_equal(c, ref)

    check()
