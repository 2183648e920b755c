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
