"""numpy restatement of the reference's likelihood metrics (utils.py:141-178).  TEST INFRASTRUCTURE ONLY (see
oracle/__init__.py): pinned against the reference's own utils.py through tests/golden/likelihood.npz, and the checker
of the device reductions in seq_recommendations_b200/likelihood.py."""
import numpy as np

_EPSILON = 1e-7


def compute_likelihood_cut(predictions, train_percent, orig_lengths=None, count_first_prob=False):
    """utils.py:145-163 (per-sequence mean NLL with a within-sequence cut; orig_lengths overrides count_first_prob)."""
    assert train_percent <= 1.0
    train_lls, val_lls = [], []
    for i, pred in enumerate(predictions):
        sort_pred = pred[:]
        if not count_first_prob:
            sort_pred = sort_pred[1:]
        if orig_lengths is not None:
            sort_pred = pred[-int(orig_lengths[i]):]
        seq_length = len(sort_pred)
        train_elems = int(np.ceil(train_percent * seq_length))
        val_elems = int(np.floor((1.0 - train_percent) * seq_length))
        if train_elems > 0:
            train_lls.append(-np.sum(np.log(sort_pred[0:train_elems])) / train_elems)
        if val_elems > 0:
            val_lls.append(-np.sum(np.log(sort_pred[-val_elems:])) / val_elems)
    return np.sum(train_lls) / len(train_lls), np.sum(val_lls) / len(val_lls)


def compute_likelihood(predictions, count_first_prob=False):
    """utils.py:166-178."""
    lls = []
    for pred in predictions:
        sort_pred = pred[:]
        if not count_first_prob:
            sort_pred = sort_pred[1:]
        sort_pred = np.clip(sort_pred, _EPSILON, 1.0 - _EPSILON)
        if len(sort_pred) > 0:
            lls.append(-np.sum(np.log(sort_pred)) / len(sort_pred))
    return np.mean(lls)
