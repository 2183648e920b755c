"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own host-side code.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What runs unmodified from /root/reference:  preprocessor.py (FullModelPreprocessor / BaselinePreprocessor
.transform_data), sampler.py (MCSampler.random_init + gen_sequence), utils.py (compute_likelihood,
compute_likelihood_cut, transition_matrix, multinomial_probabilities) and the `build_xs` function of datasets.py
(the module itself has Python-2 print statements, so only that function's source text is exec'd).

What is stubbed: the third-party modules those files import but this image lacks.  `keras.preprocessing.sequence
.pad_sequences` and `keras.utils.np_utils.to_categorical` are restated below from the Keras-2.0.x utilities (they
are ~20 lines of numpy); matplotlib is an empty stub (only plotting uses it).  Keras/Theano numerics are NOT
exercised here -- that part of the oracle stays unpinned (oracle/__init__.py).
"""
import os
import random
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- third-party stubs (Keras-2.0.x utils)
def pad_sequences(sequences, maxlen=None, dtype="int32", padding="pre", truncating="pre", value=0.0):
    lengths = [len(s) for s in sequences]
    nb_samples = len(sequences)
    if maxlen is None:
        maxlen = np.max(lengths)
    sample_shape = tuple()
    for s in sequences:
        if len(s) > 0:
            sample_shape = np.asarray(s).shape[1:]
            break
    x = (np.ones((nb_samples, maxlen) + sample_shape) * value).astype(dtype)
    for idx, s in enumerate(sequences):
        if len(s) == 0:
            continue
        if truncating == "pre":
            trunc = s[-maxlen:]
        elif truncating == "post":
            trunc = s[:maxlen]
        else:
            raise ValueError(truncating)
        trunc = np.asarray(trunc, dtype=dtype)
        if padding == "post":
            x[idx, : len(trunc)] = trunc
        elif padding == "pre":
            x[idx, -len(trunc):] = trunc
        else:
            raise ValueError(padding)
    return x


def to_categorical(y, num_classes=None):
    y = np.array(y, dtype="int").ravel()
    if not num_classes:
        num_classes = np.max(y) + 1
    n = y.shape[0]
    categorical = np.zeros((n, num_classes))
    categorical[np.arange(n), y] = 1
    return categorical


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    keras = mod("keras")
    pre = mod("keras.preprocessing")
    seq = mod("keras.preprocessing.sequence", pad_sequences=pad_sequences)
    np_utils = mod("keras.utils.np_utils", to_categorical=to_categorical)
    utils = mod("keras.utils", np_utils=np_utils)
    keras.preprocessing, keras.utils, pre.sequence = pre, utils, seq
    mpl = mod("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = mod("matplotlib.pyplot")
    mpl.animation = mod("matplotlib.animation")


def load_build_xs():
    src = open(os.path.join(REF, "datasets.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("def build_xs"))
    end = next(i for i in range(start + 1, len(src)) if src[i].startswith("def "))
    ns = {}
    exec("\n".join(src[start:end]), ns)
    return ns["build_xs"]


def ragged(seqs):
    flat = np.concatenate([np.asarray(s, dtype=np.int64) for s in seqs]) if seqs else np.zeros(0, np.int64)
    offs = np.cumsum([0] + [len(s) for s in seqs]).astype(np.int64)
    return flat, offs


def history_fixture(ref_pre, build_xs):
    """History features through the reference's own build_xs + FullModelPreprocessor, with the drivers' log(x + 1)
    (experiments_server.py:33-36); ragged lengths, a length-1 and a length-2 sequence, truncation that drops counted
    rows.  Written to its own file so that the other fixtures stay byte-identical."""
    rng = np.random.RandomState(5)
    V = 7
    vocab = dict(zip(range(V), range(V)))
    seqs = [[int(v) for v in rng.randint(0, V, size=n)] for n in (12, 1, 2, 9, 5, 3, 17, 8)]
    flat, offs = ragged(seqs)
    out = {"flat": flat, "offs": offs, "V": np.int64(V)}
    for freq in (False, True):
        xs = build_xs(seqs, vocab, freq=freq)
        xs_log = [[[np.log(x + 1) for x in row] for row in rows] for rows in xs]
        for tag, L in (("full", None), ("trunc", 6)):
            for name, feats in (("raw", xs), ("log", xs_log)):
                p = ref_pre.FullModelPreprocessor(vocab=vocab, pad_value=0.0, seq_length=L)
                _, _, c = p.transform_data(seqs, xs=feats)
                out["c_%s_%s_%s" % ("freq" if freq else "bin", name, tag)] = c
                out["T_" + tag] = np.int64(p.seq_length)
    np.savez_compressed(os.path.join(OUT, "history_features.npz"), **out)


def main():
    install_stubs()
    sys.path.insert(0, REF)
    import preprocessor as ref_pre
    if sys.argv[1:] == ["history"]:
        history_fixture(ref_pre, load_build_xs())
        print("wrote history_features.npz")
        return
    import sampler as ref_sampler
    import utils as ref_utils

    build_xs = load_build_xs()

    # ---- (1) batch format: the commented smoke sequences of model.py:413-415 plus ragged edge cases ----------
    V = 4
    vocab = dict(zip(range(V), range(V)))
    seqs = [[3, 1, 0, 2, 3, 2, 3, 1, 3, 2], [3, 1, 2, 2, 1, 1, 1, 2], [3, 1, 3, 3, 1], [2], [0, 0, 0], [1, 0]]
    xs = build_xs(seqs, vocab)
    xs_freq = build_xs(seqs, vocab, freq=True)
    flat, offs = ragged(seqs)
    out = {"flat": flat, "offs": offs, "V": np.int64(V)}
    for tag, seq_length in (("full", None), ("trunc", 4)):
        p = ref_pre.FullModelPreprocessor(vocab=vocab, pad_value=0.0, seq_length=seq_length)
        x, y, c = p.transform_data(seqs, xs=xs)
        out["x_" + tag], out["y_" + tag], out["c_" + tag] = x, y, c
        out["T_" + tag] = np.int64(p.seq_length)
    p = ref_pre.FullModelPreprocessor(vocab=vocab, pad_value=0.0, seq_length=None, sparse=True)
    x, y, c = p.transform_data(seqs, xs=xs)
    out["x_sparse"], out["y_sparse"], out["c_sparse"] = x, y, c
    p = ref_pre.FullModelPreprocessor(vocab=vocab, pad_value=0.0, seq_length=None)
    _, _, c = p.transform_data(seqs, xs=xs_freq)
    out["c_freq"] = c
    pb = ref_pre.BaselinePreprocessor(vocab=vocab, pad_value=0.0, seq_length=None)
    xb, yb = pb.transform_data(seqs, xs=xs)
    out["xb_xs"], out["yb_xs"] = xb, yb
    pb = ref_pre.BaselinePreprocessor(vocab=vocab, pad_value=0.0, seq_length=None)
    xb, yb = pb.transform_data(seqs, xs=None)
    out["xb_plain"], out["yb_plain"] = xb, yb
    np.savez_compressed(os.path.join(OUT, "batch_format.npz"), **out)

    # ---- (2) MSNBC-like sequences from the reference's own generator (config 1 input) ----------
    np.random.seed(0)
    random.seed(0)
    # random_init(n, use_end_token=True) trips its own shape assert (it draws an n x n alpha, sampler.py:94 vs :61),
    # so draw the same quantities by its recipe with the end-token column added: 17 page categories + end state.
    n = 17
    gamma = np.random.rand(n)
    gamma = gamma / np.sum(gamma)
    alpha = np.random.rand(n, n + 1)
    np.fill_diagonal(alpha, 0)
    alpha = alpha / np.sum(alpha, axis=1).reshape((n, 1))
    s = ref_sampler.MCSampler(alpha, gamma, beta=0.9, use_end_token=True)
    mc = []
    while len(mc) < 400:
        q = [int(v) for v in s.gen_sequence()]
        if len(q) >= 2:
            mc.append(q)
    flat, offs = ragged(mc)
    np.savez_compressed(os.path.join(OUT, "mc_sequences.npz"), flat=flat, offs=offs, n_states=np.int64(17),
                        alpha=s.alpha, gamma=s.gamma)

    # ---- (3) likelihood metrics and count baselines (utils.py:79-178) ----------------------------------------
    rng = np.random.RandomState(1)
    preds = [rng.uniform(0.01, 0.99, size=n).tolist() for n in (1, 2, 5, 9, 12)]
    padded = np.zeros((len(preds), 12))
    for i, pr in enumerate(preds):
        padded[i, -len(pr):] = pr
    padded = np.clip(padded, 1e-7, 1 - 1e-7)
    lengths = np.array([len(pr) for pr in preds])
    ll = ref_utils.compute_likelihood(preds, count_first_prob=False)
    ll_first = ref_utils.compute_likelihood(preds, count_first_prob=True)
    cut_tr, cut_va = ref_utils.compute_likelihood_cut(preds, 0.7, count_first_prob=False)
    cut_tr_l, cut_va_l = ref_utils.compute_likelihood_cut(padded, 0.7, orig_lengths=lengths)
    T_alpha, T_gamma = ref_utils.transition_matrix(seqs, V, 1.0, freq=False, end_state=False)
    multi = ref_utils.multinomial_probabilities(seqs, V, 1.0, True)
    flatp, offp = ragged([np.arange(len(pr)) for pr in preds])
    np.savez_compressed(os.path.join(OUT, "likelihood.npz"), preds=np.concatenate(preds), offs=offp, padded=padded,
                        lengths=lengths, ll=ll, ll_first=ll_first, cut_tr=cut_tr, cut_va=cut_va, cut_tr_l=cut_tr_l,
                        cut_va_l=cut_va_l, T_alpha=T_alpha, T_gamma=T_gamma, multi=multi)
    history_fixture(ref_pre, build_xs)
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
