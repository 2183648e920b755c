"""The history-feature / skip-branch models against the float64 oracle (oracle.keras_semantics.SkipModel): every
recurrent variant of experiments_server.py:106-191 (ytoz_ytoy, ytoz_xtoz, ytoz_ytoy_xtoz, ytoz_ytoy_xtoy, ...) and the
NoRecurrenceModel variants of :63-103 (ytoy, ytoy_xtoy, xtoy), through engine_dense.DensePath and through the
reference-facing classes of model.py.  Run with -m gpu on a B200."""
import numpy as np
import pytest
import torch

from oracle import keras_semantics as ks
from seq_recommendations_b200 import synthetic
from seq_recommendations_b200.engine_dense import DensePath

from gpu_util import as_t, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def make_inputs(V, T, B, seed, x_dim=None):
    """ids / targets like the reference's preprocessor, and the cumulative-count history features log(1 + count) of
    experiments_server.py:33-36 (x_t counts the items seen up to and including y_{t-1}; pads are all-zero rows)."""
    ids, tgt = synthetic.make_batch(V, T, B, seed=seed, min_len=1, zipf_s=0.6)
    Fx = x_dim or V
    x = np.zeros((B, T, Fx), dtype=np.float32)
    for b in range(B):
        cnt = np.zeros(Fx)
        for t in range(T):
            if ids[b, t] >= 0:
                cnt[ids[b, t] % Fx] += 1
                x[b, t] = np.log(cnt + 1)
    return ids, tgt, x


def make_weights(rng, cell, V, Fx, H, y_to_z, x_to_z, x_to_y, y_to_y, biases=True):
    G = {"simpleRNN": 1, "LSTM": 4, "GRU": 3, None: 0}[cell]
    ws = {}
    r = lambda *s: (rng.standard_normal(s) * 0.3).astype(np.float32)
    if cell:
        ws.update(W_in=r((V if y_to_z else 0) + (Fx if x_to_z else 0), G * H), U=r(H, G * H), b=r(G * H))
    if cell or x_to_y:
        ws["W_toy"] = r((H if cell else 0) + (Fx if x_to_y else 0), V)
        if biases:
            ws["b_out"] = r(V)
    if y_to_y:
        ws["A"] = r(V, V)
        if biases:
            ws["a_bias"] = r(V)
    return ws


VARIANTS = [
    # cell, act, y_to_z, x_to_z, x_to_y, y_to_y
    ("LSTM", "relu", True, False, False, True),      # ytoz_ytoy
    ("LSTM", "relu", True, True, False, False),      # ytoz_xtoz
    ("LSTM", "relu", True, True, False, True),       # ytoz_ytoy_xtoz
    ("LSTM", "relu", True, False, True, True),       # ytoz_ytoy_xtoy
    ("GRU", "tanh", True, True, True, True),         # everything at once
    ("simpleRNN", "relu", False, True, True, False), # x only into z
    (None, "linear", False, False, False, True),     # NoRecurrence: ytoy
    (None, "linear", False, False, True, True),      # NoRecurrence: ytoy_xtoy
    (None, "linear", False, False, True, False),     # NoRecurrence: xtoy
]


@pytest.mark.parametrize("cell,act,y_to_z,x_to_z,x_to_y,y_to_y", VARIANTS)
def test_skip_branch_loss_and_gradients_match_oracle(cell, act, y_to_z, x_to_z, x_to_y, y_to_y):
    V, H, T, B = 17, 12, 9, 14
    rng = np.random.default_rng(5)
    ws = make_weights(rng, cell, V, V, H, y_to_z, x_to_z, x_to_y, y_to_y)
    ids, tgt, x = make_inputs(V, T, B, seed=6)
    hot = DensePath(cell, act, V, V, H, ws, y_to_z=y_to_z, x_to_z=x_to_z, x_to_y=x_to_y, y_to_y=y_to_y)
    ora = ks.SkipModel(cell, act, ws, y_to_z=y_to_z, x_to_z=x_to_z, x_to_y=x_to_y, y_to_y=y_to_y)
    uy, ux = hot.uses_y, hot.uses_x
    t_ids = as_t(ids) if uy else None
    t_x = torch.tensor(x, dtype=torch.float64) if ux else None
    loss, grads, _ = hot.grad_batch(ids if uy else None, tgt, x if ux else None)
    rl, rg = ora.grads(t_ids, t_x, as_t(tgt))
    assert abs(loss - float(rl)) <= TOL * abs(float(rl)), (loss, float(rl))
    assert sorted(grads) == sorted(rg)
    for n in grads:
        assert rel_err(grads[n], rg[n].numpy()) <= TOL, (n, rel_err(grads[n], rg[n].numpy()))
    # model.predict and p(true item)
    probs = hot.predict_batch(ids if uy else None, x if ux else None).cpu().numpy()
    ref = ora.predict_proba(t_ids, t_x).numpy()
    assert np.abs(probs - ref).max() <= TOL
    ti, _ = hot.topk_batch(ids if uy else None, 5, last_step_only=True, x_dense=x if ux else None)
    assert np.array_equal(ti.cpu().numpy(), ks.topk_items(torch.tensor(ref[:, -1]), 5))


@pytest.mark.parametrize("cell,act,y_to_z,x_to_z,x_to_y,y_to_y", [VARIANTS[3], VARIANTS[4], VARIANTS[7]])
def test_skip_branch_training_steps_with_diagonal_constraint_and_frozen_transition_kernel(cell, act, y_to_z, x_to_z,
                                                                                         x_to_y, y_to_y):
    """Three optimisation steps: global-norm clip + Adagrad over every weight, OnlyNonZeroDiagonal applied to the
    UPDATED x rows of the output kernel (model.py:48-66, :379), and the `*_fixed` variants of experiments_server.py
    (y_to_y_trainable=False: the Markov-initialised transition kernel stays frozen and leaves the norm)."""
    V, H, T, B = 17, 10, 7, 12
    rng = np.random.default_rng(8)
    ws = make_weights(rng, cell, V, V, H, y_to_z, x_to_z, x_to_y, y_to_y, biases=False)
    for frozen in ((), ("A",)):
        hot = DensePath(cell, act, V, V, H, ws, y_to_z=y_to_z, x_to_z=x_to_z, x_to_y=x_to_y, y_to_y=y_to_y)
        ora = ks.SkipModel(cell, act, ws, y_to_z=y_to_z, x_to_z=x_to_z, x_to_y=x_to_y, y_to_y=y_to_y)
        hot.set_optimizer("adagrad", lr=0.05, epsilon=1e-8, clipnorm=1.0)
        for n in frozen:
            hot.trainable[n] = False
        for step in range(3):
            ids, tgt, x = make_inputs(V, T, B, seed=20 + step)
            loss = float(hot.train_batch(ids, tgt, x).item())
            rl, _ = ora.train_step(as_t(ids), torch.tensor(x, dtype=torch.float64), as_t(tgt), lr=0.05, epsilon=1e-8,
                                   clipnorm=1.0, frozen=frozen)
            assert abs(loss - float(rl)) <= TOL * abs(float(rl)), (step, loss, float(rl))
        for n in hot.weight_names():
            assert rel_err(hot.get_weight(n), ora.p[n].numpy()) <= 2e-4, (n, rel_err(hot.get_weight(n), ora.p[n].numpy()))
        Wx = hot.get_weight("W_toy")[hot.H:]
        assert not (Wx - np.diag(np.diag(Wx))).any()              # only the diagonal of the x block survives
        if frozen:
            assert np.array_equal(hot.get_weight("A"), ws["A"])


def test_skip_branch_dropouts_match_oracle_with_the_same_factors():
    """y->z dropout on the CONCATENATED [y ; x] input, z->y dropout and recurrent dropout in one step."""
    V, H, T, B = 17, 16, 6, 10
    rng = np.random.default_rng(11)
    ws = make_weights(rng, "LSTM", V, V, H, True, True, True, True)
    ids, tgt, x = make_inputs(V, T, B, seed=12)
    hot = DensePath("LSTM", "relu", V, V, H, ws, y_to_z=True, x_to_z=True, x_to_y=True, y_to_y=True)
    hot.dropout_in, hot.dropout_out, hot.dropout_rec = 0.2, 0.3, 0.25
    ora = ks.SkipModel("LSTM", "relu", ws, y_to_z=True, x_to_z=True, x_to_y=True, y_to_y=True)
    loss, grads, _ = hot.grad_batch(ids, tgt, x)
    w = hot.work(B, T)
    # the factors the device drew: per-token factor of the one-hot half, per-element factors of the x half
    onehot_f = w.in_scale.view(T, B).t().cpu().double()
    x_tb = torch.tensor(x, dtype=torch.float64).permute(1, 0, 2)
    xf = torch.where(x_tb != 0, w.x_drop.view(T, B, V).cpu().double() / x_tb.clamp(min=1e-30), torch.ones_like(x_tb))
    in_drop = torch.cat([onehot_f.unsqueeze(-1).expand(B, T, V), xf.permute(1, 0, 2)], dim=-1)
    out_scale = w.hscale.view(T, B, H).permute(1, 0, 2).cpu().double()
    rec = [w.rec_mask[g].cpu().double() for g in range(4)]
    rl, rg = ora.grads(as_t(ids), torch.tensor(x, dtype=torch.float64), as_t(tgt), in_drop=in_drop, out_scale=out_scale,
                       rec_masks=rec)
    assert abs(loss - float(rl)) <= TOL * abs(float(rl))
    for n in grads:
        assert rel_err(grads[n], rg[n].numpy()) <= TOL, (n, rel_err(grads[n], rg[n].numpy()))


def test_skip_branch_products_on_the_tensor_core_gemm():
    """A catalog / batch large enough that the dense products ([y ; x] input projection = K2, the logit terms and their
    gradients) run on the tcgen05 GEMM of csrc/gemm_tc.cu (3-pass split, fp32-grade)."""
    V, H, T, B = 640, 64, 12, 48
    rng = np.random.default_rng(14)
    ws = make_weights(rng, "GRU", V, V, H, True, True, True, True)
    ids, tgt, x = make_inputs(V, T, B, seed=15)
    res = {}
    for tc in ("x3", "off"):
        hot = DensePath("GRU", "tanh", V, V, H, ws, y_to_z=True, x_to_z=True, x_to_y=True, y_to_y=True, tc=tc)
        res[tc] = hot.grad_batch(ids, tgt, x)
    ora = ks.SkipModel("GRU", "tanh", ws, y_to_z=True, x_to_z=True, x_to_y=True, y_to_y=True)
    rl, rg = ora.grads(as_t(ids), torch.tensor(x, dtype=torch.float64), as_t(tgt))
    for tc in ("x3", "off"):
        loss, grads, _ = res[tc]
        assert abs(loss - float(rl)) <= TOL * abs(float(rl)), tc
        for n in grads:
            assert rel_err(grads[n], rg[n].numpy()) <= TOL, (tc, n, rel_err(grads[n], rg[n].numpy()))


@pytest.mark.parametrize("M,N,K", [(128, 64, 32), (1000, 384, 34), (4096, 1024, 17), (300, 70, 513), (129, 129, 129)])
def test_gemm_tc_matches_float64_product(M, N, K):
    """seqrec_gemm_tc directly: C = A . Bt^T (+ bias), accumulate, ragged M / N / K edges, x3 (1e-5) and bf16 (2e-2)."""
    import ctypes
    from seq_recommendations_b200._lib import call, ptr
    rng = np.random.default_rng(M + N + K)
    A = torch.tensor(rng.standard_normal((M, K)).astype(np.float32)).cuda()
    Bt = torch.tensor(rng.standard_normal((N, K)).astype(np.float32)).cuda()
    bias = torch.tensor(rng.standard_normal(N).astype(np.float32)).cuda()
    C0 = torch.tensor(rng.standard_normal((M, N)).astype(np.float32)).cuda()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    Kp = (K + 63) // 64 * 64
    bf = torch.bfloat16
    a_hi, a_lo = torch.zeros((M, Kp), dtype=bf, device="cuda"), torch.zeros((M, Kp), dtype=bf, device="cuda")
    b_hi, b_lo = torch.zeros((N, Kp), dtype=bf, device="cuda"), torch.zeros((N, Kp), dtype=bf, device="cuda")
    call("seqrec_split_bf16", ptr(A), None, ptr(a_hi), ptr(a_lo), M, K, Kp, 0, st)
    call("seqrec_split_bf16", ptr(Bt), None, ptr(b_hi), ptr(b_lo), N, K, Kp, 0, st)
    ref = A.double() @ Bt.double().t()
    for x3, tol in ((1, 1e-5), (0, 2e-2)):
        C = torch.empty((M, N), dtype=torch.float32, device="cuda")
        call("seqrec_gemm_tc", ptr(a_hi), ptr(a_lo), ptr(b_hi), ptr(b_lo), ptr(bias), ptr(C), M, N, K, Kp, Kp, N, 0, x3, st)
        want = ref + bias.double()
        assert float((C.double() - want).norm() / want.norm()) <= tol, (x3, "plain")
        C = C0.clone()
        call("seqrec_gemm_tc", ptr(a_hi), ptr(a_lo), ptr(b_hi), ptr(b_lo), None, ptr(C), M, N, K, Kp, Kp, N, 1, x3, st)
        want = ref + C0.double()
        assert float((C.double() - want).norm() / want.norm()) <= tol, (x3, "accumulate")


def test_reference_facing_classes_run_the_experiment_variants():
    """RNNFullModel / NoRecurrenceModel as experiments_methods.py builds them (run_model_with_recurrence :202-216,
    run_model_no_recurrence :142-158): list inputs [y one-hot, xs], ArrayInitializer with a Markov log-transition matrix
    (experiments_server.py:60-68), y_to_y layer frozen through set_layer_weights_trainable, fit / evaluate / predict."""
    from seq_recommendations_b200.model import ArrayInitializer, NoRecurrenceModel, RNNFullModel
    from seq_recommendations_b200.optimizers import Adagrad
    from seq_recommendations_b200.preprocessor import FullModelPreprocessor
    rng = np.random.default_rng(3)
    V, T = 9, 8
    seqs = []
    for _ in range(40):                                # a learnable chain: next = cur + 1 (mod V) with probability 0.8
        s = [int(rng.integers(0, V))]
        for _ in range(int(rng.integers(2, T + 1))):
            s.append((s[-1] + 1) % V if rng.random() < 0.8 else int(rng.integers(0, V)))
        seqs.append(s)
    xs = []
    for s in seqs:                                     # datasets.build_xs(freq=True) + log(x + 1)
        cnt, rows = np.zeros(V), []
        for it in s:
            cnt[it] += 1
            rows.append(np.log(cnt + 1).tolist())
        xs.append(rows)
    pre = FullModelPreprocessor(vocab=dict(zip(range(V), range(V))), seq_length=T)
    x, y, c = pre.transform_data(seqs, xs)
    trans = np.full((V, V), 1e-6)
    for s in seqs:
        for a, b in zip(s[:-1], s[1:]):
            trans[a, b] += 1
    init = np.log(trans / trans.sum(1, keepdims=True)).astype(np.float32)
    for build, inputs, markov in (
            (lambda: RNNFullModel(T, V, V, z_dim=8, rnn_type="LSTM", y_to_z=True, y_to_y=True, x_to_y=True,
                                  x_to_z=False, y_to_y_w_initializer=ArrayInitializer(init), z_to_y_dropout=0.3,
                                  seed=1), [x, c], True),
            (lambda: RNNFullModel(T, V, V, z_dim=8, rnn_type="LSTM", y_to_z=True, y_to_y=False, x_to_y=False,
                                  x_to_z=True, z_to_z_dropout=0.2, seed=1), [x, c], False),
            (lambda: NoRecurrenceModel(T, V, V, y_to_y_w_initializer=ArrayInitializer(init), connect_x=True,
                                       connect_y=True, seed=1), [x, c], True),
            (lambda: NoRecurrenceModel(T, V, V, connect_x=False, connect_y=True, seed=1), x, False)):
        m = build()
        names = [l.name for l in m.model.layers]
        ytoy = "y_to_y_output" if "y_to_y_output" in names else ("y_output" if "y_output" in names else None)
        frozen = ytoy is not None and markov          # the `*_fixed` variants freeze the Markov-initialised kernel
        if ytoy:
            A0 = m.get_layer_weights(ytoy)[0]
            assert np.allclose(A0, init) == markov
        if frozen:
            m.set_layer_weights_trainable(ytoy, trainable=False)
        m.compile_model(loss="categorical_crossentropy", metrics=[], optimizer=Adagrad(lr=0.05, epsilon=1e-8, clipnorm=1.))
        h = m.fit_model(inputs, y, validation_data=(inputs, y), n_epochs=4, batch_size=16, verbose=0)
        assert len(h.history["loss"]) == 4
        if markov:
            # starts AT the optimum of the Markov kernel: the remaining gradients are near zero and Adagrad's first
            # steps are sign-like, so the loss hovers (its last digits follow the summation order of the kernels that
            # happen to serve the shape) -- it must not drift away
            assert h.history["loss"][-1] < 1.05 * h.history["loss"][0]
        else:
            assert h.history["loss"][-1] < h.history["loss"][0]
        if frozen:
            assert np.array_equal(m.get_layer_weights(ytoy)[0], A0)       # frozen layer untouched
        elif ytoy:
            assert not np.array_equal(m.get_layer_weights(ytoy)[0], A0)
        names_, scores = m.evaluate(inputs, y, batch_size=16)
        assert abs(scores[0] - h.history["val_loss"][-1]) <= 1e-5 * abs(scores[0])
        p = m.predict(inputs, batch_size=16)
        assert p.shape == (len(seqs), T, V) and np.abs(p.sum(-1) - 1).max() < 1e-5
