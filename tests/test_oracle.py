"""Known-answer tests of the CPU oracle (oracle/keras_semantics.py): every Keras-2.0.x semantic of SURVEY §8(c) that the
CUDA path is later compared against is first pinned here against hand-computed scalars."""
import math

import numpy as np
import pytest
import torch

from oracle import keras_semantics as ks


def hs(a):
    return min(max(0.2 * a + 0.5, 0.0), 1.0)


def test_hard_sigmoid_and_clip_constants():
    a = torch.tensor([-3.0, -2.5, 0.0, 1.0, 2.5, 4.0])
    assert torch.allclose(ks.hard_sigmoid(a), torch.tensor([0.0, 0.0, 0.5, 0.7, 1.0, 1.0]))
    assert ks.EPS32 == float(np.float32(1e-7))
    assert ks.ONE_MINUS_EPS32 < 1.0 and ks.ONE_MINUS_EPS32 == float(np.float32(1.0 - 1e-7))


def test_lstm_two_steps_by_hand():
    # H=1, V=2, relu; gate order i,f,c,o
    W_in = np.array([[0.5, -0.4, 0.8, 0.3], [-0.2, 0.6, -0.5, 0.9]], dtype=np.float64)
    U = np.array([[0.1, 0.2, -0.3, 0.4]], dtype=np.float64)
    b = np.array([0.0, 1.0, 0.1, -0.1], dtype=np.float64)
    ids = [0, 1]
    h = c = 0.0
    outs = []
    for t in ids:
        x = W_in[t] + b
        i, f, o = hs(x[0] + h * U[0, 0]), hs(x[1] + h * U[0, 1]), hs(x[3] + h * U[0, 3])
        g = max(x[2] + h * U[0, 2], 0.0)
        c = f * c + i * g
        h = o * max(c, 0.0)
        outs.append(h)
    xp = ks.input_projection(torch.tensor(W_in), torch.tensor(b), ids=torch.tensor([ids]),
                             mask=torch.ones(1, 2, dtype=torch.bool))
    H = ks.rnn_forward(xp, torch.tensor(U), torch.ones(1, 2, dtype=torch.bool), "LSTM", "relu")
    assert np.allclose(H[0, :, 0].numpy(), outs, atol=1e-12)


def test_gru_reset_before_matmul_by_hand():
    # H=2 so that (r*h).U_h differs from r*(h.U_h) -- the Keras-2.0.x GRU, not the cuDNN one
    rng = np.random.default_rng(0)
    V, H = 3, 2
    W_in, U, b = rng.normal(size=(V, 3 * H)), rng.normal(size=(H, 3 * H)), rng.normal(size=3 * H) * 0.1
    ids = [2, 0, 1]
    h = np.zeros(H)
    outs = []
    for t in ids:
        x = W_in[t] + b
        z = np.array([hs(v) for v in x[0:H] + h @ U[:, 0:H]])
        r = np.array([hs(v) for v in x[H:2 * H] + h @ U[:, H:2 * H]])
        hh = np.tanh(x[2 * H:] + (r * h) @ U[:, 2 * H:])
        h = z * h + (1 - z) * hh
        outs.append(h.copy())
    m = torch.ones(1, 3, dtype=torch.bool)
    xp = ks.input_projection(torch.tensor(W_in), torch.tensor(b), ids=torch.tensor([ids]), mask=m)
    Hh = ks.rnn_forward(xp, torch.tensor(U), m, "GRU", "tanh")
    assert np.allclose(Hh[0].numpy(), np.array(outs), atol=1e-12)
    # and it is NOT the reset-after variant
    h2 = np.zeros(H)
    for t in ids:
        x = W_in[t] + b
        z = np.array([hs(v) for v in x[0:H] + h2 @ U[:, 0:H]])
        r = np.array([hs(v) for v in x[H:2 * H] + h2 @ U[:, H:2 * H]])
        h2 = z * h2 + (1 - z) * np.tanh(x[2 * H:] + r * (h2 @ U[:, 2 * H:]))
    assert not np.allclose(h2, outs[-1])


def test_simple_rnn_by_hand():
    W_in = np.array([[0.3], [-0.7]])
    U = np.array([[0.5]])
    b = np.array([0.1])
    h, outs = 0.0, []
    for t in [0, 0, 1]:
        h = max(W_in[t, 0] + b[0] + h * 0.5, 0.0)
        outs.append(h)
    m = torch.ones(1, 3, dtype=torch.bool)
    xp = ks.input_projection(torch.tensor(W_in), torch.tensor(b), ids=torch.tensor([[0, 0, 1]]), mask=m)
    assert np.allclose(ks.rnn_forward(xp, torch.tensor(U), m, "simpleRNN", "relu")[0, :, 0].numpy(), outs)


@pytest.mark.parametrize("cell", ["simpleRNN", "LSTM", "GRU"])
def test_mask_holds_state_and_repeats_output(cell):
    """Theano K.rnn switch: pre-padded rows start from the zero state at their first real token, pad outputs are 0,
    and a masked step in the MIDDLE repeats the previous output and holds the state."""
    rng = np.random.default_rng(1)
    V, H, T = 5, 4, 6
    ws = ks.init_weights(rng, cell, V, H, V)
    mod = ks.Model(cell, "tanh", ws, dtype=torch.float64)
    seq = torch.tensor([[1, 3, 2, 4]])
    full_mask = torch.ones(1, 4, dtype=torch.bool)
    ref = mod.hidden_states(ids=seq, mask=full_mask)[0]
    ids = torch.tensor([[-1, -1, 1, 3, 2, 4]])
    mask = ids >= 0
    out = mod.hidden_states(ids=ids, mask=mask)[0]
    assert torch.all(out[:2] == 0)
    assert torch.allclose(out[2:], ref, atol=1e-14)
    # masked step in the middle
    ids2 = torch.tensor([[1, 3, -1, 2, 4, 0]])
    out2 = mod.hidden_states(ids=ids2, mask=ids2 >= 0)[0]
    assert torch.equal(out2[2], out2[1])
    ref2 = mod.hidden_states(ids=torch.tensor([[1, 3, 2, 4, 0]]), mask=torch.ones(1, 5, dtype=torch.bool))[0]
    assert torch.allclose(out2[[0, 1, 3, 4, 5]], ref2, atol=1e-14)


def test_one_hot_matmul_equals_gather():
    """the RNN input kernel times a one-hot row IS the embedding lookup (value-equal; -0.0 aside)."""
    rng = np.random.default_rng(2)
    V, GH = 7, 12
    W = torch.tensor(rng.normal(size=(V, GH)).astype(np.float32))
    b = torch.tensor(rng.normal(size=GH).astype(np.float32))
    ids = torch.tensor([[3, 0, 6, -1]])
    mask = ids >= 0
    onehot = torch.zeros(1, 4, V)
    for t, i in enumerate(ids[0].tolist()):
        if i >= 0:
            onehot[0, t, i] = 1.0
    a = ks.input_projection(W, b, ids=ids, mask=mask)
    d = ks.input_projection(W, b, x_dense=onehot)
    assert torch.equal(a, d)
    assert torch.equal(ks.derive_mask(onehot), mask)


def test_masked_loss_by_hand_and_clip():
    z = torch.tensor([[[1.0, 2.0, 0.5], [0.0, 0.0, 0.0], [30.0, 0.0, 0.0]]], dtype=torch.float64)
    tgt = torch.tensor([[1, -1, 1]])
    mask = torch.tensor([[True, False, True]])
    loss, ce, py = ks.masked_loss(z, tgt, mask)
    p0 = math.exp(2.0) / (math.exp(1.0) + math.exp(2.0) + math.exp(0.5))
    p2 = max(math.exp(-30.0) / (1 + 2 * math.exp(-30.0)), ks.EPS32)  # clipped from ~9e-14 up to 1e-7
    assert abs(float(ce[0, 0]) + math.log(p0)) < 1e-12
    assert abs(float(ce[0, 2]) + math.log(p2)) < 1e-12
    assert float(ce[0, 1]) == 0.0
    assert abs(float(loss) - (-(math.log(p0) + math.log(p2)) / 2)) < 1e-12


def test_clip_saturation_kills_the_gradient():
    """Theano's clip passes no gradient on the saturated side: a token whose p_y < 1e-7 contributes loss but no grad."""
    z = torch.tensor([[[30.0, 0.0, 0.0], [1.0, 2.0, 0.5]]], dtype=torch.float64, requires_grad=True)
    tgt = torch.tensor([[1, 1]])
    mask = torch.ones(1, 2, dtype=torch.bool)
    loss, _, _ = ks.masked_loss(z, tgt, mask)
    g, = torch.autograd.grad(loss, z)
    assert torch.all(g[0, 0] == 0)
    s = torch.softmax(z[0, 1].detach(), dim=0)
    expect = (s - torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)) / 2
    assert torch.allclose(g[0, 1], expect, atol=1e-12)


def test_all_pad_batch_is_nan():
    z = torch.zeros(1, 2, 3)
    loss, _, _ = ks.masked_loss(z, torch.tensor([[-1, -1]]), torch.zeros(1, 2, dtype=torch.bool))
    assert torch.isnan(loss)


def test_clipnorm_and_adagrad_by_hand():
    g = [torch.tensor([3.0, 0.0]), torch.tensor([[0.0, 4.0]])]
    gc, n = ks.clip_by_global_norm(g, 1.0)
    assert float(n) == 5.0
    assert torch.allclose(gc[0], torch.tensor([0.6, 0.0])) and torch.allclose(gc[1], torch.tensor([[0.0, 0.8]]))
    gs, n = ks.clip_by_global_norm([torch.tensor([0.3, 0.4])], 1.0)  # below the threshold: untouched
    assert torch.equal(gs[0], torch.tensor([0.3, 0.4]))
    p, a = ks.adagrad_update(torch.tensor([1.0]), torch.tensor([0.6]), torch.tensor([0.0]), 0.01, 1e-8)
    assert abs(float(a) - 0.36) < 1e-7 and abs(float(p) - (1.0 - 0.01 * 0.6 / (0.6 + 1e-8))) < 1e-7
    p2, a2 = ks.adagrad_update(p, torch.tensor([0.8]), a, 0.01, 1e-8)
    assert abs(float(p2) - (float(p) - 0.01 * 0.8 / (1.0 + 1e-8))) < 1e-7


def test_train_step_matches_manual_composition():
    rng = np.random.default_rng(3)
    V, H, B, T = 6, 5, 3, 4
    ws = ks.init_weights(rng, "GRU", V, H, V)
    mod = ks.Model("GRU", "tanh", ws, dtype=torch.float64)
    ids = torch.tensor(rng.integers(0, V, size=(B, T)))
    ids[0, :2] = -1
    tgt = torch.tensor(rng.integers(0, V, size=(B, T)))
    mask = ids >= 0
    loss, gs = mod.grads(ids, tgt, mask)
    norm = math.sqrt(sum(float((g * g).sum()) for g in gs))
    before = [p.clone() for p in mod.params()]
    loss2, gcl, n2 = mod.train_step(ids, tgt, mask, lr=0.05, clipnorm=0.01)
    assert abs(float(n2) - norm) < 1e-12 and norm > 0.01
    for p0, p1, g in zip(before, mod.params(), gs):
        gc = g * (0.01 / norm)
        assert torch.allclose(p1, p0 - 0.05 * gc / (gc.abs() + 1e-8), atol=1e-12)
    # embedding rows that were never looked up keep weight and accumulator (row-sparse equivalence)
    unused = sorted(set(range(V)) - set(ids[mask].tolist()))
    if unused:
        assert torch.equal(mod.W_in[unused], before[0][unused]) and torch.all(mod.accum[0][unused] == 0)


def test_topk_ties_lower_id_first():
    p = torch.tensor([[0.1, 0.3, 0.3, 0.1, 0.2]])
    assert ks.topk_items(p, 3).tolist() == [[1, 2, 4]]
    u = torch.full((1, 6), 1.0 / 6)
    assert ks.topk_items(u, 4).tolist() == [[0, 1, 2, 3]]


def test_target_prob_pads_clip_to_eps():
    probs = torch.tensor([[[0.2, 0.8], [0.5, 0.5]]])
    p = ks.target_prob(probs, torch.tensor([[1, -1]]), torch.tensor([[True, False]]))
    assert abs(float(p[0, 0]) - 0.8) < 1e-7 and abs(float(p[0, 1]) - ks.EPS32) < 1e-12
