"""The oracle against INDEPENDENT implementations that exist in this image (torch.nn), where the two are defined to
compute the same thing.  Keras/Theano cannot run here (parity unpinned), so these are the strongest
outside checks available of the oracle's structure: gate order and cell update of the LSTM, the SimpleRNN recurrence,
back-propagation through time, and the masked-mean cross-entropy.

What is and is not covered:
  * SimpleRNN: Keras' `h = act(x.W + b + h.U)` IS torch.nn.RNN (weights transposed) -> forward and gradients compared
    as they are.
  * LSTM: Keras-2.0.x uses hard_sigmoid gates, torch.nn.LSTM logistic ones.  With the oracle's gate function swapped
    for the logistic one (monkeypatch) both are the same network with the same gate order (i, f, c|g, o), so wiring,
    cell update and BPTT are checked; the hard_sigmoid itself is pinned by its own known-answer test
    (test_oracle.py::test_hard_sigmoid_and_clip_constants).
  * GRU: NOT comparable -- torch.nn.GRU applies the reset gate AFTER the recurrent matmul, Keras-2.0.x before it
    (test_oracle.py::test_gru_reset_before_matmul_by_hand pins the Keras form by hand).
  * Loss: away from the 1e-7 clip, sum(ce*m)/sum(m) IS F.cross_entropy(..., ignore_index, reduction="mean").
"""
import numpy as np
import pytest
import torch

from oracle import keras_semantics as ks

D = torch.float64


def _weights(rng, F, H, G):
    W = torch.tensor(rng.standard_normal((F, G * H)) * 0.4, dtype=D, requires_grad=True)
    U = torch.tensor(rng.standard_normal((H, G * H)) * 0.4, dtype=D, requires_grad=True)
    b = torch.tensor(rng.standard_normal(G * H) * 0.2, dtype=D, requires_grad=True)
    return W, U, b


def _load(mod, W, U, b):
    with torch.no_grad():
        mod.weight_ih_l0.copy_(W.t())
        mod.weight_hh_l0.copy_(U.t())
        mod.bias_ih_l0.copy_(b)
        mod.bias_hh_l0.zero_()


@pytest.mark.parametrize("act", ["relu", "tanh"])
def test_simple_rnn_equals_torch_nn_rnn(act):
    rng = np.random.default_rng(3)
    B, T, F, H = 5, 9, 7, 6
    W, U, b = _weights(rng, F, H, 1)
    x = torch.tensor(rng.standard_normal((B, T, F)), dtype=D)
    mask = torch.ones(B, T, dtype=torch.bool)
    h_or = ks.rnn_forward(x @ W + b, U, mask, "simpleRNN", act)
    ref = torch.nn.RNN(F, H, nonlinearity=act, batch_first=True).to(D)
    _load(ref, W, U, b)
    h_ref, _ = ref(x)
    assert torch.allclose(h_or, h_ref, rtol=1e-12, atol=1e-12)
    # BPTT: gradients of a scalar of the outputs w.r.t. kernel, recurrent kernel and bias
    probe = torch.tensor(rng.standard_normal((B, T, H)), dtype=D)
    gW, gU, gb = torch.autograd.grad((h_or * probe).sum(), [W, U, b])
    (h_ref * probe).sum().backward()
    assert torch.allclose(gW, ref.weight_ih_l0.grad.t(), rtol=1e-10, atol=1e-12)
    assert torch.allclose(gU, ref.weight_hh_l0.grad.t(), rtol=1e-10, atol=1e-12)
    assert torch.allclose(gb, ref.bias_ih_l0.grad, rtol=1e-10, atol=1e-12)


def test_lstm_wiring_equals_torch_nn_lstm_with_logistic_gates(monkeypatch):
    monkeypatch.setattr(ks, "hard_sigmoid", torch.sigmoid)
    rng = np.random.default_rng(4)
    B, T, F, H = 4, 11, 5, 8
    W, U, b = _weights(rng, F, H, 4)
    x = torch.tensor(rng.standard_normal((B, T, F)), dtype=D)
    mask = torch.ones(B, T, dtype=torch.bool)
    h_or = ks.rnn_forward(x @ W + b, U, mask, "LSTM", "tanh")
    ref = torch.nn.LSTM(F, H, batch_first=True).to(D)
    _load(ref, W, U, b)                         # same block order along 4H: i, f, c (torch: g), o
    h_ref, _ = ref(x)
    assert torch.allclose(h_or, h_ref, rtol=1e-12, atol=1e-12)
    probe = torch.tensor(rng.standard_normal((B, T, H)), dtype=D)
    gW, gU, gb = torch.autograd.grad((h_or * probe).sum(), [W, U, b])
    (h_ref * probe).sum().backward()
    assert torch.allclose(gW, ref.weight_ih_l0.grad.t(), rtol=1e-10, atol=1e-12)
    assert torch.allclose(gU, ref.weight_hh_l0.grad.t(), rtol=1e-10, atol=1e-12)
    assert torch.allclose(gb, ref.bias_ih_l0.grad, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("cell", ["simpleRNN", "LSTM"])
def test_left_padding_equals_running_the_unpadded_sequence(cell, monkeypatch):
    """Keras masking with pre-padding: state held at zero through the pad steps, so the outputs on the real steps are
    those of the unpadded sequence run by torch.nn on its own, and the outputs on pad steps are zero vectors."""
    if cell == "LSTM":
        monkeypatch.setattr(ks, "hard_sigmoid", torch.sigmoid)
    rng = np.random.default_rng(5)
    T, F, H, G = 10, 4, 5, ks.GATES[cell]
    W, U, b = _weights(rng, F, H, G)
    lengths = [10, 6, 1]
    x = torch.zeros(len(lengths), T, F, dtype=D)
    mask = torch.zeros(len(lengths), T, dtype=torch.bool)
    for i, n in enumerate(lengths):
        x[i, T - n:] = torch.tensor(rng.standard_normal((n, F)), dtype=D)
        mask[i, T - n:] = True
    h_or = ks.rnn_forward(x @ W + b, U, mask, cell, "tanh")
    ref = (torch.nn.RNN(F, H, nonlinearity="tanh", batch_first=True) if cell == "simpleRNN"
           else torch.nn.LSTM(F, H, batch_first=True)).to(D)
    _load(ref, W, U, b)
    for i, n in enumerate(lengths):
        h_ref, _ = ref(x[i:i + 1, T - n:])
        assert torch.allclose(h_or[i, T - n:], h_ref[0], rtol=1e-12, atol=1e-12)
        assert torch.count_nonzero(h_or[i, :T - n]) == 0


def test_masked_loss_equals_cross_entropy_with_ignore_index():
    rng = np.random.default_rng(6)
    B, T, V = 6, 7, 13
    z = torch.tensor(rng.standard_normal((B, T, V)) * 2.0, dtype=D, requires_grad=True)
    tgt = torch.tensor(rng.integers(0, V, size=(B, T)))
    mask = torch.tensor(rng.random((B, T)) > 0.35)
    mask[0, 0] = True
    loss, ce, p_y = ks.masked_loss(z, tgt, mask)
    ref_t = torch.where(mask, tgt, torch.full_like(tgt, -100))
    ref = torch.nn.functional.cross_entropy(z.reshape(-1, V), ref_t.reshape(-1), ignore_index=-100, reduction="mean")
    assert float(p_y.detach().min()) > 1e-6 and float(p_y.detach().max()) < 1 - 1e-6    # the 1e-7 clip is inactive
    assert torch.allclose(loss, ref, rtol=1e-12)
    g_or, = torch.autograd.grad(loss, z, retain_graph=True)
    g_ref, = torch.autograd.grad(ref, z)
    assert torch.allclose(g_or, g_ref, rtol=1e-10, atol=1e-14)


def test_one_model_step_equals_torch_modules_end_to_end(monkeypatch):
    """Embedding row lookup + LSTM + output projection + masked CE, gradients for every weight, against
    torch.nn.Embedding / LSTM / Linear / cross_entropy wired the same way (logistic gates on both sides)."""
    monkeypatch.setattr(ks, "hard_sigmoid", torch.sigmoid)
    rng = np.random.default_rng(7)
    B, T, V, H = 5, 8, 11, 6
    ws = [rng.standard_normal((V, 4 * H)) * 0.3, rng.standard_normal((H, 4 * H)) * 0.3,
          rng.standard_normal(4 * H) * 0.1, rng.standard_normal((H, V)) * 0.3]
    ora = ks.Model("LSTM", "tanh", [w.astype(np.float64) for w in ws], dtype=D)
    ids = torch.tensor(rng.integers(0, V, size=(B, T)))
    tgt = torch.tensor(rng.integers(0, V, size=(B, T)))
    mask = torch.ones(B, T, dtype=torch.bool)
    loss, grads = ora.grads(ids, tgt, mask)
    emb = torch.nn.Embedding(V, 4 * H).to(D)
    lstm = torch.nn.LSTM(4 * H, H, batch_first=True).to(D)
    out = torch.nn.Linear(H, V, bias=False).to(D)
    with torch.no_grad():
        emb.weight.copy_(torch.tensor(ws[0]))
        lstm.weight_ih_l0.copy_(torch.eye(4 * H, dtype=D))          # the lookup already IS x.W: identity input kernel
        lstm.weight_hh_l0.copy_(torch.tensor(ws[1]).t())
        lstm.bias_ih_l0.copy_(torch.tensor(ws[2]))
        lstm.bias_hh_l0.zero_()
        out.weight.copy_(torch.tensor(ws[3]).t())
    h, _ = lstm(emb(ids))
    ref = torch.nn.functional.cross_entropy(out(h).reshape(-1, V), tgt.reshape(-1), reduction="mean")
    ref.backward()
    assert abs(float(loss) - float(ref.detach())) <= 1e-12 * abs(float(ref.detach()))
    assert torch.allclose(grads[0], emb.weight.grad, rtol=1e-9, atol=1e-13)
    assert torch.allclose(grads[1], lstm.weight_hh_l0.grad.t(), rtol=1e-9, atol=1e-13)
    assert torch.allclose(grads[2], lstm.bias_ih_l0.grad, rtol=1e-9, atol=1e-13)
    assert torch.allclose(grads[3], out.weight.grad.t(), rtol=1e-9, atol=1e-13)
