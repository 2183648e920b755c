"""Torch-CPU restatement of the Keras-2.0.x / Theano arithmetic under `model.fit` / `model.predict`
for the reference's recurrent next-item models.  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
PARITY UNPINNED (no Keras/Theano in this image; semantics follow SURVEY.md §8(c)).

Reference call sites restated here (all paths into /root/reference):
  model.py:334-336, 246      Masking(mask_value=0.0)              -> `derive_mask`
  model.py:345-352, 248-254  SimpleRNN / LSTM (+ GRU of the same Keras generation) -> `rnn_forward`
  model.py:362-363, 371-372  Dropout on y->z and z->y              -> `dropout_*` arguments
  model.py:382-384, 257      TimeDistributed(Dense)               -> `logits`
  model.py:397 + experiments_methods.py:42  softmax + categorical_crossentropy -> `masked_loss`
  experiments_methods.py:41  Adagrad(lr, epsilon=1e-8, clipnorm=1.) -> `clip_by_global_norm`, `adagrad_update`
  model.py:194-195, 106-112  predict / p(true item)               -> `predict_proba`, `target_prob`
Weight list order follows Keras `get_weights()`:  [W_in (F,G*H), U (H,G*H), b (G*H), W_out (H,V) (, b_out (V))].
"""
import math

import numpy as np
import torch

GATES = {"simpleRNN": 1, "LSTM": 4, "GRU": 3}

# Keras `_EPSILON = 10e-8`; Theano casts both clip bounds to floatx.
EPS32 = float(np.float32(1e-7))
ONE_MINUS_EPS32 = float(np.float32(1.0 - 1e-7))


def hard_sigmoid(a):
    """Theano `T.nnet.hard_sigmoid`: clip(0.2*a + 0.5, 0, 1) (Keras-2.0.x recurrent_activation default)."""
    return torch.clamp(0.2 * a + 0.5, 0.0, 1.0)


def activation(name):
    if name == "relu":
        return torch.relu
    if name == "tanh":
        return torch.tanh
    if name == "linear":
        return lambda a: a
    raise ValueError("unsupported activation %r" % (name,))


def derive_mask(x_dense):
    """Masking(mask_value=0.0): a timestep is kept iff any feature differs from 0 (model.py:335)."""
    return (x_dense != 0).any(axis=-1)


def input_projection(W_in, b, ids=None, mask=None, x_dense=None, in_scale=None):
    """Time-distributed `x . kernel + bias` that Keras precomputes for all t (implementation 0).

    ids path: a one-hot row times the kernel is the kernel row; a masked (all-zero) row gives the
    bias alone.  `in_scale` (B,T) is the y->z dropout factor applied to the one-hot input (0 or 1/(1-p)).
    """
    if x_dense is not None:
        xp = x_dense @ W_in
    else:
        safe = ids.clamp(min=0)
        xp = W_in[safe] * mask.unsqueeze(-1).to(W_in.dtype)
        if in_scale is not None:
            xp = xp * in_scale.unsqueeze(-1)
    if b is not None:
        xp = xp + b
    return xp


def rnn_forward(xp, U, mask, cell, act_name, rec_masks=None):
    """Theano `K.rnn` with a mask: state and output are held on masked steps, zero initial state.

    xp (B,T,G*H) precomputed input projection, U (H,G*H), mask (B,T) bool.
    rec_masks: optional list of G tensors (B,H) -- `recurrent_dropout` masks applied to h_{t-1}.
    Returns Hout (B,T,H).
    """
    B, T, GH = xp.shape
    G = GATES[cell]
    H = GH // G
    act = activation(act_name)
    h = xp.new_zeros(B, H)
    c = xp.new_zeros(B, H)
    outs = []
    rm = rec_masks if rec_masks is not None else [None] * G

    def hm(g):
        return h if rm[g] is None else h * rm[g]

    for t in range(T):
        m = mask[:, t].unsqueeze(1)
        x = xp[:, t]
        if cell == "simpleRNN":
            h_new = act(x + hm(0) @ U)
        elif cell == "LSTM":  # gate blocks i, f, c, o
            i = hard_sigmoid(x[:, 0:H] + hm(0) @ U[:, 0:H])
            f = hard_sigmoid(x[:, H:2 * H] + hm(1) @ U[:, H:2 * H])
            g = act(x[:, 2 * H:3 * H] + hm(2) @ U[:, 2 * H:3 * H])
            o = hard_sigmoid(x[:, 3 * H:4 * H] + hm(3) @ U[:, 3 * H:4 * H])
            c_new = f * c + i * g
            h_new = o * act(c_new)
            c = torch.where(m, c_new, c)
        elif cell == "GRU":  # gate blocks z, r, h; reset applied BEFORE the recurrent matmul
            z = hard_sigmoid(x[:, 0:H] + hm(0) @ U[:, 0:H])
            r = hard_sigmoid(x[:, H:2 * H] + hm(1) @ U[:, H:2 * H])
            hh = act(x[:, 2 * H:3 * H] + (r * hm(2)) @ U[:, 2 * H:3 * H])
            h_new = z * h + (1.0 - z) * hh
        else:
            raise ValueError(cell)
        h = torch.where(m, h_new, h)
        outs.append(h)
    return torch.stack(outs, dim=1)


def logits(Hout, W_out, b_out=None):
    z = Hout @ W_out
    if b_out is not None:
        z = z + b_out
    return z


def softmax_probs(z):
    """Keras softmax (max-subtracted) followed by the backend's renormalisation in categorical_crossentropy."""
    e = torch.exp(z - z.max(dim=-1, keepdim=True).values)
    s = e / e.sum(dim=-1, keepdim=True)
    return s


def masked_loss(z, targets, mask):
    """categorical_crossentropy (Theano backend) + Keras' masked objective: sum(ce*m)/sum(m).

    Returns (loss, ce (B,T), p_target (B,T) after clipping).
    """
    s = softmax_probs(z)
    p = s / s.sum(dim=-1, keepdim=True)
    p = torch.clamp(p, EPS32, ONE_MINUS_EPS32)
    safe = targets.clamp(min=0).unsqueeze(-1)
    p_y = torch.gather(p, -1, safe).squeeze(-1)
    mf = mask.to(z.dtype)
    ce = -torch.log(p_y) * mf
    loss = ce.sum() / mf.sum()
    return loss, ce, p_y


class Model:
    """Weights + hyper-parameters of one reference model (ytoz RNNFullModel or RNNBaseline)."""

    def __init__(self, cell, act_name, weights, dtype=torch.float32):
        self.cell = cell
        self.act = act_name
        self.dtype = dtype
        ws = [torch.tensor(np.asarray(w), dtype=dtype) for w in weights]
        self.W_in, self.U, self.b = ws[0], ws[1], ws[2]
        self.W_out = ws[3]
        self.b_out = ws[4] if len(ws) > 4 else None
        self.accum = None

    def params(self):
        ps = [self.W_in, self.U, self.b, self.W_out]
        if self.b_out is not None:
            ps.append(self.b_out)
        return ps

    def set_params(self, ps):
        self.W_in, self.U, self.b, self.W_out = ps[:4]
        if self.b_out is not None:
            self.b_out = ps[4]

    def numpy_weights(self):
        return [p.detach().numpy().copy() for p in self.params()]

    # -- forward -------------------------------------------------------------------------------------
    def hidden(self, ids=None, mask=None, x_dense=None, in_scale=None, rec_masks=None):
        if x_dense is not None and mask is None:
            mask = derive_mask(x_dense)
        xp = input_projection(self.W_in, self.b, ids=ids, mask=mask, x_dense=x_dense, in_scale=in_scale)
        return rnn_forward(xp, self.U, mask, self.cell, self.act, rec_masks=rec_masks), mask

    def loss(self, ids, targets, mask, x_dense=None, out_scale=None, in_scale=None, rec_masks=None):
        Hout, mask = self.hidden(ids=ids, mask=mask, x_dense=x_dense, in_scale=in_scale, rec_masks=rec_masks)
        if out_scale is not None:  # z->y Dropout: inverted-dropout factors (B,T,H)
            Hout = Hout * out_scale
        z = logits(Hout, self.W_out, self.b_out)
        return masked_loss(z, targets, mask)

    def predict_proba(self, ids=None, mask=None, x_dense=None):
        """model.predict: full (B,T,V) softmax probabilities (model.py:194-195)."""
        with torch.no_grad():
            Hout, mask = self.hidden(ids=ids, mask=mask, x_dense=x_dense)
            return softmax_probs(logits(Hout, self.W_out, self.b_out))

    def hidden_states(self, ids=None, mask=None, x_dense=None):
        with torch.no_grad():
            return self.hidden(ids=ids, mask=mask, x_dense=x_dense)[0]

    # -- one training step ---------------------------------------------------------------------------
    def grads(self, ids, targets, mask, **kw):
        ps = [p.detach().clone().requires_grad_(True) for p in self.params()]
        saved = self.params()
        self.set_params(ps)
        try:
            loss, _, _ = self.loss(ids, targets, mask, **kw)
            gs = torch.autograd.grad(loss, ps, allow_unused=True)
        finally:
            self.set_params(saved)
        gs = [g if g is not None else torch.zeros_like(p) for g, p in zip(gs, ps)]
        return loss.detach(), gs

    def train_step(self, ids, targets, mask, lr=0.01, epsilon=1e-8, clipnorm=1.0, trainable=None, **kw):
        """fwd + bwd + global-norm clip + Adagrad (experiments_methods.py:41).  Returns (loss, clipped grads, norm)."""
        loss, gs = self.grads(ids, targets, mask, **kw)
        ps = self.params()
        if trainable is None:
            trainable = [True] * len(ps)
        gs_t = [g for g, t in zip(gs, trainable) if t]
        gs_c, norm = clip_by_global_norm(gs_t, clipnorm)
        it = iter(gs_c)
        gs_full = [next(it) if t else torch.zeros_like(p) for p, t in zip(ps, trainable)]
        if self.accum is None:
            self.accum = [torch.zeros_like(p) for p in ps]
        new_ps = []
        for p, g, a, t in zip(ps, gs_full, self.accum, trainable):
            if t:
                p2, a2 = adagrad_update(p, g, a, lr, epsilon)
                a.copy_(a2)
                new_ps.append(p2)
            else:
                new_ps.append(p)
        self.set_params(new_ps)
        return loss, gs_full, norm


def clip_by_global_norm(grads, clipnorm):
    """Keras `clip_norm`: n = sqrt(sum_all sum(g^2)); g <- g*c/n if n >= c."""
    norm = torch.sqrt(sum((g * g).sum() for g in grads))
    if clipnorm is not None and clipnorm > 0 and float(norm) >= clipnorm:
        grads = [g * (clipnorm / norm) for g in grads]
    return grads, norm


def adagrad_update(p, g, a, lr, epsilon=1e-8):
    """Keras-2.0.x Adagrad: a += g^2; p -= lr*g/(sqrt(a)+eps)."""
    a2 = a + g * g
    p2 = p - lr * g / (torch.sqrt(a2) + epsilon)
    return p2, a2


def target_prob(probs, targets, mask):
    """model.py:108-110: max(pred * y_onehot, axis=2) clipped to [eps, 1-eps]; 0 rows (pads) clip to eps."""
    p = torch.gather(probs, -1, targets.clamp(min=0).unsqueeze(-1)).squeeze(-1)
    p = p * mask.to(p.dtype)
    return torch.clamp(p, EPS32, ONE_MINUS_EPS32)


def topk_items(probs, k):
    """Top-k oracle: stable argsort of -p, ties broken by lower item id."""
    p = probs.detach().numpy()
    order = np.argsort(-p, axis=-1, kind="stable")
    return order[..., :k].astype(np.int32)


# ---- weight initialisers with Keras-2.0.x defaults (model.py:345-352, 382-384) ------------------------------------
def glorot_uniform(rng, shape):
    limit = math.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


def glorot_normal(rng, shape):
    std = math.sqrt(2.0 / (shape[0] + shape[1]))
    return (rng.standard_normal(size=shape) * std).astype(np.float32)


def orthogonal(rng, shape):
    a = rng.standard_normal(size=shape)
    u, _, v = np.linalg.svd(a, full_matrices=False)
    q = u if u.shape == tuple(shape) else v
    return q.astype(np.float32)


def init_weights(rng, cell, F, H, V, out_bias=False):
    G = GATES[cell]
    W_in = glorot_uniform(rng, (F, G * H))
    U = np.concatenate([orthogonal(rng, (H, H)) for _ in range(G)], axis=1)
    b = np.zeros(G * H, dtype=np.float32)
    if cell == "LSTM":
        b[H:2 * H] = 1.0  # unit_forget_bias
    W_out = glorot_uniform(rng, (H, V))
    ws = [W_in, U, b, W_out]
    if out_bias:
        ws.append(np.zeros(V, dtype=np.float32))
    return ws


# ---- history-feature / skip branches ----------------------------------------------------------------------------------
class SkipModel:
    """RNNFullModel with ANY branch combination (model.py:322-403) and NoRecurrenceModel (model.py:264-319).

    weights: dict with the arrays the chosen branches own --
      W_in (F_in, G*H), U (H, G*H), b (G*H)        recurrent layer `z_to_z_output`; F_in = [V if y_to_z] + [Fx if x_to_z],
                                                   the y rows first (model.py:354 concatenate([masked_y, masked_x]))
      W_toy ((H if cell) + (Fx if x_to_y), V), b_out (V)   `to_y_output` / `x_to_y_output`; rows [z ; x] (model.py:377)
      A (V, V), a_bias (V)                         `y_to_y_output` / `y_output`
    Keras semantics restated: Masking per input; `concatenate` and `add` AND the masks of their masked
    inputs (the y_to_y Dense of RNNFullModel reads the RAW y_input, model.py:389, and carries no mask); Dropout factors
    are passed in; OnlyNonZeroDiagonal (model.py:48-66) multiplies the UPDATED kernel by [ones(skip, dim); eye(dim)].
    cell=None is NoRecurrenceModel: softmax(x.B + A[y_{t-1}])."""

    NAMES = ["W_in", "U", "b", "W_toy", "b_out", "A", "a_bias"]

    def __init__(self, cell, act_name, weights, y_to_z=True, x_to_z=False, x_to_y=False, y_to_y=False, diag_b=True,
                 dtype=torch.float64):
        self.cell, self.act, self.dtype = cell, act_name, dtype
        self.y_to_z, self.x_to_z, self.x_to_y, self.y_to_y, self.diag_b = y_to_z, x_to_z, x_to_y, y_to_y, diag_b
        self.p = {k: torch.tensor(np.asarray(v), dtype=dtype) for k, v in weights.items() if v is not None}
        self.accum = None

    def forward(self, p, ids, x, in_drop=None, out_scale=None, rec_masks=None):
        """Returns logits (B,T,V) and the loss mask (B,T).  ids (B,T) int64 (pad < 0) or None; x (B,T,Fx) or None."""
        V = (p["A"].shape[0] if "A" in p else p["W_toy"].shape[1])
        m_y = m_x = Y = None
        if ids is not None:
            m_y = ids >= 0
            Y = torch.nn.functional.one_hot(ids.clamp(min=0), V).to(self.dtype) * m_y.unsqueeze(-1).to(self.dtype)
        if x is not None:
            m_x = (x != 0).any(dim=-1)
            x = x * m_x.unsqueeze(-1).to(self.dtype)
        z = 0.0
        masks = []
        if self.cell is not None:
            parts, zm = [], []
            if self.y_to_z:
                parts.append(Y); zm.append(m_y)
            if self.x_to_z:
                parts.append(x); zm.append(m_x)
            z_in = torch.cat(parts, dim=-1)
            m_z = zm[0] if len(zm) == 1 else (zm[0] & zm[1])
            if in_drop is not None:
                z_in = z_in * in_drop
            xp = z_in @ p["W_in"]
            if "b" in p:
                xp = xp + p["b"]
            Hout = rnn_forward(xp, p["U"], m_z, self.cell, self.act, rec_masks=rec_masks)
            if out_scale is not None:
                Hout = Hout * out_scale
            toy_in = torch.cat([Hout, x], dim=-1) if self.x_to_y else Hout
            masks.append(m_z)
            if self.x_to_y:
                masks.append(m_x)
            z = toy_in @ p["W_toy"]
            if "b_out" in p:
                z = z + p["b_out"]
            if self.y_to_y:
                z = z + Y @ p["A"] + (p["a_bias"] if "a_bias" in p else 0.0)      # raw y_input: no mask of its own
        else:
            if self.y_to_y:
                z = z + Y @ p["A"] + (p["a_bias"] if "a_bias" in p else 0.0)
                masks.append(m_y)
            if self.x_to_y:
                z = z + x @ p["W_toy"] + (p["b_out"] if "b_out" in p else 0.0)
                masks.append(m_x)
        m_o = masks[0]
        for m in masks[1:]:
            m_o = m_o & m
        return z, m_o

    def loss(self, ids, x, targets, p=None, **kw):
        z, m_o = self.forward(self.p if p is None else p, ids, x, **kw)
        return masked_loss(z, targets, m_o)

    def predict_proba(self, ids, x):
        with torch.no_grad():
            z, _ = self.forward(self.p, ids, x)
            return softmax_probs(z)

    def grads(self, ids, x, targets, **kw):
        p = {k: v.detach().clone().requires_grad_(True) for k, v in self.p.items()}
        loss, _, _ = self.loss(ids, x, targets, p=p, **kw)
        names = [n for n in self.NAMES if n in p]
        gs = torch.autograd.grad(loss, [p[n] for n in names], allow_unused=True)
        return loss.detach(), {n: (g if g is not None else torch.zeros_like(p[n])) for n, g in zip(names, gs)}

    def constrain(self):
        """OnlyNonZeroDiagonal on the x rows of `to_y_output` / `x_to_y_output` (model.py:300-301, :379)."""
        if self.x_to_y and self.diag_b:
            W = self.p["W_toy"]
            dim = W.shape[1]
            skip = W.shape[0] - dim
            mask = torch.cat([torch.ones(skip, dim, dtype=self.dtype), torch.eye(dim, dtype=self.dtype)], dim=0)
            self.p["W_toy"] = W * mask

    def train_step(self, ids, x, targets, lr=0.01, epsilon=1e-8, clipnorm=1.0, frozen=(), **kw):
        loss, g = self.grads(ids, x, targets, **kw)
        names = [n for n in self.NAMES if n in self.p and n not in frozen]
        gs, norm = clip_by_global_norm([g[n] for n in names], clipnorm)
        if self.accum is None:
            self.accum = {n: torch.zeros_like(v) for n, v in self.p.items()}
        for n, gc in zip(names, gs):
            self.p[n], self.accum[n] = adagrad_update(self.p[n], gc, self.accum[n], lr, epsilon)
        self.constrain()
        return loss, norm
