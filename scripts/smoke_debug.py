import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from oracle import keras_semantics as ks
from seq_recommendations_b200 import synthetic
from seq_recommendations_b200.engine import HotPath
V, H, T, B = 300, 32, 12, 24
for tc in ("x3", "off"):
  for cell, act in (("GRU", "tanh"), ("LSTM", "relu")):
    ws = synthetic.make_weights(cell, V, H, seed=1)
    ids, tgt = synthetic.make_batch(V, T, B, seed=2)
    hot = HotPath(cell, act, V, H, V, weights=ws, tc=tc)
    hot.set_optimizer("adagrad", lr=0.05, epsilon=1e-8, clipnorm=1.0)
    ora = ks.Model(cell, act, ws, dtype=torch.float64)
    ti, tt = torch.tensor(ids.astype(np.int64)), torch.tensor(tgt.astype(np.int64))
    loss, grads, _ = hot.grad_batch(ids, tgt)
    rl, rg = ora.grads(ti, tt, ti >= 0)
    ge = [float(np.linalg.norm(g - r.numpy()) / np.linalg.norm(r.numpy())) for g, r in zip(grads, rg)]
    for step in range(2):
        l = float(hot.train_batch(ids, tgt).item())
        ref, _, _ = ora.train_step(ti, tt, ti >= 0, lr=0.05, epsilon=1e-8, clipnorm=1.0)
    we = [float(np.linalg.norm(m - r) / np.linalg.norm(r)) for m, r in zip(hot.get_weights(), ora.numpy_weights())]
    print(tc, cell, "grad errs", ["%.1e" % e for e in ge], "weight errs", ["%.1e" % e for e in we])
