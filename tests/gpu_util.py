"""Shared helpers of the -m gpu parity tests: the CUDA path (through the C-ABI) and the oracle on identical inputs."""
import numpy as np
import torch

from oracle import keras_semantics as ks
from seq_recommendations_b200 import synthetic
from seq_recommendations_b200.engine import HotPath


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_pair(cell, act, V, H, seed=0, out_bias=False, dtype=torch.float64, bias_scale=0.0, tc="off"):
    ws = synthetic.make_weights(cell, V, H, seed=seed, out_bias=out_bias)
    if bias_scale:
        rng = np.random.default_rng(seed + 7)
        ws[2] = ws[2] + (rng.standard_normal(ws[2].shape) * bias_scale).astype(np.float32)
        if out_bias:
            ws[4] = (rng.standard_normal(ws[4].shape) * bias_scale).astype(np.float32)
    hot = HotPath(cell, act, V, H, V, out_bias=out_bias, weights=ws, tc=tc)
    ora = ks.Model(cell, act, ws, dtype=dtype)
    return hot, ora, ws


def as_t(a):
    return torch.tensor(np.asarray(a).astype(np.int64))
