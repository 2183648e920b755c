"""Parity of the tensor-core recurrent scan (csrc/rnn_tc.cu: thread-block clusters, tcgen05 M=64 atoms, hidden state
all-gathered through distributed shared memory) against the fp32 SIMT scan of the same library and the float64 oracle.
Tolerance: 1e-4 relative (north_star's fp32 bound; the recurrent products are 3-pass bf16 splits)."""
import numpy as np
import pytest
import torch

from seq_recommendations_b200 import synthetic
from seq_recommendations_b200.engine import HotPath

from gpu_util import as_t, make_pair, rel_err

pytestmark = pytest.mark.gpu

# (cell, act, H, T, B): ragged last cluster (B % 64 != 0), one and several clusters, both cluster sizes (H/32 = 4, 8)
CASES = [("LSTM", "tanh", 256, 1, 70), ("GRU", "tanh", 256, 1, 9), ("GRU", "tanh", 128, 2, 64),
         ("LSTM", "tanh", 256, 12, 70), ("LSTM", "relu", 256, 7, 64), ("GRU", "tanh", 256, 9, 130),
         ("GRU", "relu", 256, 5, 33), ("LSTM", "tanh", 128, 10, 96), ("GRU", "tanh", 128, 11, 50)]


def _pair(cell, act, V, H, seed, tc_scan):
    ws = synthetic.make_weights(cell, V, H, seed=seed)
    rng = np.random.default_rng(seed + 3)
    ws[2] = ws[2] + (rng.standard_normal(ws[2].shape) * 0.1).astype(np.float32)
    hot = HotPath(cell, act, V, H, V, weights=ws, tc="off")
    assert hot.Ut_hi is not None
    hot.rnn_tc = tc_scan
    return hot, ws


@pytest.mark.parametrize("cell,act,H,T,B", CASES)
def test_tc_scan_matches_simt_scan(cell, act, H, T, B):
    V = 500
    ids, tgt = synthetic.make_batch(V, T, B, seed=21, min_len=1)
    out = {}
    for tc_scan in (True, False):
        hot, _ = _pair(cell, act, V, H, 17, tc_scan)
        hid = hot.hidden_batch(ids).cpu().numpy()
        w = hot.work(B, T)
        out[tc_scan] = (hid, w.xg.cpu().numpy().copy(), w.cst.cpu().numpy().copy())
    assert rel_err(out[True][0], out[False][0]) <= 2e-5            # hidden outputs
    assert rel_err(out[True][1], out[False][1]) <= 2e-5            # saved gates
    if cell == "LSTM":
        assert rel_err(out[True][2], out[False][2]) <= 2e-5        # cell states
    # padded steps hold the state: the first valid step of every row starts from zero, outputs before it are zero
    pad = ids < 0
    assert np.all(out[True][0][pad & (np.cumsum(~pad, axis=1) == 0)] == 0.0)


@pytest.mark.parametrize("cell,act,H,T,B", CASES[:7])
def test_tc_scan_loss_and_gradients_match_oracle(cell, act, H, T, B):
    V = 400
    hot, ora, _ = make_pair(cell, act, V, H, seed=5, bias_scale=0.1, tc="off")
    hot.rnn_tc = True                          # (GRU-128 defaults to the register-resident scan)
    ids, tgt = synthetic.make_batch(V, T, B, seed=6, min_len=1)
    loss, grads, _ = hot.grad_batch(ids, tgt)
    rl, rg = ora.grads(as_t(ids), as_t(tgt), as_t(ids) >= 0)
    assert abs(loss - float(rl)) <= 1e-4 * abs(float(rl))
    for name, g, r in zip(["W_in", "U", "b", "W_out"], grads, rg):
        assert rel_err(g, r.numpy()) <= 1e-4, (name, rel_err(g, r.numpy()))


@pytest.mark.parametrize("cell,act,H,T,B", CASES)
def test_tc_backward_scan_matches_simt_scan(cell, act, H, T, B):
    """Same saved gates and dL/dhout into both backward scans: dxp (and the GRU's r*h operand) must agree."""
    from seq_recommendations_b200._lib import ACT, CELL, call, ptr
    V = 500
    ids, tgt = synthetic.make_batch(V, T, B, seed=22, min_len=1)
    out = {}
    for tc_scan in (True, False):
        hot, _ = _pair(cell, act, V, H, 18, False)          # the SAME (SIMT) forward feeds both backward scans:
        hot.hidden_batch(ids)                                # activation kinks must not flip between the two runs
        hot.rnn_tc = tc_scan
        w = hot.work(B, T)
        g = torch.Generator(device="cpu").manual_seed(5)
        w.dh.copy_(torch.randn(w.dh.shape, generator=g).to(w.dh.device) * w.mask.view(T, B, 1))
        hot._rnn_backward(w)
        torch.cuda.synchronize()
        out[tc_scan] = (w.xg.cpu().numpy().copy(), w.cst.cpu().numpy().copy())
    assert rel_err(out[True][0], out[False][0]) <= 5e-5, rel_err(out[True][0], out[False][0])
    if cell == "GRU":
        assert rel_err(out[True][1], out[False][1]) <= 5e-5


def test_tc_scan_long_sequence_many_clusters():
    """T = 200 (cfg5's length), 5 clusters: error does not grow with the number of exchange rounds."""
    V, H, T, B = 300, 256, 200, 300
    ids, _ = synthetic.make_batch(V, T, B, seed=4)
    hid = {}
    for tc_scan in (True, False):
        hot, _ = _pair("GRU", "tanh", V, H, 9, tc_scan)
        hid[tc_scan] = hot.hidden_batch(ids).cpu().numpy()
    assert rel_err(hid[True], hid[False]) <= 5e-5


@pytest.mark.parametrize("cell,H,T,B", [("LSTM", 256, 9, 70), ("GRU", 256, 12, 33), ("GRU", 128, 50, 256), ("LSTM", 128, 3, 5)])
def test_tc_weight_gradient_matches_simt_gemm(cell, H, T, B):
    """dU / db from the split-K tcgen05 GEMM (MN-major operands, wgrad_tc.cu) vs the fp32 SIMT GEMM on the same dxp."""
    V = 300
    ids, tgt = synthetic.make_batch(V, T, B, seed=31, min_len=1)
    out = {}
    for tc_w in (True, False):
        hot, _ = _pair(cell, "tanh", V, H, 23, False)
        hot.wgrad_tc = tc_w
        _, grads, _ = hot.grad_batch(ids, tgt)
        out[tc_w] = (grads[1], grads[2])
    assert rel_err(out[True][0], out[False][0]) <= 2e-5, rel_err(out[True][0], out[False][0])
    assert rel_err(out[True][1], out[False][1]) <= 1e-6
