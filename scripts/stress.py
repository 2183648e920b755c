"""Randomised shape sweep: tensor-core path (x3) vs the exact-fp32 SIMT path of the same library on loss, every gradient
and the top-k ranking.  python scripts/stress.py [N_CASES] [SEED]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from seq_recommendations_b200 import synthetic
from seq_recommendations_b200.engine import HotPath


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))


n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for case in range(n_cases):
    cell = ["GRU", "LSTM"][int(rng.integers(2))]
    act = ["tanh", "relu"][int(rng.integers(2))] if cell == "LSTM" else "tanh"
    H = int(rng.choice([32, 64, 96, 100, 128, 160, 192, 256]))
    V = int(rng.integers(260, 30000))
    T = int(rng.integers(1, 24))
    B = int(rng.integers(max(1, 130 // T), 260))
    ws = synthetic.make_weights(cell, V, H, seed=case)
    ws[3] = ws[3] * float(rng.uniform(1, 6))
    ids, tgt = synthetic.make_batch(V, T, B, seed=100 + case, min_len=1)
    out = {}
    for tc in ("x3", "off"):
        os.environ["SEQREC_RNN_TC"] = "" if tc == "x3" else "0"
        os.environ["SEQREC_WGRAD_TC"] = "1" if tc == "x3" else "0"
        hot = HotPath(cell, act, V, H, V, weights=ws, tc=tc)
        loss, grads, _ = hot.grad_batch(ids, tgt)
        k = min(20, V)
        ti, tp = hot.topk_batch(ids, k, last_step_only=True)
        out[tc] = (loss, grads, ti.cpu().numpy(), tp.cpu().numpy())
    errs = [rel(a, b) for a, b in zip(out["x3"][1], out["off"][1])]
    lerr = abs(out["x3"][0] / out["off"][0] - 1)
    same_top = float(np.mean(out["x3"][2] == out["off"][2]))
    ok = max(errs) <= 1e-4 and lerr <= 1e-5 and same_top >= 0.999
    bad += not ok
    print("%2d %s/%s V=%5d H=%3d T=%2d B=%3d  loss %.1e grads %s topk-agree %.4f %s" % (
        case, cell, act, V, H, T, B, lerr, ["%.0e" % e for e in errs], same_top, "ok" if ok else "MISMATCH"), flush=True)
print("stress:", "all ok" if bad == 0 else "%d MISMATCH" % bad)
sys.exit(1 if bad else 0)
