"""Time the recurrent scans alone (CUDA events): python scripts/time_rnn.py CELL H B T [tc|simt] [bwd]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from seq_recommendations_b200 import synthetic
from seq_recommendations_b200.engine import HotPath

cell, H, B, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
mode = sys.argv[5] if len(sys.argv) > 5 else "tc"
V = 1000
ws = synthetic.make_weights(cell, V, H, seed=0)
hot = HotPath(cell, "tanh", V, H, V, weights=ws, tc="off")
hot.rnn_tc = mode == "tc"
ids, tgt = synthetic.make_batch(V, T, B, seed=0)
reps = int(os.environ.get("REPS", "5"))
for _ in range(2):
    hot.hidden_batch(ids)
torch.cuda.synchronize()
w = hot.work(B, T)
hot.prof = []
for _ in range(reps):
    hot._forward_hidden(w, training=False)
    hot._mark("end")
torch.cuda.synchronize()
ph = hot.phase_times_ms()
print(cell, H, B, T, mode, "rnn_fwd ms", round(ph["rnn_fwd"] / reps, 4), "us/step", round(1e3 * ph["rnn_fwd"] / reps / T, 3))

if len(sys.argv) > 6 and sys.argv[6] == "bwd":
    w.dh.normal_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hot._rnn_backward(w)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        hot._forward_hidden(w, training=False)
        e0.record()
        hot._rnn_backward(w)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    print(cell, H, B, T, mode, "rnn_bwd ms", round(tot / reps, 4), "us/step", round(1e3 * tot / reps / T, 3))

if os.environ.get("TIMELINE"):
    import ctypes
    from seq_recommendations_b200 import _lib
    lib = _lib.load()
    buf = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
    lib.seqrec_rnn_tc_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
    names = {"fwd": ["hfull", "mma_issued", "acc_seen", "staged", "xw_staged", "xw_free", "xw_issued"],
             "bwd": ["enter_x", "acc_seen", "rfree_ok", "sent", "rfull_ok", "summed"]}
    for which in ("fwd", "bwd"):
        buf.zero_()
        if which == "fwd":
            hot._forward_hidden(w, training=False)
        else:
            hot._rnn_backward(w)
        torch.cuda.synchronize()
        a = buf.cpu().numpy().reshape(64, 8)
        t0 = a[2][a[2] > 0].min()
        print(which, names[which])
        for r in range(2, 8):
            print(which, "round", r, [int(x - t0) if x > 0 else None for x in a[r][:len(names[which])]])
    lib.seqrec_rnn_tc_debug_buffer(None)
