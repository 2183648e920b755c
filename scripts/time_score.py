"""Phase times of the scoring path at cfg5 size: python scripts/time_score.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seq_recommendations_b200 import synthetic
from seq_recommendations_b200.engine import HotPath
from seq_recommendations_b200._lib import call, ptr

cfg = synthetic.CONFIGS["cfg5_score_gru256_100k"]
V, H, T = cfg["V"], cfg["H"], cfg["T"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["B"]
hot = HotPath(cfg["cell"], cfg["act"], V, H, V, weights=synthetic.make_weights(cfg["cell"], V, H, seed=0))
ids, tgt = synthetic.make_batch(V, T, B, seed=0)
di = torch.from_numpy(ids).cuda()


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


w = hot.work(B, T)
print("stage+format", timed(lambda: hot._stage(w, di, None)))
hot.prof = []
hot._forward_hidden(w, training=False); hot._mark("end"); torch.cuda.synchronize()
print("hidden phases", hot.phase_times_ms()); hot.prof = None
print("topk total", timed(lambda: hot.topk_batch(di, 20, last_step_only=True)))
wl = hot.work(B, 1)
print("last-step stats (ce fwd)", timed(lambda: hot._forward_ce(wl, with_targets=False)))
out_i = torch.empty((B, 20), dtype=torch.int32, device="cuda"); out_p = torch.empty((B, 20), device="cuda")
print("topk kernel", timed(lambda: call("seqrec_topk", ptr(wl.hout[0]), ptr(hot.W_out), None, ptr(wl.m), ptr(wl.s), ptr(out_i),
                                          ptr(out_p), B, H, V, 20, hot.stream)))
