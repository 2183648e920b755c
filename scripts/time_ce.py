"""clock64 timeline of CTA 0 of the logits forward kernel at cfg2 size: python scripts/time_ce.py"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seq_recommendations_b200 import synthetic, _lib
from seq_recommendations_b200.engine import HotPath

cfg = synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2_reddit_gru128"]
V, H, T, B = cfg["V"], cfg["H"], cfg["T"], cfg["B"]
hot = HotPath(cfg["cell"], cfg["act"], V, H, V, weights=synthetic.make_weights(cfg["cell"], V, H, seed=0))
ids, tgt = synthetic.make_batch(V, T, B, seed=0)
for _ in range(3):
    hot.loss_batch(ids, tgt)
torch.cuda.synchronize()
lib = _lib.load()
buf = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
lib.seqrec_ce_tc_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
hot.loss_batch(ids, tgt)
torch.cuda.synchronize()
lib.seqrec_ce_tc_debug_buffer(None)
a = buf.cpu().numpy().reshape(64, 8)
names = ["tempty_ok", "full0_ok", "mmas_issued", "stage_commit", "tile_commit", "after_seg_last", "-", "before_tempty"]
t0 = a[8][a[8] > 0].min()
print(names)
for r in range(8, 22):
    print(r, [int(x - t0) if x > 0 else None for x in a[r][:8]])
