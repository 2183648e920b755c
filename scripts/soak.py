"""Soak run: many training steps on rotating batches (protocol races / rare deadlocks surface as a trap or a NaN).
python scripts/soak.py CONFIG STEPS"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from seq_recommendations_b200 import synthetic
from seq_recommendations_b200.engine import HotPath

name, steps = sys.argv[1], int(sys.argv[2])
cfg = synthetic.CONFIGS[name]
V, H, T, B = cfg["V"], cfg["H"], cfg["T"], cfg["B"]
hot = HotPath(cfg["cell"], cfg["act"], V, H, V, weights=synthetic.make_weights(cfg["cell"], V, H, seed=0))
hot.set_optimizer("adagrad", lr=0.01, epsilon=1e-8, clipnorm=1.0)
batches = [tuple(torch.from_numpy(a).cuda() for a in synthetic.make_batch(V, T, B, seed=s)) for s in range(8)]
t0 = time.time()
losses = []
for s in range(steps):
    loss = hot.train_batch(*batches[s % 8])
    if s % max(1, steps // 10) == 0 or s == steps - 1:
        losses.append(float(loss.item()))
torch.cuda.synchronize()
print(name, "steps", steps, "wall %.1fs" % (time.time() - t0), "losses", [round(x, 4) for x in losses])
assert all(np.isfinite(losses)) and losses[-1] < losses[0]
print("soak ok")
