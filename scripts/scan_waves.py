"""Wave structure of the tensor-core scan on this device: the co-resident cluster count (cudaOccupancyMaxActiveClusters
through seqrec_rnn_tc_max_clusters) and the forward scan timed at batch sizes either side of it.

    python scripts/scan_waves.py [CELL H T]        (default LSTM 256 100)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from seq_recommendations_b200 import _lib, synthetic  # noqa: E402
from seq_recommendations_b200.engine import HotPath  # noqa: E402

cell = sys.argv[1] if len(sys.argv) > 1 else "LSTM"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
T = int(sys.argv[3]) if len(sys.argv) > 3 else 100
V = 1000
resident = int(_lib.load().seqrec_rnn_tc_max_clusters(_lib.CELL[cell], H))
out = {"cell": cell, "H": H, "T": T, "clusters_resident": resident, "rnn_fwd_ms": {}}
hot = HotPath(cell, "tanh", V, H, V, weights=synthetic.make_weights(cell, V, H, seed=0), tc="off")
hot.rnn_tc = True
for clusters in sorted({1, resident - 1, resident, resident + 1, 16, 2 * resident, 2 * resident + 1}):
    if clusters < 1:
        continue
    B = 64 * clusters
    ids, _ = synthetic.make_batch(V, T, B, seed=0)
    for _ in range(2):
        hot.hidden_batch(ids)
    w = hot.work(B, T)
    torch.cuda.synchronize()
    hot.prof = []
    for _ in range(5):
        hot._forward_hidden(w, training=False)
        hot._mark("end")
    torch.cuda.synchronize()
    out["rnn_fwd_ms"]["%d clusters (B=%d)" % (clusters, B)] = round(hot.phase_times_ms()["rnn_fwd"] / 5, 4)
    hot.prof = None
print(json.dumps(out))
