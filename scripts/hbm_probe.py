#!/usr/bin/env python
"""Gather / scatter-add against the HBM roofline, standalone (no model): the two kernels at a BASELINE config's table
shape with Zipf or uniform ids.  Run once plainly for the CUDA-event numbers, and once under
`ncu --set full -k regex:gather_rows|scatter_add_rows` for dram__bytes / dram__throughput (profiles/).

    python scripts/hbm_probe.py --config cfg4_gru256_1m [--uniform] [--iters 20]
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seq_recommendations_b200 import synthetic  # noqa: E402
from seq_recommendations_b200._lib import call, ptr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg4_gru256_1m")
    ap.add_argument("--uniform", action="store_true", help="uniform ids over the catalog instead of Zipf(1.1)")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    cfg = synthetic.CONFIGS[args.config]
    G = {"simpleRNN": 1, "LSTM": 4, "GRU": 3}[cfg["cell"]]
    V, GH, T, B = cfg["V"], G * cfg["H"], cfg["T"], cfg["B"]
    N = B * T
    dev = torch.device("cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(0)
    ids_np = rng.integers(0, V, size=N).astype(np.int32) if args.uniform else synthetic.zipf_items(rng, V, N)
    ids = torch.from_numpy(ids_np).to(dev)
    mask = torch.ones(N, dtype=torch.uint8, device=dev)
    W = torch.randn((V, GH), dtype=torch.float32, device=dev)
    b = torch.zeros(GH, dtype=torch.float32, device=dev)
    xp = torch.empty((N, GH), dtype=torch.float32, device=dev)
    dW = torch.zeros((V, GH), dtype=torch.float32, device=dev)
    touched = torch.zeros(V, dtype=torch.int32, device=dev)
    rows = torch.empty(V, dtype=torch.int32, device=dev)
    n_rows = torch.zeros(1, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def gather():
        call("seqrec_gather_rows", ptr(W), ptr(b), ptr(ids), ptr(mask), None, ptr(xp), N, V, GH, st)

    def scatter():
        call("seqrec_scatter_add_rows", ptr(xp), ptr(ids), ptr(mask), None, ptr(dW), ptr(touched), ptr(rows), ptr(n_rows),
             N, V, GH, st)

    out = {"config": args.config, "ids": "uniform" if args.uniform else "zipf(1.1)", "V": V, "GH": GH, "N": N,
           "unique_rows": int(len(np.unique(ids_np)))}
    for name, fn, alg in (("gather", gather, N * (4 + 2 * GH * 4)), ("scatter_add", scatter, N * (4 + 3 * GH * 4))):
        for _ in range(3):
            fn()
        ms = []
        for _ in range(args.iters):
            flush.zero_()
            if name == "scatter_add":          # keep the touched-row bookkeeping of a real step: flags start at zero
                touched.zero_()
                n_rows.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = float(np.median(ms))
        out[name] = {"ms": t, "algorithmic_bytes": alg, "algorithmic_gbs": alg / (t * 1e-3) / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
