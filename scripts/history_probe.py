#!/usr/bin/env python
"""The history-feature kernel (seqrec_history_features) against the HBM roofline, standalone: algorithmic bytes =
n_seqs * T * V * 4 written (+ the ragged corpus read once).  Run once plainly for the CUDA-event numbers and once under
`ncu --set full -k regex:history_features` for dram__bytes (profiles/).

    python scripts/history_probe.py [--iters 10]
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
from seq_recommendations_b200._lib import call, ptr  # noqa: E402

SHAPES = {  # name: (n_seqs, T, V, lengths)
    "msnbc_like (V=17, T=50, 200k sequences)": (200000, 50, 17, (2, 52)),
    "cfg2_like (V=10k, T=50, 512 sequences)": (512, 50, 10000, (26, 52)),
    "mid (V=1000, T=100, 4096 sequences)": (4096, 100, 1000, (51, 102)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak = 6538.6
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = {"hbm_peak_gbs": peak, "shapes": {}}
    rng = np.random.default_rng(0)
    for name, (n, T, V, (lo, hi)) in SHAPES.items():
        lens = rng.integers(lo, hi, size=n)
        offs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        items = torch.from_numpy(rng.integers(0, V, size=int(offs[-1])).astype(np.int32)).to(dev)
        d_offs = torch.from_numpy(offs).to(dev)
        table = torch.from_numpy(np.log(np.arange(hi + 1, dtype=np.float64) + 1.0).astype(np.float32)).to(dev)
        c = torch.empty((n, T, V), dtype=torch.float32, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)

        def run():
            call("seqrec_history_features", ptr(items), ptr(d_offs), ptr(c), n, T, V, 1, ptr(table), hi + 1, ptr(err), st)

        for _ in range(3):
            run()
        ms = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        alg = n * T * V * 4 + int(offs[-1]) * 4 + (n + 1) * 8
        t = float(np.median(ms))
        out["shapes"][name] = {"ms": round(t, 4), "algorithmic_bytes": alg, "algorithmic_gbs": round(alg / t / 1e6, 1),
                               "frac_of_hbm_peak": round(alg / t / 1e6 / peak, 3), "err_flag": int(err.item())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
