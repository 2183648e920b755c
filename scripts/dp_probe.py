"""Why do the persistent logits kernels of a data-parallel cfg2 step take longer at N >= 4?  (torchrun, one rank per GPU)

Legs, rank 0's phase times of the same cfg2 step:
  A  process group initialised, no NCCL traffic yet, HotPath without a communicator
  B  the same after NCCL has been used (communicator exists)
  C  HotPath with the communicator (the data-parallel step)
and a CUPTI timeline (torch.profiler) of two steps of leg C: every kernel with its stream, start and duration, so what
is resident next to the logits kernels can be read off."""
import json
import os
import sys

import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seq_recommendations_b200 import dist, synthetic          # noqa: E402
from seq_recommendations_b200.engine import HotPath           # noqa: E402


def phases(hot, resident, steps=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for s in range(4):
        hot.train_batch(*resident[s % len(resident)])
    torch.cuda.synchronize()
    hot.use_graphs = False
    hot.prof = []
    for s in range(steps):
        flush.zero_()
        hot.train_batch(*resident[s % len(resident)])
    torch.cuda.synchronize()
    out = {k: round(v / steps, 4) for k, v in hot.phase_times_ms().items()}
    hot.prof = None
    return out


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist.init_from_env("nccl")
    cfg = synthetic.CONFIGS[os.environ.get("PROBE_CFG", "cfg2_reddit_gru128")]
    V, H, T, B = cfg["V"], cfg["H"], cfg["T"], cfg["B"]
    ws = synthetic.make_weights(cfg["cell"], V, H, seed=0)
    host = [synthetic.make_batch(V, T, B, seed=100 * rank + i) for i in range(4)]
    resident = [(torch.from_numpy(i).cuda(), torch.from_numpy(t).cuda()) for i, t in host]
    solo = dist.Comm.__new__(dist.Comm)
    solo.enabled, solo.group, solo.rank, solo.world = False, None, 0, 1

    def make(comm):
        h = HotPath(cfg["cell"], cfg["act"], V, H, V, weights=ws, comm=comm, seed=rank)
        h.set_optimizer("adagrad", lr=0.01, epsilon=1e-8, clipnorm=1.0)
        return h

    res = {}
    hot = make(solo)
    res["A_solo_before_nccl"] = phases(hot, resident)
    t = torch.ones(1, device="cuda")
    td.all_reduce(t)
    td.barrier()
    torch.cuda.synchronize()
    res["B_solo_after_nccl"] = phases(hot, resident)
    del hot
    comm = dist.Comm()
    hot = make(comm)
    res["C_dp"] = phases(hot, resident)
    # leg D: the same, every collective issued synchronously on the compute stream order (no overlap)
    os.environ["SEQREC_DP_SERIAL"] = "1"
    hot.dp_serial = True
    res["D_dp_serial"] = phases(hot, resident)
    hot.dp_serial = False

    # ---- timeline of two eager steps (rank 0)
    timeline = None
    try:
        from torch.profiler import profile, ProfilerActivity
        hot.use_graphs = False
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for s in range(3):
                hot.train_batch(*resident[s % 4])
            torch.cuda.synchronize()
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        timeline = sorted(((e.time_range.start, e.time_range.end - e.time_range.start, e.name[:70]) for e in ev))
        t0 = timeline[0][0] if timeline else 0
        timeline = [(round(a - t0, 1), round(d, 1), n) for a, d, n in timeline]
    except Exception as exc:                                       # CUPTI closed on the box: keep the phase legs
        timeline = "profiler unavailable: %r" % (exc,)
    td.barrier()
    if rank == 0:
        print(json.dumps({"world": world, "legs": res}, indent=1))
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/dp_probe_timeline_n%d.json" % world, "w") as f:
            json.dump(timeline, f)
    td.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
