"""Turn ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches_r1c.csv           -> markdown table on stdout
    python scripts/summarize_ncu.py full gpurun_out/prof_x.ncu-rep [workload]      -> markdown table; with a workload
                                                                                      name also updates
                                                                                      profiles/traffic.json
Runs in the build container (ncu -i reads reports without a GPU)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
    "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name.split("(")[0][:70]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if r[ui] == "ns":
            v /= 1e3
        elif r[ui] == "ms":
            v *= 1e3
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, n, t, t / n, 100 * t / tot))


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def full(path, workload=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    cols = {m: hdr.index(m) for m in METRICS if m in hdr}
    names = [short(r[ki]) for r in data]
    print("| metric | " + " | ".join("`%s`" % n for n in names) + " |\n|---|" + "---|" * len(names))
    for m, i in cols.items():
        print("| %s (%s) | " % (m, units[i]) + " | ".join(r[i] for r in data) + " |")
    if workload:
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        t = json.load(open(tpath)) if os.path.exists(tpath) else {}
        entry = {}
        for r, n in zip(data, names):
            rd = to_bytes(r[cols["dram__bytes_read.sum"]], units[cols["dram__bytes_read.sum"]])
            wr = to_bytes(r[cols["dram__bytes_write.sum"]], units[cols["dram__bytes_write.sum"]])
            entry[n] = {"dram_bytes_per_launch": rd + wr, "source": os.path.basename(path)}
        t[workload] = entry
        json.dump(t, open(tpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
