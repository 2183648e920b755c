"""profiles/traffic.json from the raw CSV pages (`ncu -i x.ncu-rep --page raw --csv`) that scripts/profile_r2.sh brings
back in gpurun_out/; copies the pages to profiles/r2_<name>.raw.csv.  Where a capture holds hardware-counter sections only
(no dram__bytes_read/write.sum), the bytes are dram__bytes.sum.per_second x gpu__time_duration.sum and say so.

    python scripts/traffic_from_raw.py
"""
import csv
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = {"cfg4_gru256_1m": ["prof3_ce_cfg4", "prof3_dw_cfg4"], "cfg3_lstm256_50k": ["prof3_ce_cfg3", "prof3_cfg3_rnn"],
       "cfg2_reddit_gru128": ["prof3_cfg2"], "cfg1_msnbc_lstm100": ["prof3_cfg1"]}
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
RATE = {k + "/s": v for k, v in BYTES.items()}
TIME_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    out = {}
    for workload, files in SRC.items():
        entry = {}
        for f in files:
            path = os.path.join(ROOT, "gpurun_out", f + ".raw.csv")
            if not os.path.exists(path):
                continue
            rows = list(csv.reader(open(path)))
            hdr, units, data = rows[0], rows[1], rows[2:]

            def get(r, name, table=None):
                if name not in hdr:
                    return None
                i = hdr.index(name)
                if r[i] == "":
                    return None
                v = float(r[i].replace(",", ""))
                return v * table[units[i]] if table else v

            for r in data:
                name = r[hdr.index("Kernel Name")].replace("void ", "").replace("<unnamed>::", "").split("(")[0][:70]
                us = get(r, "gpu__time_duration.sum", TIME_US)
                rd, wr = get(r, "dram__bytes_read.sum", BYTES), get(r, "dram__bytes_write.sum", BYTES)
                d = {"duration_us": round(us, 2),
                     "tensor_pipe_active_pct_of_elapsed":
                         get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                     "dram_throughput_pct_of_peak": get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                     "l2_hit_rate_pct": get(r, "lts__t_sector_hit_rate.pct"),
                     "source": "ncu --clock-control none, scripts/profile_r2.sh (round 2, final kernels), raw page "
                               "profiles/r2_%s.raw.csv" % f}
                if rd is not None and wr is not None:
                    d.update(dram_bytes_per_launch=rd + wr, dram_read_bytes=rd, dram_write_bytes=wr)
                else:
                    d["dram_bytes_per_launch"] = get(r, "dram__bytes.sum.per_second", RATE) * us * 1e-6
                    d["dram_bytes_from"] = "dram__bytes.sum.per_second x gpu__time_duration.sum (counter sections only)"
                if name in entry:
                    # the same kernel in a second capture: the --set full one (read/write split) stays, the other is kept
                    # under a suffixed key that load_traffic's prefix match ignores
                    keep_new = "dram_read_bytes" in d and "dram_read_bytes" not in entry[name]
                    if keep_new:
                        entry["(also) " + name + " [" + f + "]"], entry[name] = entry[name], d
                    else:
                        entry["(also) " + name + " [" + f + "]"] = d
                else:
                    entry[name] = d
            shutil.copy(path, os.path.join(ROOT, "profiles", "r2_%s.raw.csv" % f))
        if entry:
            out[workload] = entry
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1, sort_keys=True)
    for workload, entry in out.items():
        for k, v in entry.items():
            print("| %s | `%s` | %.1f | %s | %.3g | %s | %s |" % (
                workload, k, v["duration_us"],
                "-" if v["tensor_pipe_active_pct_of_elapsed"] is None else "%.1f" % v["tensor_pipe_active_pct_of_elapsed"],
                v["dram_bytes_per_launch"] / 1e9,
                "-" if v["dram_throughput_pct_of_peak"] is None else "%.1f" % v["dram_throughput_pct_of_peak"],
                "-" if v["l2_hit_rate_pct"] is None else "%.1f" % v["l2_hit_rate_pct"]))


if __name__ == "__main__":
    main()
