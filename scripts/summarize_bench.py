"""Markdown tables from bench.py JSON lines (gpurun_out/*.log) for profiles/*_summary.md.

    python scripts/summarize_bench.py gpurun_out/bench_r2_full.log [more logs ...]
"""
import json
import sys


def lines(path):
    for l in open(path):
        l = l.strip()
        if l.startswith("{"):
            try:
                yield json.loads(l)
            except ValueError:
                pass


def row(name, d):
    rf = d.get("roofline") or {}
    clk = d.get("clocks") or {}
    e2e = d["e2e"]["value"] if isinstance(d.get("e2e"), dict) else d.get("e2e")
    plug = (d.get("e2e_plugin") or {}).get("value")
    return "| %s | %d | %.4g | %.4g | %s | %s | %.3f | %s | %s %s |" % (
        name, d.get("n_gpus", 1), d["ms_per_step"], d["value"], ("%.4g" % e2e) if e2e else "-",
        ("%.4g" % plug) if plug else "-", rf.get("frac", float("nan")), d.get("gpu_launches", "-"),
        clk.get("sm_mhz"), ",".join(clk.get("reasons", [])) or "none")


def phases(name, d):
    p = d.get("phases_ms") or {}
    tot = sum(p.values())
    return "| %s | " % name + " | ".join("%s %.3f" % (k, v) for k, v in p.items()) + " | sum %.3f |" % tot


def main():
    print("| workload | N | ms/step | seq/s | e2e seq/s | plugin seq/s | roofline frac | launches | SM MHz, reasons |")
    print("|---|---|---|---|---|---|---|---|---|")
    ph = []
    extra = []
    for path in sys.argv[1:]:
        for d in lines(path):
            if "config" not in d:
                continue
            name = d["config"].get("workload", "?")
            print(row(name, d))
            ph.append(phases(name, d))
            for k, v in (d.get("configs") or {}).items():
                v = dict(v)
                v.setdefault("n_gpus", d.get("n_gpus", 1))
                print(row(k, v))
                if v.get("phases_ms"):
                    ph.append(phases(k, v))
            if d.get("strong"):
                s = d["strong"]
                extra.append("strong scaling %s N=%d: global batch %d, %.4g ms/step, %.4g seq/s (e2e %.4g)" % (
                    name, d["n_gpus"], s["global_batch"], s["ms_per_step"], s["value"], s["e2e"]))
            if d.get("parity_check"):
                extra.append("parity_check %s N=%d: %s" % (name, d["n_gpus"], json.dumps(d["parity_check"])))
            if d.get("cpu_baseline"):
                c = d["cpu_baseline"]
                extra.append("cpu_baseline %s: %.4g %s on %d cores (%s): %s" % (name, c["value"], c["unit"], c["cores"],
                                                                              c["kind"], c["sample"]))
    print("\nPhase times (ms per step, eager launches with an event at every boundary):\n")
    for p in ph:
        print(p)
    print()
    for e in extra:
        print("* " + e)


if __name__ == "__main__":
    main()
