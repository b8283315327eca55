"""Summarise an ncu CSV of one graph replay (ncu --graph-profiling node --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum[,sm__pipe_tensor_cycles_active...] --csv) into per-kernel-family totals:

    python tools/ncu_graph_traffic.py gpurun_out/r02_graph_nodes.csv profiles/r02_ncu_traffic.json [replays]

bench_roofline.py reads the JSON for `roofline.traffic` (DRAM bytes of the dominant family's launches in ONE step)."""
import csv
import json
import re
import sys
from collections import defaultdict


def family(name):
    if re.search(r"gemm_(f16x3|tf32x3)", name):
        return "gemm_family"
    m = re.search(r"epnet::(\w+)", name)
    return m.group(1) if m else re.sub(r"\(.*", "", name).strip()[:60]


def main(src, dst, replays=1):
    rows = []
    with open(src, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    # long format: one row per (launch ID, metric)
    per = defaultdict(dict)
    names = {}
    for r in rows:
        i = r.get("ID")
        names[i] = r.get("Kernel Name", "")
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = r.get("Metric Unit", "")
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit, 1.0)
        per[i][r["Metric Name"]] = v * scale
    fam = defaultdict(lambda: {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "seconds": 0.0})
    kernels = defaultdict(lambda: {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "seconds": 0.0})
    for i, m in per.items():
        for table, key in ((fam, family(names[i])), (kernels, re.sub(r"\(.*", "", names[i]).strip())):
            e = table[key]
            e["launches"] += 1
            e["dram_read"] += m.get("dram__bytes_read.sum", 0.0)
            e["dram_write"] += m.get("dram__bytes_write.sum", 0.0)
            e["seconds"] += m.get("gpu__time_duration.sum", 0.0)
    out = {"source": "%s (ncu --graph-profiling node, cold-cache serialised kernel nodes of %d graph replay(s); per replay below)" % (src, replays)}
    for table in (fam, kernels):
        for k, e in table.items():
            out[k] = {"launches": e["launches"] // replays, "dram_read": round(e["dram_read"] / replays), "dram_write": round(e["dram_write"] / replays),
                      "us": round(e["seconds"] / replays * 1e6, 1)}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    tot = sum(e["seconds"] for e in kernels.values()) / replays
    print("%-60s %8s %9s %9s %9s" % ("kernel", "launches", "us", "dramR_MB", "dramW_MB"))
    for k, e in sorted(kernels.items(), key=lambda kv: -kv[1]["seconds"]):
        print("%-60s %8d %9.1f %9.1f %9.1f" % (k[:60], e["launches"] // replays, e["seconds"] / replays * 1e6, e["dram_read"] / replays / 1e6,
                                                   e["dram_write"] / replays / 1e6))
    print("total device time per replay (serialised, cold): %.1f us" % (tot * 1e6))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1)
