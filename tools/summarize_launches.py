"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
agg = defaultdict(lambda: [0.0, 0])
for k, ns in rows:
    agg[k][0] += ns
    agg[k][1] += 1
total = sum(v[0] for v in agg.values())
print("launches: %d   total device time: %.3f ms (cold-cache, serialised: compare shares, not absolutes)" % (len(rows), total / 1e6))
print("%9s %7s %9s %7s  kernel" % ("total_us", "count", "avg_us", "share"))
for k, (ns, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print("%9.1f %7d %9.2f %6.1f%%  %s" % (ns / 1e3, n, ns / n / 1e3, 100 * ns / total, k[:110]))
ours = sum(v[0] for k, v in agg.items() if "epnet::" in k)
print("\nkernels of libepnet_b200.so: %.1f%% of device time, %d launches" % (100 * ours / total, sum(v[1] for k, v in agg.items() if "epnet::" in k)))
