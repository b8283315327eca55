"""CUPTI (torch.profiler) trace of the TIMED schedule -- the pipelined graph replays bench.py measures -- summarised per kernel:
launches, summed duration, and the UNION of the intervals during which at least one launch of the kernel (or of the GEMM family)
was running, per step.  The union is the kernel's share of wall time in the real schedule (<= ms_per_step by construction),
unlike durations of launches run alone.

    python tools/pipeline_timeline.py [depth=8] [steps=16] > profiles/r02_pipeline_timeline.txt
"""
import json
import os
import re
import sys
import tempfile
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
model = bench.build_model(dev)
runner = model.make_runner(2, 16384, dev, pipeline=depth)
pool = [{k: v.to(dev) for k, v in b.items()} for b in bench.make_pool(12, 1000)]


def run(n):
    for i in range(n):
        b = pool[i % len(pool)]
        runner(b["points"], b["image"], b["xy"])
    if depth > 1:
        runner.join()
    torch.cuda.synchronize()


run(2 * depth)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
run(steps)
e.record()
torch.cuda.synchronize()
plain_ms = s.elapsed_time(e) / steps
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(steps)
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [x for x in json.load(open(path))["traceEvents"] if x.get("cat") == "kernel"]
t0, t1 = min(x["ts"] for x in ev), max(x["ts"] + x["dur"] for x in ev)
span_ms = (t1 - t0) / 1e3


def union(intervals):
    tot, end = 0.0, -1.0
    for a, b in sorted(intervals):
        if a > end:
            tot += b - a
            end = b
        elif b > end:
            tot += b - end
            end = b
    return tot


def fam(name):
    if re.search(r"gemm_(f16x3|tf32x3)", name):
        return "GEMM family"
    m = re.search(r"epnet::(\w+)", name)
    return m.group(1) if m else name[:50]


groups = defaultdict(list)
for x in ev:
    groups[re.sub(r"\(.*", "", x["name"])[:70]].append((x["ts"], x["ts"] + x["dur"]))
fams = defaultdict(list)
for x in ev:
    fams[fam(x["name"])].append((x["ts"], x["ts"] + x["dur"]))
print("pipelined replay, depth %d, %d steps: %.3f ms/step plain (CUDA events), %.3f ms/step under CUPTI (kernel span)" % (depth, steps, plain_ms, span_ms / steps))
print("per step: launches, summed kernel duration (overlapping launches counted separately), union of busy intervals")
print("%-72s %8s %10s %10s %7s" % ("kernel", "launches", "sum_us", "union_us", "union%"))
for k, iv in sorted(list(fams.items()) + [("  " + k, v) for k, v in groups.items()], key=lambda kv: (not kv[0].startswith("  ") and -1e18 or 0) - union(kv[1])):
    u = union(iv)
    print("%-72s %8.1f %10.1f %10.1f %6.1f%%" % (k[:72], len(iv) / steps, sum(b - a for a, b in iv) / steps, u / steps, 100.0 * u / (t1 - t0)))
print("any kernel running: %.1f us/step of %.1f" % (union([(x["ts"], x["ts"] + x["dur"]) for x in ev]) / steps, (t1 - t0) / steps))
