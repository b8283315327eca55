"""Profiler view of the end-to-end loop of bench.py (host-pinned inputs -> H2D -> graph replay -> D2H): per-category device time,
memcpy durations and the busy fraction of the GPU, to see what separates `e2e` from the device-resident `value`."""
import json
import os
import sys
import tempfile

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda:0")
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model = bench.build_model(dev)
runner = model.make_runner(2, 16384, dev, pipeline=depth)
pinned = [{k: v.pin_memory() for k, v in b.items()} for b in bench.make_pool(4, 1000)]
out_host, done = [None] * depth, [None] * depth


def step(i):
    slot = i % depth
    hb = pinned[i % len(pinned)]
    if done[slot] is not None:
        done[slot].synchronize()
    xyz, feats = runner(hb["points"], hb["image"], hb["xy"])
    st = runner.stream_of_last_call()
    if out_host[slot] is None:
        out_host[slot] = (torch.empty(xyz.shape).pin_memory(), torch.empty(feats.shape).pin_memory())
    with torch.cuda.stream(st):
        out_host[slot][0].copy_(xyz, non_blocking=True)
        out_host[slot][1].copy_(feats, non_blocking=True)
        done[slot] = torch.cuda.Event()
        done[slot].record(st)


for i in range(24):
    step(i)
runner.join()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(24, 64):
        step(i)
    runner.join()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
gpu = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
t0, t1 = min(e["ts"] for e in gpu), max(e["ts"] + e["dur"] for e in gpu)
print("40 steps: GPU span %.2f ms -> %.3f ms/step" % ((t1 - t0) / 1e3, (t1 - t0) / 1e3 / 40))
for cat in ("kernel", "gpu_memcpy"):
    es = [e for e in gpu if e["cat"] == cat]
    print("  %-10s %5d events, %.2f ms total" % (cat, len(es), sum(e["dur"] for e in es) / 1e3))
mc = sorted((e for e in gpu if e["cat"] == "gpu_memcpy"), key=lambda e: -e["dur"])[:6]
for e in mc:
    print("   memcpy %-40s %8.1f us  %s" % (e["name"][:40], e["dur"], e["args"].get("bytes", "")))
cpu = [e for e in ev if e.get("cat") in ("cuda_runtime", "cuda_driver")]
agg = {}
for e in cpu:
    agg[e["name"]] = agg.get(e["name"], 0) + e["dur"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print("   host %-40s %8.2f ms" % (k[:40], v / 1e3))
