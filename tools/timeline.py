"""Per-stream timeline of one runner step (eager schedule, same streams as the captured graph):
total busy time and span per CUDA stream, and the top kernels per stream."""
import json
import os
import sys
import tempfile
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
model = bench.build_model(dev)
runner = model.make_runner(2, 16384, dev, use_graph=False)
pool = [{k: v.to(dev) for k, v in b.items()} for b in bench.make_pool(2, 1000)]
for i in range(3):
    runner.eager(pool[i % 2]["points"], pool[i % 2]["image"], pool[i % 2]["xy"])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    runner.eager(pool[0]["points"], pool[0]["image"], pool[0]["xy"])
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
by_stream = defaultdict(list)
for e in ev:
    by_stream[e["args"].get("stream", e.get("tid"))].append(e)
t0 = min(e["ts"] for e in ev)
for s, es in sorted(by_stream.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    busy = sum(e["dur"] for e in es)
    print(f"\n== stream {s}: {len(es)} kernels, busy {busy/1e3:.3f} ms, span {min(e['ts'] for e in es)-t0:.0f}..{max(e['ts']+e['dur'] for e in es)-t0:.0f} us")
    agg = defaultdict(lambda: [0.0, 0])
    for e in es:
        agg[e["name"][:90]][0] += e["dur"]
        agg[e["name"][:90]][1] += 1
    for name, (d, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"   {d:9.1f} us  x{n:3d}  {name}")
