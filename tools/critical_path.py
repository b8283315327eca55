"""Kernel-by-kernel timeline (start/end offsets, stream) of one eager multi-stream step of the runner, to read the critical path."""
import json, os, sys, tempfile
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
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
path = os.path.join(tempfile.mkdtemp(), "t.json"); prof.export_chrome_trace(path)
ev = sorted([e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")], key=lambda e: e["ts"])
t0 = ev[0]["ts"]
streams = {}
for e in ev:
    s = streams.setdefault(e["args"].get("stream"), len(streams))
    if e["dur"] >= 8:
        print(f"{e['ts']-t0:8.0f} {e['ts']-t0+e['dur']:8.0f} s{s} {e['dur']:7.1f}us {e['name'][:60]}")
print("end", max(e["ts"] + e["dur"] for e in ev) - t0)
