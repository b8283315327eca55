"""Where a training step of the module path (forward + backward + Adam, train-mode BN, strict fp32) spends its device time."""
import os
import sys
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
model.train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
pool = [{k: v.to(dev) for k, v in b.items()} for b in bench.make_pool(2, 1000)]


def step(b):
    opt.zero_grad(set_to_none=True)
    _, feats = model(b["points"], b["image"], b["xy"].clone())
    (feats * feats).mean().backward()
    opt.step()


for i in range(4):
    step(pool[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        step(pool[i % 2])
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name[:100]][0] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
        agg[e.name[:100]][1] += 1
tot = sum(v[0] for v in agg.values())
print("device time per step: %.2f ms" % (tot / 3e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print("%9.2f ms  x%4d  %s" % (v[0] / 3e3, v[1] // 3, k))
