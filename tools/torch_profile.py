"""One eager backbone step under torch.profiler: which CUDA kernels (ours, cuDNN, ATen) the time goes to."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

impl = sys.argv[1] if len(sys.argv) > 1 else "ours"
tf32 = len(sys.argv) > 2 and sys.argv[2] == "tf32"
torch.backends.cudnn.allow_tf32 = tf32
torch.backends.cuda.matmul.allow_tf32 = tf32
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
model = bench.build_model(dev)  # reference arm: bench.py --impl reference (baseline/ref_arm.py)
model.auto_fast_inference = False
pool = [{k: v.to(dev) for k, v in b.items()} for b in bench.make_pool(2, 1000)]
with torch.no_grad():
    for i in range(3):
        model(pool[i % 2]["points"], pool[i % 2]["image"], pool[i % 2]["xy"].clone())
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        model(pool[0]["points"], pool[0]["image"], pool[0]["xy"].clone())
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
