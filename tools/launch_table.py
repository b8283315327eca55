"""Every product-kernel launch of one backbone step (eager, single stream, CUDA events around each launch), slowest first, with
the integer arguments of the C-ABI call (sizes), so the per-shape cost of the GEMM / conv / deconv launches is visible."""
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from epnet_b200 import pointnet2_cuda as pc  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
model = bench.build_model(dev)
from epnet_b200.runner import BackboneRunner  # noqa: E402
tiles = sys.argv[2] if len(sys.argv) > 2 else "latency"  # "throughput" = the tile policy of the pipelined (benchmarked) runners
runner = BackboneRunner(model, 2, 16384, dev, use_graph=False, tiles=tiles)
pool = [{k: v.to(dev) for k, v in b.items()} for b in bench.make_pool(2, 1000)]
for i in range(3):
    runner.eager(pool[i % 2]["points"], pool[i % 2]["image"], pool[i % 2]["xy"], single_stream=True)
torch.cuda.synchronize()
pc.PROFILE = []
runner.eager(pool[0]["points"], pool[0]["image"], pool[0]["xy"], single_stream=True)
torch.cuda.synchronize()
rec, pc.PROFILE = pc.PROFILE, None
rows = [(e0.elapsed_time(e1) * 1e3, name, args) for name, args, e0, e1 in rec]
tot = defaultdict(float)
for us, name, args in rows:
    tot[name] += us
print("total %.0f us over %d launches" % (sum(r[0] for r in rows), len(rows)))
for name, us in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"  {us:8.1f} us  {name}")
for us, name, args in sorted(rows, key=lambda r: -r[0])[: int(sys.argv[1]) if len(sys.argv) > 1 else 50]:
    print(f"{us:8.1f} us  {name:28s} {args}")
