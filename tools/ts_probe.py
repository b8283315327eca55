"""The narrow-tile kernel at the set-abstraction level-1 shapes (K = 32): pooled and plain epilogue, timed alone; run under ncu for the
source-level stall picture."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epnet_b200 import gemm  # noqa: E402
from epnet_b200.gemm import PackedLinear  # noqa: E402

dev = torch.device("cuda")
x = torch.randn(262144, 32, device=dev)
cases = [(64, 32), (32, 1), (64, 1), (16, 1)]
with gemm.tile_policy("throughput"):
    for n, pool in cases:
        lin = PackedLinear(torch.randn(n, 32, device=dev) / 6, torch.zeros(n, device=dev))
        out = torch.empty(262144 // pool, n, device=dev)
        for _ in range(2):
            lin(x, relu=True, pool=pool, out=out)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            lin(x, relu=True, pool=pool, out=out)
        e.record()
        torch.cuda.synchronize()
        print("L=262144 K=32 N=%d pool=%d: %.1f us" % (n, pool, s.elapsed_time(e) * 200))
