"""Few-tile GEMMs with a long k-loop (the coarse levels of the backbone: 128-2048 rows, K = 512-2048) on the narrow-tile kernel: time
against K, N (number of CTAs) and rows, inside a CUDA graph of 20 launches so that the Python launch cost is not in the number."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epnet_b200 import gemm  # noqa: E402
from epnet_b200.gemm import PackedLinear  # noqa: E402

dev = torch.device("cuda")
cases = [(128, 2048, 1024), (128, 2048, 64), (128, 1024, 1024), (128, 512, 1024), (128, 256, 1024), (128, 128, 1024), (128, 32, 1024),
         (512, 1536, 512), (2048, 768, 512), (2048, 2048, 1024), (16384, 2048, 1024)]
for policy in ("latency", "throughput"):
    with gemm.tile_policy(policy):
        for L, K, N in cases:
            x = torch.randn(L, K, device=dev)
            lin = PackedLinear(torch.randn(N, K, device=dev) / K ** 0.5, torch.zeros(N, device=dev))
            out = torch.empty(L, N, device=dev)
            lin(x, relu=True, out=out)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(20):
                    lin(x, relu=True, out=out)
            g.replay()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            g.replay()
            e.record()
            torch.cuda.synchronize()
            us = s.elapsed_time(e) * 1e3 / 20
            bn, _ = lin.for_rows(L)
            print("%-10s L=%6d K=%5d N=%5d BN=%3d kblocks=%3d ctas=%4d: %7.1f us  (%.2f us per k-block, %.1f TFLOP/s fp32-equiv)" %
                  (policy, L, K, N, bn, (K + 31) // 32, ((L + 127) // 128) * ((N + bn - 1) // bn), us, us / ((K + 31) // 32), 2.0 * L * K * N / us / 1e6))
