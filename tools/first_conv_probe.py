"""The first convolution (3 -> 64 channels at 384 x 1280, batch 2) alone: SIMT kernel (csrc/first_conv.cu) vs the tensor-core path, timed in a
graph of 10 launches; run under ncu (-k regex:first_conv) for the pipe picture."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epnet_b200.gemm import PackedConv3x3  # noqa: E402

dev = torch.device("cuda")
x = torch.zeros(2, 384, 1280, 4, device=dev)
x[..., :3] = torch.randn(2, 384, 1280, 3, device=dev)
conv = PackedConv3x3(torch.randn(64, 3, 3, 3, device=dev) / 5, torch.randn(64, device=dev), stride=1)
for label, w_c3 in (("simt", conv.w_c3), ("tensor-core", None)):
    conv.w_c3 = w_c3
    conv(x, relu=True, planes_out=True, f32_out=False)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            conv(x, relu=True, planes_out=True, f32_out=False)
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    print("first convolution, %s path, planes out: %.1f us" % (label, s.elapsed_time(e) * 100))
