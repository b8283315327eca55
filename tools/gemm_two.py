"""Two launches of the tcgen05 GEMM for an ncu capture: the first image convolution (3->64, 384x1280) and the 1x1 fusion conv
shape (983040 x 64 -> 32).  cudaProfilerStart/Stop bracket exactly one launch of each."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epnet_b200.gemm import PackedConv3x3, PackedLinear  # noqa: E402

conv = PackedConv3x3(torch.randn(64, 3, 3, 3, device="cuda") / 5, torch.randn(64, device="cuda"), stride=1)
x = torch.randn(2, 384, 1280, 4, device="cuda")
lin = PackedLinear(torch.randn(32, 64, device="cuda") / 8, torch.randn(32, device="cuda"))
X = torch.randn(983040, 64, device="cuda")
out = torch.empty(983040, 32, device="cuda")
for _ in range(3):
    conv(x, relu=True)
    lin(X, relu=True, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
conv(x, relu=True)
lin(X, relu=True, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
