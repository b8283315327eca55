import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epnet_b200.gemm import PackedLinear, PackedConv3x3
X = torch.randn(65536, 512, device="cuda")
lin = PackedLinear(torch.randn(256, 512, device="cuda") / 22.6, torch.zeros(256, device="cuda"))
out = torch.empty(65536, 256, device="cuda")
for _ in range(3):
    lin(X, relu=True, out=out)
x = torch.randn(2, 96, 320, 128, device="cuda")
conv = PackedConv3x3(torch.randn(256, 128, 3, 3, device="cuda") / 34, torch.zeros(256, device="cuda"), stride=1)
for _ in range(3):
    conv(x, relu=True)
torch.cuda.synchronize()
print("ok")
