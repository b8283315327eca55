"""The seven TMA-fed convolutions of the image stream (planes in), alone on the GPU: us per launch for the current column-tile policy
(EPNET_TMA_BN_MAX = 256 | 128)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epnet_b200.gemm import PackedConv3x3, Planes, tma_bn  # noqa: E402

dev = torch.device("cuda")
shapes = [(384, 1280, 64, 64, 2), (192, 640, 64, 128, 1), (192, 640, 128, 128, 2), (96, 320, 128, 256, 1), (96, 320, 256, 256, 2),
          (48, 160, 256, 512, 1), (48, 160, 512, 512, 2)]
tot = 0.0
for H, W, ci, co, st in shapes:
    x = torch.randn(2, H, W, ci, device=dev)
    h1 = x.half()
    px = Planes(h1, ((x - h1.float()) * 2048.0).half())
    conv = PackedConv3x3(torch.randn(co, ci, 3, 3, device=dev) / (3 * ci ** 0.5), torch.zeros(co, device=dev), stride=st)
    for _ in range(3):
        conv(px, relu=True, planes_out=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        conv(px, relu=True, planes_out=True)
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 100
    ho, wo = (H - 1) // st + 1, (W - 1) // st + 1
    gf = 2.0 * 2 * ho * wo * 9 * ci * co / 1e9
    tot += us
    print("conv %4dx%-4d %3d->%-3d s%d  BN=%3d  %7.1f us  %6.1f TFLOP/s fp32-equivalent" % (H, W, ci, co, st, tma_bn(co, 2 * ho * wo), us, gf / us * 1e3))
print("EPNET_TMA_BN_MAX=%s total %.1f us" % (os.environ.get("EPNET_TMA_BN_MAX", "256"), tot))
