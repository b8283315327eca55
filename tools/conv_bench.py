"""Image-stream 3x3 convolutions: implicit tcgen05 3xTF32 GEMM on NHWC vs cuDNN fp32 (NCHW), batch 2."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epnet_b200.gemm import PackedConv3x3  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True


def t(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / it * 1e3


tot_o = tot_c = 0
for (ci, co, h, w, s) in [(3, 64, 384, 1280, 1), (64, 64, 384, 1280, 2), (64, 128, 192, 640, 1), (128, 128, 192, 640, 2), (128, 256, 96, 320, 1),
                          (256, 256, 96, 320, 2), (256, 512, 48, 160, 1), (512, 512, 48, 160, 2)]:
    x = torch.randn(2, ci, h, w, device="cuda")
    wt = torch.randn(co, ci, 3, 3, device="cuda") / (9 * ci) ** 0.5
    conv = PackedConv3x3(wt, torch.zeros(co, device="cuda"), stride=s)
    xn = torch.zeros(2, h, w, conv.cin_p, device="cuda")
    xn[..., :ci] = x.permute(0, 2, 3, 1)
    out = conv(xn)
    to = t(lambda: conv(xn, relu=True, out=out))
    with torch.no_grad():
        tc = t(lambda: torch.nn.functional.conv2d(x, wt, None, stride=s, padding=1))
    fl = 2.0 * out.numel() * 9 * ci
    tot_o += to
    tot_c += tc
    print(f"conv {ci:3d}->{co:3d} {h}x{w} s{s}: tcgen05 {to:8.1f} us ({fl/to/1e6:6.1f} TF fp32-equiv) | cuDNN {tc:8.1f} us ({fl/tc/1e6:5.1f} TF)")
print(f"sum: tcgen05 {tot_o:.0f} us, cuDNN {tot_c:.0f} us")
