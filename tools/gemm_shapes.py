"""Times the tcgen05 GEMM alone at the backbone's GEMM / conv / deconv shapes (B=2): one line per shape, microseconds."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epnet_b200.gemm import PackedConv3x3, PackedDeconv, PackedLinear  # noqa: E402


def t(fn, it=20):
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


shapes = [
    (131072, 3, 16, 1), (131072, 16, 16, 1), (131072, 16, 32, 16), (262144, 3, 32, 1), (262144, 32, 32, 1), (262144, 32, 64, 32),
    (32768, 99, 64, 1), (32768, 64, 64, 1), (32768, 64, 128, 16), (65536, 99, 64, 1), (65536, 64, 96, 1), (65536, 96, 128, 32),
    (8192, 259, 128, 1), (8192, 128, 196, 1), (8192, 196, 256, 16), (16384, 259, 128, 1), (16384, 196, 256, 32),
    (2048, 515, 256, 1), (2048, 256, 256, 1), (2048, 256, 512, 16), (4096, 515, 256, 1), (4096, 256, 384, 1), (4096, 384, 512, 32),
    (512, 1536, 512, 1), (512, 512, 512, 1), (2048, 768, 512, 1), (8192, 608, 256, 1), (32768, 256, 128, 1), (32768, 128, 128, 1),
    (32768, 256, 128, 1), (8192, 192, 96, 1), (2048, 512, 256, 1), (983040, 64, 32, 1), (128, 2048, 1024, 1), (512, 1024, 512, 1), (128, 1024, 256, 1), (128, 512, 256, 1),
]
tot = 0.0
for (L, K, N, pool) in shapes:
    x = torch.randn(L, (K + 3) // 4 * 4, device="cuda")[:, :K]
    lin = PackedLinear(torch.randn(N, K, device="cuda") / K ** 0.5, torch.randn(N, device="cuda"))
    us = t(lambda: lin(x, relu=True, pool=pool))
    tot += us
    print(f"gemm L={L:7d} K={K:5d} N={N:4d} pool={pool:2d} BN={lin.for_rows(L)[0]:3d}: {us:7.1f} us  {2.0 * L * K * N / us / 1e6:7.1f} TF")
print(f"gemm sum {tot:.0f} us")
tot = 0.0
for (cin, cout, H, W, stride) in [(3, 64, 384, 1280, 1), (64, 64, 384, 1280, 2), (64, 128, 192, 640, 1), (128, 128, 192, 640, 2),
                                  (128, 256, 96, 320, 1), (256, 256, 96, 320, 2), (256, 512, 48, 160, 1), (512, 512, 48, 160, 2)]:
    conv = PackedConv3x3(torch.randn(cout, cin, 3, 3, device="cuda") / (9 * cin) ** 0.5, torch.randn(cout, device="cuda"), stride=stride)
    x = torch.randn(2, H, W, conv.cin_p, device="cuda")
    us = t(lambda: conv(x, relu=True))
    tot += us
    ho, wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    print(f"conv {cin:3d}->{cout:3d} {H}x{W} s{stride} BN={conv.lin.for_rows(2 * ho * wo)[0]:3d}: {us:7.1f} us  {2.0 * 2 * ho * wo * 9 * cin * cout / us / 1e6:7.1f} TF")
print(f"conv sum {tot:.0f} us")
tot = 0.0
cat = torch.empty(2, 384, 1280, 64, device="cuda")
for i, (cin, k) in enumerate([(64, 2), (128, 4), (256, 8), (512, 16)]):
    de = PackedDeconv(torch.randn(cin, 16, k, k, device="cuda") / cin ** 0.5, None)
    x = torch.randn(2, 384 // k, 1280 // k, cin, device="cuda")
    us = t(lambda: de(x, cat[..., 16 * i:16 * i + 16]))
    tot += us
    print(f"deconv {cin:3d}->16 k={k:2d} BN={de.lin.for_rows(2 * (384 // k) * (1280 // k))[0]:3d}: {us:7.1f} us  {2.0 * 2 * 384 * 1280 * cin * 16 / us / 1e6:7.1f} TF")
print(f"deconv sum {tot:.0f} us")
