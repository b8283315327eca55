"""tcgen05 3xTF32 GEMM vs cuBLAS sgemm (+ separate bias/ReLU pass) at the backbone's shared-MLP shapes (B=2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epnet_b200.gemm import PackedLinear  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False


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


shapes = [  # (rows L, K, N, pool)   SA levels: L = B*M*ns
    (131072, 3, 16, 1), (131072, 16, 16, 1), (131072, 16, 32, 16), (262144, 3, 32, 1), (262144, 32, 32, 1), (262144, 32, 64, 32),
    (32768, 99, 64, 1), (32768, 64, 64, 1), (32768, 64, 128, 16), (65536, 99, 64, 1), (65536, 64, 96, 1), (65536, 96, 128, 32),
    (8192, 259, 128, 1), (8192, 128, 196, 1), (8192, 196, 256, 16), (16384, 259, 128, 1), (16384, 196, 256, 32),
    (2048, 515, 256, 1), (2048, 256, 256, 1), (2048, 256, 512, 16), (4096, 515, 256, 1), (4096, 256, 384, 1), (4096, 384, 512, 32),
    (512, 1536, 512, 1), (512, 512, 512, 1), (2048, 768, 512, 1), (8192, 608, 256, 1), (32768, 256, 128, 1), (32768, 128, 128, 1),
    (32768, 256, 128, 1), (8192, 192, 96, 1), (2048, 512, 256, 1),
]
tot_o = tot_c = 0.0
for (L, K, N, pool) in shapes:
    x = torch.randn(L, (K + 3) // 4 * 4, device="cuda")[:, :K]
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    lin = PackedLinear(w, b)
    to = t(lambda: lin(x, relu=True, pool=pool))
    xc = x.t().contiguous()  # channel-major operand for the cuBLAS arm: (K, L)

    def cub():
        y = torch.matmul(w, xc)
        y.add_(b[:, None]).relu_()
        if pool > 1:
            y = y.view(N, L // pool, pool).max(-1).values
        return y
    tc = t(cub)
    flops = 2.0 * L * K * N
    byts = 4.0 * (L * K + L // pool * N)
    tot_o += to
    tot_c += tc
    print(f"L={L:7d} K={K:5d} N={N:4d} pool={pool:2d}: tcgen05 {to:8.1f} us ({flops/to/1e6:7.2f} TF, {byts/to/1e3:7.1f} GB/s) | cuBLAS+passes {tc:8.1f} us")
print(f"sum: tcgen05 {tot_o:.0f} us, cuBLAS {tot_c:.0f} us")
