"""Few-tile long-K GEMM (128 x 2048 -> 1024) on the narrow-tile kernel with 64-, 32- and 16-column tiles (2 / 3 / 3 operand stages in
TMEM), and on the TMA-fed FP16-split kernel from pre-split planes: which part of the per-k-block chain bounds it?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epnet_b200 import pointnet2_cuda as pc  # noqa: E402
from epnet_b200._lib import LIB  # noqa: E402
from epnet_b200.gemm import PackedLinear, Planes  # noqa: E402

dev = torch.device("cuda")


def timed(fn):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / 20


for L, K, N in ((128, 2048, 1024), (512, 1536, 512), (2048, 768, 512), (128, 512, 1024)):
    x = torch.randn(L, K, device=dev)
    lin = PackedLinear(torch.randn(N, K, device=dev) / K ** 0.5, torch.zeros(N, device=dev))
    out = torch.empty(L, N, device=dev)
    for bn in (64, 32, 16):
        wp = lin._pack(bn)
        us = timed(lambda: pc._call("gemm_tf32x3", LIB.epnet_gemm_tf32x3, x, L, K, N, x.data_ptr(), K, wp.data_ptr(), bn, lin.bias.data_ptr(), 1, 1,
                                    out.data_ptr(), N))
        print("L=%5d K=%5d N=%5d narrow-tile BN=%2d: %6.1f us" % (L, K, N, bn, us))
    h1 = x.half()
    px = Planes(h1, ((x - h1.float()) * 2048.0).half())
    us = timed(lambda: lin.from_planes(px, relu=True, out=out))
    print("L=%5d K=%5d N=%5d TMA-fed from planes: %6.1f us" % (L, K, N, us))
