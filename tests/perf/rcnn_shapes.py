"""Times the point ops at the RCNN stage's shapes (200 RoIs x 512 points: FPS 512->128->32, ball query ns=64, grouping C=128/256)
against the reference's kernels on the same GPU: many small problems instead of two big ones."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from epnet_b200 import pointnet2_cuda as ours  # noqa: E402
from oracle import ref_cuda  # noqa: E402

dev = torch.device("cuda:0")
R = 200


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


g = torch.Generator().manual_seed(0)
for n, m, radius, c in ((512, 128, 0.2, 128), (128, 32, 0.4, 128)):
    xyz = ((torch.rand(R, n, 3, generator=g) - 0.5) * torch.tensor([2.0, 2.0, 4.4])).to(dev)
    temp = torch.empty(R, n, device=dev)
    idx = torch.empty(R, m, dtype=torch.int32, device=dev)

    def fps(be):
        temp.fill_(1e10)
        be.furthest_point_sampling_wrapper(R, n, m, xyz, temp, idx)
    a, b = t(lambda: fps(ours)), t(lambda: fps(ref_cuda))
    print(f"fps        {R}x{n}->{m}: ours {a:7.1f} us  ref {b:7.1f} us  x{b / a:.1f}")
    new_xyz = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(R, m, 3)).contiguous()
    bidx = torch.zeros(R, m, 64, dtype=torch.int32, device=dev)
    a = t(lambda: ours.ball_query_wrapper(R, n, m, radius, 64, new_xyz, xyz, bidx))
    b = t(lambda: ref_cuda.ball_query_wrapper(R, n, m, radius, 64, new_xyz, xyz, bidx))
    print(f"ball_query {R}x{n}/{m} r={radius} ns=64: ours {a:7.1f} us  ref {b:7.1f} us  x{b / a:.1f}")
    feats = torch.randn(R, c, n, device=dev)
    out = torch.empty(R, c, m, 64, device=dev)
    a = t(lambda: ours.group_points_wrapper(R, c, n, m, 64, feats, bidx, out))
    b = t(lambda: ref_cuda.group_points_wrapper(R, c, n, m, 64, feats, bidx, out))
    by = 4.0 * R * (c * n + m * 64 + c * m * 64)
    print(f"group      {R}x{c}x{n} -> {m}x64: ours {a:7.1f} us ({by / a / 1e3:6.0f} GB/s)  ref {b:7.1f} us  x{b / a:.1f}")
