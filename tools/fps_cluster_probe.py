"""FPS 16384 -> 4096 (B = 2): one CTA per scene vs a thread-block cluster of 2 / 4 CTAs per scene (EPNET_FPS_CLUSTER, read once per
process -> run once per setting); checks the indices against the C oracle and prints ns per iteration."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from epnet_b200 import pointnet2_cuda as pc, scenes  # noqa: E402

B, N, M = 2, 16384, 4096
pts = torch.stack([scenes.lidar_scene(1000 + i, N) for i in range(B)])
want = oracle.furthest_point_sampling(pts.numpy(), M)
x = pts.cuda()
temp = torch.full((B, N), 1e10, device="cuda")
idx = torch.empty(B, M, dtype=torch.int32, device="cuda")
pc.furthest_point_sampling_wrapper(B, N, M, x, temp, idx)
torch.cuda.synchronize()
ok = np.array_equal(idx.cpu().numpy(), want)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(10):
    temp.fill_(1e10)
    s.record()
    pc.furthest_point_sampling_wrapper(B, N, M, x, temp, idx)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
print("EPNET_FPS_CLUSTER=%s exact=%s  %.1f us  %.1f ns/iteration" % (os.environ.get("EPNET_FPS_CLUSTER", "-"), ok, min(ts) * 1e3, min(ts) * 1e6 / (M - 1)))
