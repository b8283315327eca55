"""The memory-bound ops at the bench's sweep shapes (bench_roofline.sweep_rooflines), plus a few more channel-major shapes, one line each.
    python tools/op_roofline_probe.py [tag]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_roofline as br  # noqa: E402
from epnet_b200 import pointnet2_cuda as pc  # noqa: E402

dev = torch.device("cuda:0")
peak, src = br.hbm_peak()
tag = sys.argv[1] if len(sys.argv) > 1 else ""
rows = br.sweep_rooflines(dev, peak)
g = torch.Generator().manual_seed(1)


def extra(name, ints, fn):
    t = br._event_time(fn)
    by = br.algorithmic_bytes(name, ints)
    rows.append({"op": name, "shape": list(ints), "us": round(t * 1e6, 1), "algorithmic_mb": round(by / 1e6, 1),
                 "achieved_gbs": round(by / t / 1e9, 1), "frac": round(by / t / 1e9 / peak, 3)})


for (C, N, M, ns) in ((512, 65536, 16384, 32), (512, 131072, 16384, 32), (256, 16384, 4096, 32), (128, 131072, 4096, 32)):
    pts = torch.randn(1, C, N, device=dev)
    idx = torch.randint(0, N, (1, M, ns), generator=g).int().to(dev)
    out = torch.empty(1, C, M, ns, device=dev)
    extra("group_points", (1, C, N, M, ns), lambda: pc.group_points_wrapper(1, C, N, M, ns, pts, idx, out))
    del pts, out
for (C, m, n) in ((256, 4096, 131072), (512, 16384, 65536), (512, 1024, 131072)):
    pts = torch.randn(1, C, m, device=dev)
    idx3 = torch.randint(0, m, (1, n, 3), generator=g).int().to(dev)
    w = torch.rand(1, n, 3, generator=g).to(dev)
    out = torch.empty(1, C, n, device=dev)
    extra("three_interpolate", (1, C, m, n), lambda: pc.three_interpolate_wrapper(1, C, m, n, pts, idx3, w, out))
    del pts, out
for (B, C, N, M, ns) in ((2, 64, 16384, 4096, 32), (2, 128, 4096, 1024, 32), (1, 512, 16384, 4096, 32)):  # training shapes: gradient
    go = torch.randn(B, C, M, ns, device=dev)
    idx = torch.randint(0, N, (B, M, ns), generator=g).int().to(dev)
    gp = torch.zeros(B, C, N, device=dev)
    extra("group_points_grad", (B, C, N, M, ns), lambda: pc.group_points_grad_wrapper(B, C, N, M, ns, go, idx, gp))
for (B, C, m, n) in ((2, 256, 4096, 16384), (2, 512, 1024, 4096)):  # feature-propagation shapes, forward and gradient
    pts = torch.randn(B, C, m, device=dev)
    idx3 = torch.randint(0, m, (B, n, 3), generator=g).int().to(dev)
    w = torch.rand(B, n, 3, generator=g).to(dev)
    out = torch.empty(B, C, n, device=dev)
    extra("three_interpolate", (B, C, m, n), lambda: pc.three_interpolate_wrapper(B, C, m, n, pts, idx3, w, out))
    gp = torch.zeros(B, C, m, device=dev)
    extra("three_interpolate_grad", (B, C, n, m), lambda: pc.three_interpolate_grad_wrapper(B, C, n, m, out, idx3, w, gp))
for (C, H, W, n) in ((64, 192, 640, 131072), (32, 384, 1280, 131072)):
    fmap = torch.randn(1, C, H, W, device=dev)
    xy = (torch.rand(1, n, 2, generator=g) * 2 - 1).to(dev)
    out = torch.empty(1, C, n, device=dev)
    extra("grid_gather_bilinear", (1, C, H, W, n), lambda: pc.grid_gather_bilinear_wrapper(1, C, H, W, n, fmap, xy, False, out))
    del fmap, out
print("# %s  peak %.0f GB/s (%s)  EPNET_STAGED_LONG_ROWS=%s EPNET_CM_ROWS=%s" % (tag, peak, src, os.environ.get("EPNET_STAGED_LONG_ROWS", ""), os.environ.get("EPNET_CM_ROWS", "")))
for r in rows:
    print("%-30s %-34s %9.1f us %8.1f MB %8.1f GB/s  frac %.3f" % (r["op"], r["shape"], r["us"], r["algorithmic_mb"], r["achieved_gbs"], r["frac"]))
