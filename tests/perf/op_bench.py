"""Per-op timing at the backbone's shapes: the B200 kernels vs the reference's own kernels (oracle/_ref).
CUDA-event timing, inputs resident, L2 left warm (shapes here are far below L2; these are latency numbers).
Usage: python tests/perf/op_bench.py [--iters 20] [--json out.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from epnet_b200 import pointnet2_cuda as ours  # noqa: E402
from epnet_b200 import scenes  # noqa: E402


def timeit(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(iters))
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--json", default=None)
    ap.add_argument("--no-ref", action="store_true")
    a = ap.parse_args()
    backends = {"ours": ours}
    if not a.no_ref:
        from oracle import ref_cuda
        if ref_cuda.available():
            backends["ref"] = ref_cuda
    B = a.batch
    data = scenes.batch(1000, B)
    xyz0 = data["points"].cuda()
    rows = []

    def rec(name, shape, fn_of_backend):
        r = {"op": name, "shape": shape}
        for k, be in backends.items():
            r[k + "_us"] = round(timeit(lambda: fn_of_backend(be), a.iters), 2)
        if "ref_us" in r:
            r["speedup"] = round(r["ref_us"] / r["ours_us"], 2)
        rows.append(r)
        print(r, flush=True)

    levels = [(16384, 4096, (0.1, 0.5), 0), (4096, 1024, (0.5, 1.0), 96), (1024, 256, (1.0, 2.0), 256), (256, 64, (2.0, 4.0), 512)]
    xyz = xyz0
    xyz_levels = [xyz0]
    for (N, M, radii, C) in levels:
        temp = torch.empty(B, N, device="cuda")
        idx = torch.empty(B, M, dtype=torch.int32, device="cuda")

        def fps(be):
            temp.fill_(1e10)
            be.furthest_point_sampling_wrapper(B, N, M, xyz, temp, idx)
        rec("fps(+fill)", f"B{B} N{N} M{M}", fps)
        ours.furthest_point_sampling_wrapper(B, N, M, xyz, temp.fill_(1e10), idx)
        new_xyz = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(B, M, 3)).contiguous()
        for r, ns in zip(radii, (16, 32)):
            bidx = torch.zeros(B, M, ns, dtype=torch.int32, device="cuda")
            rec("ball_query", f"B{B} N{N} M{M} r{r} ns{ns}", lambda be: be.ball_query_wrapper(B, N, M, r, ns, new_xyz, xyz, bidx))
            for c in sorted({3, C} - {0}):
                feats = torch.randn(B, c, N, device="cuda")
                out = torch.empty(B, c, M, ns, device="cuda")
                rec("group_points", f"B{B} C{c} N{N} M{M} ns{ns}", lambda be: be.group_points_wrapper(B, c, N, M, ns, feats, bidx, out))
                rec("group_points_grad", f"B{B} C{c} N{N} M{M} ns{ns}",
                    lambda be: be.group_points_grad_wrapper(B, c, N, M, ns, out, bidx, feats))
        xyz = new_xyz
        xyz_levels.append(xyz)
    for lvl, C in ((3, 1024), (2, 512), (1, 512), (0, 256)):
        unknown, known = xyz_levels[lvl], xyz_levels[lvl + 1]
        n, m = unknown.shape[1], known.shape[1]
        d2 = torch.empty(B, n, 3, device="cuda")
        idx = torch.empty(B, n, 3, dtype=torch.int32, device="cuda")
        rec("three_nn", f"B{B} n{n} m{m}", lambda be: be.three_nn_wrapper(B, n, m, unknown, known, d2, idx))
        w = torch.rand(B, n, 3, device="cuda")
        feats = torch.randn(B, C, m, device="cuda")
        out = torch.empty(B, C, n, device="cuda")
        rec("three_interpolate", f"B{B} C{C} m{m} n{n}", lambda be: be.three_interpolate_wrapper(B, C, m, n, feats, idx, w, out))
        rec("three_interpolate_grad", f"B{B} C{C} m{m} n{n}",
            lambda be: be.three_interpolate_grad_wrapper(B, C, n, m, out, idx, w, feats))
    # LI-Fusion gather vs ATen
    for (C, H, W, n) in ((64, 192, 640, 4096), (128, 96, 320, 1024), (256, 48, 160, 256), (512, 24, 80, 64), (32, 384, 1280, 16384)):
        fmap = torch.randn(B, C, H, W, device="cuda")
        xy = torch.rand(B, n, 2, device="cuda") * 2 - 1
        out = torch.empty(B, C, n, device="cuda")
        r = {"op": "grid_gather", "shape": f"B{B} C{C} {H}x{W} n{n}"}
        r["ours_us"] = round(timeit(lambda: ours.grid_gather_bilinear_wrapper(B, C, H, W, n, fmap, xy, False, out), a.iters), 2)
        r["ref_us"] = round(timeit(lambda: torch.nn.functional.grid_sample(fmap, xy.unsqueeze(1), align_corners=False), a.iters), 2)
        r["speedup"] = round(r["ref_us"] / r["ours_us"], 2)
        rows.append(r)
        print(r, flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
