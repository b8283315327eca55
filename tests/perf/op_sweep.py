"""BASELINE.json configs[4]: op-level sweep, N in {4k, 16k, 64k, 128k} x npoint M in {1k, 4k, 16k} (M < N) x nsample in {16, 32, 64}
x C in {64, 128, 256, 512}, B = 1 per GPU (B*C*M*ns*4 bytes capped at 3 GB), the B200 kernels next to the reference's own kernels
on the same GPU.  One JSON row per (op, shape): microseconds, speed-up, algorithmic GB/s and fraction of the measured HBM peak for
the memory-bound ops.

    python tests/perf/op_sweep.py out.json                      # 1 GPU, full grid, ours + reference kernels
    torchrun --nproc-per-node N tests/perf/op_sweep.py out.json # N GPUs: every rank runs the grid on its own cloud (the ops are
                                                                # per-scene: data parallel, no collective); times are the MAX over
                                                                # ranks, `gpus` and the aggregate rate are recorded; reference skipped"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench_roofline as br  # noqa: E402
from epnet_b200 import pointnet2_cuda as ours  # noqa: E402
from oracle import ref_cuda  # noqa: E402

world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
peak, _ = br.hbm_peak()
g = torch.Generator().manual_seed(rank)
rows = []


def t(fn, it=5):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / it * 1e3


def rec(op, shape, f_ours, f_ref, by=None):
    us = t(f_ours)
    if world > 1:  # slowest rank defines the job's time; every rank processed one cloud
        tt = torch.tensor([us], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        us = tt.item()
        f_ref = None
    r = {"op": op, "shape": shape, "ours_us": round(us, 1), "gpus": world, "clouds_per_s": round(world / us * 1e6, 1)}
    if f_ref is not None:
        r["ref_us"] = round(t(f_ref, it=2), 1)
        r["speedup"] = round(r["ref_us"] / r["ours_us"], 2)
    if by:
        r["algorithmic_gbs"] = round(by / r["ours_us"] / 1e3, 1)
        r["frac_hbm"] = round(by / r["ours_us"] / 1e3 / peak, 3)
    rows.append(r)
    if rank == 0:
        print(r, flush=True)


for N in (4096, 16384, 65536, 131072):
  for M in (1024, 4096, 16384):
    if M >= N:
        continue
    xyz = (torch.rand(1, N, 3, generator=g) * torch.tensor([80.0, 4.0, 70.0])).to(dev)
    temp = torch.empty(1, N, device=dev)
    idx = torch.empty(1, M, dtype=torch.int32, device=dev)

    def fps(be):
        temp.fill_(1e10)
        be.furthest_point_sampling_wrapper(1, N, M, xyz, temp, idx)
    rec("fps", [1, N, M], lambda: fps(ours), lambda: fps(ref_cuda))
    rows[-1]["ours_ns_per_iter"] = round(rows[-1]["ours_us"] * 1e3 / (M - 1), 1)
    new_xyz = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(1, M, 3)).contiguous()
    d2 = torch.empty(1, N, 3, device=dev)
    i3 = torch.empty(1, N, 3, dtype=torch.int32, device=dev)
    rec("three_nn", [1, N, M], lambda: ours.three_nn_wrapper(1, N, M, xyz, new_xyz, d2, i3), lambda: ref_cuda.three_nn_wrapper(1, N, M, xyz, new_xyz, d2, i3))
    for ns in (16, 32, 64):
        radius = {16: 1.0, 32: 2.0, 64: 3.0}[ns]
        bidx = torch.zeros(1, M, ns, dtype=torch.int32, device=dev)
        rec("ball_query", [1, N, M, radius, ns], lambda: ours.ball_query_wrapper(1, N, M, radius, ns, new_xyz, xyz, bidx),
            lambda: ref_cuda.ball_query_wrapper(1, N, M, radius, ns, new_xyz, xyz, bidx))
        if ours.SORTED_QUERY_MIN_N <= N <= ours.SORTED_QUERY_MAX_N:  # the runner's path: one sort per cloud, queries through the buckets
            bk = ours.bucket_cloud(xyz)
            rec("ball_query_sorted (sort excluded)", [1, N, M, radius, ns], lambda: ours.ball_query_sorted_wrapper(1, M, radius, ns, new_xyz, bk, bidx), None)
            if ns == 16:
                rec("bucket_cloud (the sort)", [1, N], lambda: ours.bucket_cloud(xyz), None)
        for C in (64, 128, 256, 512):
            if C * M * ns * 4 > 3e9:
                continue
            feats = torch.randn(1, C, N, device=dev)
            out = torch.empty(1, C, M, ns, device=dev)
            by = br.algorithmic_bytes("group_points", (1, C, N, M, ns))
            rec("group_points", [1, C, N, M, ns], lambda: ours.group_points_wrapper(1, C, N, M, ns, feats, bidx, out),
                lambda: ref_cuda.group_points_wrapper(1, C, N, M, ns, feats, bidx, out), by)
            gp = torch.zeros(1, C, N, device=dev)
            rec("group_points_grad", [1, C, N, M, ns], lambda: ours.group_points_grad_wrapper(1, C, N, M, ns, out, bidx, gp),
                lambda: ref_cuda.group_points_grad_wrapper(1, C, N, M, ns, out, bidx, gp), br.algorithmic_bytes("group_points_grad", (1, C, N, M, ns)))
            fpm = feats.transpose(1, 2).contiguous()
            kp = (C + 6) // 4 * 4
            opm = torch.empty(M * ns, kp, device=dev)
            rec("group_concat_pm", [1, C, N, M, ns], lambda: ours.group_concat_pm_wrapper(1, C, N, M, ns, xyz, new_xyz, fpm, bidx, opm), None,
                br.algorithmic_bytes("group_concat_pm", (1, C, N, M, ns, C, kp)))
            del feats, out, fpm, opm, gp
    for C in (64, 128, 256, 512):
        kf = torch.randn(1, C, M, device=dev)
        w = torch.rand(1, N, 3, device=dev)
        out = torch.empty(1, C, N, device=dev)
        by = br.algorithmic_bytes("three_interpolate", (1, C, M, N))
        rec("three_interpolate", [1, C, M, N], lambda: ours.three_interpolate_wrapper(1, C, M, N, kf, i3, w, out),
            lambda: ref_cuda.three_interpolate_wrapper(1, C, M, N, kf, i3, w, out), by)
        kpm = kf.transpose(1, 2).contiguous()
        opm = torch.empty(N, C, device=dev)
        rec("three_interpolate_concat_pm", [1, C, M, N], lambda: ours.three_interpolate_concat_pm_wrapper(1, C, M, N, 0, kpm, i3, w, None, opm), None, by)
        gidx = idx
        pts = torch.randn(1, C, N, device=dev)
        og = torch.empty(1, C, M, device=dev)
        rec("gather_points", [1, C, N, M], lambda: ours.gather_points_wrapper(1, C, N, M, pts, gidx, og),
            lambda: ref_cuda.gather_points_wrapper(1, C, N, M, pts, gidx, og), br.algorithmic_bytes("gather_points", (1, C, N, M)))
        del kf, w, out, kpm, opm, pts, og
  if N in (16384, 131072):  # LI-Fusion gather at the image-stream map sizes, n = N points
    for (C, H, W) in ((64, 192, 640), (128, 96, 320), (32, 384, 1280), (128, 384, 1280)):
        fmap = torch.randn(1, C, H, W, device=dev)
        xyn = (torch.rand(1, N, 2, generator=g) * 2 - 1).to(dev)
        og = torch.empty(1, C, N, device=dev)
        rec("grid_gather_bilinear", [1, C, H, W, N], lambda: ours.grid_gather_bilinear_wrapper(1, C, H, W, N, fmap, xyn, False, og),
            lambda: torch.nn.functional.grid_sample(fmap, xyn.unsqueeze(1), align_corners=False), br.algorithmic_bytes("grid_gather_bilinear", (1, C, H, W, N)))
        del fmap, og
if rank != 0:
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/op_sweep.json"
json.dump(rows, open(path, "w"), indent=1)
nan = float("nan")
with open(os.path.splitext(path)[0] + ".txt", "w") as f:
    f.write("# op-level sweep (BASELINE.json configs[4]), one cloud per GPU, %d GPU(s); tests/perf/op_sweep.py, CUDA events (max over ranks); reference "
            "kernels = oracle/_ref (unmodified sources built for sm_100a; ATen grid_sample for the LI-Fusion gather); frac = algorithmic bytes / time / "
            "measured HBM peak of ONE GPU\n" % world)
    f.write(f"{'op':28s} {'shape':34s} {'ours_us':>9s} {'ref_us':>10s} {'speedup':>8s} {'alg GB/s':>9s} {'frac':>6s}\n")
    for r in rows:
        f.write(f"{r['op']:28s} {str(r['shape']):34s} {r['ours_us']:9.1f} {r.get('ref_us', nan):10.1f} {r.get('speedup', nan):8.2f} "
                f"{r.get('algorithmic_gbs', nan):9.1f} {r.get('frac_hbm', nan):6.3f}\n")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
