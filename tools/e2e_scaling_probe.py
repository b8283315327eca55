"""Which part of the end-to-end step limits multi-GPU scaling?  Under torchrun (one rank per GPU) the pipelined runner is timed with the
host->device input copies and the device->host result copies switched on and off independently (max over ranks, ms per step).

    python -m torch.distributed.run --nproc-per-node N tools/e2e_scaling_probe.py [steps]
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
model = bench.build_model(dev)
depth = 8
runner = model.make_runner(2, 16384, dev, pipeline=depth)
host = bench.make_pool(4, 1000 + 1000 * rank, with_u8=True)
pinned = [{k: v.pin_memory() for k, v in b.items()} for b in host]
resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
out_host = [None] * depth
done = [None] * depth


def sync_ranks():
    if world > 1:
        dist.barrier()


def run(h2d, d2h, image_key="image"):
    def step(i):
        slot = i % depth
        src = pinned if h2d else resident
        b = src[i % len(src)]
        if done[slot] is not None:
            done[slot].synchronize()
        xyz, feats = runner(b["points"], b[image_key], b["xy"])
        st = runner.stream_of_last_call()
        with torch.cuda.stream(st):
            if d2h:
                if out_host[slot] is None:
                    out_host[slot] = (torch.empty(xyz.shape).pin_memory(), torch.empty(feats.shape).pin_memory())
                out_host[slot][0].copy_(xyz, non_blocking=True)
                out_host[slot][1].copy_(feats, non_blocking=True)
            done[slot] = torch.cuda.Event()
            done[slot].record(st)
    for i in range(depth + 2):
        step(i)
    runner.join()
    ms = bench.timed_region(step, steps, sync_ranks, runner.join)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() / steps


rows = [("resident inputs, results stay on the device", False, False, "image"), ("H2D only", True, False, "image"), ("D2H only", False, True, "image"),
        ("H2D + D2H (= e2e)", True, True, "image"), ("H2D (uint8 frame) + D2H", True, True, "image_u8")]
for name, h2d, d2h, key in rows:
    ms = run(h2d, d2h, key)
    if rank == 0:
        print("%d GPU(s)  %-46s %.4f ms/step  %.1f scenes/s" % (world, name, ms, 2 * world / ms * 1e3), flush=True)


def copy_rate(direction, mb=17.17, reps=40):
    """the link alone: `reps` pinned copies of one step's result size, all ranks at once -> GB/s per GPU (slowest rank) and in aggregate"""
    n = int(mb * 1e6 / 4)
    d = torch.empty(n, device=dev)
    h = torch.empty(n).pin_memory()
    src, dst = (d, h) if direction == "d2h" else (h, d)
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    sync_ranks()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gbs = reps * n * 4 / (t.item() * 1e-3) / 1e9
    if rank == 0:
        print("%d GPU(s)  %s alone, %.2f MB pinned copies on every rank at once: %.1f GB/s per GPU, %.1f GB/s in aggregate (%.3f ms per copy)" %
              (world, direction.upper(), mb, gbs, gbs * world, t.item() / reps), flush=True)


copy_rate("d2h")
copy_rate("h2d", mb=3.45)
copy_rate("h2d", mb=12.45)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
