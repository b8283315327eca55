"""bench.py --impl reference: the UNMODIFIED reference, through its own public API.

The network is `Pointnet2MSG` of the reference's own lib/net/pointnet2_msg.py (staged byte for byte under baseline/_ref),
configured by the reference's own lib/config.py + tools/cfgs/LI_Fusion_with_attention_use_ce_loss.yaml, built from the
reference's own pointnet2_modules / pointnet2_utils / pytorch_utils, calling the reference's own CUDA kernels
(oracle/_ref/libpointnet2_ref.so = the four *_gpu.cu files compiled unmodified for sm_100a) and ATen's grid_sample.
This process never imports epnet_b200 or the oracle package, so libepnet_b200.so / liboracle.so are not mapped
(`native_so_loaded` in the driver's record shows libpointnet2_ref.so only).

Two precisions are timed: "stock" = torch 2.11 defaults (cuDNN convolutions may use TF32, matmul fp32) -- what a user of
the reference gets on this box, and the line's headline `value`; "strict_fp32" = allow_tf32 False everywhere, the precision
the product's 1e-5 parity is stated at.  Same inputs, weights (torch.manual_seed(0)), batch, steps and warm-up as the
product arm."""
import contextlib
import json
import os
import sys

import torch

from . import ref_env

ROOT = ref_env.ROOT


def _set_precision(stock):
    torch.backends.cudnn.allow_tf32 = bool(stock)
    torch.backends.cuda.matmul.allow_tf32 = False  # torch's own default


def run(args, world, rank, local_rank, distributed, device, helpers):
    """helpers: bench.py's ClockSampler, timed_region, rank_sync, constants"""
    scenes = ref_env.load_by_path("_epnet_scenes", os.path.join(ROOT, "epnet_b200", "scenes.py"))  # pure torch, no dlopen
    shard = ref_env.load_by_path("_epnet_shard", os.path.join(ROOT, "epnet_b200", "shard.py"))
    B, N, POOL = helpers["BATCH_PER_GPU"], helpers["NPOINTS"], helpers["POOL"]
    rank_sync, timed_region = helpers["rank_sync"], helpers["timed_region"]

    with contextlib.redirect_stdout(sys.stderr):  # the reference's constructors print banners; stdout carries the JSON line only
        ref = ref_env.import_reference("reference")
        torch.manual_seed(0)
        model = ref.pointnet2_msg.Pointnet2MSG(input_channels=int(ref.cfg.RPN.USE_INTENSITY) + 3 * int(ref.cfg.RPN.USE_RGB),
                                                use_xyz=True).to(device).eval()  # lib/net/rpn.py:19-21

    host_pool = [scenes.batch(1000 + 1000 * rank + 10 * i, B, N) for i in range(POOL)]
    dev_pool = [{k: v.to(device) for k, v in b.items()} for b in host_pool]
    pinned = [{k: v.pin_memory() for k, v in b.items()} for b in host_pool[:4]]
    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned[0].values())
    out_host = [None]

    train = args.mode == "train"
    if train:  # BASELINE.json configs[2]: forward + backward + Adam, train-mode BN, DDP all-reduce (the reference uses nn.DataParallel)
        model.train()
        train_model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank]) if distributed else model
        opt = torch.optim.Adam(train_model.parameters(), lr=1e-4)

    def train_step(points, image, xy):
        opt.zero_grad(set_to_none=True)
        _, feats = train_model(points, image, xy)
        loss = (feats * feats).mean()
        loss.backward()
        opt.step()
        return loss

    def step_resident(i):
        b = dev_pool[i % POOL]
        if train:
            return train_step(b["points"], b["image"], b["xy"].clone())
        with torch.no_grad():
            return model(b["points"], b["image"], b["xy"].clone())  # lib/net/pointnet2_msg.py:208-210 normalises xy in place

    def step_e2e(i):
        hb = pinned[i % len(pinned)]
        pts = hb["points"].to(device, non_blocking=True)
        img = hb["image"].to(device, non_blocking=True)
        xy = hb["xy"].to(device, non_blocking=True)
        if train:
            loss = train_step(pts, img, xy)
            if out_host[0] is None:
                out_host[0] = (torch.empty((), dtype=torch.float32).pin_memory(),)
            out_host[0][0].copy_(loss.detach(), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return
        with torch.no_grad():
            xyz, feats = model(pts, img, xy)
        if out_host[0] is None:
            out_host[0] = (torch.empty(xyz.shape, dtype=xyz.dtype).pin_memory(), torch.empty(feats.shape, dtype=feats.dtype).pin_memory())
        out_host[0][0].copy_(xyz, non_blocking=True)
        out_host[0][1].copy_(feats, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller owns the result before the next step

    def measure(stock):
        _set_precision(stock)
        torch.backends.cudnn.benchmark = True
        for i in range(max(args.warmup, 3)):
            step_resident(i)
        torch.cuda.synchronize()
        ms = timed_region(step_resident, args.steps, rank_sync)
        for i in range(3):
            step_e2e(i)
        ms_e2e = timed_region(step_e2e, args.steps, rank_sync)
        return shard.max_over_ranks([ms, ms_e2e], device)

    sampler = helpers["ClockSampler"](local_rank)
    if rank == 0:
        sampler.start()
    ms, ms_e2e = measure(stock=True)
    clocks = sampler.stop() if rank == 0 else None
    ms_s, ms_e2e_s = measure(stock=False)
    d2h_bytes = sum(t.numel() * t.element_size() for t in out_host[0])
    total = B * world * args.steps

    def sps(t):
        return round(total / (t / 1e3), 3)

    loaded = sorted({ln.split()[-1] for ln in open("/proc/self/maps") if ROOT in ln and ".so" in ln})
    line = {
        "impl": "reference",
        "metric": "RPN backbone scenes/s (forward, 16384 pts + 384x1280 image, LI-Fusion with attention)" if not train else
                  "RPN backbone training scenes/s (forward+backward+Adam, train-mode BN, DDP all-reduce, 16384 pts + 384x1280 image)",
        "value": sps(ms), "unit": "scenes/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (torch 2.11 defaults: cuDNN convolutions may run TF32; strict fp32 in `strict_fp32`)", "data": "synthetic",
        "config": {"workload": helpers["WORKLOAD"], "batch_per_gpu": B, "npoints": N, "tf32": "stock torch defaults (cudnn.allow_tf32=True)",
                   "l2": "inputs rotate over a %d-batch resident pool (%.0f MB > 126 MB L2)" % (POOL, POOL * h2d_bytes / 1e6),
                   "parallelism": "dp%d (independent scenes per GPU, no collective in the forward)" % world,
                   "cuda_graph": False, "layout": "reference module path", "batches_in_flight": 1,
                   "reference": "lib/net/pointnet2_msg.py::Pointnet2MSG, unmodified, on the reference's own kernels + ATen grid_sample"},
        "e2e": {"value": sps(ms_e2e), "unit": "scenes/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": round(ms_e2e / args.steps, 4)},
        "strict_fp32": {"value": sps(ms_s), "ms_per_step": round(ms_s / args.steps, 4), "e2e_value": sps(ms_e2e_s),
                        "e2e_ms_per_step": round(ms_e2e_s / args.steps, 4), "note": "cudnn.allow_tf32=False, matmul.allow_tf32=False"},
        "gpu_launches": 0,
        "clocks": clocks,
        "cpu_baseline": {"value": sps(ms), "unit": "scenes/s", "cores": 0, "kind": "reference",
                         "sample": "not a CPU run: the reference's own Python (baseline/_ref, unmodified) on the reference's own pointnet2 "
                                   "CUDA kernels (oracle/_ref, unmodified sources) + ATen grid_sample, same B200, same inputs/weights -- "
                                   "BASELINE.json's second baseline; the reference has no CPU implementation of this path"},
        "repo_native_libraries_mapped": [os.path.relpath(p, ROOT) for p in loaded],
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
