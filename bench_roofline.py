"""Roofline leg of bench.py: which of OUR kernels dominates the timed step, its algorithmic bytes (SURVEY.md
section 8d formulas) over its device time measured live with CUDA events, against the measured HBM peak;
plus the same figure for the four memory-bound ops at sweep sizes where they leave L2 (BASELINE.json
configs[4]) and the FPS latency per iteration."""
import json
import os

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(name, a):
    """SURVEY.md 8(d).  `a` = the integer arguments of the C-ABI call, in order."""
    if name == "furthest_point_sampling":
        b, n, m = a[:3]
        return b * (12 * n + 4 * m)
    if name in ("gather_points", "gather_points_grad"):
        b, c, n, m = a[:4]
        return b * (4 * c * min(n, m) + 4 * m + 4 * c * m) + (4 * b * c * n if name.endswith("grad") else 0)
    if name == "ball_query":
        b, n, m, ns = a[0], a[1], a[2], a[3]
        return b * (12 * n + 12 * m + 4 * m * ns)
    if name in ("group_points", "group_points_grad"):
        b, c, n, m, ns = a[:5]
        return b * (4 * c * min(n, m * ns) + 4 * m * ns + 4 * c * m * ns) + (4 * b * c * n if name.endswith("grad") else 0)
    if name == "three_nn":
        b, n, m = a[:3]
        return b * (12 * n + 12 * m + 24 * n)
    if name == "three_interpolate":
        b, c, m, n = a[:4]
        return b * (4 * c * min(m, 3 * n) + 24 * n + 4 * c * n)
    if name == "three_interpolate_grad":
        b, c, n, m = a[:4]
        return b * (4 * c * min(m, 3 * n) + 24 * n + 4 * c * n) + 4 * b * c * m
    if name.startswith("grid_gather_bilinear"):
        b, c, h, w, n = a[:5]
        return b * (4 * c * min(h * w, 4 * n) + 8 * n + 4 * c * n)
    if name == "fps_sample":
        b, n, m, aux = a[:4]
        return b * (12 * n + 4 * m + 12 * m + 8 * aux * m)
    if name == "group_concat_pm":  # (b, c, n, m, ns, ldf, ldo)
        b, c, n, m, ns, _, ldo = a[:7]
        return b * ((4 * c + 12) * min(n, m * ns) + 4 * m * ns + 4 * ldo * m * ns)
    if name == "three_interpolate_concat_pm":  # (b, c2, m, n, c1, ldk, lds, ldo)
        b, c2, m, n, c1 = a[:5]
        return b * (4 * c2 * min(m, 3 * n) + 24 * n + 4 * c1 * n + 4 * (c1 + c2) * n)
    if name == "grid_gather_pm":  # (b, c, h, w, n, align, ldo)
        b, c, h, w, n = a[:5]
        return b * (4 * c * min(h * w, 4 * n) + 8 * n + 4 * c * n)
    if name == "gemm_tf32x3":  # (L, K, N, ldx, BN, relu, pool, ldy)
        L, K, N = a[:3]
        pool = a[6]
        return 4 * (L * K + 2 * N * K + (L // max(pool, 1)) * N)
    if name == "bias_relu":
        b, c, l = a[:3]
        return 8 * b * c * l
    if name == "attention_scale_pm":  # (rows, rc, c, ld1, ld2, ldx, ldo): r1, r2, x in, out
        rows, rc, c = a[:3]
        return 4 * rows * (2 * rc + 2 * c)
    if name == "three_nn_weights":  # three_nn + the (n, 3) weights
        b, n, m = a[:3]
        return b * (12 * n + 12 * m + 24 * n + 12 * n)
    if name == "bucket_cloud":  # (b, n, npad): cloud in, (x, y, z, index) records + one box per 64 of them out
        b, n, npad = a[:3]
        return b * (12 * n + 16 * npad + 32 * (npad // 64))
    if name == "ball_query_sorted":  # (b, npad, m, nsample): records + boxes + centres in, idx out
        b, npad, m, ns = a[:4]
        return b * (16 * npad + 32 * (npad // 64) + 12 * m + 4 * m * ns)
    if name == "grid_gather_nhwc_pm":  # (b, c, h, w, n, ldc, align, ldo)
        b, c, h, w, n = a[:5]
        return b * (4 * c * min(h * w, 4 * n) + 8 * n + 4 * c * n)
    return 0


def _event_time(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3  # seconds


def sweep_rooflines(device, peak):
    """Memory-bound ops at sizes whose output alone exceeds L2 (126 MB), B=1."""
    from epnet_b200 import pointnet2_cuda as pc
    g = torch.Generator(device="cpu").manual_seed(0)
    rows = []

    def add(name, ints, fn):
        t = _event_time(fn)
        by = algorithmic_bytes(name, ints)
        rows.append({"op": name, "shape": list(ints), "us": round(t * 1e6, 1), "algorithmic_mb": round(by / 1e6, 1),
                     "achieved_gbs": round(by / t / 1e9, 1), "frac": round(by / t / 1e9 / peak, 3)})

    C, N, M, ns = 128, 65536, 16384, 32
    pts = torch.randn(1, C, N, device=device)
    idx = torch.randint(0, N, (1, M, ns), generator=g).int().to(device)
    out = torch.empty(1, C, M, ns, device=device)
    add("group_points", (1, C, N, M, ns), lambda: pc.group_points_wrapper(1, C, N, M, ns, pts, idx, out))
    gp = torch.zeros(1, C, N, device=device)
    add("group_points_grad", (1, C, N, M, ns), lambda: pc.group_points_grad_wrapper(1, C, N, M, ns, out, idx, gp))
    C, N, M = 512, 131072, 65536
    pts = torch.randn(1, C, N, device=device)
    gidx = torch.randint(0, N, (1, M), generator=g).int().to(device)
    out = torch.empty(1, C, M, device=device)
    add("gather_points", (1, C, N, M), lambda: pc.gather_points_wrapper(1, C, N, M, pts, gidx, out))
    C, m, n = 256, 32768, 131072
    pts = torch.randn(1, C, m, device=device)
    idx3 = torch.randint(0, m, (1, n, 3), generator=g).int().to(device)
    w = torch.rand(1, n, 3, generator=g).to(device)
    out = torch.empty(1, C, n, device=device)
    add("three_interpolate", (1, C, m, n), lambda: pc.three_interpolate_wrapper(1, C, m, n, pts, idx3, w, out))
    C, H, W, n = 128, 384, 1280, 131072
    fmap = torch.randn(1, C, H, W, device=device)
    xy = (torch.rand(1, n, 2, generator=g) * 2 - 1).to(device)
    out = torch.empty(1, C, n, device=device)
    add("grid_gather_bilinear", (1, C, H, W, n), lambda: pc.grid_gather_bilinear_wrapper(1, C, H, W, n, fmap, xy, False, out))
    # the same gather from an NHWC map (the layout of the runner's image stream): every tap is one contiguous, 128-bit vectorised channel run
    fmap_nhwc = fmap.permute(0, 2, 3, 1).contiguous()
    out_pm = torch.empty(n, C, device=device)
    t = _event_time(lambda: pc.grid_gather_nhwc_pm_wrapper(1, C, H, W, n, fmap_nhwc, xy, False, out_pm))
    by = algorithmic_bytes("grid_gather_nhwc_pm", (1, C, H, W, n))
    rows.append({"op": "grid_gather_nhwc_pm", "shape": [1, C, H, W, n], "us": round(t * 1e6, 1), "algorithmic_mb": round(by / 1e6, 1),
                 "achieved_gbs": round(by / t / 1e9, 1), "frac": round(by / t / 1e9 / peak, 3)})
    del fmap, fmap_nhwc, out_pm
    # point-major fused kernels (the runner's path): a gathered point is one contiguous row
    C, N, M, ns = 128, 65536, 16384, 32
    xyz = torch.randn(1, N, 3, device=device)
    new_xyz = xyz[:, :M].contiguous()
    feats_pm = torch.randn(1, N, C, device=device)
    idx = torch.randint(0, N, (1, M, ns), generator=g).int().to(device)
    kp = (C + 3 + 3) // 4 * 4
    out = torch.empty(M * ns, kp, device=device)
    t = _event_time(lambda: pc.group_concat_pm_wrapper(1, C, N, M, ns, xyz, new_xyz, feats_pm, idx, out))
    by = 4 * C * min(N, M * ns) + 4 * M * ns + 4 * kp * M * ns
    rows.append({"op": "group_concat_pm", "shape": [1, C, N, M, ns], "us": round(t * 1e6, 1), "algorithmic_mb": round(by / 1e6, 1),
                 "achieved_gbs": round(by / t / 1e9, 1), "frac": round(by / t / 1e9 / peak, 3)})
    C2, m, n = 256, 32768, 131072
    known = torch.randn(1, m, C2, device=device)
    idx3 = torch.randint(0, m, (1, n, 3), generator=g).int().to(device)
    d2 = torch.rand(1, n, 3, generator=g).to(device)
    out = torch.empty(n, C2, device=device)
    d2 = d2 / d2.sum(-1, keepdim=True)  # normalised weights
    t = _event_time(lambda: pc.three_interpolate_concat_pm_wrapper(1, C2, m, n, 0, known, idx3, d2, None, out))
    by = 4 * C2 * min(m, 3 * n) + 24 * n + 4 * C2 * n
    rows.append({"op": "three_interpolate_concat_pm", "shape": [1, C2, m, n], "us": round(t * 1e6, 1), "algorithmic_mb": round(by / 1e6, 1),
                 "achieved_gbs": round(by / t / 1e9, 1), "frac": round(by / t / 1e9 / peak, 3)})
    return rows


def gemm_tensor_rate(device):
    """tcgen05 3xTF32 GEMM at a large shared-MLP shape: fp32-equivalent TFLOP/s (2*L*K*N / time); the tensor pipe executes 3x that in TF32."""
    from epnet_b200.gemm import PackedLinear
    L, K, N = 65536, 512, 256
    x = torch.randn(L, K, device=device)
    lin = PackedLinear(torch.randn(N, K, device=device) / K ** 0.5, torch.zeros(N, device=device))
    out = torch.empty(L, N, device=device)
    t = _event_time(lambda: lin(x, relu=True, out=out))
    return {"kernel": "gemm_tf32x3", "shape": [L, K, N], "us": round(t * 1e6, 1), "fp32_equiv_tflops": round(2.0 * L * K * N / t / 1e12, 1),
            "tf32_mma_tflops": round(6.0 * L * K * N / t / 1e12, 1)}


def measure(model, runner, dev_pool, device, world, ms_per_step=None, batches_in_flight=1):
    from epnet_b200 import pointnet2_cuda as pc
    peak, peak_src = hbm_peak()
    # instrumented EAGER passes of the same step: CUDA events around every C-ABI launch on its own stream
    reps = 3
    with torch.no_grad():  # untimed eager pass: the eager schedule allocates its activations outside the graph pool the first time
        b = dev_pool[0]
        if runner is not None:
            runner.eager(b["points"], b["image"], b["xy"].clone(), single_stream=True)
        else:
            model(b["points"], b["image"], b["xy"].clone())
    pc.PROFILE = []
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        s.record()
        for r in range(reps):
            b = dev_pool[r % len(dev_pool)]
            if runner is not None:  # same kernels, one stream: per-kernel durations are not inflated by the concurrent branches
                runner.eager(b["points"], b["image"], b["xy"].clone(), single_stream=True)
            else:
                model(b["points"], b["image"], b["xy"].clone())
        e.record()
    torch.cuda.synchronize()
    prof, pc.PROFILE = pc.PROFILE, None
    step_ms = s.elapsed_time(e) / reps
    agg = {}
    for name, ints, e0, e1 in prof:
        key = (name, tuple(ints))
        t = e0.elapsed_time(e1) * 1e-3
        tot, cnt = agg.get(key, (0.0, 0))
        agg[key] = (tot + t, cnt + 1)
    ranked = sorted(agg.items(), key=lambda kv: -kv[1][0])
    # shares are of the summed device time of the product launches of one step (the wall time of this instrumented pass is
    # dominated by the host: an event pair and a Python call per launch)
    device_s = sum(v[0] for v in agg.values()) / reps
    kernels = [{"kernel": k[0], "args": list(k[1]), "launches_per_step": v[1] // reps, "us_per_launch": round(v[0] / v[1] * 1e6, 2),
                "share_of_device_time": round(v[0] / reps / device_s, 4)} for k, v in ranked[:12]]
    # per kernel NAME: total time, algorithmic bytes / flops over all its launches of one step
    by_name = {}
    for (name, ints), (tot, cnt) in agg.items():
        # plain, implicit-convolution and transposed-convolution launches are the same device code: gemm_tf32x3_kernel
        # (tiles wider than 64 columns) or gemm_tf32x3_ts_kernel (A operand through TMEM, tiles of <= 64 columns)
        family = ("gemm_tf32x3", "conv3x3_nhwc_tf32x3", "deconv_nhwc_tf32x3", "gemm_tf32x3_grouped", "gemm_tf32x3_cm", "conv3x3_planes_tma",
                  "gemm_planes_tma", "deconv_planes_tma", "gemm_tf32x3_rows")
        e = by_name.setdefault("gemm_tf32x3_kernel" if name in family else name,
                               {"time": 0.0, "launches": 0, "bytes": 0.0, "flops": 0.0})
        e["time"] += tot / reps
        e["launches"] += cnt // reps
        e["bytes"] += algorithmic_bytes(name, ints) * (cnt // reps)
        if name in ("ball_query", "three_nn", "three_nn_weights"):  # exhaustive searches: (b, n, m, ...) -> b*n*m distance tests
            e["tests"] = e.get("tests", 0.0) + float(ints[0]) * ints[1] * ints[2] * (cnt // reps)
        before = e["flops"]
        if name == "gemm_tf32x3":
            e["flops"] += 2.0 * ints[0] * ints[1] * ints[2] * (cnt // reps)
        if name == "conv3x3_nhwc_tf32x3":  # (b, h, w, cin, cout, stride, BN, relu, ldy)
            b_, h_, w_, ci_, co_, st_ = ints[:6]
            ho, wo = (h_ - 1) // st_ + 1, (w_ - 1) // st_ + 1
            e["flops"] += 2.0 * b_ * ho * wo * 9 * ci_ * co_ * (cnt // reps)
            e["bytes"] += 4.0 * (b_ * h_ * w_ * ci_ + 2 * 9 * ci_ * co_ + b_ * ho * wo * co_) * (cnt // reps)
        if name == "conv3x3_planes_tma":  # (b, h, w, cin, cout, stride, ldx, BN, relu, ldy, ldh): FP16 planes in (2 x 2 bytes per element)
            b_, h_, w_, ci_, co_, st_ = ints[:6]
            ho, wo = (h_ - 1) // st_ + 1, (w_ - 1) // st_ + 1
            e["flops"] += 2.0 * b_ * ho * wo * 9 * ci_ * co_ * (cnt // reps)
            e["bytes"] += 4.0 * (b_ * h_ * w_ * ci_ + 9 * ci_ * co_ + b_ * ho * wo * co_) * (cnt // reps)
        if name == "gemm_planes_tma":  # (L, K, N, ldx, BN, relu, pool, ldy, ldh)
            e["flops"] += 2.0 * ints[0] * ints[1] * ints[2] * (cnt // reps)
            e["bytes"] += 4.0 * (ints[0] * ints[1] + ints[1] * ints[2] + ints[0] // max(ints[6], 1) * ints[2]) * (cnt // reps)
        if name == "gemm_tf32x3_rows":  # (Lmax, K, N, ldx, phase_k, BN, relu, ldy): sparse image tail; Lmax = static upper bound of the rows
            e["flops"] += 2.0 * ints[0] * ints[1] * ints[2] * (cnt // reps)
            e["bytes"] += 4.0 * (ints[0] * ints[1] + ints[0] * ints[2]) * (cnt // reps)
        if name == "gemm_tf32x3_cm":  # (L, K, N, pts, ldx, BN, relu)
            e["flops"] += 2.0 * ints[0] * ints[1] * ints[2] * (cnt // reps)
            e["bytes"] += 4.0 * (ints[0] * ints[1] + 2 * ints[1] * ints[2] + ints[0] * ints[2]) * (cnt // reps)
        if name == "gemm_tf32x3_grouped":  # (scenes, n, m, ns, c, ldf, BN, N, relu, pool, ldy): rows gathered, never materialised
            rows_, k_, n_ = ints[0] * ints[2] * ints[3], ints[4] + 3, ints[7]
            e["flops"] += 2.0 * rows_ * k_ * n_ * (cnt // reps)
            e["bytes"] += 4.0 * (ints[0] * min(ints[1], ints[2] * ints[3]) * k_ + rows_ + 2 * k_ * n_ + rows_ // max(ints[9], 1) * n_) * (cnt // reps)
        if name == "deconv_nhwc_tf32x3":  # (b, h, w, cin, k, co, ldx, BN, relu, ldo): every input pixel -> a k x k x co patch
            b_, h_, w_, ci_, k_, co_ = ints[:6]
            e["flops"] += 2.0 * b_ * h_ * w_ * ci_ * k_ * k_ * co_ * (cnt // reps)
            e["bytes"] += 4.0 * (b_ * h_ * w_ * ci_ + 2 * ci_ * k_ * k_ * co_ + b_ * h_ * w_ * k_ * k_ * co_) * (cnt // reps)
        # tiles wider than 64 columns run the FP16 two-term split (kind::f16 MMAs: the bf16 peak), the others the TF32 split (half
        # of it); either way three MMAs per fp32-equivalent product
        bn_at = {"gemm_tf32x3": 4, "conv3x3_nhwc_tf32x3": 6, "deconv_nhwc_tf32x3": 7}.get(name)
        if e["flops"] > before:
            from epnet_b200 import gemm as _g
            f16 = (bn_at is not None and _g.F16_WIDE and ints[bn_at] > 64) or name.endswith("_planes_tma")
            e["ideal_tf32_flops"] = e.get("ideal_tf32_flops", 0.0) + 3.0 * (e["flops"] - before) * (0.5 if f16 else 1.0)
    bf16_peak = 1383.2
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            bf16_peak = float(json.load(f).get("bf16_tflops_sustained", bf16_peak))
    except Exception:  # noqa: BLE001
        pass
    per_kernel = []
    for name, e in sorted(by_name.items(), key=lambda kv: -kv[1]["time"]):
        row = {"kernel": name, "launches_per_step": e["launches"], "us_per_step": round(e["time"] * 1e6, 1),
               "share_of_device_time": round(e["time"] / device_s, 4)}
        if name == "gemm_tf32x3_kernel":
            row["device_kernels"] = ["gemm_tf32x3_ts_kernel (tiles <= 64 columns)", "gemm_f16x3_kernel (wider tiles)",
                                     "gemm_tf32x3_kernel (wider tiles with EPNET_F16_WIDE=0)"]
        if e["flops"] > 0:
            # frac = ALGORITHMIC flops (2*L*K*N) / time / dense-TF32 peak; the MMA work actually issued (3 MMAs per product, an
            # FP16-split MMA counting half) is tensor_pipe_frac.  Times here are of launches run ALONE (eager, single stream).
            work = e.get("ideal_tf32_flops", 3 * e["flops"])
            row.update({"bound": "tensor", "achieved": round(e["flops"] / e["time"] / 1e12, 1), "peak": round(bf16_peak / 2, 1),
                        "unit": "TFLOP/s", "frac": round(e["flops"] / e["time"] / 1e12 / (bf16_peak / 2), 4),
                        "tensor_pipe_frac": round(work / e["time"] / 1e12 / (bf16_peak / 2), 4)})
        else:
            row.update({"bound": "hbm", "achieved": round(e["bytes"] / e["time"] / 1e9, 1), "peak": peak, "unit": "GB/s",
                        "frac": round(e["bytes"] / e["time"] / 1e9 / peak, 5)})
        if e.get("tests"):  # the real bound of the search kernels is the fp32 distance-test rate, not bytes
            row["distance_tests_per_s"] = round(e["tests"] / e["time"], -9)
        # the bytes roofline is reported for every non-GEMM kernel, but it is not what limits these
        real_bound = {"fps_sample": "latency per iteration (M-1 serially dependent arg-max steps on one SM per scene): see fps_ns_per_iteration",
                      "furthest_point_sampling": "latency per iteration: see fps_ns_per_iteration",
                      "bucket_cloud": "latency (one CTA per scene sorts the cloud in shared memory, beside the FPS chain)",
                      "ball_query": "fp32 distance tests: see distance_tests_per_s", "ball_query_sorted": "fp32 distance tests on the overlapping buckets",
                      "three_nn": "fp32 distance tests: see distance_tests_per_s",
                      "three_nn_weights": "fp32 distance tests: see distance_tests_per_s"}.get(name)
        if real_bound:
            row["real_bound"] = real_bound
        per_kernel.append(row)
    top = dict(per_kernel[0])
    gemm_row = next((r for r in per_kernel if r["kernel"] == "gemm_tf32x3_kernel"), None)
    fam = by_name.get("gemm_tf32x3_kernel")
    if gemm_row is not None and fam is not None and ms_per_step:
        # SURVEY 8(d) / VERDICT r01 item 3: achieved = ALGORITHMIC flops (2*L*K*N summed over the family's launches of one step) over
        # time measured with CUDA events over the TIMED region -- the whole step, because inside the captured, pipelined schedule the
        # family's launches overlap each other and the non-GEMM kernels (no per-kernel event can be recorded inside a graph replay).
        # The family therefore gets the whole step's time charged: a lower bound on its own rate.  The tensor pipe executes
        # 3 MMAs per product (TF32 split) or 3 half-cost MMAs (FP16 split): that figure is `tensor_pipe_frac`, kept apart.
        step_s = ms_per_step * 1e-3
        dense_tf32 = bf16_peak / 2
        top = {"kernel": "GEMM family of the step: gemm_f16x3_kernel (tiles > 64 columns), gemm_tf32x3_ts_kernel (tiles <= 64 columns)",
               "launches_per_step": fam["launches"], "bound": "tensor", "unit": "TFLOP/s",
               "achieved": round(fam["flops"] / step_s / 1e12, 1), "peak": round(dense_tf32, 1),
               "frac": round(fam["flops"] / step_s / 1e12 / dense_tf32, 4),
               "algorithmic_gflop_per_step": round(fam["flops"] / 1e9, 1),
               "time_basis": "ms_per_step of the timed region (%.4f ms, CUDA events, %d batches in flight): the family is charged the WHOLE "
                             "step, which it shares with the non-GEMM kernels" % (ms_per_step, batches_in_flight),
               "frac_of_bf16_peak": round(fam["flops"] / step_s / 1e12 / bf16_peak, 4),
               "tensor_pipe_frac": round(fam.get("ideal_tf32_flops", 3 * fam["flops"]) / step_s / 1e12 / dense_tf32, 4),
               "tensor_pipe_note": "MMA work issued (3 MMAs per fp32-equivalent product; an FP16-split MMA counts half a TF32 MMA) over the "
                                   "same time and peak: the share of the step during which the tensor pipe would be busy at peak rate",
               "isolated": {"us_per_step": gemm_row["us_per_step"], "frac": round(fam["flops"] / fam["time"] / 1e12 / dense_tf32, 4),
                            "note": "each launch alone on the GPU (eager, single stream, CUDA events per launch): under-filled grids, "
                                    "NOT the timed schedule -- diagnostic only"}}
    traffic = None
    try:  # DRAM bytes of the family's launches in ONE graph replay from the committed ncu capture (tools/ncu_graph_traffic.py)
        with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as f:
            t = json.load(f)
        fam_t = t.get("gemm_family") if top["bound"] == "tensor" else t.get(top["kernel"])
        if fam_t:
            traffic = {"bytes_per_step": fam_t["dram_read"] + fam_t["dram_write"], "dram_read": fam_t["dram_read"], "dram_write": fam_t["dram_write"],
                       "launches": fam_t["launches"], "algorithmic_bytes_per_step": round(fam["bytes"]) if fam is not None else None,
                       "source": t.get("source")}
    except Exception:  # noqa: BLE001
        pass
    top.update({"traffic": traffic, "peak_source": peak_src if top["bound"] == "hbm" else
                "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (dense TF32 runs at half the bf16 rate; no TF32 figure is measured)",
                "algorithmic": "sum over the kernel's launches in one step of SURVEY.md 8(d) bytes, or 2*L*K*N flops"})
    out = {"roofline": top, "rooflines_by_kernel": per_kernel, "kernel_breakdown": kernels,
           "product_device_time_per_step_ms": round(device_s * 1e3, 3), "instrumented_pass_wall_ms": round(step_ms, 3)}
    fps = [(k, v) for k, v in agg.items() if k[0] in ("furthest_point_sampling", "fps_sample")]
    if fps:
        out["fps_ns_per_iteration"] = {"%d->%d" % (k[1][1], k[1][2]): round(v[0] / v[1] / max(k[1][2] - 1, 1) * 1e9, 1) for k, v in fps}
    if world == 1:
        out["op_rooflines"] = sweep_rooflines(device, peak)
        out["gemm_tensor_rate"] = gemm_tensor_rate(device)
    return out
