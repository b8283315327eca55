"""Times rotated / axis-aligned BEV NMS at the proposal layer's sizes (lib/config.py:188-190,203-205: 9000 -> 300 at test time,
12000 -> 2048 in training, split 70/30 over two distance bands, lib/rpn/proposal_layer.py:66-71) against the reference's mask
kernel on the same GPU.  The reference then copies the mask to the host and runs the greedy loop there (iou3d.cpp:95-113); the
copy is timed, the C++ loop is not (it is replayed in Python by the tests), so the reference column is a LOWER bound."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_iou3d import proposals  # noqa: E402
from epnet_b200 import iou3d_utils  # noqa: E402
from oracle import ref_cuda  # noqa: E402


def dev_us(fn, it=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / it * 1e3


def wall_us(fn, it=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(it):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / it * 1e6


lib = ref_cuda._load()
for label, n, post, thresh in (("test near", 6300, 210, 0.7), ("test far", 2700, 90, 0.7), ("train near", 8400, 1433, 0.85), ("train far", 3600, 615, 0.85)):
    b = torch.from_numpy(proposals(7, n, objects=60)).cuda()
    cb = (n + 63) // 64
    mask = torch.zeros((n, cb), dtype=torch.int64, device="cuda")
    host = torch.empty((n, cb), dtype=torch.int64).pin_memory()
    ws = torch.empty(iou3d_utils.nms_workspace_bytes(1, n) // 8, dtype=torch.int64, device="cuda")
    for rotated in (True, False):
        fn = lib.ref_nms_mask if rotated else lib.ref_nms_normal_mask
        ref_kernel = dev_us(lambda: fn(b.data_ptr(), mask.data_ptr(), n, thresh))
        ref_copy = wall_us(lambda: (fn(b.data_ptr(), mask.data_ptr(), n, thresh), host.copy_(mask)))
        full = dev_us(lambda: iou3d_utils.nms_batched(b.unsqueeze(0), thresh, rotated=rotated, workspace=ws))
        top = dev_us(lambda: iou3d_utils.nms_batched(b.unsqueeze(0), thresh, max_out=post, rotated=rotated, workspace=ws))
        keep, num = iou3d_utils.nms_batched(b.unsqueeze(0), thresh, rotated=rotated, workspace=ws)
        print(f"{label:10s} n={n:5d} {'rotated' if rotated else 'normal ':7s} thr={thresh}: kept {int(num[0]):5d} | ref mask kernel {ref_kernel:8.1f} us, "
              f"+D2H {ref_copy:8.1f} us (host loop not included) | ours mask+scan {full:8.1f} us, stop at {post}: {top:8.1f} us | "
              f"x{ref_copy / full:.1f} / x{ref_copy / top:.1f}")

# two scenes x two bands in one launch (what a batched proposal layer issues)
boxes = torch.from_numpy(np.stack([proposals(20 + s, 6300, objects=60) for s in range(4)])).cuda()
counts = torch.tensor([6300, 2700, 6300, 2700], dtype=torch.int32, device="cuda")
ws = torch.empty(iou3d_utils.nms_workspace_bytes(4, 6300) // 8, dtype=torch.int64, device="cuda")
t4 = dev_us(lambda: iou3d_utils.nms_batched(boxes, 0.7, max_out=210, counts=counts, workspace=ws))
print(f"batched 2 scenes x 2 bands (6300/2700), rotated, stop at 210: {t4:8.1f} us for all four problems, no host sync")

# pairwise IoU: RoIs against ground truth (lib/rpn/proposal_target_layer.py) and a dense 4096 x 4096 matrix
for m, k in ((512, 64), (4096, 4096)):
    a, c = torch.from_numpy(proposals(3, m, objects=20)).cuda(), torch.from_numpy(proposals(3, k, objects=20, jitter_seed=9)).cuda()
    out = torch.zeros((m, k), device="cuda")
    ref = dev_us(lambda: lib.ref_boxes_iou_bev(m, a.data_ptr(), k, c.data_ptr(), out.data_ptr()))
    ours = dev_us(lambda: iou3d_utils.boxes_iou_bev(a, c))
    print(f"boxes_iou_bev {m}x{k}: ours {ours:8.1f} us  ref {ref:8.1f} us  x{ref / ours:.1f}")
