"""One rotated and one axis-aligned NMS at the proposal layer's test-time size (6300 boxes, threshold 0.7), for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:nms_ -o gpurun_out/nms python tests/perf/nms_once.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_iou3d import proposals  # noqa: E402
from epnet_b200 import iou3d_utils  # noqa: E402

b = torch.from_numpy(proposals(7, 6300, objects=60)).cuda().unsqueeze(0)
for rotated in (True, False):
    keep, num = iou3d_utils.nms_batched(b, 0.7, rotated=rotated)
    torch.cuda.synchronize()
    print("rotated" if rotated else "normal", "kept", int(num[0]))
