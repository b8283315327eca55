"""Generates the data behind tests/golden/iou3d_ref_b200.npz: pairwise rotated BEV overlaps and IoUs of one seeded case computed
by the REFERENCE'S OWN, unmodified kernels (oracle/_ref, built from /root/reference/lib/utils/iou3d/src/iou3d_kernel.cu) on a
B200.  Run on the GPU box:

    gpurun -- python tests/golden/make_iou3d_golden.py      # writes gpurun_out/iou3d_dump.npz (also holds the product's outputs)

then keep a, b, ov_ref, iou_ref:  np.savez_compressed("tests/golden/iou3d_ref_b200.npz", ...).  The fixture lets the CPU suite
check the C oracle against real reference outputs, and the GPU suite check the product bit for bit without oracle/_ref.
"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from test_iou3d import proposals, special_boxes
from epnet_b200 import iou3d_utils
from oracle import ref_cuda

a = np.concatenate([proposals(1, 257, objects=4), special_boxes()])
b = np.concatenate([special_boxes(), proposals(1, 130, objects=4, jitter_seed=51)])
ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
out = dict(a=a, b=b)
out["ov_ours"] = iou3d_utils.boxes_overlap_bev(ta, tb).cpu().numpy()
out["ov_ref"] = ref_cuda.boxes_pairwise_bev(ta, tb, False).cpu().numpy()
out["iou_ours"] = iou3d_utils.boxes_iou_bev(ta, tb).cpu().numpy()
out["iou_ref"] = ref_cuda.boxes_pairwise_bev(ta, tb, True).cpu().numpy()
ang = torch.linspace(-8, 8, 4001, device="cuda")
out["ang"], out["cos"], out["sin"] = ang.cpu().numpy(), torch.cos(ang).cpu().numpy(), torch.sin(ang).cpu().numpy()
np.savez(os.path.join("gpurun_out", "iou3d_dump.npz"), **out)
print("mismatch overlap", (out["ov_ours"].view(np.uint32) != out["ov_ref"].view(np.uint32)).sum(), "iou", (out["iou_ours"].view(np.uint32) != out["iou_ref"].view(np.uint32)).sum())
