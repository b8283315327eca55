"""Diagnostic: dump the product's and the reference kernels' pairwise BEV overlaps for one seeded case (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
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
