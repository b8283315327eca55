"""TEST INFRASTRUCTURE: run the product's Python layers (epnet_b200.pointnet2_utils / pointnet2_modules / pointnet2_msg /
proposal_select) on ANOTHER backend -- the C oracle on CPU tensors -- to check the host logic without a GPU.  The product has no
backend parameter (no multi-backend dispatch); the swap is a monkeypatch of the three module attributes its layers look up at call
time, undone on exit."""
import contextlib

import torch


@contextlib.contextmanager
def swapped(pointnet2_backend=None, feature_gather=None, nms_batched=None):
    from epnet_b200 import iou3d_utils, li_fusion, pointnet2_utils
    saved = (pointnet2_utils._BACKEND, li_fusion.feature_gather, iou3d_utils.nms_batched)
    if pointnet2_backend is not None:
        pointnet2_utils._BACKEND = pointnet2_backend
    if feature_gather is not None:
        li_fusion.feature_gather = feature_gather
    if nms_batched is not None:
        iou3d_utils.nms_batched = nms_batched
    try:
        yield
    finally:
        pointnet2_utils._BACKEND, li_fusion.feature_gather, iou3d_utils.nms_batched = saved


def cpu_oracle():
    """the product's layers on the C oracle + torch's CPU grid_sample"""
    from oracle import cpu_backend

    def gather(fm, xy, align_corners=False):
        return torch.nn.functional.grid_sample(fm, xy.unsqueeze(1), align_corners=bool(align_corners)).squeeze(2)

    return swapped(cpu_backend, gather)
