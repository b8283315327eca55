"""`iou3d_cuda` -- the extension module /root/reference/lib/utils/iou3d/iou3d_utils.py:2 imports, on libepnet_b200.so.

Same four function names and positional signatures as the reference's pybind table (lib/utils/iou3d/src/iou3d.cpp:172-177),
so the reference's own iou3d_utils.py / proposal layers run unchanged on top of it (epnet_b200.install()).  The reference's
`nms_gpu(boxes, keep, thresh)` fills a HOST int64 tensor and returns the count, so this form has to read the result back;
the sync-free forms live in epnet_b200.iou3d_utils (nms_batched / nms_fixed)."""
import torch

from . import iou3d_utils as _u
from . import pointnet2_cuda as _pc
from ._lib import LIB


def _pairwise(name, fn, boxes_a, boxes_b, out):
    if out.shape != (boxes_a.shape[0], boxes_b.shape[0]):
        raise ValueError("output must be (%d, %d)" % (boxes_a.shape[0], boxes_b.shape[0]))
    if out.numel():
        _pc._call(name, fn, boxes_a, boxes_a.shape[0], _pc._f(boxes_a, "boxes_a"), boxes_b.shape[0], _pc._f(boxes_b, "boxes_b"),
                  _pc._f(out, "out"))
    return 1


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    """iou3d.cpp:34-52"""
    return _pairwise("boxes_overlap_bev", LIB.epnet_boxes_overlap_bev, boxes_a, boxes_b, ans_overlap)


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    """iou3d.cpp:54-72"""
    return _pairwise("boxes_iou_bev", LIB.epnet_boxes_iou_bev, boxes_a, boxes_b, ans_iou)


def _nms(boxes, keep, thresh, rotated):
    if keep.is_cuda or keep.dtype != torch.int64 or keep.numel() < boxes.shape[0]:
        raise ValueError("keep must be a host int64 tensor with one slot per box (iou3d.cpp:79-86)")
    if boxes.shape[0] == 0:
        return 0
    kept, num_out = _u.nms_batched(boxes.unsqueeze(0), thresh, rotated=rotated)
    n = int(num_out.item())
    keep[:n] = kept[0, :n].cpu()
    return n


def nms_gpu(boxes, keep, nms_overlap_thresh):
    """iou3d.cpp:74-121: boxes (N,5) sorted by descending score, keep (N) host int64 -> number kept"""
    return _nms(boxes, keep, nms_overlap_thresh, True)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    """iou3d.cpp:124-170"""
    return _nms(boxes, keep, nms_overlap_thresh, False)
