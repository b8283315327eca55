"""Rotated BEV IoU and NMS -- mirror of /root/reference/lib/utils/iou3d/iou3d_utils.py on libepnet_b200.so
(SURVEY.md section 8(f), rank 2).  Same function names, arguments and returns; differences, all on the host side:

* nothing is allocated or freed by the native call, and the greedy scan runs on the device: the reference copies an
  N x ceil(N/64) mask to the host and loops there (iou3d.cpp:95-113).  `nms_gpu` / `nms_normal_gpu` still read ONE int
  (the number of kept boxes) because their return value has a data-dependent length;
* `nms_fixed` / `nms_batched` are the sync-free forms the proposal layer wants (lib/rpn/proposal_layer.py:101-115): a fixed
  `max_out`, the keep list padded, the count left on the device, several problems (scenes x distance bands) in one launch.
"""
import ctypes

import torch

from . import pointnet2_cuda as pc
from ._lib import LIB, check


def boxes3d_to_bev_torch(boxes3d):
    """lib/utils/kitti_utils.py:137-150: (N,7) [x, y, z, h, w, l, ry] -> (N,5) [x1, y1, x2, y2, ry] in the x-z plane."""
    boxes_bev = boxes3d.new_empty((boxes3d.shape[0], 5))
    cu, cv = boxes3d[:, 0], boxes3d[:, 2]
    half_l, half_w = boxes3d[:, 5] / 2, boxes3d[:, 4] / 2
    boxes_bev[:, 0], boxes_bev[:, 1] = cu - half_l, cv - half_w
    boxes_bev[:, 2], boxes_bev[:, 3] = cu + half_l, cv + half_w
    boxes_bev[:, 4] = boxes3d[:, 6]
    return boxes_bev


def _boxes5(t, name):
    if t.dim() < 2 or t.shape[-1] != 5:
        raise ValueError("%s must be (..., 5) [x1, y1, x2, y2, ry], got %s" % (name, tuple(t.shape)))
    return t.contiguous()


def _pairwise(name, fn, boxes_a, boxes_b):
    boxes_a, boxes_b = _boxes5(boxes_a, "boxes_a"), _boxes5(boxes_b, "boxes_b")
    out = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    if out.numel():
        pc._call(name, fn, boxes_a, boxes_a.shape[0], pc._f(boxes_a, "boxes_a"), boxes_b.shape[0], pc._f(boxes_b, "boxes_b"), out.data_ptr())
    return out


def boxes_overlap_bev(boxes_a, boxes_b):
    """(M,5), (N,5) -> (M,N) intersection areas of the rotated rectangles (iou3d_cuda.boxes_overlap_bev_gpu, iou3d.cpp:34-52)."""
    return _pairwise("boxes_overlap_bev", LIB.epnet_boxes_overlap_bev, boxes_a, boxes_b)


def boxes_iou_bev(boxes_a, boxes_b):
    """iou3d_utils.py:6-18: (M,5), (N,5) -> (M,N) rotated BEV IoU."""
    return _pairwise("boxes_iou_bev", LIB.epnet_boxes_iou_bev, boxes_a, boxes_b)


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """iou3d_utils.py:21-53: (N,7), (M,7) [x, y, z, h, w, l, ry] -> (N,M) 3D IoU = BEV overlap x height overlap / union volume."""
    overlaps_bev = boxes_overlap_bev(boxes3d_to_bev_torch(boxes_a), boxes3d_to_bev_torch(boxes_b))
    a_min, a_max = (boxes_a[:, 1] - boxes_a[:, 3]).view(-1, 1), boxes_a[:, 1].view(-1, 1)
    b_min, b_max = (boxes_b[:, 1] - boxes_b[:, 3]).view(1, -1), boxes_b[:, 1].view(1, -1)
    overlaps_h = torch.clamp(torch.min(a_max, b_max) - torch.max(a_min, b_min), min=0)
    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    return overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-7)


def nms_workspace_bytes(s, n):
    out = ctypes.c_ulonglong(0)
    check(LIB.epnet_nms_workspace_bytes(int(s), int(n), ctypes.byref(out)), "nms_workspace_bytes")
    return int(out.value)


def nms_batched(boxes, thresh, max_out=0, counts=None, rotated=True, workspace=None):
    """boxes (S,N,5), every problem sorted by descending score -> keep (S,N) int64 (first num_out[s] entries valid, the rest
    -1), num_out (S) int32, both on the device; no host synchronisation.  counts (S) int32 = boxes present per problem."""
    boxes = _boxes5(boxes, "boxes")
    if boxes.dim() != 3:
        raise ValueError("boxes must be (S, N, 5)")
    s, n = boxes.shape[0], boxes.shape[1]
    keep = torch.full((s, n), -1, dtype=torch.int64, device=boxes.device)
    num_out = torch.zeros((s,), dtype=torch.int32, device=boxes.device)
    if s == 0 or n == 0:
        return keep, num_out
    need = nms_workspace_bytes(s, n)
    if workspace is None:
        workspace = torch.empty((need + 7) // 8, dtype=torch.int64, device=boxes.device)
    elif workspace.device != boxes.device or workspace.numel() * workspace.element_size() < need or workspace.data_ptr() % 8:
        raise ValueError("workspace must be an 8-byte aligned buffer of >= %d bytes on %s" % (need, boxes.device))
    cptr = 0
    if counts is not None:
        if counts.shape != (s,):
            raise ValueError("counts must be (S,)")
        cptr = pc._i(counts, "counts")
    fn = LIB.epnet_nms_rotated if rotated else LIB.epnet_nms_normal
    pc._call("nms_rotated" if rotated else "nms_normal", fn, boxes, s, n, pc._f(boxes, "boxes"), cptr, float(thresh), int(max_out),
             workspace.data_ptr(), keep.data_ptr(), num_out.data_ptr())
    pc.LAUNCHES[0] += 1  # mask kernel + scan kernel
    return keep, num_out


def nms_fixed(boxes, scores, thresh, max_out, rotated=True):
    """One problem, sync-free: -> (indices into `boxes` of the kept boxes, padded with -1 to max_out; count on the device)."""
    order = scores.sort(0, descending=True)[1]
    keep, num_out = nms_batched(boxes[order].unsqueeze(0), thresh, max_out=max_out, rotated=rotated)
    keep = keep[0, :max_out]
    idx = torch.where(keep >= 0, order[keep.clamp(min=0)], keep) if keep.numel() else keep
    if idx.numel() < max_out:    # fewer boxes than max_out: pad to the promised length
        idx = torch.cat([idx, idx.new_full((max_out - idx.numel(),), -1)])
    return idx, num_out[0]


def _nms(boxes, scores, thresh, rotated):
    order = scores.sort(0, descending=True)[1]
    keep, num_out = nms_batched(boxes[order].unsqueeze(0), thresh, rotated=rotated)
    return order[keep[0, :int(num_out.item())]].contiguous()


def nms_gpu(boxes, scores, thresh):
    """iou3d_utils.py:56-70: boxes (N,5) [x1, y1, x2, y2, ry], scores (N) -> indices of the kept boxes, best score first."""
    return _nms(boxes, scores, thresh, True)


def nms_normal_gpu(boxes, scores, thresh):
    """iou3d_utils.py:73-87: the same with the axis-aligned IoU of the extents."""
    return _nms(boxes, scores, thresh, False)
