"""oracle/ref_cuda.py -- TEST INFRASTRUCTURE.  The reference's OWN, unmodified CUDA kernels
(oracle/_ref/libpointnet2_ref.so, built by oracle/build_ref.sh) behind the nine `*_wrapper` names of the
reference's pybind table, so tests can pin the oracle and the product kernels against the real thing on
a B200, and bench.py --impl reference can time it.  Never imported by epnet_b200/.
"""
import ctypes
import os

import torch

# no relative import: baseline/ref_env.py loads this file by path (as `pointnet2_cuda`) without importing the oracle package,
# so that the reference arm's process maps the reference's kernels and nothing else of this repo
REF_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libpointnet2_ref.so")

_c_int, _c_float, _c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_void_p


def available():
    return os.path.exists(REF_LIB_PATH)


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libpointnet2_ref.so not built; run oracle/build_ref.sh where /root/reference exists")
        lib = ctypes.CDLL(REF_LIB_PATH)
        sigs = {
            "ref_furthest_point_sampling": [_c_int] * 3 + [_c_void_p] * 4,
            "ref_gather_points": [_c_int] * 4 + [_c_void_p] * 4,
            "ref_gather_points_grad": [_c_int] * 4 + [_c_void_p] * 4,
            "ref_ball_query": [_c_int] * 3 + [_c_float, _c_int] + [_c_void_p] * 4,
            "ref_group_points": [_c_int] * 5 + [_c_void_p] * 4,
            "ref_group_points_grad": [_c_int] * 5 + [_c_void_p] * 4,
            "ref_three_nn": [_c_int] * 3 + [_c_void_p] * 5,
            "ref_three_interpolate": [_c_int] * 4 + [_c_void_p] * 5,
            "ref_three_interpolate_grad": [_c_int] * 4 + [_c_void_p] * 5,
            "ref_roipool3d": [_c_int] * 5 + [_c_void_p] * 5,
            "ref_boxes_overlap_bev": [_c_int, _c_void_p, _c_int, _c_void_p, _c_void_p],
            "ref_boxes_iou_bev": [_c_int, _c_void_p, _c_int, _c_void_p, _c_void_p],
            "ref_nms_mask": [_c_void_p, _c_void_p, _c_int, _c_float],
            "ref_nms_normal_mask": [_c_void_p, _c_void_p, _c_int, _c_float],
        }
        for name, argtypes in sigs.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = None
        for name in ("ref_nms_gpu", "ref_nms_normal_gpu"):
            fn = getattr(lib, name)
            fn.argtypes = [_c_void_p, _c_void_p, _c_int, _c_float]
            fn.restype = _c_int
        _lib = lib
    return _lib


def _p(t):
    assert t.is_cuda and t.is_contiguous()
    return t.data_ptr()


def _s(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def furthest_point_sampling_wrapper(b, n, m, points_tensor, temp_tensor, idx_tensor):
    _load().ref_furthest_point_sampling(b, n, m, _p(points_tensor), _p(temp_tensor), _p(idx_tensor), _s(points_tensor))
    return 1


def gather_points_wrapper(b, c, n, npoints, points_tensor, idx_tensor, out_tensor):
    _load().ref_gather_points(b, c, n, npoints, _p(points_tensor), _p(idx_tensor), _p(out_tensor), _s(points_tensor))
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out_tensor, idx_tensor, grad_points_tensor):
    _load().ref_gather_points_grad(b, c, n, npoints, _p(grad_out_tensor), _p(idx_tensor), _p(grad_points_tensor),
                                   _s(grad_out_tensor))
    return 1


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz_tensor, xyz_tensor, idx_tensor):
    _load().ref_ball_query(b, n, m, float(radius), nsample, _p(new_xyz_tensor), _p(xyz_tensor), _p(idx_tensor),
                           _s(xyz_tensor))
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points_tensor, idx_tensor, out_tensor):
    _load().ref_group_points(b, c, n, npoints, nsample, _p(points_tensor), _p(idx_tensor), _p(out_tensor),
                             _s(points_tensor))
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out_tensor, idx_tensor, grad_points_tensor):
    _load().ref_group_points_grad(b, c, n, npoints, nsample, _p(grad_out_tensor), _p(idx_tensor),
                                  _p(grad_points_tensor), _s(grad_out_tensor))
    return 1


def three_nn_wrapper(b, n, m, unknown_tensor, known_tensor, dist2_tensor, idx_tensor):
    _load().ref_three_nn(b, n, m, _p(unknown_tensor), _p(known_tensor), _p(dist2_tensor), _p(idx_tensor),
                         _s(unknown_tensor))


def three_interpolate_wrapper(b, c, m, n, points_tensor, idx_tensor, weight_tensor, out_tensor):
    _load().ref_three_interpolate(b, c, m, n, _p(points_tensor), _p(idx_tensor), _p(weight_tensor), _p(out_tensor),
                                  _s(points_tensor))


def three_interpolate_grad_wrapper(b, c, n, m, grad_out_tensor, idx_tensor, weight_tensor, grad_points_tensor):
    _load().ref_three_interpolate_grad(b, c, n, m, _p(grad_out_tensor), _p(idx_tensor), _p(weight_tensor),
                                       _p(grad_points_tensor), _s(grad_out_tensor))


def roipool3d(pts, boxes3d, pts_feature, sampled=512):
    """The reference's roipool3dLauncher on already-enlarged boxes (it allocates and frees its own scratch, default stream)."""
    B, N, _ = pts.shape
    M, C = boxes3d.shape[1], pts_feature.shape[2]
    out = torch.zeros((B, M, sampled, 3 + C), dtype=torch.float32, device=pts.device)
    flag = torch.zeros((B, M), dtype=torch.int32, device=pts.device)
    torch.cuda.synchronize()
    _load().ref_roipool3d(B, N, M, C, sampled, _p(pts), _p(boxes3d), _p(pts_feature), _p(out), _p(flag))
    torch.cuda.synchronize()
    return out, flag


def boxes_pairwise_bev(boxes_a, boxes_b, iou):
    """The reference's boxesoverlapLauncher (iou=False) / boxesioubevLauncher (iou=True) on (N,5) [x1,y1,x2,y2,ry] boxes."""
    out = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    torch.cuda.synchronize()
    fn = _load().ref_boxes_iou_bev if iou else _load().ref_boxes_overlap_bev
    fn(boxes_a.shape[0], _p(boxes_a), boxes_b.shape[0], _p(boxes_b), _p(out))
    torch.cuda.synchronize()
    return out


def nms(boxes, thresh, rotated=True):
    """The reference's nms_gpu / nms_normal_gpu on score-sorted boxes: its mask kernel, then the host loop of
    iou3d.cpp:100-113 replayed on the copied mask.  Returns the kept indices (int64, host)."""
    import numpy as np
    n = boxes.shape[0]
    col_blocks = (n + 63) // 64
    mask = torch.zeros((n, col_blocks), dtype=torch.int64, device=boxes.device)
    torch.cuda.synchronize()
    (_load().ref_nms_mask if rotated else _load().ref_nms_normal_mask)(_p(boxes), _p(mask), n, float(thresh))
    torch.cuda.synchronize()
    m = mask.cpu().numpy().view(np.uint64)
    remv = np.zeros(col_blocks, dtype=np.uint64)
    keep = []
    for i in range(n):
        nblock, inblock = divmod(i, 64)
        if not (int(remv[nblock]) >> inblock) & 1:
            keep.append(i)
            remv[nblock:] |= m[i, nblock:]
    return torch.tensor(keep, dtype=torch.int64)


# ---- the pybind tables of the two next-row extensions, on the reference's own kernels (used by baseline/ref_env.py) ----

class iou3d_cuda:  # lib/utils/iou3d/src/iou3d.cpp:172-177
    @staticmethod
    def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
        torch.cuda.current_stream().synchronize()  # the reference launches on the legacy default stream
        _load().ref_boxes_overlap_bev(boxes_a.shape[0], _p(boxes_a), boxes_b.shape[0], _p(boxes_b), _p(ans_overlap))
        return 1

    @staticmethod
    def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
        torch.cuda.current_stream().synchronize()
        _load().ref_boxes_iou_bev(boxes_a.shape[0], _p(boxes_a), boxes_b.shape[0], _p(boxes_b), _p(ans_iou))
        return 1

    @staticmethod
    def nms_gpu(boxes, keep, thresh):
        assert not keep.is_cuda and keep.dtype == torch.int64 and keep.is_contiguous()
        torch.cuda.current_stream().synchronize()
        return _load().ref_nms_gpu(_p(boxes), keep.data_ptr(), boxes.shape[0], float(thresh))

    @staticmethod
    def nms_normal_gpu(boxes, keep, thresh):
        assert not keep.is_cuda and keep.dtype == torch.int64 and keep.is_contiguous()
        torch.cuda.current_stream().synchronize()
        return _load().ref_nms_normal_gpu(_p(boxes), keep.data_ptr(), boxes.shape[0], float(thresh))


class roipool3d_cuda:  # lib/utils/roipool3d/src/roipool3d.cpp:40-70
    @staticmethod
    def forward(xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag):
        b, n, m = xyz.shape[0], xyz.shape[1], boxes3d.shape[1]
        torch.cuda.current_stream().synchronize()
        _load().ref_roipool3d(b, n, m, pts_feature.shape[2], pooled_features.shape[2], _p(xyz), _p(boxes3d), _p(pts_feature),
                              _p(pooled_features), _p(pooled_empty_flag))
        return 1
