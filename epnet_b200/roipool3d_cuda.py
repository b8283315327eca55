"""`roipool3d_cuda` -- the extension module /root/reference/lib/utils/roipool3d/roipool3d_utils.py:2 imports, on
libepnet_b200.so: `forward` with the positional signature of roipool3d.cpp:40-70 (GPU path only; the reference's CPU helpers
`pts_in_boxes3d_cpu` / `roipool3d_cpu` serve its dataset code, which is out of scope)."""
from . import pointnet2_cuda as _pc
from ._lib import LIB


def forward(xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag):
    """xyz (B,N,3), boxes3d (B,M,7) already enlarged, pts_feature (B,N,C) -> pooled_features (B,M,S,3+C), pooled_empty_flag (B,M),
    both arriving zeroed (roipool3d_utils.py:20-22)"""
    b, n = xyz.shape[0], xyz.shape[1]
    m, c, s = boxes3d.shape[1], pts_feature.shape[2], pooled_features.shape[2]
    if pooled_features.shape != (b, m, s, 3 + c) or pooled_empty_flag.shape != (b, m):
        raise ValueError("roipool3d_cuda.forward: inconsistent shapes")
    _pc._call("roipool3d", LIB.epnet_roipool3d, xyz, b, n, m, c, s, _pc._f(xyz, "xyz"), _pc._f(boxes3d, "boxes3d"),
              _pc._f(pts_feature, "pts_feature"), _pc._f(pooled_features, "pooled_features"), _pc._i(pooled_empty_flag, "pooled_empty_flag"))
    return 1


def pts_in_boxes3d_cpu(*_a, **_k):
    raise NotImplementedError("epnet_b200 has no CPU path (roipool3d.cpp:113-140 serves the reference's dataset code only)")


roipool3d_cpu = pts_in_boxes3d_cpu
