"""3D RoI point pooling -- mirror of the GPU path of /root/reference/lib/utils/roipool3d/roipool3d_utils.py:7-28 on
libepnet_b200.so."""
import torch

from . import pointnet2_cuda as pc
from ._lib import LIB


def enlarge_box3d(boxes3d, extra_width):
    """lib/utils/kitti_utils.py:153-163: boxes3d (N,7) [x, y, z, h, w, l, ry]; sizes grow by 2*extra, the bottom-centre y by extra."""
    large = boxes3d.clone()
    large[:, 3:6] += extra_width * 2
    large[:, 1] += extra_width
    return large


def roipool3d_gpu(pts, pts_feature, boxes3d, pool_extra_width, sampled_pt_num=512):
    """pts (B,N,3), pts_feature (B,N,C), boxes3d (B,M,7) -> pooled_features (B,M,sampled,3+C), pooled_empty_flag (B,M) int32."""
    batch_size, boxes_num, feature_len = pts.shape[0], boxes3d.shape[1], pts_feature.shape[2]
    pooled_boxes3d = enlarge_box3d(boxes3d.reshape(-1, 7), pool_extra_width).view(batch_size, -1, 7)
    pooled_features = torch.zeros((batch_size, boxes_num, sampled_pt_num, 3 + feature_len), dtype=torch.float32, device=pts.device)
    pooled_empty_flag = torch.zeros((batch_size, boxes_num), dtype=torch.int32, device=pts.device)
    pts, pooled_boxes3d, pts_feature = pts.contiguous(), pooled_boxes3d.contiguous(), pts_feature.contiguous()
    pc._call("roipool3d", LIB.epnet_roipool3d, pts, batch_size, pts.shape[1], boxes_num, feature_len, sampled_pt_num,
             pc._f(pts, "pts"), pc._f(pooled_boxes3d, "boxes3d"), pc._f(pts_feature, "pts_feature"), pooled_features.data_ptr(),
             pooled_empty_flag.data_ptr())
    return pooled_features, pooled_empty_flag
