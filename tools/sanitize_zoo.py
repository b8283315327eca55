"""Every product kernel once at SMALL shapes, for compute-sanitizer (memcheck / racecheck / synccheck are 10-1000x slower than a
plain run).  Covers the shared-memory, DSMEM-cluster, mbarrier/bulk-copy/TMA and tcgen05 kernels the sanitizers are for.

    compute-sanitizer --tool memcheck  python tools/sanitize_zoo.py
    compute-sanitizer --tool racecheck python tools/sanitize_zoo.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epnet_b200 import image_prep, iou3d_utils, li_fusion, roipool3d_utils, scenes  # noqa: E402
from epnet_b200 import pointnet2_cuda as pc  # noqa: E402
from epnet_b200 import pointnet2_utils as pu  # noqa: E402
from epnet_b200.gemm import PackedConv3x3, PackedDeconv, PackedLinear, Planes, grouped_first_layer  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B = 2
# FPS: bucket kernel (4096), warp kernel (<= 512), cluster/DSMEM kernel (> 16384), streaming is > 131072 (skipped: slow under sanitizer)
for n, m in ((4096, 512), (512, 128), (300, 64), (20000, 256)):
    pts = torch.stack([scenes.lidar_scene(1000 + i, n) for i in range(B)]).to(dev)
    idx = pu.furthest_point_sample(pts, m)
pts = torch.stack([scenes.lidar_scene(1000 + i, 8192) for i in range(B)]).to(dev)
temp = torch.full((B, 8192), 1e10, device=dev)
idx = torch.empty(B, 1024, dtype=torch.int32, device=dev)
new_xyz = torch.empty(B, 1024, 3, device=dev)
xy = torch.rand(B, 8192, 2, device=dev)
new_xy = torch.empty(B, 1024, 2, device=dev)
pc.fps_sample_wrapper(B, 8192, 1024, pts, temp, idx, new_xyz, xy, new_xy)
ball = pu.ball_query(0.8, 16, pts, new_xyz)
buckets = pc.bucket_cloud(pts)
bidx = torch.zeros(B, 1024, 32, dtype=torch.int32, device=dev)
pc.ball_query_sorted_wrapper(B, 1024, 1.0, 32, new_xyz, buckets, bidx)
feats = torch.randn(B, 24, 8192, device=dev, requires_grad=True)
grouped = pu.grouping_operation(feats, ball)
gathered = pu.gather_operation(feats, idx)
dist, nn_idx = pu.three_nn(pts, new_xyz)
w = torch.softmax(-dist, dim=2).contiguous()
up = pu.three_interpolate(gathered, nn_idx, w)
(grouped.sum() + up.sum()).backward()  # the three gradient kernels
fmap = torch.randn(B, 16, 24, 80, device=dev, requires_grad=True)
li_fusion.feature_gather(fmap, torch.rand(B, 500, 2, device=dev) * 2 - 1, False).sum().backward()
# staged-row kernels want rows that fit shared memory and 16-byte alignment
big = torch.randn(1, 8, 16384, device=dev)
pu.grouping_operation(big, torch.randint(0, 16384, (1, 2048, 16), generator=g).int().to(dev))
# point-major fused kernels
feats_pm = torch.randn(B, 8192, 32, device=dev)
out = torch.empty(B * 1024 * 16, 40, device=dev)
pc.group_concat_pm_wrapper(B, 32, 8192, 1024, 16, pts, new_xyz, feats_pm, ball, out)
wts = torch.empty(B, 8192, 3, device=dev)
d2 = torch.empty(B, 8192, 3, device=dev)
i3 = torch.empty(B, 8192, 3, dtype=torch.int32, device=dev)
pc.three_nn_weights_wrapper(B, 8192, 1024, pts, new_xyz, d2, i3, wts)
known_pm = torch.randn(B, 1024, 64, device=dev)
o = torch.empty(B * 8192, 96, device=dev)
pc.three_interpolate_concat_pm_wrapper(B, 64, 1024, 8192, 32, known_pm, i3, wts, feats_pm, o)
img = torch.randn(B, 24, 80, 64, device=dev)
pc.grid_gather_nhwc_pm_wrapper(B, 64, 24, 80, 1024, img, torch.rand(B, 1024, 2, device=dev) * 2 - 1, False, torch.empty(B * 1024, 64, device=dev))
pc.attention_scale_pm_wrapper(torch.randn(2048, 24, device=dev), torch.randn(2048, 24, device=dev), torch.randn(24, device=dev),
                              torch.zeros(1, device=dev), torch.randn(2048, 96, device=dev), torch.empty(2048, 96, device=dev))
# tcgen05 GEMM family: narrow (A through TMEM), wide FP16 split, wide TF32 split, grouped operand, channel-major epilogue, pooled
x = torch.randn(3000, 96, device=dev)
PackedLinear(torch.randn(64, 96, device=dev) / 10, torch.zeros(64, device=dev))(x, relu=True)
PackedLinear(torch.randn(128, 96, device=dev) / 10, torch.zeros(128, device=dev))(x[:2048], relu=True, pool=16)
xw = torch.randn(2000, 512, device=dev)
from epnet_b200 import gemm  # noqa: E402
with gemm.tile_policy("throughput"):
    PackedLinear(torch.randn(256, 512, device=dev) / 22, None)(xw, relu=True)
    with gemm.f16_split(False):
        PackedLinear(torch.randn(256, 512, device=dev) / 22, None)(xw, relu=True)
lin0 = PackedLinear(torch.randn(32, 35, device=dev) / 6, torch.zeros(32, device=dev))
grouped_first_layer(lin0, pts, new_xyz, feats_pm, ball, relu=True)
PackedLinear(torch.randn(128, 96, device=dev) / 10, None)(x[:2048], relu=True, out_cm=torch.empty(2, 128, 1024, device=dev))
# convolutions: SIMT-producer kernels, planes out, TMA-fed kernel (stride 1 and 2), transposed convolution
xi = torch.zeros(B, 32, 64, 4, device=dev)
xi[..., :3] = torch.randn(B, 32, 64, 3, device=dev)
c0 = PackedConv3x3(torch.randn(64, 3, 3, 3, device=dev) / 5, torch.zeros(64, device=dev), stride=1)
_, p0 = c0(xi, relu=True, planes_out=True, f32_out=False)
c1 = PackedConv3x3(torch.randn(64, 64, 3, 3, device=dev) / 24, None, stride=2)
y1, p1 = c1(p0, relu=False, planes_out=True)
c2 = PackedConv3x3(torch.randn(128, 64, 3, 3, device=dev) / 24, torch.zeros(128, device=dev), stride=1)
c2(p1, relu=True)
c2(y1, relu=True)
PackedLinear(torch.randn(48, 64, device=dev) / 8, None).from_planes(Planes(p1.h1.view(-1, 64), p1.h2.view(-1, 64)), relu=True)
cat = torch.empty(B, 32, 64, 16, device=dev)
PackedDeconv(torch.randn(64, 16, 2, 2, device=dev) / 8, None)(y1, cat)
# image preparation
image_prep.normalise_pad(torch.randint(0, 256, (B, 30, 70, 3), dtype=torch.uint8, device=dev), out_hw=(32, 72))
image_prep.nchw_to_nhwc4(torch.randn(B, 3, 32, 64, device=dev))
# next rows: RoI pooling, rotated IoU / NMS
boxes = torch.zeros(B, 16, 7, device=dev)
boxes[..., 0] = torch.linspace(-20, 20, 16, device=dev); boxes[..., 1] = 1.8; boxes[..., 2] = torch.linspace(5, 60, 16, device=dev)
boxes[..., 3:6] = torch.tensor([1.6, 1.7, 4.0], device=dev); boxes[..., 6] = 0.3
roipool3d_utils.roipool3d_gpu(pts, torch.randn(B, 8192, 16, device=dev), boxes, 1.0, sampled_pt_num=64)
c = torch.rand(700, 2, device=dev) * 12
bev = torch.cat([c - 1.0, c + torch.tensor([2.9, 0.6], device=dev), torch.rand(700, 1, device=dev) * 6 - 3], dim=1)
iou3d_utils.boxes_iou_bev(bev[:100], bev)
iou3d_utils.nms_gpu(bev, torch.rand(700, device=dev), 0.5)
iou3d_utils.nms_normal_gpu(bev, torch.rand(700, device=dev), 0.5)
torch.cuda.synchronize()
print("sanitize zoo ok")
