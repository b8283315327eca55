"""the same ops at the RCNN stage's many-small-problems shapes (lib/net/rcnn_net.py:28-41 with
cfg.RCNN.SA_CONFIG: 200 RoIs x 512 points, FPS 512 -> 128 -> 32, ball queries r = 0.2 / 0.4 with nsample 64, then GroupAll).
Index outputs bit-exact against the oracle and against the reference's own kernels."""
import numpy as np
import pytest
import torch

import oracle
from gpu_util import OpRunner

pytestmark = pytest.mark.gpu
ROIS = 200


def roi_clouds(seed, n):
    """RoI-local coordinates as roipool3d hands them over: points of a ~4 x 2 x 2 m box, canonical frame, many duplicates
    (roipool3d repeats points when a box holds fewer than 512)."""
    rng = np.random.RandomState(seed)
    pts = (rng.rand(ROIS, n, 3).astype(np.float32) - 0.5) * np.array([2.0, 2.0, 4.4], np.float32)
    for r in range(ROIS):  # a third of the RoIs hold few distinct points, repeated cyclically (roipool3d_kernel.cu:152-158)
        if r % 3 == 0:
            k = 1 + (r * 7) % 60
            pts[r] = pts[r, np.arange(n) % k]
    return pts


@pytest.fixture(scope="module")
def ours():
    from epnet_b200 import pointnet2_cuda
    return OpRunner(pointnet2_cuda)


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_cuda
    return OpRunner(ref_cuda) if ref_cuda.available() else None


@pytest.mark.parametrize("n,m,radius", [(512, 128, 0.2), (128, 32, 0.4)])
def test_rcnn_sa_level(ours, ref, n, m, radius):
    xyz = roi_clouds(n, n)
    idx = ours.fps(xyz, m)
    np.testing.assert_array_equal(idx, oracle.furthest_point_sampling(xyz, m))
    new_xyz = np.stack([xyz[b][idx[b]] for b in range(ROIS)])
    ball = ours.ball_query(radius, 64, xyz, new_xyz)
    np.testing.assert_array_equal(ball, oracle.ball_query(radius, 64, xyz, new_xyz))
    feats = np.random.RandomState(1).randn(ROIS, 128, n).astype(np.float32)
    grouped = ours.group(feats, ball)
    np.testing.assert_array_equal(grouped, oracle.group_points(feats, ball))
    if ref is not None:
        np.testing.assert_array_equal(idx, ref.fps(xyz, m))
        np.testing.assert_array_equal(ball, ref.ball_query(radius, 64, xyz, new_xyz))
        np.testing.assert_array_equal(grouped, ref.group(feats, ball))


def test_rcnn_group_all_module(ours):
    """GroupAll (pointnet2_utils.py:267-287): the last RCNN level pools all 32 points of every RoI."""
    from epnet_b200 import pointnet2_utils as pu
    xyz = torch.from_numpy(roi_clouds(3, 32)).cuda()
    feats = torch.randn(ROIS, 256, 32, device="cuda")
    out = pu.GroupAll(use_xyz=True)(xyz, None, feats)
    want = torch.cat([xyz.transpose(1, 2).unsqueeze(2), feats.unsqueeze(2)], dim=1)
    assert out.shape == (ROIS, 259, 1, 32) and torch.equal(out, want)
