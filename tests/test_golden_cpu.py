"""CPU checks against the golden fixtures produced by the reference's own Python layers (tests/golden/make_golden.py):
  * the oracle's ops reproduce the recorded op outputs;
  * this repo's mirror modules (pointnet2_utils / pointnet2_modules / pointnet2_msg), run on the CPU oracle backend,
    reproduce the reference modules' outputs -- i.e. the mirror composes the ops exactly as the reference does, with
    identical state-dict keys.  No CUDA involved; the GPU counterpart is tests/test_golden_gpu.py."""
import numpy as np
import torch

import oracle
from backend_swap import cpu_oracle
from golden_util import load, small_backbone_config, state_dict


def test_oracle_reproduces_recorded_op_outputs():
    g = load("ops_lidar2048.npz")
    pts = g["points"]
    np.testing.assert_array_equal(oracle.furthest_point_sampling(pts, 512), g["fps_idx"])
    np.testing.assert_array_equal(oracle.ball_query(0.8, 16, pts, g["new_xyz"]), g["ball_idx"])
    np.testing.assert_array_equal(oracle.group_points(g["feats"], g["ball_idx"]), g["grouped"])
    d2, i3 = oracle.three_nn(pts, g["new_xyz"])
    np.testing.assert_array_equal(i3, g["nn_idx"])
    np.testing.assert_allclose(np.sqrt(d2), g["nn_dist"], rtol=2e-7, atol=0)  # torch.sqrt (pointnet2_utils.py:98) vs libm: <= 1 ulp
    pooled = g["grouped"].max(axis=3)
    np.testing.assert_array_equal(oracle.three_interpolate(pooled, g["nn_idx"], g["weight"]), g["interpolated"])


def test_mirror_op_surface_matches_reference_python():
    g = load("ops_lidar2048.npz")
    from epnet_b200 import pointnet2_utils as ops
    from epnet_b200.pointnet2_utils import QueryAndGroup
    pts, feats = torch.from_numpy(g["points"]), torch.from_numpy(g["feats"])
    with cpu_oracle():
        _op_surface_checks(g, ops, QueryAndGroup, pts, feats)


def _op_surface_checks(g, ops, QueryAndGroup, pts, feats):
    idx = ops.furthest_point_sample(pts, 512)
    assert idx.dtype == torch.int32 and torch.equal(idx, torch.from_numpy(g["fps_idx"]))
    new_xyz = ops.gather_operation(pts.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    assert torch.equal(new_xyz, torch.from_numpy(g["new_xyz"]))
    qg = QueryAndGroup(0.8, 16, use_xyz=True)(pts, new_xyz, feats)
    assert torch.equal(qg, torch.from_numpy(g["query_and_group"]))
    dist, nn_idx = ops.three_nn(pts, new_xyz)
    assert torch.equal(dist, torch.from_numpy(g["nn_dist"])) and torch.equal(nn_idx, torch.from_numpy(g["nn_idx"]))


def test_mirror_sa_and_fp_modules_match_reference_python():
    g = load("modules_sa_fp.npz")
    from epnet_b200.pointnet2_modules import PointnetFPModule, PointnetSAModuleMSG
    sa = PointnetSAModuleMSG(npoint=128, radii=[0.8, 1.6], nsamples=[16, 32], mlps=[[6, 8, 16], [6, 8, 24]], use_xyz=True, bn=True)
    fp = PointnetFPModule(mlp=[40 + 6, 32, 16])
    sa.load_state_dict(state_dict(g, "sa."), strict=True)  # identical key names = checkpoint compatibility
    fp.load_state_dict(state_dict(g, "fp."), strict=True)
    sa.eval(); fp.eval()
    pts, feats = torch.from_numpy(g["points"]), torch.from_numpy(g["feats"])
    with torch.no_grad(), cpu_oracle():
        new_xyz, new_feat, idx = sa(pts, feats)
        out = fp(pts, new_xyz, feats, new_feat)
    assert torch.equal(idx, torch.from_numpy(g["sa_idx"]))
    assert torch.equal(new_xyz, torch.from_numpy(g["sa_new_xyz"]))
    torch.testing.assert_close(new_feat, torch.from_numpy(g["sa_features"]), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(out, torch.from_numpy(g["fp_out"]), rtol=1e-6, atol=1e-6)


def test_mirror_backbone_matches_reference_python():
    g = load("backbone_small.npz")
    from epnet_b200 import Pointnet2MSG
    net = Pointnet2MSG(config=small_backbone_config())
    net.load_state_dict(state_dict(g, "w."), strict=True)
    net.eval()
    xy = torch.from_numpy(g["xy"]).clone()
    with torch.no_grad(), cpu_oracle():
        xyz, feat = net(torch.from_numpy(g["points"]), torch.from_numpy(g["image"]), xy)
    assert torch.equal(xy, torch.from_numpy(g["xy_after_call"]))  # the in-place normalisation of pointnet2_msg.py:209-210
    assert torch.equal(xyz, torch.from_numpy(g["out_xyz"]))
    torch.testing.assert_close(feat, torch.from_numpy(g["out_features"]), rtol=1e-5, atol=1e-6)
