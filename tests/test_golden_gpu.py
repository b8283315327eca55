"""The CUDA product against fixtures recorded from the reference's own Python layers (tests/golden/, made in the build
container where /root/reference exists; this file never reads /root/reference)."""
import numpy as np
import pytest
import torch

from golden_util import load, small_backbone_config, state_dict

pytestmark = pytest.mark.gpu


def _strict():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def test_op_surface_on_cuda():
    g = load("ops_lidar2048.npz")
    from epnet_b200 import pointnet2_utils as pu
    pts, feats = torch.from_numpy(g["points"]).cuda(), torch.from_numpy(g["feats"]).cuda()
    idx = pu.furthest_point_sample(pts, 512)
    np.testing.assert_array_equal(idx.cpu().numpy(), g["fps_idx"])
    new_xyz = pu.gather_operation(pts.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    np.testing.assert_array_equal(new_xyz.cpu().numpy(), g["new_xyz"])
    ball = pu.ball_query(0.8, 16, pts, new_xyz)
    np.testing.assert_array_equal(ball.cpu().numpy(), g["ball_idx"])
    np.testing.assert_array_equal(pu.grouping_operation(feats, ball).cpu().numpy(), g["grouped"])
    qg = pu.QueryAndGroup(0.8, 16, use_xyz=True)(pts, new_xyz, feats)
    np.testing.assert_array_equal(qg.cpu().numpy(), g["query_and_group"])
    dist, nn_idx = pu.three_nn(pts, new_xyz)
    np.testing.assert_array_equal(nn_idx.cpu().numpy(), g["nn_idx"])
    np.testing.assert_allclose(dist.cpu().numpy(), g["nn_dist"], rtol=1e-6, atol=0)
    up = pu.three_interpolate(torch.from_numpy(g["grouped"].max(axis=3)).cuda(), nn_idx, torch.from_numpy(g["weight"]).cuda())
    np.testing.assert_allclose(up.cpu().numpy(), g["interpolated"], rtol=1e-5, atol=1e-7)


def test_sa_fp_modules_on_cuda():
    _strict()
    g = load("modules_sa_fp.npz")
    from epnet_b200.pointnet2_modules import PointnetFPModule, PointnetSAModuleMSG
    sa = PointnetSAModuleMSG(npoint=128, radii=[0.8, 1.6], nsamples=[16, 32], mlps=[[6, 8, 16], [6, 8, 24]], use_xyz=True, bn=True)
    fp = PointnetFPModule(mlp=[40 + 6, 32, 16])
    sa.load_state_dict(state_dict(g, "sa.")); fp.load_state_dict(state_dict(g, "fp."))
    sa.cuda().eval(); fp.cuda().eval()
    pts, feats = torch.from_numpy(g["points"]).cuda(), torch.from_numpy(g["feats"]).cuda()
    with torch.no_grad():
        new_xyz, new_feat, idx = sa(pts, feats)
        out = fp(pts, new_xyz, feats, new_feat)
    np.testing.assert_array_equal(idx.cpu().numpy(), g["sa_idx"])
    np.testing.assert_array_equal(new_xyz.cpu().numpy(), g["sa_new_xyz"])
    np.testing.assert_allclose(new_feat.cpu().numpy(), g["sa_features"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(out.cpu().numpy(), g["fp_out"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("path", ["module", "runner"])
def test_small_backbone_on_cuda(path):
    _strict()
    g = load("backbone_small.npz")
    from epnet_b200 import Pointnet2MSG
    net = Pointnet2MSG(config=small_backbone_config())
    net.load_state_dict(state_dict(g, "w."), strict=True)
    net.cuda().eval()
    net.auto_fast_inference = False
    pts, img = torch.from_numpy(g["points"]).cuda(), torch.from_numpy(g["image"]).cuda()
    xy = torch.from_numpy(g["xy"]).cuda()
    with torch.no_grad():
        if path == "module":
            xyz, feat = net(pts, img, xy.clone())
        else:
            runner = net.make_runner(2, 1024, torch.device("cuda"), image_hw=(96, 320))
            xyz, feat = runner(pts, img, xy)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(xyz.cpu().numpy(), g["out_xyz"])
    want = g["out_features"]
    err = np.abs(feat.cpu().numpy() - want).max()
    assert err <= 2e-5 * np.abs(want).max(), (err, np.abs(want).max())
