"""Generates tests/golden/*.npz by running the REFERENCE'S OWN Python layers, imported from /root/reference in this
(GPU-less) container:  pointnet2_lib/pointnet2/pointnet2_utils.py, pointnet2_modules.py and lib/net/pointnet2_msg.py.
Their CUDA extension `pointnet2_cuda` is replaced by the C oracle (oracle/cpu_backend.py) and the legacy
`torch.cuda.FloatTensor/IntTensor` constructors they allocate with are pointed at CPU constructors; nothing in the
reference files is modified.  The fixtures record inputs, weights and outputs so that the GPU box (which has no
/root/reference) can check the product against the reference's own composition.

    python tests/golden/make_golden.py        # needs /root/reference
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def import_reference():
    """-> (pointnet2_utils, pointnet2_modules, pointnet2_msg, cfg) of the reference, running on the CPU oracle."""
    from oracle import cpu_backend
    sys.modules["pointnet2_cuda"] = cpu_backend

    class EasyDict(dict):  # the reference needs `easydict`, which is not installed here
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                setattr(self, k, v)

        def __setattr__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setitem__(k, v)

        __setitem__ = __setattr__

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

    mod = types.ModuleType("easydict")
    mod.EasyDict = EasyDict
    sys.modules["easydict"] = mod
    torch.cuda.FloatTensor = lambda *s: torch.empty(*s, dtype=torch.float32)
    torch.cuda.IntTensor = lambda *s: torch.empty(*s, dtype=torch.int32)
    for p in (REF, os.path.join(REF, "lib", "net")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from pointnet2_lib.pointnet2 import pointnet2_modules, pointnet2_utils
    from lib.config import cfg
    import lib.net.pointnet2_msg as pointnet2_msg
    return pointnet2_utils, pointnet2_modules, pointnet2_msg, cfg


def small_cfg(cfg):
    """The published LI-Fusion-with-attention structure at a size whose weights fit a fixture."""
    cfg.RPN.USE_INTENSITY = False
    cfg.RPN.SA_CONFIG.NPOINTS = [256, 64, 16, 4]
    cfg.RPN.SA_CONFIG.RADIUS = [[0.8, 2.0], [2.0, 4.0], [4.0, 8.0], [8.0, 16.0]]
    cfg.RPN.SA_CONFIG.NSAMPLE = [[16, 32], [16, 32], [16, 32], [16, 32]]
    cfg.RPN.SA_CONFIG.MLPS = [[[8, 8, 16], [8, 8, 16]], [[16, 16, 32], [16, 24, 32]], [[32, 40, 48], [32, 40, 48]],
                              [[48, 48, 64], [48, 56, 64]]]
    cfg.RPN.FP_MLPS = [[32, 32], [48, 48], [64, 64], [64, 64]]
    cfg.LI_FUSION.ENABLED = True
    cfg.LI_FUSION.ADD_Image_Attention = True
    cfg.LI_FUSION.IMG_CHANNELS = [3, 8, 16, 24, 32]
    cfg.LI_FUSION.POINT_CHANNELS = [32, 64, 96, 128]
    cfg.LI_FUSION.IMG_FEATURES_CHANNEL = 32
    cfg.LI_FUSION.DeConv_Reduce = [4, 4, 4, 4]
    return cfg


def tnp(t):
    return t.detach().cpu().numpy()


def main():
    pu, pm, msg, cfg = import_reference()
    from epnet_b200 import scenes
    torch.manual_seed(0)
    out = {}

    # ---- op level: the six autograd Functions of pointnet2_utils.py ----
    pts = torch.stack([scenes.lidar_scene(1000 + i, 2048) for i in range(2)])
    feats = torch.randn(2, 6, 2048)
    idx = pu.furthest_point_sample(pts, 512)
    new_xyz = pu.gather_operation(pts.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    ball = pu.ball_query(0.8, 16, pts, new_xyz)
    grouped = pu.grouping_operation(feats, ball)
    dist, nn_idx = pu.three_nn(pts, new_xyz)
    recip = 1.0 / (dist + 1e-8)
    weight = recip / recip.sum(dim=2, keepdim=True)
    pooled = grouped.max(dim=3)[0].contiguous()
    up = pu.three_interpolate(pooled, nn_idx, weight.contiguous())
    qg = pu.QueryAndGroup(0.8, 16, use_xyz=True)(pts, new_xyz, feats)
    np.savez_compressed(os.path.join(HERE, "ops_lidar2048.npz"), points=tnp(pts), feats=tnp(feats), fps_idx=tnp(idx),
                        new_xyz=tnp(new_xyz), ball_idx=tnp(ball), grouped=tnp(grouped), nn_dist=tnp(dist), nn_idx=tnp(nn_idx),
                        weight=tnp(weight), interpolated=tnp(up), query_and_group=tnp(qg))

    # ---- module level: one SA-MSG and one FP module (train-mode BN = batch statistics, and eval-mode) ----
    sa = pm.PointnetSAModuleMSG(npoint=128, radii=[0.8, 1.6], nsamples=[16, 32], mlps=[[6, 8, 16], [6, 8, 24]], use_xyz=True, bn=True)
    fp = pm.PointnetFPModule(mlp=[40 + 6, 32, 16])
    sa.eval(); fp.eval()
    with torch.no_grad():
        for m in list(sa.modules()) + list(fp.modules()):
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5); m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.1)
        sa_xyz, sa_feat, sa_idx = sa(pts, feats)
        fp_out = fp(pts, sa_xyz, feats, sa_feat)
    mod = {"points": tnp(pts), "feats": tnp(feats), "sa_new_xyz": tnp(sa_xyz), "sa_features": tnp(sa_feat), "sa_idx": tnp(sa_idx),
           "fp_out": tnp(fp_out)}
    mod.update({"sa." + k: tnp(v) for k, v in sa.state_dict().items()})
    mod.update({"fp." + k: tnp(v) for k, v in fp.state_dict().items()})
    np.savez_compressed(os.path.join(HERE, "modules_sa_fp.npz"), **mod)

    # ---- backbone level: Pointnet2MSG.forward of lib/net/pointnet2_msg.py with LI-Fusion + attention, small config ----
    small_cfg(cfg)
    net = msg.Pointnet2MSG(input_channels=0, use_xyz=True).eval()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5); m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.1)
        data = scenes.batch(2000, 2, 1024)
        image = data["image"][:, :, ::4, ::4].contiguous()  # 96 x 320 canvas; xy stays in 1280x384 pixel units
        xy = data["xy"].clone()
        xyz_out, feat_out = net(data["points"], image, xy)
    bb = {"points": tnp(data["points"]), "image": tnp(image), "xy": tnp(data["xy"]), "xy_after_call": tnp(xy), "out_xyz": tnp(xyz_out),
          "out_features": tnp(feat_out)}
    bb.update({"w." + k: tnp(v) for k, v in net.state_dict().items()})
    np.savez_compressed(os.path.join(HERE, "backbone_small.npz"), **bb)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
