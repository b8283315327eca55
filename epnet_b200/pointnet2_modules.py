"""Set-abstraction (SA) and feature-propagation (FP) modules behind the interface of the reference's
pointnet2_lib/pointnet2/pointnet2_modules.py: same constructor keywords, the attribute names checkpoints are keyed by
(`groupers`, `mlps`, `mlp`), same return values -- including the FPS indices EPNet's fork returns as a third value
(pointnet2_modules.py:72), which LI-Fusion uses to carry pixel coordinates along (lib/net/pointnet2_msg.py:217-219).

These modules are the op-by-op ("module") path on the B200 kernels (epnet_b200.pointnet2_utils): training and the drop-in use;
inference goes through runner.py, which fuses most of this.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils
from . import pytorch_utils as pt_utils

_POOLS = {"max_pool": F.max_pool2d, "avg_pool": F.avg_pool2d}


class _PointnetSAModuleBase(nn.Module):
    """Sample centres (FPS unless the caller brings them), then per scale: group, shared MLP, pool over the group."""

    def __init__(self):
        super().__init__()
        self.npoint, self.groupers, self.mlps = None, None, None
        self.pool_method = "max_pool"

    def _centres(self, xyz):
        """(centres (B,npoint,3), FPS indices (B,npoint)); (None, None) for a GroupAll level (pointnet2_modules.py:36-45)"""
        if self.npoint is None:
            return None, None
        picked = pointnet2_utils.furthest_point_sample(xyz, self.npoint)
        as_rows = pointnet2_utils.gather_operation(xyz.transpose(1, 2).contiguous(), picked)
        return as_rows.transpose(1, 2).contiguous(), picked

    def forward(self, xyz, features=None, new_xyz=None):
        """xyz (B,N,3), features (B,C,N) -> (new_xyz (B,npoint,3), (B, sum of the scales' last widths, npoint), FPS idx | None)"""
        if self.pool_method not in _POOLS:
            raise NotImplementedError
        fps_idx = None
        if new_xyz is None:
            new_xyz, fps_idx = self._centres(xyz)
        per_scale = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            y = mlp(grouper(xyz, new_xyz, features))                      # (B, width, npoint, nsample)
            y = _POOLS[self.pool_method](y, kernel_size=[1, y.size(3)])   # pool the nsample axis (pointnet2_modules.py:59-66)
            per_scale.append(y.squeeze(-1))
        return new_xyz, torch.cat(per_scale, dim=1), fps_idx


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """SA level with multi-scale grouping: one (radius, nsample, MLP) triple per scale (pointnet2_modules.py:75-109)."""

    def __init__(self, *, npoint, radii, nsamples, mlps, bn=True, use_xyz=True, pool_method="max_pool", instance_norm=False):
        super().__init__()
        if not (len(radii) == len(nsamples) == len(mlps)):
            raise AssertionError("radii, nsamples and mlps describe the same scales")
        self.npoint, self.pool_method = npoint, pool_method
        self.groupers, self.mlps = nn.ModuleList(), nn.ModuleList()
        for radius, nsample, widths in zip(radii, nsamples, mlps):
            if npoint is None:
                self.groupers.append(pointnet2_utils.GroupAll(use_xyz))
            else:
                self.groupers.append(pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz))
            if use_xyz:
                widths[0] += 3  # the caller's list is widened in place, as the reference does (:105-106) and its callers expect
            self.mlps.append(pt_utils.SharedMLP(widths, bn=bn, instance_norm=instance_norm))


class PointnetSAModule(PointnetSAModuleMSG):
    """The single-scale special case (pointnet2_modules.py:112-130)."""

    def __init__(self, *, mlp, npoint=None, radius=None, nsample=None, bn=True, use_xyz=True, pool_method="max_pool",
                 instance_norm=False):
        super().__init__(npoint=npoint, radii=[radius], nsamples=[nsample], mlps=[mlp], bn=bn, use_xyz=use_xyz,
                         pool_method=pool_method, instance_norm=instance_norm)


class PointnetFPModule(nn.Module):
    """FP level (pointnet2_modules.py:133-173): every fine point takes the inverse-distance blend of its three nearest coarse
    points' features, the skip features are stacked underneath, a shared MLP follows."""

    def __init__(self, *, mlp, bn=True, activation=nn.ReLU(inplace=True)):
        super().__init__()
        self.mlp = pt_utils.SharedMLP(mlp, bn=bn, activation=activation)

    def forward(self, unknown, known, unknow_feats, known_feats):
        """unknown (B,n,3), known (B,m,3) | None, unknow_feats (B,C1,n) | None, known_feats (B,C2,m) -> (B, mlp[-1], n)"""
        if known is None:  # a GroupAll level above: its single feature column is broadcast to every point
            carried = known_feats.expand(known_feats.size(0), known_feats.size(1), unknown.size(1))
        else:
            dist, nearest = pointnet2_utils.three_nn(unknown, known)
            inv = 1.0 / (dist + 1e-8)                                   # pointnet2_modules.py:157-159
            carried = pointnet2_utils.three_interpolate(known_feats, nearest, inv / inv.sum(dim=2, keepdim=True))
        stacked = carried if unknow_feats is None else torch.cat([carried, unknow_feats], dim=1)
        return self.mlp(stacked.unsqueeze(-1)).squeeze(-1)
