"""Set-abstraction and feature-propagation modules with the interface, attribute names and return
values of the reference's pointnet2_lib/pointnet2/pointnet2_modules.py, on the B200 ops.

`ops` selects the op set (epnet_b200.pointnet2_utils.make_ops); None = the product kernels.
"""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils
from . import pytorch_utils as pt_utils


class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None
        self.pool_method = "max_pool"
        self._ops = pointnet2_utils.OPS

    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None, new_xyz=None):
        """pointnet2_modules.py:19-72.
        xyz (B,N,3), features (B,C,N) -> (new_xyz (B,npoint,3), new_features (B,sum_k mlps[k][-1],npoint),
        idx (B,npoint) int32 FPS indices or None).  EPNet's fork returns the FPS indices as a third value
        (:72); LI-Fusion uses them to carry pixel coordinates along (lib/net/pointnet2_msg.py:217-219)."""
        ops = self._ops
        new_features_list = []
        idx = None
        if new_xyz is None and self.npoint is not None:
            xyz_flipped = xyz.transpose(1, 2).contiguous()
            idx = ops.furthest_point_sample(xyz, self.npoint)
            new_xyz = ops.gather_operation(xyz_flipped, idx).transpose(1, 2).contiguous()

        for grouper, mlp in zip(self.groupers, self.mlps):
            new_features = grouper(xyz, new_xyz, features)  # (B, C, npoint, nsample)
            new_features = mlp(new_features)  # (B, mlp[-1], npoint, nsample)
            if self.pool_method == "max_pool":
                new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
            elif self.pool_method == "avg_pool":
                new_features = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)])
            else:
                raise NotImplementedError
            new_features_list.append(new_features.squeeze(-1))  # (B, mlp[-1], npoint)

        return new_xyz, torch.cat(new_features_list, dim=1), idx


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Set abstraction with multi-scale grouping (pointnet2_modules.py:75-109)."""

    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int], mlps: List[List[int]], bn: bool = True,
                 use_xyz: bool = True, pool_method="max_pool", instance_norm=False, ops=None):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        if ops is not None:
            self._ops = ops
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for radius, nsample, mlp_spec in zip(radii, nsamples, mlps):
            self.groupers.append(pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz, ops=self._ops)
                                 if npoint is not None else pointnet2_utils.GroupAll(use_xyz))
            if use_xyz:
                mlp_spec[0] += 3  # in place, like the reference (:105-106): callers see the widened spec
            self.mlps.append(pt_utils.SharedMLP(mlp_spec, bn=bn, instance_norm=instance_norm))
        self.pool_method = pool_method


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set abstraction (pointnet2_modules.py:112-130)."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None, bn: bool = True,
                 use_xyz: bool = True, pool_method="max_pool", instance_norm=False, ops=None):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz,
                         pool_method=pool_method, instance_norm=instance_norm, ops=ops)


class PointnetFPModule(nn.Module):
    """Feature propagation (pointnet2_modules.py:133-173): inverse-distance interpolation of the coarser
    level's features onto the finer level's points, concatenation with the skip features, shared MLP."""

    def __init__(self, *, mlp: List[int], bn: bool = True, activation=nn.ReLU(inplace=True), ops=None):
        super().__init__()
        self.mlp = pt_utils.SharedMLP(mlp, bn=bn, activation=activation)
        self._ops = ops or pointnet2_utils.OPS

    def forward(self, unknown: torch.Tensor, known: torch.Tensor, unknow_feats: torch.Tensor,
                known_feats: torch.Tensor) -> torch.Tensor:
        """unknown (B,n,3), known (B,m,3), unknow_feats (B,C1,n), known_feats (B,C2,m) -> (B,mlp[-1],n)."""
        ops = self._ops
        if known is not None:
            dist, idx = ops.three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            norm = torch.sum(dist_recip, dim=2, keepdim=True)
            weight = dist_recip / norm
            interpolated_feats = ops.three_interpolate(known_feats, idx, weight)
        else:
            interpolated_feats = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))

        if unknow_feats is not None:
            new_features = torch.cat([interpolated_feats, unknow_feats], dim=1)
        else:
            new_features = interpolated_feats
        return self.mlp(new_features.unsqueeze(-1)).squeeze(-1)
