"""Python op surface: the six autograd Functions plus QueryAndGroup / GroupAll of the reference's
pointnet2_lib/pointnet2/pointnet2_utils.py (same names, argument order and return values), running on
the B200 kernels of libepnet_b200.so.

The Functions are produced by `make_ops(backend)` where `backend` is any module exposing the nine
`*_wrapper` functions of the reference's pybind table; the module-level names below are bound to
epnet_b200.pointnet2_cuda.  tests/ and bench.py's reference arm build a second set on the reference's
own kernels (oracle/ref_cuda.py) through the same factory -- the product never does.
"""
from types import SimpleNamespace
from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from . import pointnet2_cuda as _default_backend


def make_ops(backend):
    """Build the op set of pointnet2_utils.py on `backend` (reference line numbers in each docstring)."""

    class FurthestPointSampling(Function):
        @staticmethod
        def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
            """pointnet2_utils.py:12-29.  xyz (B,N,3) -> (B,npoint) int32 indices; bit-exact."""
            assert xyz.is_contiguous()
            B, N, _ = xyz.size()
            output = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
            temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
            backend.furthest_point_sampling_wrapper(B, N, npoint, xyz, temp, output)
            ctx.mark_non_differentiable(output)
            return output

        @staticmethod
        def backward(ctx, a=None):
            return None, None

    class GatherOperation(Function):
        @staticmethod
        def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
            """pointnet2_utils.py:42-63.  features (B,C,N), idx (B,npoint) -> (B,C,npoint)."""
            assert features.is_contiguous()
            assert idx.is_contiguous()
            B, npoint = idx.size()
            _, C, N = features.size()
            output = torch.empty((B, C, npoint), dtype=torch.float32, device=features.device)
            backend.gather_points_wrapper(B, C, N, npoint, features, idx, output)
            ctx.for_backwards = (idx, C, N)
            return output

        @staticmethod
        def backward(ctx, grad_out):
            idx, C, N = ctx.for_backwards
            B, npoint = idx.size()
            grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
            backend.gather_points_grad_wrapper(B, C, N, npoint, grad_out.contiguous(), idx, grad_features)
            return grad_features, None

    class ThreeNN(Function):
        @staticmethod
        def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
            """pointnet2_utils.py:79-98.  Returns (sqrt of the 3 smallest squared distances, their indices)."""
            assert unknown.is_contiguous()
            assert known.is_contiguous()
            B, N, _ = unknown.size()
            m = known.size(1)
            dist2 = torch.empty((B, N, 3), dtype=torch.float32, device=unknown.device)
            idx = torch.empty((B, N, 3), dtype=torch.int32, device=unknown.device)
            backend.three_nn_wrapper(B, N, m, unknown, known, dist2, idx)
            ctx.mark_non_differentiable(idx)
            return torch.sqrt(dist2), idx

        @staticmethod
        def backward(ctx, a=None, b=None):
            return None, None

    class ThreeInterpolate(Function):
        @staticmethod
        def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
            """pointnet2_utils.py:111-131.  features (B,C,m), idx/weight (B,n,3) -> (B,C,n)."""
            assert features.is_contiguous()
            assert idx.is_contiguous()
            assert weight.is_contiguous()
            B, c, m = features.size()
            n = idx.size(1)
            ctx.three_interpolate_for_backward = (idx, weight, m)
            output = torch.empty((B, c, n), dtype=torch.float32, device=features.device)
            backend.three_interpolate_wrapper(B, c, m, n, features, idx, weight, output)
            return output

        @staticmethod
        def backward(ctx, grad_out: torch.Tensor):
            idx, weight, m = ctx.three_interpolate_for_backward
            B, c, n = grad_out.size()
            grad_features = torch.zeros((B, c, m), dtype=torch.float32, device=grad_out.device)
            backend.three_interpolate_grad_wrapper(B, c, n, m, grad_out.contiguous(), idx, weight, grad_features)
            return grad_features, None, None

    class GroupingOperation(Function):
        @staticmethod
        def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
            """pointnet2_utils.py:159-179.  features (B,C,N), idx (B,npoint,nsample) -> (B,C,npoint,nsample)."""
            assert features.is_contiguous()
            assert idx.is_contiguous()
            B, nfeatures, nsample = idx.size()
            _, C, N = features.size()
            output = torch.empty((B, C, nfeatures, nsample), dtype=torch.float32, device=features.device)
            backend.group_points_wrapper(B, C, N, nfeatures, nsample, features, idx, output)
            ctx.for_backwards = (idx, N)
            return output

        @staticmethod
        def backward(ctx, grad_out: torch.Tensor):
            idx, N = ctx.for_backwards
            B, C, npoint, nsample = grad_out.size()
            grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
            backend.group_points_grad_wrapper(B, C, N, npoint, nsample, grad_out.contiguous(), idx, grad_features)
            return grad_features, None

    class BallQuery(Function):
        @staticmethod
        def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
            """pointnet2_utils.py:203-221.  -> (B,npoint,nsample) int32, bit-exact incl. padding order."""
            assert new_xyz.is_contiguous()
            assert xyz.is_contiguous()
            B, N, _ = xyz.size()
            npoint = new_xyz.size(1)
            idx = torch.zeros((B, npoint, nsample), dtype=torch.int32, device=xyz.device)
            backend.ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx)
            ctx.mark_non_differentiable(idx)
            return idx

        @staticmethod
        def backward(ctx, a=None):
            return None, None, None, None

    ops = SimpleNamespace(
        backend=backend,
        FurthestPointSampling=FurthestPointSampling, furthest_point_sample=FurthestPointSampling.apply,
        GatherOperation=GatherOperation, gather_operation=GatherOperation.apply,
        ThreeNN=ThreeNN, three_nn=ThreeNN.apply,
        ThreeInterpolate=ThreeInterpolate, three_interpolate=ThreeInterpolate.apply,
        GroupingOperation=GroupingOperation, grouping_operation=GroupingOperation.apply,
        BallQuery=BallQuery, ball_query=BallQuery.apply,
    )
    return ops


OPS = make_ops(_default_backend)

FurthestPointSampling = OPS.FurthestPointSampling
furthest_point_sample = OPS.furthest_point_sample
GatherOperation = OPS.GatherOperation
gather_operation = OPS.gather_operation
ThreeNN = OPS.ThreeNN
three_nn = OPS.three_nn
ThreeInterpolate = OPS.ThreeInterpolate
three_interpolate = OPS.three_interpolate
GroupingOperation = OPS.GroupingOperation
grouping_operation = OPS.grouping_operation
BallQuery = OPS.BallQuery
ball_query = OPS.ball_query


class QueryAndGroup(nn.Module):
    """pointnet2_utils.py:231-264: ball query, group xyz (re-centred on the query) and features."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True, ops=None):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self._ops = ops or OPS

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None) -> torch.Tensor:
        """xyz (B,N,3), new_xyz (B,npoint,3), features (B,C,N) -> (B,3+C,npoint,nsample)."""
        ops = self._ops
        idx = ops.ball_query(self.radius, self.nsample, xyz, new_xyz)
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = ops.grouping_operation(xyz_trans, idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            grouped_features = ops.grouping_operation(features, idx)
            if self.use_xyz:
                return torch.cat([grouped_xyz, grouped_features], dim=1)
            return grouped_features
        assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
        return grouped_xyz


class GroupAll(nn.Module):
    """pointnet2_utils.py:267-290: a single group holding every point."""

    def __init__(self, use_xyz: bool = True, ops=None):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            if self.use_xyz:
                return torch.cat([grouped_xyz, grouped_features], dim=1)
            return grouped_features
        return grouped_xyz
