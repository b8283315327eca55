"""LI-Fusion point-wise image-feature gather: the B200 replacement for the `grid_sample` call site of
Feature_Gather (/root/reference/lib/net/pointnet2_msg.py:107-120).

`grid_sample(feature_map, xy)` below accepts what that call site passes -- feature_map (B,C,H,W) and a
grid of shape (B,1,N,2) -- and returns (B,C,1,N), so it can be bound over the module-global name
`grid_sample` of the reference's pointnet2_msg (see epnet_b200.install()).  `align_corners` defaults to
None = torch's own default (False on torch >= 1.3), so the unchanged reference call computes what it
computes with stock torch on the same box; pass True for the torch-1.2 semantics the published
checkpoints were trained with.
"""
import torch
from torch.autograd import Function

from . import pointnet2_cuda as _backend


class GridGatherBilinear(Function):
    @staticmethod
    def forward(ctx, feature_map: torch.Tensor, xy: torch.Tensor, align_corners: bool) -> torch.Tensor:
        """feature_map (B,C,H,W), xy (B,N,2) normalised to [-1,1] -> (B,C,N)."""
        feature_map = feature_map.contiguous()
        xy = xy.contiguous()
        B, C, H, W = feature_map.shape
        N = xy.shape[1]
        out = torch.empty((B, C, N), dtype=torch.float32, device=feature_map.device)
        _backend.grid_gather_bilinear_wrapper(B, C, H, W, N, feature_map, xy, align_corners, out)
        ctx.save_for_backward(xy)
        ctx.meta = (B, C, H, W, N, bool(align_corners))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (xy,) = ctx.saved_tensors
        B, C, H, W, N, align_corners = ctx.meta
        grad_map = torch.zeros((B, C, H, W), dtype=torch.float32, device=grad_out.device)
        _backend.grid_gather_bilinear_grad_wrapper(B, C, H, W, N, grad_out.contiguous(), xy, align_corners, grad_map)
        return grad_map, None, None


def feature_gather(feature_map: torch.Tensor, xy: torch.Tensor, align_corners: bool = False) -> torch.Tensor:
    """Feature_Gather (pointnet2_msg.py:107-120): (B,C,H,W) x (B,N,2) -> (B,C,N)."""
    return GridGatherBilinear.apply(feature_map, xy, bool(align_corners))


def grid_sample(input, grid, mode="bilinear", padding_mode="zeros", align_corners=None):  # noqa: A002
    """Drop-in for torch.nn.functional.grid_sample restricted to the LI-Fusion use: bilinear, zero padding,
    grid (B,1,N,2).  Anything else is refused (there is no fallback to ATen)."""
    if mode != "bilinear" or padding_mode != "zeros":
        raise NotImplementedError("epnet_b200.grid_sample: only mode='bilinear', padding_mode='zeros'")
    if grid.dim() != 4 or grid.shape[1] != 1 or grid.shape[-1] != 2:
        raise NotImplementedError("epnet_b200.grid_sample: grid must have shape (B,1,N,2), got %s" % (tuple(grid.shape),))
    return feature_gather(input, grid[:, 0], bool(align_corners)).unsqueeze(2)
