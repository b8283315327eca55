"""Python op surface of pointnet2_lib/pointnet2/pointnet2_utils.py on the B200 kernels: the six autograd ops
(FurthestPointSampling, GatherOperation, ThreeNN, ThreeInterpolate, GroupingOperation, BallQuery and their lower-case
`.apply` aliases) plus the QueryAndGroup / GroupAll modules -- same names, argument order and return values.

Design: every op is described once by a pair of plain functions (launch, optional gradient launch) over the nine `*_wrapper`
entry points of the reference's pybind table, served by epnet_b200.pointnet2_cuda (libepnet_b200.so) -- the only backend; a table
turns them into `torch.autograd.Function` classes.  There is no backend parameter anywhere in the product: the CPU tests swap the
module-level `_BACKEND` for the C oracle from tests/backend_swap.py.  Buffers the reference expects pre-filled (FPS scratch 1e10, zeroed ball-query indices, zeroed gradients) are
prepared here, exactly where the reference's Python layer prepares them (pointnet2_utils.py:26, :218, :67/:146/:190).
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import pointnet2_cuda as _BACKEND  # the one backend; looked up at call time (tests/backend_swap.py replaces it on CPU)


def _buf(like, *shape, dtype=torch.float32, fill=None):
    """output / scratch tensor on `like`'s device; fill=None leaves it uninitialised"""
    if fill is None:
        return torch.empty(shape, dtype=dtype, device=like.device)
    return torch.full(shape, fill, dtype=dtype, device=like.device)


def _dense(*tensors):
    for t in tensors:
        assert t.is_contiguous(), "pointnet2 ops take contiguous tensors (as the reference asserts)"


# ---- op table: name -> (forward launch, gradient launch or None, number of forward arguments) -------------------------
# forward(be, ctx, *args) returns the op's outputs; it may stash what the gradient needs on ctx.saved
# gradient(be, ctx, *grads) returns the gradient w.r.t. the FIRST forward argument (the only differentiable one)

def _fps(be, ctx, cloud, count):  # (B,N,3) -> (B,count) int32 sample indices, bit-exact (pointnet2_utils.py:12-29)
    _dense(cloud)
    scenes, n = cloud.shape[0], cloud.shape[1]
    picked = _buf(cloud, scenes, count, dtype=torch.int32)
    be.furthest_point_sampling_wrapper(scenes, n, count, cloud, _buf(cloud, scenes, n, fill=1e10), picked)
    ctx.mark_non_differentiable(picked)
    return picked


def _gather(be, ctx, feats, which):  # (B,C,N), (B,M) -> (B,C,M)   (pointnet2_utils.py:42-63)
    _dense(feats, which)
    scenes, chans, n = feats.shape
    m = which.shape[1]
    ctx.saved = (which, chans, n)
    picked = _buf(feats, scenes, chans, m)
    be.gather_points_wrapper(scenes, chans, n, m, feats, which, picked)
    return picked


def _gather_grad(be, ctx, g):  # scatter-add back (pointnet2_utils.py:65-73)
    which, chans, n = ctx.saved
    scenes, m = which.shape
    acc = _buf(g, scenes, chans, n, fill=0.0)
    be.gather_points_grad_wrapper(scenes, chans, n, m, g.contiguous(), which, acc)
    return acc


def _nn3(be, ctx, queries, anchors):  # -> (distances (B,n,3), indices (B,n,3))   (pointnet2_utils.py:79-98)
    _dense(queries, anchors)
    scenes, n = queries.shape[0], queries.shape[1]
    sq = _buf(queries, scenes, n, 3)
    which = _buf(queries, scenes, n, 3, dtype=torch.int32)
    be.three_nn_wrapper(scenes, n, anchors.shape[1], queries, anchors, sq, which)
    ctx.mark_non_differentiable(which)
    return sq.sqrt(), which  # the kernel returns squared distances; the reference takes the root in Python too (:98)


def _interp(be, ctx, feats, which, w):  # (B,C,m), (B,n,3) x2 -> (B,C,n)   (pointnet2_utils.py:111-131)
    _dense(feats, which, w)
    scenes, chans, m = feats.shape
    n = which.shape[1]
    ctx.saved = (which, w, m)
    mixed = _buf(feats, scenes, chans, n)
    be.three_interpolate_wrapper(scenes, chans, m, n, feats, which, w, mixed)
    return mixed


def _interp_grad(be, ctx, g):  # (pointnet2_utils.py:133-153)
    which, w, m = ctx.saved
    scenes, chans, n = g.shape
    acc = _buf(g, scenes, chans, m, fill=0.0)
    be.three_interpolate_grad_wrapper(scenes, chans, n, m, g.contiguous(), which, w, acc)
    return acc


def _group(be, ctx, feats, which):  # (B,C,N), (B,M,S) -> (B,C,M,S)   (pointnet2_utils.py:159-179)
    _dense(feats, which)
    scenes, chans, n = feats.shape
    m, s = which.shape[1], which.shape[2]
    ctx.saved = (which, n)
    grouped = _buf(feats, scenes, chans, m, s)
    be.group_points_wrapper(scenes, chans, n, m, s, feats, which, grouped)
    return grouped


def _group_grad(be, ctx, g):  # (pointnet2_utils.py:181-197)
    which, n = ctx.saved
    scenes, chans, m, s = g.shape
    acc = _buf(g, scenes, chans, n, fill=0.0)
    be.group_points_grad_wrapper(scenes, chans, n, m, s, g.contiguous(), which, acc)
    return acc


def _ball(be, ctx, radius, per_ball, cloud, centres):  # -> (B,M,per_ball) int32, reference padding rule (pointnet2_utils.py:203-221)
    _dense(centres, cloud)
    scenes, n, m = cloud.shape[0], cloud.shape[1], centres.shape[1]
    members = _buf(cloud, scenes, m, per_ball, dtype=torch.int32, fill=0)  # an empty ball keeps these zeros
    be.ball_query_wrapper(scenes, n, m, radius, per_ball, centres, cloud, members)
    ctx.mark_non_differentiable(members)
    return members


_TABLE = {
    "FurthestPointSampling": (_fps, None, 2),
    "GatherOperation": (_gather, _gather_grad, 2),
    "ThreeNN": (_nn3, None, 2),
    "ThreeInterpolate": (_interp, _interp_grad, 3),
    "GroupingOperation": (_group, _group_grad, 2),
    "BallQuery": (_ball, None, 4),
}
_ALIASES = {"FurthestPointSampling": "furthest_point_sample", "GatherOperation": "gather_operation", "ThreeNN": "three_nn",
            "ThreeInterpolate": "three_interpolate", "GroupingOperation": "grouping_operation", "BallQuery": "ball_query"}


def _function_class(name, launch, grad, arity):
    def forward(ctx, *args):
        return launch(_BACKEND, ctx, *args)

    def backward(ctx, *grads):
        first = grad(_BACKEND, ctx, grads[0]) if grad is not None else None
        return (first,) + (None,) * (arity - 1)

    return type(name, (Function,), {"forward": staticmethod(forward), "backward": staticmethod(backward),
                                    "__doc__": "%s of pointnet2_utils.py on libepnet_b200.so" % name})


for _name, (_launch, _grad, _arity) in _TABLE.items():  # module-level names, as the reference exposes them
    globals()[_name] = _function_class(_name, _launch, _grad, _arity)
    globals()[_ALIASES[_name]] = globals()[_name].apply
del _name, _launch, _grad, _arity


class QueryAndGroup(nn.Module):
    """Ball query around every centre, then the members' coordinates (relative to the centre) and features, stacked on the
    channel axis (pointnet2_utils.py:231-264).  cloud (B,N,3), centres (B,M,3), feats (B,C,N) -> (B, 3+C, M, nsample)."""

    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        members = ball_query(self.radius, self.nsample, xyz, new_xyz)
        offsets = grouping_operation(xyz.transpose(1, 2).contiguous(), members) - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return offsets
        picked = grouping_operation(features, members)
        return torch.cat([offsets, picked], dim=1) if self.use_xyz else picked


class GroupAll(nn.Module):
    """One group holding the whole cloud (pointnet2_utils.py:267-290): (B,N,3), _, (B,C,N) -> (B, 3+C, 1, N)."""

    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        coords = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return coords
        feats = features.unsqueeze(2)
        return torch.cat([coords, feats], dim=1) if self.use_xyz else feats
