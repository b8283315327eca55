"""oracle/cpu_backend.py -- TEST INFRASTRUCTURE.  The C restatement (liboracle.so) behind the nine `*_wrapper`
names of the reference's pybind table, on CPU torch tensors (zero-copy numpy views).  It lets the reference's own
Python layers run in the GPU-less build container (tests/golden/make_golden.py) and lets the mirror modules be
checked against them on CPU."""
import torch

from . import _ball, _fps, _gather, _gather_grad, _group, _group_grad, _interp, _interp_grad, _nn


def _np(t):
    assert not t.is_cuda and t.is_contiguous()
    return t.detach().numpy()


def furthest_point_sampling_wrapper(b, n, m, points_tensor, temp_tensor, idx_tensor):
    _fps(b, n, m, _np(points_tensor), _np(temp_tensor), _np(idx_tensor))
    return 1


def gather_points_wrapper(b, c, n, npoints, points_tensor, idx_tensor, out_tensor):
    _gather(b, c, n, npoints, _np(points_tensor), _np(idx_tensor), _np(out_tensor))
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out_tensor, idx_tensor, grad_points_tensor):
    _gather_grad(b, c, n, npoints, _np(grad_out_tensor), _np(idx_tensor), _np(grad_points_tensor))
    return 1


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz_tensor, xyz_tensor, idx_tensor):
    _ball(b, n, m, float(radius), nsample, _np(new_xyz_tensor), _np(xyz_tensor), _np(idx_tensor))
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points_tensor, idx_tensor, out_tensor):
    _group(b, c, n, npoints, nsample, _np(points_tensor), _np(idx_tensor), _np(out_tensor))
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out_tensor, idx_tensor, grad_points_tensor):
    _group_grad(b, c, n, npoints, nsample, _np(grad_out_tensor), _np(idx_tensor), _np(grad_points_tensor))
    return 1


def three_nn_wrapper(b, n, m, unknown_tensor, known_tensor, dist2_tensor, idx_tensor):
    _nn(b, n, m, _np(unknown_tensor), _np(known_tensor), _np(dist2_tensor), _np(idx_tensor))


def three_interpolate_wrapper(b, c, m, n, points_tensor, idx_tensor, weight_tensor, out_tensor):
    _interp(b, c, m, n, _np(points_tensor), _np(idx_tensor), _np(weight_tensor), _np(out_tensor))


def three_interpolate_grad_wrapper(b, c, n, m, grad_out_tensor, idx_tensor, weight_tensor, grad_points_tensor):
    _interp_grad(b, c, n, m, _np(grad_out_tensor), _np(idx_tensor), _np(weight_tensor), _np(grad_points_tensor))
