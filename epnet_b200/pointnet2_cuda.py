"""`pointnet2_cuda` -- the extension module the reference imports at
pointnet2_lib/pointnet2/pointnet2_utils.py:7, re-implemented on libepnet_b200.so.

Same nine function names and positional signatures as the reference's pybind table
(pointnet2_lib/pointnet2/src/pointnet2_api.cpp:11-23), so the reference's own pointnet2_utils.py runs
unchanged on top of it (see epnet_b200.install()).  Like the reference wrappers (sampling.cpp:17) work
is enqueued on the calling thread's current CUDA stream and nothing synchronises; unlike them, bad
inputs raise instead of being trusted, and a failed launch raises instead of exit(-1).
"""
import torch

from ._lib import LIB, check


def _dev_ptr(t, dtype, name):
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (epnet_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t.data_ptr()


def _f(t, name):
    return _dev_ptr(t, torch.float32, name)


def _i(t, name):
    return _dev_ptr(t, torch.int32, name)


def _stream(t):
    # current stream of the tensor's device, like THCState_getCurrentStream in the reference wrappers;
    # the device is made current so launches land on it even from a thread that never set it.
    return torch.cuda.current_stream(t.device).cuda_stream


class _on:
    """device guard: the reference relies on the caller having selected the device; we do it."""

    def __init__(self, t):
        self.guard = torch.cuda.device(t.device)

    def __enter__(self):
        self.guard.__enter__()

    def __exit__(self, *a):
        return self.guard.__exit__(*a)


def furthest_point_sampling_wrapper(b, n, m, points_tensor, temp_tensor, idx_tensor):
    with _on(points_tensor):
        check(LIB.epnet_furthest_point_sampling(b, n, m, _f(points_tensor, "xyz"), _f(temp_tensor, "temp"),
                                                _i(idx_tensor, "idx"), _stream(points_tensor)),
              "furthest_point_sampling")
    return 1


def gather_points_wrapper(b, c, n, npoints, points_tensor, idx_tensor, out_tensor):
    with _on(points_tensor):
        check(LIB.epnet_gather_points(b, c, n, npoints, _f(points_tensor, "points"), _i(idx_tensor, "idx"),
                                      _f(out_tensor, "out"), _stream(points_tensor)), "gather_points")
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out_tensor, idx_tensor, grad_points_tensor):
    with _on(grad_out_tensor):
        check(LIB.epnet_gather_points_grad(b, c, n, npoints, _f(grad_out_tensor, "grad_out"), _i(idx_tensor, "idx"),
                                           _f(grad_points_tensor, "grad_points"), _stream(grad_out_tensor)),
              "gather_points_grad")
    return 1


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz_tensor, xyz_tensor, idx_tensor):
    with _on(xyz_tensor):
        check(LIB.epnet_ball_query(b, n, m, float(radius), nsample, _f(new_xyz_tensor, "new_xyz"), _f(xyz_tensor, "xyz"),
                                   _i(idx_tensor, "idx"), _stream(xyz_tensor)), "ball_query")
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points_tensor, idx_tensor, out_tensor):
    with _on(points_tensor):
        check(LIB.epnet_group_points(b, c, n, npoints, nsample, _f(points_tensor, "points"), _i(idx_tensor, "idx"),
                                     _f(out_tensor, "out"), _stream(points_tensor)), "group_points")
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out_tensor, idx_tensor, grad_points_tensor):
    with _on(grad_out_tensor):
        check(LIB.epnet_group_points_grad(b, c, n, npoints, nsample, _f(grad_out_tensor, "grad_out"),
                                          _i(idx_tensor, "idx"), _f(grad_points_tensor, "grad_points"),
                                          _stream(grad_out_tensor)), "group_points_grad")
    return 1


def three_nn_wrapper(b, n, m, unknown_tensor, known_tensor, dist2_tensor, idx_tensor):
    with _on(unknown_tensor):
        check(LIB.epnet_three_nn(b, n, m, _f(unknown_tensor, "unknown"), _f(known_tensor, "known"),
                                 _f(dist2_tensor, "dist2"), _i(idx_tensor, "idx"), _stream(unknown_tensor)), "three_nn")


def three_interpolate_wrapper(b, c, m, n, points_tensor, idx_tensor, weight_tensor, out_tensor):
    with _on(points_tensor):
        check(LIB.epnet_three_interpolate(b, c, m, n, _f(points_tensor, "points"), _i(idx_tensor, "idx"),
                                          _f(weight_tensor, "weight"), _f(out_tensor, "out"), _stream(points_tensor)),
              "three_interpolate")


def three_interpolate_grad_wrapper(b, c, n, m, grad_out_tensor, idx_tensor, weight_tensor, grad_points_tensor):
    with _on(grad_out_tensor):
        check(LIB.epnet_three_interpolate_grad(b, c, n, m, _f(grad_out_tensor, "grad_out"), _i(idx_tensor, "idx"),
                                               _f(weight_tensor, "weight"), _f(grad_points_tensor, "grad_points"),
                                               _stream(grad_out_tensor)), "three_interpolate_grad")


# ---- LI-Fusion gather (not part of the reference's pybind table; see epnet_b200.li_fusion) ----

def grid_gather_bilinear_wrapper(b, c, h, w, n, fmap_tensor, xy_tensor, align_corners, out_tensor):
    with _on(fmap_tensor):
        check(LIB.epnet_grid_gather_bilinear(b, c, h, w, n, _f(fmap_tensor, "feature_map"), _f(xy_tensor, "xy"),
                                             int(bool(align_corners)), _f(out_tensor, "out"), _stream(fmap_tensor)),
              "grid_gather_bilinear")


def grid_gather_bilinear_grad_wrapper(b, c, h, w, n, grad_out_tensor, xy_tensor, align_corners, grad_fmap_tensor):
    with _on(grad_out_tensor):
        check(LIB.epnet_grid_gather_bilinear_grad(b, c, h, w, n, _f(grad_out_tensor, "grad_out"), _f(xy_tensor, "xy"),
                                                  int(bool(align_corners)), _f(grad_fmap_tensor, "grad_feature_map"),
                                                  _stream(grad_out_tensor)), "grid_gather_bilinear_grad")
