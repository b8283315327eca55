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


def _dev_ptr(t, dtype, name, contiguous=True):
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (epnet_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if contiguous and not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    if not contiguous and t.stride(-1) != 1:
        raise ValueError("%s must have unit stride along its last axis" % name)
    return t.data_ptr()


def _f(t, name, contiguous=True):
    return _dev_ptr(t, torch.float32, name, contiguous)


def _i(t, name):
    return _dev_ptr(t, torch.int32, name)


LAUNCHES = [0]   # kernels launched through this module since it was last reset (bench.py's gpu_launches)
PROFILE = None   # set to a list to record (name, int args, start_event, end_event) around every launch


def _call(name, fn, ref, *args):
    """Launch `fn(*args, stream)` on the current stream of `ref`'s device.  The reference wrappers rely on the
    caller having selected the device (sampling.cpp:17); here it is selected when it is not current."""
    device = ref.device
    guard = None
    if torch.cuda.current_device() != device.index:
        guard = torch.cuda.device(device)
        guard.__enter__()
    try:
        stream = torch.cuda.current_stream(device)
        if PROFILE is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            code = fn(*args, stream.cuda_stream)
            e1.record(stream)
            PROFILE.append((name, tuple(a for a in args if isinstance(a, int) and abs(a) < (1 << 31)), e0, e1))
        else:
            code = fn(*args, stream.cuda_stream)
    finally:
        if guard is not None:
            guard.__exit__(None, None, None)
    LAUNCHES[0] += 1
    check(code, name)


def furthest_point_sampling_wrapper(b, n, m, points_tensor, temp_tensor, idx_tensor):
    _call("furthest_point_sampling", LIB.epnet_furthest_point_sampling, points_tensor, b, n, m, _f(points_tensor, "xyz"),
          _f(temp_tensor, "temp"), _i(idx_tensor, "idx"))
    return 1


def gather_points_wrapper(b, c, n, npoints, points_tensor, idx_tensor, out_tensor):
    _call("gather_points", LIB.epnet_gather_points, points_tensor, b, c, n, npoints, _f(points_tensor, "points"),
          _i(idx_tensor, "idx"), _f(out_tensor, "out"))
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out_tensor, idx_tensor, grad_points_tensor):
    _call("gather_points_grad", LIB.epnet_gather_points_grad, grad_out_tensor, b, c, n, npoints,
          _f(grad_out_tensor, "grad_out"), _i(idx_tensor, "idx"), _f(grad_points_tensor, "grad_points"))
    return 1


# clouds for which the runner sorts once per level and answers both radii through the buckets (a sort costs ~200 us on one SM
# per scene: it pays when it runs beside the FPS of the level, not in front of a single query -- the drop-in wrapper below
# therefore keeps the exhaustive scan: 226 us sorted vs 89 us exhaustive for ONE cold query at 16384 x 4096)
SORTED_QUERY_MIN_N, SORTED_QUERY_MAX_N, SORTED_QUERY_MAX_NSAMPLE = 8192, 16384, 64


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz_tensor, xyz_tensor, idx_tensor):
    _call("ball_query", LIB.epnet_ball_query, xyz_tensor, b, n, m, float(radius), nsample, _f(new_xyz_tensor, "new_xyz"),
          _f(xyz_tensor, "xyz"), _i(idx_tensor, "idx"))
    return 1


def bucket_cloud(xyz_tensor):
    """(sorted (B, npad, 4), boxes (B, npad/64, 8)) of a cloud (B, n, 3), n <= 16384: see epnet_bucket_cloud.  One sort serves
    every ball query (any radius, nsample <= 64) against that cloud."""
    import torch
    b, n = xyz_tensor.shape[0], xyz_tensor.shape[1]
    npad = 64
    while npad < n:
        npad *= 2
    srt = torch.empty((b, npad, 4), dtype=torch.float32, device=xyz_tensor.device)
    boxes = torch.empty((b, npad // 64, 8), dtype=torch.float32, device=xyz_tensor.device)
    _call("bucket_cloud", LIB.epnet_bucket_cloud, xyz_tensor, b, n, npad, _f(xyz_tensor, "xyz"), srt.data_ptr(), boxes.data_ptr())
    return srt, boxes


def ball_query_sorted_wrapper(b, m, radius, nsample, new_xyz_tensor, buckets, idx_tensor):
    srt, boxes = buckets
    _call("ball_query_sorted", LIB.epnet_ball_query_sorted, new_xyz_tensor, b, srt.shape[1], m, float(radius), nsample,
          _f(new_xyz_tensor, "new_xyz"), srt.data_ptr(), boxes.data_ptr(), _i(idx_tensor, "idx"))
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points_tensor, idx_tensor, out_tensor):
    _call("group_points", LIB.epnet_group_points, points_tensor, b, c, n, npoints, nsample, _f(points_tensor, "points"),
          _i(idx_tensor, "idx"), _f(out_tensor, "out"))
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out_tensor, idx_tensor, grad_points_tensor):
    _call("group_points_grad", LIB.epnet_group_points_grad, grad_out_tensor, b, c, n, npoints, nsample,
          _f(grad_out_tensor, "grad_out"), _i(idx_tensor, "idx"), _f(grad_points_tensor, "grad_points"))
    return 1


def three_nn_wrapper(b, n, m, unknown_tensor, known_tensor, dist2_tensor, idx_tensor):
    _call("three_nn", LIB.epnet_three_nn, unknown_tensor, b, n, m, _f(unknown_tensor, "unknown"), _f(known_tensor, "known"),
          _f(dist2_tensor, "dist2"), _i(idx_tensor, "idx"))


def three_interpolate_wrapper(b, c, m, n, points_tensor, idx_tensor, weight_tensor, out_tensor):
    _call("three_interpolate", LIB.epnet_three_interpolate, points_tensor, b, c, m, n, _f(points_tensor, "points"),
          _i(idx_tensor, "idx"), _f(weight_tensor, "weight"), _f(out_tensor, "out"))


def three_interpolate_grad_wrapper(b, c, n, m, grad_out_tensor, idx_tensor, weight_tensor, grad_points_tensor):
    _call("three_interpolate_grad", LIB.epnet_three_interpolate_grad, grad_out_tensor, b, c, n, m,
          _f(grad_out_tensor, "grad_out"), _i(idx_tensor, "idx"), _f(weight_tensor, "weight"),
          _f(grad_points_tensor, "grad_points"))


# ---- LI-Fusion gather (not part of the reference's pybind table; see epnet_b200.li_fusion) ----

def grid_gather_bilinear_wrapper(b, c, h, w, n, fmap_tensor, xy_tensor, align_corners, out_tensor):
    _call("grid_gather_bilinear", LIB.epnet_grid_gather_bilinear, fmap_tensor, b, c, h, w, n, _f(fmap_tensor, "feature_map"),
          _f(xy_tensor, "xy"), int(bool(align_corners)), _f(out_tensor, "out"))


def grid_gather_bilinear_grad_wrapper(b, c, h, w, n, grad_out_tensor, xy_tensor, align_corners, grad_fmap_tensor):
    _call("grid_gather_bilinear_grad", LIB.epnet_grid_gather_bilinear_grad, grad_out_tensor, b, c, h, w, n,
          _f(grad_out_tensor, "grad_out"), _f(xy_tensor, "xy"), int(bool(align_corners)),
          _f(grad_fmap_tensor, "grad_feature_map"))


# ---- fused entry points (include/epnet_b200.h); used by epnet_b200.runner, not by the reference surface ----

def _opt(t, dtype, name):
    return None if t is None else _dev_ptr(t, dtype, name)


def fps_sample_wrapper(b, n, m, xyz, temp, idx, new_xyz=None, aux_in=None, aux_out=None):
    aux_dim = 0 if aux_in is None else aux_in.shape[-1]
    _call("fps_sample", LIB.epnet_fps_sample, xyz, b, n, m, _f(xyz, "xyz"), _f(temp, "temp"), _i(idx, "idx"),
          _opt(new_xyz, torch.float32, "new_xyz"), _opt(aux_in, torch.float32, "aux_in"),
          _opt(aux_out, torch.float32, "aux_out"), aux_dim)


def fps_prefix_check_wrapper(b, n, m, xyz, winners, flag):
    """flag[s] = 1 iff furthest-point sampling of m of the n points of xyz[s] (temp = 1e10) is exactly the identity 0..m-1, every
    arg-max unique (csrc/fps.cu); winners (B, m) float32 scratch, flag (B,) int32"""
    _call("fps_prefix_check", LIB.epnet_fps_prefix_check, xyz, b, n, m, _f(xyz, "xyz"), _f(winners, "winners"), _i(flag, "flag"))


def fps_sample_guarded_wrapper(b, n, m, xyz, temp, idx, identity, new_xyz=None, aux_in=None, aux_out=None):
    """fps_sample_wrapper, except that scenes with identity[s] != 0 (fps_prefix_check_wrapper) are answered with the prefix"""
    aux_dim = 0 if aux_in is None else aux_in.shape[-1]
    _call("fps_sample_guarded", LIB.epnet_fps_sample_guarded, xyz, b, n, m, _f(xyz, "xyz"), _f(temp, "temp"), _i(idx, "idx"),
          _opt(new_xyz, torch.float32, "new_xyz"), _opt(aux_in, torch.float32, "aux_in"),
          _opt(aux_out, torch.float32, "aux_out"), aux_dim, _i(identity, "identity"))


def group_concat_wrapper(b, c, n, m, nsample, xyz, new_xyz, features, idx, out):
    _call("group_concat", LIB.epnet_group_concat, xyz, b, c, n, m, nsample, _f(xyz, "xyz"), _f(new_xyz, "new_xyz"),
          _opt(features, torch.float32, "features"), _i(idx, "idx"), _f(out, "out"))


def attention_scale_pm_wrapper(r1, r2, w3, b3, x, out):
    """out[row] = x[row] * sigmoid(w3 . tanh(r1[row] + r2[row]) + b3); r1, r2 (rows, rc), x, out (rows, c) row-major (strided ok)"""
    rows, rc, c = r1.shape[0], r1.shape[1], x.shape[1]
    _call("attention_scale_pm", LIB.epnet_attention_scale_pm, x, rows, rc, c, _f(r1, "r1", contiguous=False), r1.stride(0),
          _f(r2, "r2", contiguous=False), r2.stride(0), _f(w3, "w3"), _f(b3, "b3"), _f(x, "x", contiguous=False), x.stride(0),
          _f(out, "out", contiguous=False), out.stride(0))


def bias_relu_wrapper(b, c, l, x, bias):
    _call("bias_relu", LIB.epnet_bias_relu, x, b, c, l, _f(x, "x"), _f(bias, "bias"))


def bias_relu_maxpool_wrapper(b, c, m, nsample, x, bias, out_ptr, out_batch_stride):
    """out_ptr: raw device address of out[0, c_offset, 0] inside a (B, C_total, M) buffer with batch stride
    out_batch_stride floats."""
    _call("bias_relu_maxpool", LIB.epnet_bias_relu_maxpool, x, b, c, m, nsample, _f(x, "x"), _f(bias, "bias"), out_ptr,
          out_batch_stride)


def three_interpolate_concat_wrapper(b, c2, m, n, c1, known_feats, idx, dist2, skip_feats, out):
    _call("three_interpolate_concat", LIB.epnet_three_interpolate_concat, known_feats, b, c2, m, n, c1,
          _f(known_feats, "known_feats"), _i(idx, "idx"), _f(dist2, "dist2"), _opt(skip_feats, torch.float32, "skip_feats"),
          _f(out, "out"))


def group_concat_pm_wrapper(b, c, n, m, nsample, xyz, new_xyz, feats_pm, idx, out):
    """out (B*M*nsample, ldo) rows = [features (C) | xyz - centre (3) | pad]; feats_pm (B,N,C) point-major or None."""
    ldf = 0 if feats_pm is None else feats_pm.stride(-2)
    _call("group_concat_pm", LIB.epnet_group_concat_pm, xyz, b, c, n, m, nsample, _f(xyz, "xyz"), _f(new_xyz, "new_xyz"),
          _opt(feats_pm, torch.float32, "feats_pm"), ldf, _i(idx, "idx"), out.data_ptr(), out.stride(0))


def three_interpolate_concat_pm_wrapper(b, c2, m, n, c1, known_pm, idx, weight, skip_pm, out, from_dist2=False):
    _call("three_interpolate_concat_pm", LIB.epnet_three_interpolate_concat_pm, known_pm, b, c2, m, n, c1, known_pm.data_ptr(),
          known_pm.stride(-2), _i(idx, "idx"), _f(weight, "weight"), int(bool(from_dist2)), None if skip_pm is None else skip_pm.data_ptr(),
          0 if skip_pm is None else skip_pm.stride(-2), out.data_ptr(), out.stride(0))


def three_nn_weights_wrapper(b, n, m, unknown, known, dist2, idx, weight):
    _call("three_nn_weights", LIB.epnet_three_nn_weights, unknown, b, n, m, _f(unknown, "unknown"), _f(known, "known"),
          _f(dist2, "dist2"), _i(idx, "idx"), _f(weight, "weight"))


def grid_gather_pm_wrapper(b, c, h, w, n, fmap, xy, align_corners, out):
    _call("grid_gather_pm", LIB.epnet_grid_gather_pm, fmap, b, c, h, w, n, _f(fmap, "feature_map"), _f(xy, "xy"),
          int(bool(align_corners)), out.data_ptr(), out.stride(0))


def grid_gather_nhwc_pm_wrapper(b, c, h, w, n, fmap_nhwc, xy, align_corners, out):
    """fmap_nhwc (B,H,W,C') with C' >= c channels-last contiguous; out (B*n, ldo)"""
    _call("grid_gather_nhwc_pm", LIB.epnet_grid_gather_nhwc_pm, fmap_nhwc, b, c, h, w, n, _f(fmap_nhwc, "feature_map"),
          fmap_nhwc.stride(-2), _f(xy, "xy"), int(bool(align_corners)), out.data_ptr(), out.stride(0))


def deconv_shuffle_nhwc_wrapper(b, h, w, k, co, y, out, col_off):
    _call("deconv_shuffle_nhwc", LIB.epnet_deconv_shuffle_nhwc, y, b, h, w, k, co, _f(y, "y"), out.data_ptr(), out.stride(-2), col_off)
