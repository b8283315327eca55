"""ctypes binding of libepnet_b200.so -- the only door from Python into the CUDA kernels.

There is deliberately no fallback: if the library is missing and cannot be built, importing raises.
"""
import ctypes
import os

from . import build as _build

_c_int, _c_float, _c_void_p, _c_ll = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_longlong

# name -> argtypes, exactly the declarations of include/epnet_b200.h
SIGNATURES = {
    "epnet_furthest_point_sampling": [_c_int] * 3 + [_c_void_p] * 4,
    "epnet_gather_points": [_c_int] * 4 + [_c_void_p] * 4,
    "epnet_gather_points_grad": [_c_int] * 4 + [_c_void_p] * 4,
    "epnet_ball_query": [_c_int] * 3 + [_c_float, _c_int] + [_c_void_p] * 4,
    "epnet_bucket_cloud": [_c_int] * 3 + [_c_void_p] * 4,
    "epnet_ball_query_sorted": [_c_int] * 3 + [_c_float, _c_int] + [_c_void_p] * 5,
    "epnet_group_points": [_c_int] * 5 + [_c_void_p] * 4,
    "epnet_group_points_grad": [_c_int] * 5 + [_c_void_p] * 4,
    "epnet_three_nn": [_c_int] * 3 + [_c_void_p] * 5,
    "epnet_three_interpolate": [_c_int] * 4 + [_c_void_p] * 5,
    "epnet_three_interpolate_grad": [_c_int] * 4 + [_c_void_p] * 5,
    "epnet_grid_gather_bilinear": [_c_int] * 5 + [_c_void_p] * 2 + [_c_int] + [_c_void_p] * 2,
    "epnet_grid_gather_bilinear_grad": [_c_int] * 5 + [_c_void_p] * 2 + [_c_int] + [_c_void_p] * 2,
    "epnet_fps_sample": [_c_int] * 3 + [_c_void_p] * 6 + [_c_int, _c_void_p],
    "epnet_fps_prefix_check": [_c_int] * 3 + [_c_void_p] * 4,
    "epnet_sa_first_level": [_c_int] * 7 + [_c_void_p] * 5 + [_c_int, _c_void_p],
    "epnet_fps_sample_guarded": [_c_int] * 3 + [_c_void_p] * 6 + [_c_int, _c_void_p, _c_void_p],
    "epnet_group_concat": [_c_int] * 5 + [_c_void_p] * 6,
    "epnet_attention_scale_pm": [_c_int] * 3 + [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int,
                                 _c_void_p],
    "epnet_bias_relu": [_c_int, _c_int, _c_ll, _c_void_p, _c_void_p, _c_void_p],
    "epnet_bias_relu_maxpool": [_c_int] * 4 + [_c_void_p] * 3 + [_c_ll, _c_void_p],
    "epnet_three_interpolate_concat": [_c_int] * 5 + [_c_void_p] * 6,
    "epnet_group_concat_pm": [_c_int] * 5 + [_c_void_p] * 3 + [_c_int, _c_void_p, _c_void_p, _c_int, _c_void_p],
    "epnet_three_interpolate_concat_pm": [_c_int] * 5 + [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int,
                                          _c_void_p],
    "epnet_three_nn_weights": [_c_int] * 3 + [_c_void_p] * 6,
    "epnet_grid_gather_pm": [_c_int] * 5 + [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p],
    "epnet_gemm_tf32x3_grouped": [_c_int] * 5 + [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_int,
                                  _c_int, _c_void_p, _c_int, _c_void_p],
    "epnet_gemm_tf32x3_cm": [_c_int] * 4 + [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p],
    "epnet_conv3x3_nhwc_tf32x3": [_c_int] * 6 + [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p],
    "epnet_conv3x3_nhwc_f16x3": [_c_int] * 6 + [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p],
    "epnet_grid_gather_nhwc_pm": [_c_int] * 5 + [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p],
    "epnet_deconv_nhwc_tf32x3": [_c_int] * 6 + [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p],
    "epnet_deconv_nhwc_f16x3": [_c_int] * 6 + [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p],
    "epnet_deconv_shuffle_nhwc": [_c_int] * 5 + [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p],
    "epnet_roipool3d": [_c_int] * 5 + [_c_void_p] * 6,
    "epnet_boxes_overlap_bev": [_c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p],
    "epnet_boxes_iou_bev": [_c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p],
    "epnet_nms_workspace_bytes": [_c_int, _c_int, ctypes.POINTER(ctypes.c_ulonglong)],
    "epnet_nms_rotated": [_c_int, _c_int, _c_void_p, _c_void_p, _c_float, _c_int] + [_c_void_p] * 4,
    "epnet_nms_normal": [_c_int, _c_int, _c_void_p, _c_void_p, _c_float, _c_int] + [_c_void_p] * 4,
    "epnet_gemm_tf32x3": [_c_int] * 3 + [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_void_p],
    "epnet_conv3x3_planes_tma": [_c_int] * 6 + [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p,
                                 _c_int, _c_void_p],
    "epnet_gemm_planes_tma": [_c_int] * 3 + [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_void_p,
                              _c_void_p, _c_int, _c_void_p],
    "epnet_deconv_planes_tma": [_c_int] * 6 + [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p],
    "epnet_conv3x3_nhwc_tf32x3_planes": [_c_int] * 6 + [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_int,
                                         _c_void_p],
    "epnet_tail_taps": [_c_int] * 5 + [_c_void_p] * 5,
    "epnet_tail_plan": [_c_void_p] * 5 + [_c_int, _c_void_p],
    "epnet_tail_scatter": [_c_int] * 4 + [_c_void_p] * 8 + [_c_int, _c_void_p],
    "epnet_tail_blend": [_c_int, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p],
    "epnet_gemm_tf32x3_rows": [_c_int] * 3 + [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int,
                               _c_void_p, _c_int, _c_void_p],
    "epnet_conv3x3_c3_planes": [_c_int] * 4 + [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_void_p],
    "epnet_image_prep_u8": [_c_int] * 3 + [_c_ll, _c_ll, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p],
    "epnet_image_nchw_to_nhwc4": [_c_int] * 3 + [_c_void_p, _c_void_p, _c_void_p],
    "epnet_gemm_overflow_read": [_c_void_p, _c_void_p],
    "epnet_gemm_overflow_reset": [_c_void_p],
    "epnet_gemm_f16x3": [_c_int] * 3 + [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_void_p],
}


def _load():
    path = _build.LIB
    if _build.stale():
        try:
            _build.build()
        except Exception as exc:  # noqa: BLE001 - re-raised below when the library is unusable
            if not os.path.exists(path):
                raise ImportError(
                    "epnet_b200: libepnet_b200.so is missing and could not be built (%s). "
                    "Run `python -m epnet_b200.build` where nvcc is available." % exc)
    lib = ctypes.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync: fail loudly
        fn.argtypes = argtypes
        fn.restype = _c_int
    lib.epnet_abi_version.restype = _c_int
    lib.epnet_error_string.argtypes = [_c_int]
    lib.epnet_error_string.restype = ctypes.c_char_p
    return lib


LIB = _load()
LIB_PATH = _build.LIB


class EpnetKernelError(RuntimeError):
    pass


def check(code, what):
    if code != 0:
        raise EpnetKernelError("%s failed: %s" % (what, LIB.epnet_error_string(code).decode()))
