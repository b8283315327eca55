"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (liboracle.so, built from pointnet2_oracle.c) of the reference's pointnet2 CUDA ops and of the
LI-Fusion bilinear gather, exposed on numpy arrays.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / reference legs may import this package; epnet_b200/ never does.

Parity status: pinned against the reference's own kernels (oracle/_ref/libpointnet2_ref.so, built by
oracle/build_ref.sh from the unmodified sources under /root/reference) run on a B200 -- see
tests/test_ref_pin.py and tests/golden/.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_LIB_PATH = os.path.join(HERE, "_ref", "libpointnet2_ref.so")


def build(force=False):
    """Compile liboracle.so and (only where /root/reference exists) oracle/_ref/libpointnet2_ref.so."""
    src = os.path.join(HERE, "pointnet2_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "liboracle.so"] + (["-B"] if force else []), check=True,
                       capture_output=True)
    shim = os.path.join(HERE, "ref_shim.cu")
    if os.path.isdir("/root/reference") and (force or not os.path.exists(REF_LIB_PATH)
                                             or os.path.getmtime(REF_LIB_PATH) < os.path.getmtime(shim)):
        subprocess.run([os.path.join(HERE, "build_ref.sh")], check=True, capture_output=True)


build()
_lib = ctypes.CDLL(LIB_PATH)

_F = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_I = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i, _f = ctypes.c_int, ctypes.c_float


def _sig(name, *argtypes):
    fn = getattr(_lib, name)
    fn.argtypes = list(argtypes)
    fn.restype = None
    return fn


_fps = _sig("oracle_furthest_point_sampling", _i, _i, _i, _F, _F, _I)
_gather = _sig("oracle_gather_points", _i, _i, _i, _i, _F, _I, _F)
_gather_grad = _sig("oracle_gather_points_grad", _i, _i, _i, _i, _F, _I, _F)
_ball = _sig("oracle_ball_query", _i, _i, _i, _f, _i, _F, _F, _I)
_group = _sig("oracle_group_points", _i, _i, _i, _i, _i, _F, _I, _F)
_group_grad = _sig("oracle_group_points_grad", _i, _i, _i, _i, _i, _F, _I, _F)
_nn = _sig("oracle_three_nn", _i, _i, _i, _F, _F, _F, _I)
_interp = _sig("oracle_three_interpolate", _i, _i, _i, _i, _F, _I, _F, _F)
_interp_grad = _sig("oracle_three_interpolate_grad", _i, _i, _i, _i, _F, _I, _F, _F)
_grid = _sig("oracle_grid_gather_bilinear", _i, _i, _i, _i, _i, _F, _F, _i, _F)
_grid_grad = _sig("oracle_grid_gather_bilinear_grad", _i, _i, _i, _i, _i, _F, _F, _i, _F)
_roi = _sig("oracle_roipool3d", _i, _i, _i, _i, _i, _F, _F, _F, _F, _I)
_pair = _sig("oracle_boxes_pairwise_bev", _i, _i, _F, _i, _F, _F)
_lib.oracle_nms_bev.argtypes = [_i, _i, _F, _f, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")]
_lib.oracle_nms_bev.restype = _i
_lib.oracle_fps_block_size.argtypes = [_i]
_lib.oracle_fps_block_size.restype = _i


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def fps_block_size(n):
    return _lib.oracle_fps_block_size(int(n))


def furthest_point_sampling(xyz, npoint, temp=None, return_temp=False):
    """xyz (B,N,3) -> idx (B,npoint) int32.  temp defaults to the 1e10 fill of pointnet2_utils.py:26."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    temp = np.full((B, N), 1e10, dtype=np.float32) if temp is None else _f32(temp).copy()
    idx = np.zeros((B, npoint), dtype=np.int32)
    _fps(B, N, npoint, xyz, temp, idx)
    return (idx, temp) if return_temp else idx


def gather_points(points, idx):
    points, idx = _f32(points), _i32(idx)
    B, C, N = points.shape
    M = idx.shape[1]
    out = np.empty((B, C, M), dtype=np.float32)
    _gather(B, C, N, M, points, idx, out)
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, M = grad_out.shape
    g = np.zeros((B, C, n), dtype=np.float32)
    _gather_grad(B, C, n, M, grad_out, idx, g)
    return g


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), dtype=np.int32)
    _ball(B, N, M, float(radius), nsample, new_xyz, xyz, idx)
    return idx


def group_points(points, idx):
    points, idx = _f32(points), _i32(idx)
    B, C, N = points.shape
    _, M, ns = idx.shape
    out = np.empty((B, C, M, ns), dtype=np.float32)
    _group(B, C, N, M, ns, points, idx, out)
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, M, ns = grad_out.shape
    g = np.zeros((B, C, n), dtype=np.float32)
    _group_grad(B, C, n, M, ns, grad_out, idx, g)
    return g


def three_nn(unknown, known):
    """-> (dist2 (B,n,3) SQUARED distances, idx (B,n,3))"""
    unknown, known = _f32(unknown), _f32(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = np.empty((B, n, 3), dtype=np.float32)
    idx = np.empty((B, n, 3), dtype=np.int32)
    _nn(B, n, m, unknown, known, dist2, idx)
    return dist2, idx


def three_interpolate(points, idx, weight):
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    B, C, m = points.shape
    n = idx.shape[1]
    out = np.empty((B, C, n), dtype=np.float32)
    _interp(B, C, m, n, points, idx, weight, out)
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    B, C, n = grad_out.shape
    g = np.zeros((B, C, m), dtype=np.float32)
    _interp_grad(B, C, n, m, grad_out, idx, weight, g)
    return g


def grid_gather_bilinear(fmap, xy, align_corners=False):
    fmap, xy = _f32(fmap), _f32(xy)
    B, C, H, W = fmap.shape
    N = xy.shape[1]
    out = np.empty((B, C, N), dtype=np.float32)
    _grid(B, C, H, W, N, fmap, xy, int(bool(align_corners)), out)
    return out


def grid_gather_bilinear_grad(grad_out, xy, h, w, align_corners=False):
    grad_out, xy = _f32(grad_out), _f32(xy)
    B, C, N = grad_out.shape
    g = np.zeros((B, C, h, w), dtype=np.float32)
    _grid_grad(B, C, h, w, N, grad_out, xy, int(bool(align_corners)), g)
    return g


def roipool3d(pts, pts_feature, boxes3d, sampled=512):
    """boxes3d already enlarged.  -> pooled_features (B,M,sampled,3+C), pooled_empty_flag (B,M)"""
    pts, pts_feature, boxes3d = _f32(pts), _f32(pts_feature), _f32(boxes3d)
    B, N, _ = pts.shape
    M, C = boxes3d.shape[1], pts_feature.shape[2]
    out = np.zeros((B, M, sampled, 3 + C), dtype=np.float32)
    flag = np.zeros((B, M), dtype=np.int32)
    _roi(B, N, M, C, sampled, pts, boxes3d, pts_feature, out, flag)
    return out, flag


def boxes_overlap_bev(boxes_a, boxes_b):
    """(M,5), (N,5) [x1,y1,x2,y2,ry] -> (M,N) intersection areas (iou3d_kernel.cu box_overlap)."""
    return _pairwise(0, boxes_a, boxes_b)


def boxes_iou_bev(boxes_a, boxes_b):
    return _pairwise(1, boxes_a, boxes_b)


def boxes_iou_normal(boxes_a, boxes_b):
    return _pairwise(2, boxes_a, boxes_b)


def _pairwise(mode, boxes_a, boxes_b):
    boxes_a, boxes_b = _f32(boxes_a).reshape(-1, 5), _f32(boxes_b).reshape(-1, 5)
    out = np.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=np.float32)
    _pair(mode, boxes_a.shape[0], boxes_a, boxes_b.shape[0], boxes_b, out)
    return out


def nms_bev(boxes, thresh, rotated=True):
    """boxes (N,5) sorted by descending score -> kept indices (int64), greedy as iou3d.cpp:100-113."""
    boxes = _f32(boxes).reshape(-1, 5)
    keep = np.zeros(max(boxes.shape[0], 1), dtype=np.int64)
    k = _lib.oracle_nms_bev(int(bool(rotated)), boxes.shape[0], boxes, float(thresh), keep)
    return keep[:k].copy()


def fps_tie_rule_bruteforce(xyz, npoint):
    """Independent second statement of FPS used to cross-check the thread/tree emulation above: per
    iteration take the maximum running distance and, among exact ties, the index minimising
    (bitreverse_L(k mod BS), k)  (SURVEY.md section 8 a1).  Pure numpy; small inputs only."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    bs = fps_block_size(N)
    L = bs.bit_length() - 1
    k = np.arange(N, dtype=np.int64)
    slot = k % bs
    rev = np.zeros(N, dtype=np.int64)
    for bit in range(L):
        rev |= ((slot >> bit) & 1) << (L - 1 - bit)
    order = rev * (N + 1) + k  # smaller = preferred among ties
    out = np.zeros((B, npoint), dtype=np.int32)
    for s in range(B):
        p = xyz[s]
        temp = np.full(N, 1e10, dtype=np.float32)
        last = 0
        for j in range(1, npoint):
            d = (p - p[last]).astype(np.float32)
            # fma(dz,dz, fma(dx,dx, dy*dy)) evaluated in float64 then rounded once per step
            t = (d[:, 1].astype(np.float64) * d[:, 1].astype(np.float64)).astype(np.float32)
            t = (d[:, 0].astype(np.float64) * d[:, 0].astype(np.float64) + t.astype(np.float64)).astype(np.float32)
            t = (d[:, 2].astype(np.float64) * d[:, 2].astype(np.float64) + t.astype(np.float64)).astype(np.float32)
            temp = np.minimum(t, temp)
            mx = temp.max()
            tied = np.nonzero(temp == mx)[0]
            last = int(tied[np.argmin(order[tied])])
            out[s, j] = last
    return out


def image_prep(img_u8, hw=(384, 1280), mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """The reference's host-side image preparation, restated: lib/datasets/kitti_dataset.py:44-55 (float64: /255, -mean, /std,
    zero-padded canvas) followed by the cast and permute of lib/net/train_functions.py:37 -> (B,3,H,W) float32.
    Parity unpinned by execution: the reference function itself needs PIL and `np.float` (gone in numpy 2.x); the restatement is
    the same four numpy statements on the same dtype.  img_u8 (B,h,w,3) uint8, or a list of (h_i,w_i,3) arrays."""
    imgs = list(img_u8)
    out = np.zeros([len(imgs), hw[0], hw[1], 3], dtype=np.float64)
    for i, im in enumerate(imgs):
        im = np.asarray(im).astype(np.float64)
        im = im / 255.0
        im -= np.asarray(mean)
        im /= np.asarray(std)
        out[i, :im.shape[0], :im.shape[1], :] = im
    return np.ascontiguousarray(out.astype(np.float32).transpose(0, 3, 1, 2))
