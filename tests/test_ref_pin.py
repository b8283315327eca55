"""Pins the oracle: the C restatement (oracle/pointnet2_oracle.c) must reproduce, bit for bit, what the
reference's OWN unmodified kernels (oracle/_ref/libpointnet2_ref.so, compiled for sm_100a from the sources
under /root/reference by oracle/build_ref.sh) compute on a B200.  Also checks the product kernels against
the reference kernels directly, without the oracle in between."""
import numpy as np
import pytest

import oracle
from cases import cloud, lidar
from gpu_util import OpRunner

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_cuda
    if not ref_cuda.available():
        pytest.skip("oracle/_ref/libpointnet2_ref.so not built")
    return OpRunner(ref_cuda)


@pytest.fixture(scope="module")
def ours():
    from epnet_b200 import pointnet2_cuda
    return OpRunner(pointnet2_cuda)


@pytest.mark.parametrize("kind,b,n,m", [("gauss", 2, 3, 3), ("gauss", 2, 257, 64), ("lattice", 2, 300, 150),
                                        ("lattice", 2, 2100, 300), ("identical", 2, 70, 20), ("uniform", 2, 1500, 300),
                                        ("lattice", 1, 16384, 500)])
def test_oracle_fps_equals_reference_kernel(ref, ours, kind, b, n, m):
    xyz = cloud(11, b, n, kind, dup_frac=0.05 if n > 100 else 0.0)
    r, rt = ref.fps(xyz, m, return_temp=True)
    o, ot = oracle.furthest_point_sampling(xyz, m, return_temp=True)
    np.testing.assert_array_equal(o, r)
    np.testing.assert_array_equal(ot, rt)
    np.testing.assert_array_equal(ours.fps(xyz, m), r)


def test_fps_lidar_full_size_three_way(ref, ours):
    xyz = lidar(1000, 2)
    r = ref.fps(xyz, 4096)
    np.testing.assert_array_equal(oracle.furthest_point_sampling(xyz, 4096), r)
    np.testing.assert_array_equal(ours.fps(xyz, 4096), r)


@pytest.mark.parametrize("n,m,radius,nsample", [(700, 100, 0.3, 16), (700, 100, 1e-4, 8), (4096, 1024, 0.4, 32)])
def test_oracle_ball_query_equals_reference_kernel(ref, ours, n, m, radius, nsample):
    xyz = cloud(21, 2, n, "gauss", dup_frac=0.05)
    new_xyz = np.ascontiguousarray(xyz[:, :m] + (0 if radius > 1e-3 else 1.0))
    r = ref.ball_query(radius, nsample, xyz, new_xyz)
    np.testing.assert_array_equal(oracle.ball_query(radius, nsample, xyz, new_xyz), r)
    np.testing.assert_array_equal(ours.ball_query(radius, nsample, xyz, new_xyz), r)


def test_ball_query_lidar_full_size_three_way(ref, ours):
    xyz = lidar(1002, 1)
    fps = ref.fps(xyz, 4096)
    new_xyz = np.stack([xyz[0][fps[0]]])
    for radius, ns in ((0.1, 16), (0.5, 32)):
        r = ref.ball_query(radius, ns, xyz, new_xyz)
        np.testing.assert_array_equal(oracle.ball_query(radius, ns, xyz, new_xyz), r)
        np.testing.assert_array_equal(ours.ball_query(radius, ns, xyz, new_xyz), r)


@pytest.mark.parametrize("kind,n,m", [("gauss", 1000, 250), ("lattice", 999, 130), ("gauss", 50, 2), ("gauss", 16384, 4096)])
def test_oracle_three_nn_equals_reference_kernel(ref, ours, kind, n, m):
    unknown, known = cloud(31, 2, n, kind), cloud(32, 2, m, kind)
    rd, ri = ref.three_nn(unknown, known)
    od, oi = oracle.three_nn(unknown, known)
    np.testing.assert_array_equal(oi, ri)
    np.testing.assert_array_equal(od, rd)
    gd, gi = ours.three_nn(unknown, known)
    np.testing.assert_array_equal(gi, ri)
    np.testing.assert_array_equal(gd, rd)


def test_oracle_float_ops_equal_reference_kernels(ref, ours):
    rng = np.random.RandomState(5)
    pts = rng.randn(2, 24, 500).astype(np.float32)
    idx = rng.randint(0, 500, size=(2, 128, 16)).astype(np.int32)
    np.testing.assert_array_equal(oracle.group_points(pts, idx), ref.group(pts, idx))
    np.testing.assert_array_equal(ours.group(pts, idx), ref.group(pts, idx))
    g1 = np.ascontiguousarray(idx[:, :, 0])
    np.testing.assert_array_equal(oracle.gather_points(pts, g1), ref.gather(pts, g1))
    w = rng.rand(2, 128, 3).astype(np.float32)
    i3 = np.ascontiguousarray(idx[:, :, :3])
    r = ref.three_interpolate(pts, i3, w)
    np.testing.assert_array_equal(oracle.three_interpolate(pts, i3, w), r)  # same FMA order -> same bits
    np.testing.assert_array_equal(ours.three_interpolate(pts, i3, w), r)
    go = rng.randn(2, 24, 128, 16).astype(np.float32)
    np.testing.assert_allclose(oracle.group_points_grad(go, idx, 500), ref.group_grad(go, idx, 500), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(ours.group_grad(go, idx, 500), ref.group_grad(go, idx, 500), rtol=1e-4, atol=1e-4)
    go3 = np.ascontiguousarray(go[..., 0])
    np.testing.assert_allclose(oracle.three_interpolate_grad(go3, i3, w, 500), ref.three_interpolate_grad(go3, i3, w, 500),
                               rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(ours.three_interpolate_grad(go3, i3, w, 500), ref.three_interpolate_grad(go3, i3, w, 500),
                               rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(oracle.gather_points_grad(go3, g1, 500), ref.gather_grad(go3, g1, 500), rtol=1e-4, atol=1e-4)
