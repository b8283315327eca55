"""Parity of the B200 kernels (through the C ABI, via epnet_b200.pointnet2_cuda) against the oracle.
Indices are compared bit-exactly; float outputs to rtol 1e-5 (BASELINE.json north_star); gradients, which
the reference accumulates with unordered fp32 atomics, to rtol 1e-4 / atol 1e-5."""
import numpy as np
import pytest
import torch

import oracle
from cases import cloud, lidar
from gpu_util import OpRunner, dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ours():
    from epnet_b200 import pointnet2_cuda
    return OpRunner(pointnet2_cuda)


# ------------------------------------------------------------------ FPS
@pytest.mark.parametrize("kind,b,n,m", [
    ("gauss", 2, 1, 1), ("gauss", 2, 3, 3), ("gauss", 3, 31, 31), ("gauss", 2, 33, 20), ("gauss", 2, 257, 64),
    ("lattice", 2, 300, 150), ("lattice", 2, 1024, 256), ("lattice", 1, 2100, 300), ("identical", 2, 70, 20),
    ("uniform", 2, 1500, 300), ("gauss", 2, 4096, 1024), ("lattice", 1, 5000, 700), ("uniform", 1, 12000, 500),
    ("lattice", 1, 16384, 600), ("gauss", 1, 20000, 300),
    # 16384 < N <= 131072: thread-block cluster of 2..8 CTAs exchanging candidates through DSMEM
    ("lattice", 1, 40000, 400), ("uniform", 2, 65536, 300), ("gauss", 1, 131072, 200), ("lattice", 2, 16385, 100),
    # beyond the cluster's capacity: streaming kernel
    ("uniform", 1, 140000, 40),
])
def test_fps_bit_exact(ours, kind, b, n, m):
    xyz = cloud(11, b, n, kind, dup_frac=0.05 if n > 100 else 0.0)
    want, want_t = oracle.furthest_point_sampling(xyz, m, return_temp=True)
    got, got_t = ours.fps(xyz, m, return_temp=True)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(got_t, want_t)  # the in/out scratch ends up identical too


def test_fps_backbone_shapes_lidar(ours):
    xyz = lidar(1000, 2)  # 2 % exact duplicates, KITTI-shaped
    cur = xyz
    for m in (4096, 1024, 256, 64):
        want = oracle.furthest_point_sampling(cur, m)
        got = ours.fps(cur, m)
        np.testing.assert_array_equal(got, want)
        cur = np.stack([cur[s][want[s]] for s in range(cur.shape[0])])


def test_fps_honours_caller_temp(ours):
    xyz = cloud(12, 2, 900, "gauss")
    temp = np.random.RandomState(0).rand(2, 900).astype(np.float32) * 3
    want = oracle.furthest_point_sampling(xyz, 100, temp=temp)
    np.testing.assert_array_equal(ours.fps(xyz, 100, temp=temp), want)


def test_fps_m_exceeds_distinct_points(ours):
    xyz = cloud(13, 2, 64, "lattice")[:, :, :] * 0 + cloud(13, 2, 64, "lattice")[:, :8].repeat(8, axis=1)  # 8 distinct
    want = oracle.furthest_point_sampling(xyz, 40)
    np.testing.assert_array_equal(ours.fps(xyz, 40), want)


# ------------------------------------------------------------------ ball query
@pytest.mark.parametrize("n,m,radius,nsample", [
    (700, 100, 0.3, 16), (700, 100, 1.0, 32), (700, 100, 1e-4, 8), (700, 100, 100.0, 5), (4096, 1024, 0.4, 16),
    (1999, 333, 0.5, 32), (5000, 70, 0.2, 64), (33, 33, 0.8, 3), (1, 1, 1.0, 4),
])
def test_ball_query_bit_exact(ours, n, m, radius, nsample):
    xyz = cloud(21, 2, n, "gauss", dup_frac=0.05 if n > 100 else 0)
    new_xyz = np.ascontiguousarray(xyz[:, :m] + (0 if radius > 1e-3 else 1.0))
    want = oracle.ball_query(radius, nsample, xyz, new_xyz)
    np.testing.assert_array_equal(ours.ball_query(radius, nsample, xyz, new_xyz), want)


def test_ball_query_backbone_sa1_lidar(ours):
    xyz = lidar(1001, 2)
    fps = oracle.furthest_point_sampling(xyz, 4096)
    new_xyz = np.stack([xyz[s][fps[s]] for s in range(2)])
    for radius, ns in ((0.1, 16), (0.5, 32)):
        want = oracle.ball_query(radius, ns, xyz, new_xyz)
        np.testing.assert_array_equal(ours.ball_query(radius, ns, xyz, new_xyz), want)


def test_ball_query_unaligned_scene_base(ours):
    # n % 4 != 0 makes scene 1's base pointer 4-byte aligned only: exercises the non-bulk staging path
    xyz = cloud(22, 3, 1001, "uniform")
    new_xyz = np.ascontiguousarray(xyz[:, ::5])
    want = oracle.ball_query(6.0, 16, xyz, new_xyz)
    np.testing.assert_array_equal(ours.ball_query(6.0, 16, xyz, new_xyz), want)


# ------------------------------------------------------------------ three_nn / interpolate
@pytest.mark.parametrize("kind,n,m", [("gauss", 1000, 250), ("lattice", 999, 130), ("gauss", 4096, 1024), ("gauss", 50, 2),
                                      ("gauss", 7, 1), ("uniform", 3000, 4001), ("lattice", 16384, 4096)])
def test_three_nn_bit_exact(ours, kind, n, m):
    unknown = cloud(31, 2, n, kind)
    known = cloud(32, 2, m, kind)
    want_d, want_i = oracle.three_nn(unknown, known)
    got_d, got_i = ours.three_nn(unknown, known)
    np.testing.assert_array_equal(got_i, want_i)
    np.testing.assert_array_equal(got_d, want_d)  # same FMA contraction -> identical bits, inf included


@pytest.mark.parametrize("c,m,n", [(8, 64, 256), (3, 10, 17), (256, 4096, 16384), (33, 100, 1001),
                                   # rows staged in shared memory: ragged last row group; one 128 KB row per CTA; partly staged 240 KB row
                                   (70, 1024, 16384), (9, 32768, 131072), (5, 60000, 240000),
                                   # through the point-major scratch copy: ragged channel quad and ragged last tile
                                   (45, 50000, 100003)])
def test_three_interpolate(ours, c, m, n):
    rng = np.random.RandomState(41)
    pts = rng.randn(2, c, m).astype(np.float32)
    idx = rng.randint(0, m, size=(2, n, 3)).astype(np.int32)
    w = rng.rand(2, n, 3).astype(np.float32)
    w /= w.sum(-1, keepdims=True)
    want = oracle.three_interpolate(pts, idx, w)
    got = ours.three_interpolate(pts, idx, w)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-7)
    np.testing.assert_array_equal(got, want)  # stronger: the reference's FMA order is reproduced
    go = rng.randn(2, c, n).astype(np.float32)
    np.testing.assert_allclose(ours.three_interpolate_grad(go, idx, w, m), oracle.three_interpolate_grad(go, idx, w, m),
                               rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------ gather / group
@pytest.mark.parametrize("c,n,m,ns", [(3, 100, 30, 16), (96, 4096, 1024, 32), (5, 77, 13, 3), (64, 16384, 4096, 16),
                                      # staged rows: ragged row group; 160 KB rows; 256 KB rows below / above the work threshold (two parts); staged
                                      # gather_points; 1 MB rows in five parts; gather_points with fewer outputs than row elements; rows too long to stage
                                      (40, 4096, 1024, 32), (20, 40000, 4096, 32), (7, 65536, 2048, 64), (9, 65536, 2048, 64), (300, 1024, 4096, 16),
                                      (5, 250000, 4096, 64), (300, 16384, 4096, 16), (3, 600000, 8192, 64),
                                      # rows longer than shared memory with enough work: through the point-major scratch copy (one / two channel chunks)
                                      (20, 65536, 4096, 32), (133, 60000, 2048, 32)])
def test_group_and_gather(ours, c, n, m, ns):
    rng = np.random.RandomState(51)
    pts = rng.randn(2, c, n).astype(np.float32)
    idx = rng.randint(0, n, size=(2, m, ns)).astype(np.int32)
    idx[:, :, ns // 2:] = idx[:, :, :1]  # ball-query style padding: many repeats of one index
    np.testing.assert_array_equal(ours.group(pts, idx), oracle.group_points(pts, idx))
    gidx = np.ascontiguousarray(idx[:, :, 0])
    np.testing.assert_array_equal(ours.gather(pts, gidx), oracle.gather_points(pts, gidx))
    go = rng.randn(2, c, m, ns).astype(np.float32)
    np.testing.assert_allclose(ours.group_grad(go, idx, n), oracle.group_points_grad(go, idx, n), rtol=1e-4, atol=1e-4)
    go1 = np.ascontiguousarray(go[..., 0])
    np.testing.assert_allclose(ours.gather_grad(go1, gidx, n), oracle.gather_points_grad(go1, gidx, n), rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------ LI-Fusion gather
@pytest.mark.parametrize("align_corners", [False, True])
@pytest.mark.parametrize("c,h,w,n", [(6, 12, 20, 200), (64, 192, 640, 4096), (32, 384, 1280, 16384), (5, 7, 9, 33)])
def test_grid_gather_vs_oracle_and_aten(c, h, w, n, align_corners):
    from epnet_b200 import li_fusion
    rng = np.random.RandomState(61)
    fmap = rng.randn(2, c, h, w).astype(np.float32)
    xy = (rng.rand(2, n, 2).astype(np.float32) * 2.3 - 1.15)  # some points outside the image
    xy[0, 0] = [-1, -1]; xy[0, 1] = [1, 1]; xy[0, 2] = [0, 0]
    f = dev(fmap).requires_grad_(True)
    g = dev(xy)
    out = li_fusion.grid_sample(f, g.unsqueeze(1), align_corners=align_corners)
    assert out.shape == (2, c, 1, n)
    want = oracle.grid_gather_bilinear(fmap, xy, align_corners)
    np.testing.assert_allclose(out.detach().squeeze(2).cpu().numpy(), want, rtol=1e-5, atol=1e-6)
    f2 = dev(fmap).requires_grad_(True)
    aten = torch.nn.functional.grid_sample(f2, g.unsqueeze(1), mode="bilinear", padding_mode="zeros", align_corners=align_corners)
    torch.testing.assert_close(out, aten, rtol=1e-5, atol=1e-6)
    go = torch.randn_like(out)
    out.backward(go)
    aten.backward(go)
    torch.testing.assert_close(f.grad, f2.grad, rtol=1e-4, atol=1e-5)


def test_grid_sample_refuses_what_it_does_not_implement():
    from epnet_b200 import li_fusion
    f = torch.zeros(1, 1, 4, 4, device="cuda")
    with pytest.raises(NotImplementedError):
        li_fusion.grid_sample(f, torch.zeros(1, 2, 3, 2, device="cuda"))
    with pytest.raises(NotImplementedError):
        li_fusion.grid_sample(f, torch.zeros(1, 1, 3, 2, device="cuda"), mode="nearest")


def test_cpu_tensors_are_refused():
    from epnet_b200 import pointnet2_cuda
    x = torch.zeros(1, 8, 3)
    with pytest.raises(ValueError):
        pointnet2_cuda.furthest_point_sampling_wrapper(1, 8, 2, x, torch.zeros(1, 8), torch.zeros(1, 2, dtype=torch.int32))


# ------------------------------------------------------------------ ball query through sorted buckets
@pytest.mark.parametrize("kind,n,m,radius,nsample", [
    ("uniform", 2048, 512, 0.15, 16), ("gauss", 4096, 1024, 0.5, 32), ("lattice", 5000, 700, 0.3, 64), ("identical", 300, 50, 0.1, 16),
    ("uniform", 16384, 4096, 0.02, 16), ("gauss", 16384, 1000, 100.0, 32), ("gauss", 3001, 333, 0.4, 1), ("uniform", 64, 64, 0.5, 7),
])
def test_ball_query_sorted_bit_exact(kind, n, m, radius, nsample):
    """The sorted-bucket search must return exactly what the exhaustive scan returns: duplicates, empty balls (tiny radius),
    balls that hold the whole cloud (huge radius), nsample 1..64, n not a power of two."""
    from epnet_b200 import pointnet2_cuda as pc
    xyz = cloud(21, 2, n, kind, dup_frac=0.05 if n > 100 else 0.0)
    new_xyz = np.ascontiguousarray(xyz[:, :m] + (0.01 if kind == "lattice" else 0.0)).astype(np.float32)
    want = oracle.ball_query(radius, nsample, xyz, new_xyz)
    x, q = dev(xyz), dev(new_xyz)
    idx = torch.zeros((2, m, nsample), dtype=torch.int32, device="cuda")
    buckets = pc.bucket_cloud(x)
    srt = buckets[0].cpu().numpy()
    orig = srt[..., 3].view(np.int32)
    for s in range(2):  # the sorted copy is a permutation of the cloud plus padding
        valid = orig[s] >= 0
        assert sorted(orig[s][valid].tolist()) == list(range(n))
        np.testing.assert_array_equal(srt[s][valid][:, :3], xyz[s][orig[s][valid]])
    pc.ball_query_sorted_wrapper(2, m, radius, nsample, q, buckets, idx)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(idx.cpu().numpy(), want)
