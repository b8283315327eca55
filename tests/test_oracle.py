"""CPU tests of the oracle itself (no GPU): the C restatement against independent numpy statements of
the same rules, against torch's CPU grid_sample, and on the edge cases the GPU parity tests rely on."""
import math

import numpy as np
import pytest
import torch

import oracle
from cases import cloud


def test_fps_block_size_matches_reference_log_formula():
    # cuda_utils.h:10-14 computes 2^int(log(n)/log(2)) capped to [1,1024] in double arithmetic
    for n in list(range(1, 5000)) + [8191, 8192, 16383, 16384, 65536, 131072, 199999]:
        pow2 = int(math.log(float(n)) / math.log(2.0))
        assert oracle.fps_block_size(n) == max(min(1 << pow2, 1024), 1), n


@pytest.mark.parametrize("kind,n,m", [("gauss", 257, 64), ("lattice", 300, 120), ("lattice", 2100, 200),
                                      ("identical", 70, 20), ("gauss", 1, 1), ("gauss", 3, 3), ("uniform", 1500, 300)])
def test_fps_thread_tree_emulation_equals_tie_rule(kind, n, m):
    xyz = cloud(3, 2, n, kind, dup_frac=0.1 if n > 50 else 0.0)
    got = oracle.furthest_point_sampling(xyz, m)
    want = oracle.fps_tie_rule_bruteforce(xyz, m)
    np.testing.assert_array_equal(got, want)


def test_fps_tie_is_not_lowest_index():
    # a tie between k=3 and k=513 at BS=1024 goes to 513 (bit-reversed slot order)
    n = 2048
    xyz = np.zeros((1, n, 3), dtype=np.float32)
    xyz[0, 3] = xyz[0, 513] = [5, 0, 0]
    idx = oracle.furthest_point_sampling(xyz, 2)
    assert idx[0, 1] == 513


def test_fps_temp_is_running_min_distance():
    xyz = cloud(5, 1, 500, "gauss")
    idx, temp = oracle.furthest_point_sampling(xyz, 50, return_temp=True)
    d = ((xyz[0][:, None, :] - xyz[0][idx[0]][None, :, :]) ** 2).sum(-1)[:, :-1].min(1)  # last sample never applied
    np.testing.assert_allclose(temp[0], d, rtol=1e-5, atol=1e-6)
    assert len(set(idx[0].tolist())) == 50


def _ball_query_numpy(radius, nsample, xyz, new_xyz):
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    out = np.zeros((B, M, nsample), dtype=np.int32)
    r2 = np.float32(radius) * np.float32(radius)
    for b in range(B):
        for j in range(M):
            d = new_xyz[b, j] - xyz[b]
            t = (d[:, 1].astype(np.float64) ** 2).astype(np.float32)
            t = (d[:, 0].astype(np.float64) ** 2 + t).astype(np.float32)
            t = (d[:, 2].astype(np.float64) ** 2 + t).astype(np.float32)
            hits = np.nonzero(t < r2)[0][:nsample]
            if len(hits):
                out[b, j, :] = hits[0]
                out[b, j, :len(hits)] = hits
    return out


@pytest.mark.parametrize("radius,nsample", [(0.3, 16), (1.0, 32), (1e-4, 8), (100.0, 5)])
def test_ball_query_order_padding_and_empty(radius, nsample):
    xyz = cloud(7, 2, 700, "gauss", dup_frac=0.05)
    new_xyz = np.ascontiguousarray(xyz[:, ::7] + (0 if radius > 1e-3 else 1.0))
    got = oracle.ball_query(radius, nsample, xyz, new_xyz)
    np.testing.assert_array_equal(got, _ball_query_numpy(radius, nsample, xyz, new_xyz))
    if radius < 1e-3:
        assert (got == 0).all()  # no neighbour anywhere: the caller's zero fill survives


def test_three_nn_lexicographic_ties_and_short_known():
    unknown = cloud(9, 2, 300, "lattice")
    known = cloud(10, 2, 90, "lattice")
    d2, idx = oracle.three_nn(unknown, known)
    full = ((unknown[:, :, None, :] - known[:, None, :, :]) ** 2).sum(-1)
    order = np.argsort(full, axis=2, kind="stable")[:, :, :3]  # stable = lower index first among ties
    np.testing.assert_array_equal(idx, order.astype(np.int32))
    np.testing.assert_array_equal(d2, np.take_along_axis(full, order, 2))
    d2s, idxs = oracle.three_nn(unknown, known[:, :2])  # m < 3: unfilled slot = +inf / index 0
    assert np.isinf(d2s[..., 2]).all() and (idxs[..., 2] == 0).all()


def test_gather_group_interpolate_definitions():
    rng = np.random.RandomState(0)
    pts = rng.randn(2, 5, 40).astype(np.float32)
    idx = rng.randint(0, 40, size=(2, 9, 4)).astype(np.int32)
    g = oracle.group_points(pts, idx)
    for b in range(2):
        np.testing.assert_array_equal(g[b], pts[b][:, idx[b]])
    np.testing.assert_array_equal(oracle.gather_points(pts, idx[:, :, 0]), g[..., 0])
    w = rng.rand(2, 9, 3).astype(np.float32)
    out = oracle.three_interpolate(pts, idx[:, :, :3], w)
    want = (g[..., :3].astype(np.float64) * w[:, None].astype(np.float64)).sum(-1)
    np.testing.assert_allclose(out, want, rtol=1e-6, atol=1e-6)


def test_grads_are_transposes_of_forward():
    rng = np.random.RandomState(1)
    pts = rng.randn(2, 3, 30).astype(np.float32)
    idx = rng.randint(0, 30, size=(2, 8, 5)).astype(np.int32)
    go = rng.randn(2, 3, 8, 5).astype(np.float32)
    lhs = (oracle.group_points(pts, idx).astype(np.float64) * go).sum()
    rhs = (oracle.group_points_grad(go, idx, 30).astype(np.float64) * pts).sum()
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs))
    w = rng.rand(2, 8, 3).astype(np.float32)
    go3 = rng.randn(2, 3, 8).astype(np.float32)
    lhs = (oracle.three_interpolate(pts, idx[:, :, :3], w).astype(np.float64) * go3).sum()
    rhs = (oracle.three_interpolate_grad(go3, idx[:, :, :3], w, 30).astype(np.float64) * pts).sum()
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs))
    lhs = (oracle.gather_points(pts, idx[:, :, 0]).astype(np.float64) * go[..., 0]).sum()
    rhs = (oracle.gather_points_grad(go[..., 0], idx[:, :, 0], 30).astype(np.float64) * pts).sum()
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs))


@pytest.mark.parametrize("align_corners", [False, True])
def test_grid_gather_matches_torch_cpu_grid_sample(align_corners):
    rng = np.random.RandomState(2)
    fmap = rng.randn(2, 6, 12, 20).astype(np.float32)
    xy = (rng.rand(2, 200, 2).astype(np.float32) * 2.4 - 1.2)  # includes points outside [-1,1]
    xy[0, 0] = [-1, -1]; xy[0, 1] = [1, 1]; xy[0, 2] = [0, 0]; xy[0, 3] = [1.0, -1.0]
    got = oracle.grid_gather_bilinear(fmap, xy, align_corners)
    want = torch.nn.functional.grid_sample(torch.from_numpy(fmap), torch.from_numpy(xy).unsqueeze(1), mode="bilinear",
                                           padding_mode="zeros", align_corners=align_corners).squeeze(2).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
    # backward against autograd
    f = torch.from_numpy(fmap).requires_grad_(True)
    out = torch.nn.functional.grid_sample(f, torch.from_numpy(xy).unsqueeze(1), align_corners=align_corners).squeeze(2)
    go = torch.from_numpy(rng.randn(*out.shape).astype(np.float32))
    out.backward(go)
    got_g = oracle.grid_gather_bilinear_grad(go.numpy(), xy, 12, 20, align_corners)
    np.testing.assert_allclose(got_g, f.grad.numpy(), rtol=1e-4, atol=1e-5)


def test_fps_of_an_fps_ordered_cloud_is_the_identity_prefix():
    """Greedy FPS is self-consistent: sampling M2 < M points from the M points an earlier FPS produced (in its output order)
    returns 0, 1, ..., M2-1 whenever the earlier maxima were unique -- the reason every coarser level of the backbone is a
    prefix of the finer one.  (With exact ties the reference's bit-reversal tie rule may pick a different index, so the
    product still runs every level)"""
    from cases import lidar
    xyz = lidar(3, 2, 4096)
    first = oracle.furthest_point_sampling(xyz, 1024)
    coarse = np.stack([xyz[b][first[b]] for b in range(2)])
    again = oracle.furthest_point_sampling(coarse, 256)
    np.testing.assert_array_equal(again, np.broadcast_to(np.arange(256, dtype=np.int32), (2, 256)))


def _prefix_condition(p, m):
    """csrc/fps.cu epnet_fps_prefix_check, restated: with T_j(s) = min(1e10, d(p_j, p_0), ..., d(p_j, p_{s-1})) in the reference's fp32
    arithmetic (the oracle's distance: fma(dz,dz, fma(dx,dx, dy*dy))), the identity 0..m-1 is certified iff T_j(s) < T_s(s) for every
    step 1 <= s < m and every j != s."""
    n = p.shape[0]
    t = np.full(n, np.float32(1e10), dtype=np.float32)
    for s in range(1, m):
        q = p[s - 1]
        dx, dy, dz = (p[:, 0] - q[0]).astype(np.float32), (p[:, 1] - q[1]).astype(np.float32), (p[:, 2] - q[2]).astype(np.float32)
        # float64 products of float32 values are exact, one rounding per fma: the same chain as the kernels
        inner = (dx.astype(np.float64) * dx + (dy * dy).astype(np.float32).astype(np.float64)).astype(np.float32)
        d = (dz.astype(np.float64) * dz + inner).astype(np.float32)
        t = np.minimum(d, t)
        others = np.delete(t, s)
        if not np.all(others < t[s]):
            return False
    return True


def test_prefix_condition_certifies_exactly_the_identity():
    """The decision rule of the FPS shortcut against the oracle's sampling: whenever the condition holds, sampling the cloud returns
    the identity (whatever the tie rule); on clouds with ties or not in furthest-point order it must refuse, and there the oracle's
    answer is indeed allowed to differ."""
    from cases import cloud, lidar
    xyz = lidar(5, 1, 2048)[0]
    first = oracle.furthest_point_sampling(xyz[None], 512)[0]
    ordered = xyz[first]
    assert _prefix_condition(ordered, 128)
    np.testing.assert_array_equal(oracle.furthest_point_sampling(ordered[None], 128)[0], np.arange(128, dtype=np.int32))
    swapped = ordered.copy()
    swapped[[3, 200]] = swapped[[200, 3]]
    assert not _prefix_condition(swapped, 128)
    assert not np.array_equal(oracle.furthest_point_sampling(swapped[None], 128)[0], np.arange(128, dtype=np.int32))
    lat = cloud(71, 1, 1024, "lattice")[0]
    lat_first = oracle.furthest_point_sampling(lat[None], 256)[0]
    assert not _prefix_condition(lat[lat_first], 64)  # 216 distinct lattice points: ties from the first steps on
    dup = ordered.copy()
    dup[40] = dup[41]
    assert not _prefix_condition(dup, 128)
    # and the subset property the runner relies on: the condition for (n, m) implies it for every prefix (n' <= n, m' <= m)
    assert _prefix_condition(ordered[:128], 32) and _prefix_condition(ordered[:32], 8)


def test_image_prep_restatement_matches_the_synthetic_loader():
    """oracle.image_prep (numpy float64, lib/datasets/kitti_dataset.py:44-55) == scenes.host_image_prep (torch float64): the two
    independent restatements of the reference's host-side preparation agree bit for bit, ragged sizes zero-padded."""
    import torch
    from epnet_b200 import scenes
    h = scenes.batch(1000, 2, 256, with_u8=True)
    np.testing.assert_array_equal(h["image"].numpy(), oracle.image_prep(h["image_u8"].numpy()))
    a = np.random.RandomState(0).randint(0, 256, size=(370, 1224, 3)).astype(np.uint8)
    out = oracle.image_prep([a])
    assert out.shape == (1, 3, 384, 1280) and out.dtype == np.float32
    assert np.all(out[0, :, 370:, :] == 0) and np.all(out[0, :, :, 1224:] == 0)
    v = a[5, 7].astype(np.float64)
    np.testing.assert_array_equal(out[0, :, 5, 7], (((v / 255.0) - np.array([0.485, 0.456, 0.406])) / np.array([0.229, 0.224, 0.225])).astype(np.float32))
