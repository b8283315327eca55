"""The set-abstraction levels sample each other's output (reference pointnet2_msg.py:141-153, pointnet2_modules.py:38-42).  A cloud in
furthest-point order samples to the identity prefix unless two of its points tie for a maximum; epnet_fps_prefix_check decides that
exactly and epnet_fps_sample_guarded answers with the prefix or runs the sampling kernel.  Everything here is bit-exact: the guarded
chain must equal the plain sampling kernels (themselves pinned to the oracle and the reference kernel in test_ops_gpu.py /
test_ref_pin.py) on clouds where the shortcut engages AND on clouds where it must not (lattices, duplicates, clouds that are not in
furthest-point order at all)."""
import numpy as np
import pytest
import torch

import oracle
from cases import cloud, lidar
from gpu_util import dev

pytestmark = pytest.mark.gpu


def _fps(pc, xyz, m, identity=None):
    B, n = xyz.shape[0], xyz.shape[1]
    temp = torch.full((B, n), 1e10, device="cuda")
    idx = torch.full((B, m), -7, dtype=torch.int32, device="cuda")
    new_xyz = torch.full((B, m, 3), float("nan"), device="cuda")
    if identity is None:
        pc.fps_sample_wrapper(B, n, m, xyz, temp, idx, new_xyz)
    else:
        pc.fps_sample_guarded_wrapper(B, n, m, xyz, temp, idx, identity, new_xyz)
    return idx, new_xyz


def _check(pc, xyz, m):
    B, n = xyz.shape[0], xyz.shape[1]
    flag = torch.full((B,), -1, dtype=torch.int32, device="cuda")
    pc.fps_prefix_check_wrapper(B, n, m, xyz, torch.empty((B, m), device="cuda"), flag)
    return flag


def _chain(pc, xyz, levels):
    """plain chain and guarded chain (one check after the first level) -> flags, [(idx, new_xyz)] of both"""
    plain, guarded = [], []
    cur_p = cur_g = xyz
    flag = None
    for li, m in enumerate(levels):
        ip, xp = _fps(pc, cur_p, m)
        ig, xg = _fps(pc, cur_g, m, identity=flag)
        plain.append((ip, xp))
        guarded.append((ig, xg))
        if li == 0:
            flag = _check(pc, xg, levels[1])
        cur_p, cur_g = xp, xg
    torch.cuda.synchronize()
    return flag.cpu().numpy(), plain, guarded


@pytest.mark.parametrize("kind,n,levels,engages", [
    ("lidar", 16384, (4096, 1024, 256, 64), True),      # the backbone's chain on KITTI-shaped scenes (2 % exact duplicates)
    ("uniform", 8192, (2048, 512, 128, 32), True),
    ("gauss", 3000, (700, 300, 50), True),
    ("lattice", 4096, (1024, 256, 64), False),           # equal spacings: ties at almost every step
    ("identical", 512, (128, 32), False),                # all distances 0
    ("gauss", 600, (600, 600), None),                    # m == n at both levels
])
def test_guarded_chain_equals_the_sampling_kernels(kind, n, levels, engages):
    from epnet_b200 import pointnet2_cuda as pc
    xyz_np = lidar(2000, 2)[:, :n] if kind == "lidar" else cloud(71, 2, n, kind, dup_frac=0.03 if kind in ("uniform", "gauss") else 0.0)
    flag, plain, guarded = _chain(pc, dev(np.ascontiguousarray(xyz_np)), levels)
    for (ip, xp), (ig, xg) in zip(plain, guarded):
        assert torch.equal(ip, ig)
        assert torch.equal(xp, xg)
    if engages is True:
        assert flag.tolist() == [1, 1]
        for ip, _ in plain[1:]:  # and the plain kernels agree that the answer is the identity
            assert torch.equal(ip.cpu(), torch.arange(ip.shape[1], dtype=torch.int32).expand(2, -1))
    elif engages is False:
        assert flag.tolist() == [0, 0]
    want = oracle.furthest_point_sampling(xyz_np, levels[0])  # the chain starts from the oracle's first level
    np.testing.assert_array_equal(plain[0][0].cpu().numpy(), want)


def test_flag_is_per_scene_and_rejects_clouds_that_are_not_in_sampling_order():
    from epnet_b200 import pointnet2_cuda as pc
    a = lidar(2100, 1)[0, :8192]
    first = oracle.furthest_point_sampling(a[None], 2048)[0]
    ordered = a[first]                                       # in furthest-point order: the identity is exact
    shuffled = ordered[np.random.RandomState(0).permutation(2048)]
    swapped = ordered.copy()
    swapped[[5, 900]] = swapped[[900, 5]]                   # one transposition breaks it
    tied = ordered.copy()
    tied[1500] = tied[3]                                     # a duplicate of an early pick: T_1500(s) ties nothing until... its distance is 0
    tied[40] = tied[41]                                      # two equal points next to each other in the order: step 40 ties with point 41
    batch = dev(np.stack([ordered, shuffled, swapped, tied]).astype(np.float32))
    flag = _check(pc, batch, 512)
    ig, xg = _fps(pc, batch, 512, identity=flag)
    ip, xp = _fps(pc, batch, 512)
    torch.cuda.synchronize()
    assert flag.cpu().tolist() == [1, 0, 0, 0]
    assert torch.equal(ig, ip) and torch.equal(xg, xp)
    np.testing.assert_array_equal(ip.cpu().numpy(), oracle.furthest_point_sampling(batch.cpu().numpy(), 512))
    assert torch.equal(ip[0].cpu(), torch.arange(512, dtype=torch.int32))


def test_auxiliary_rows_follow_the_prefix():
    from epnet_b200 import pointnet2_cuda as pc
    a = lidar(2200, 2)[:, :4096]
    first = oracle.furthest_point_sampling(a, 1024)
    ordered = dev(np.stack([a[s][first[s]] for s in range(2)]))
    aux = torch.randn(2, 1024, 2, device="cuda")
    flag = _check(pc, ordered, 256)
    out = []
    for ident in (None, flag):
        temp = torch.full((2, 1024), 1e10, device="cuda")
        idx = torch.empty((2, 256), dtype=torch.int32, device="cuda")
        new_xyz, new_aux = torch.empty((2, 256, 3), device="cuda"), torch.empty((2, 256, 2), device="cuda")
        if ident is None:
            pc.fps_sample_wrapper(2, 1024, 256, ordered, temp, idx, new_xyz, aux, new_aux)
        else:
            pc.fps_sample_guarded_wrapper(2, 1024, 256, ordered, temp, idx, ident, new_xyz, aux, new_aux)
        out.append((idx, new_xyz, new_aux))
    torch.cuda.synchronize()
    assert flag.cpu().tolist() == [1, 1]
    for p, g in zip(*out):
        assert torch.equal(p, g)


def test_check_refuses_more_steps_than_it_can_stage_and_bad_arguments():
    from epnet_b200 import pointnet2_cuda as pc
    from epnet_b200._lib import EpnetKernelError
    xyz = dev(cloud(72, 1, 6000, "gauss"))
    flag = _check(pc, xyz, 3000)  # m > 2048: no shortcut, flags cleared
    torch.cuda.synchronize()
    assert flag.cpu().tolist() == [0]
    with pytest.raises(EpnetKernelError):
        _check(pc, xyz, 7000)  # m > n


def test_runner_is_bit_identical_with_and_without_the_shortcut():
    import bench
    from epnet_b200 import scenes
    device = torch.device("cuda:0")
    model = bench.build_model(device)
    d = scenes.batch(4321, 2, 16384)
    pts, img, xy = d["points"].to(device), d["image"].to(device), d["xy"].to(device)
    outs = []
    for prefix in (True, False):
        r = model.make_runner(2, 16384, device, prefix_fps=prefix)
        with torch.no_grad():
            xyz, feats = r(pts, img, xy.clone())
        torch.cuda.synchronize()
        outs.append((xyz.clone(), feats.clone()))
        if prefix:
            assert r.fps_identity.cpu().tolist() == [1, 1]
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
