"""The first set-abstraction level as one kernel (csrc/sa_first_level.cu): group -> re-centre -> three-layer shared MLP -> max over
the ball, against a float64 restatement of the reference's layers (pointnet2_utils.py:241-264 QueryAndGroup with use_xyz and no
features, pytorch_utils.py:20-32 SharedMLP with BatchNorm(eval) folded, pointnet2_modules.py:52-56 max_pool2d) and against the
tcgen05 GEMM path it replaces in the runner.  Tolerance: BASELINE.json north_star's 1e-5 relative, stated on the output scale."""
import numpy as np
import pytest
import torch

from cases import lidar

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _layers(seed, widths, gain=1.0):
    from epnet_b200.gemm import PackedLinear
    g = torch.Generator().manual_seed(seed)
    lins, k = [], 3
    for n in widths:
        w = torch.randn(n, k, generator=g) * gain / k ** 0.5
        b = torch.randn(n, generator=g) * 0.1
        lins.append(PackedLinear(w.cuda(), b.cuda()))
        k = n
    return lins


def _want(xyz, new_xyz, idx, lins):
    B, m, ns = idx.shape
    x = (xyz.unsqueeze(1).expand(B, m, -1, 3).gather(2, idx.long().unsqueeze(-1).expand(B, m, ns, 3)) - new_xyz.unsqueeze(2)).double()  # fp32 subtraction, as the reference
    for lin in lins:
        x = torch.relu(x @ lin._w.double().t() + lin.bias.double())
    return x.max(dim=2).values.reshape(B * m, -1)


@pytest.mark.parametrize("widths,ns,B,m", [((32, 32, 64), 32, 2, 4096), ((16, 16, 32), 16, 2, 4096), ((32, 32, 64), 32, 1, 37),
                                           ((16, 16, 32), 16, 3, 13), ((32, 32, 64), 32, 1, 1)])
def test_fused_level_equals_float64_layers_and_the_gemm_path(widths, ns, B, m):
    from epnet_b200 import pointnet2_cuda as pc
    from epnet_b200.gemm import FusedFirstLevel, grouped_first_layer
    n = 16384 if m > 100 else 500
    xyz = torch.from_numpy(lidar(3000, B, 16384)[:, :n].copy()).cuda()
    new_xyz = xyz[:, :m].contiguous()
    idx = torch.zeros(B, m, ns, dtype=torch.int32, device="cuda")
    pc.ball_query_wrapper(B, n, m, 1.0, ns, new_xyz, xyz, idx)
    lins = _layers(7, widths)
    assert FusedFirstLevel.supports(lins, ns)
    wide = torch.full((B * m, widths[2] + 5), 123.0, device="cuda")  # a column slice of a wider buffer, as the runner's concat buffer
    FusedFirstLevel(lins, ns)(xyz, new_xyz, idx, wide[:, 3:3 + widths[2]])
    got = wide[:, 3:3 + widths[2]]
    want = _want(xyz, new_xyz, idx, lins)
    scale = want.abs().max().item()
    assert (got.double() - want).abs().max().item() <= TOL * scale
    assert torch.all(wide[:, :3] == 123.0) and torch.all(wide[:, 3 + widths[2]:] == 123.0)  # nothing outside the slice
    x = grouped_first_layer(lins[0], xyz, new_xyz, None, idx, relu=True)
    x = lins[1](x, relu=True)
    ref = lins[2](x, relu=True, pool=ns)
    assert (got - ref).abs().max().item() <= TOL * scale


def test_unsupported_widths_are_refused_and_the_guard_fires():
    from epnet_b200 import pointnet2_cuda as pc
    from epnet_b200._lib import EpnetKernelError, LIB
    from epnet_b200.gemm import FusedFirstLevel, OverflowFlag
    assert not FusedFirstLevel.supports(_layers(1, (32, 32, 64)), 16)
    assert not FusedFirstLevel.supports(_layers(1, (64, 64, 128)), 32)
    xyz = torch.from_numpy(lidar(3100, 1, 16384)[:, :2000].copy()).cuda()
    new_xyz = xyz[:, :64].contiguous()
    idx = torch.zeros(1, 64, 32, dtype=torch.int32, device="cuda")
    pc.ball_query_wrapper(1, 2000, 64, 2.0, 32, new_xyz, xyz, idx)
    out = torch.empty(64, 64, device="cuda")
    with pytest.raises(EpnetKernelError):  # widths the kernel is not instantiated for
        pc._call("sa_first_level", LIB.epnet_sa_first_level, xyz, 1, 2000, 64, 32, 32, 32, 128, xyz.data_ptr(), new_xyz.data_ptr(),
                 idx.data_ptr(), xyz.data_ptr(), out.data_ptr(), 64)
    flag = OverflowFlag(torch.device("cuda:0"))
    flag.reset()
    FusedFirstLevel(_layers(2, (32, 32, 64)), 32)(xyz, new_xyz, idx, out)
    flag.read_async()
    torch.cuda.synchronize()
    assert flag.value() == 0
    FusedFirstLevel(_layers(2, (32, 32, 64), gain=300.0), 32)(xyz, new_xyz, idx, out)  # activations far beyond 6e4
    flag.read_async()
    torch.cuda.synchronize()
    assert flag.value() != 0 and out.abs().max().item() > 6e4
    flag.reset()


def test_runner_with_and_without_the_fused_level():
    import bench
    from epnet_b200 import scenes
    device = torch.device("cuda:0")
    model = bench.build_model(device)
    d = scenes.batch(4400, 2, 16384)
    pts, img, xy = d["points"].to(device), d["image"].to(device), d["xy"].to(device)
    outs = []
    for fused in (True, False):
        r = model.make_runner(2, 16384, device, fused_first_level=fused)
        assert bool(r._fused_sa) == fused  # the published configuration's first level qualifies
        with torch.no_grad():
            xyz, feats = r(pts, img, xy.clone())
        torch.cuda.synchronize()
        assert not r.overflowed()
        outs.append((xyz.clone(), feats.clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    scale = outs[1][1].abs().max().item()
    assert (outs[0][1] - outs[1][1]).abs().max().item() <= TOL * scale
