"""Unit parity of the fused / point-major kernels (include/epnet_b200.h "fused entry points") against the composition of
reference-surface ops they replace (oracle for indices, plain torch for the glue)."""
import numpy as np
import pytest
import torch

import oracle
from cases import cloud

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_fps_sample_emits_new_xyz_and_payload():
    from epnet_b200 import pointnet2_cuda as pc
    xyz = cloud(3, 2, 3000, "uniform", dup_frac=0.05)
    aux = np.random.RandomState(0).rand(2, 3000, 2).astype(np.float32)
    want = oracle.furthest_point_sampling(xyz, 700)
    x, a = _dev(xyz), _dev(aux)
    temp = torch.full((2, 3000), 1e10, device="cuda")
    idx = torch.empty(2, 700, dtype=torch.int32, device="cuda")
    new_xyz = torch.empty(2, 700, 3, device="cuda")
    aux_out = torch.empty(2, 700, 2, device="cuda")
    pc.fps_sample_wrapper(2, 3000, 700, x, temp, idx, new_xyz, a, aux_out)
    np.testing.assert_array_equal(idx.cpu().numpy(), want)
    for s in range(2):
        np.testing.assert_array_equal(new_xyz[s].cpu().numpy(), xyz[s][want[s]])
        np.testing.assert_array_equal(aux_out[s].cpu().numpy(), aux[s][want[s]])


@pytest.mark.parametrize("c,ns", [(0, 16), (96, 32), (10, 4)])
def test_group_concat_variants(c, ns):
    from epnet_b200 import pointnet2_cuda as pc
    rng = np.random.RandomState(c + ns)
    B, N, M = 2, 500, 60
    xyz = cloud(5, B, N, "gauss")
    new_xyz = np.ascontiguousarray(xyz[:, :M])
    feats = rng.randn(B, max(c, 1), N).astype(np.float32)[:, :c]
    idx = rng.randint(0, N, size=(B, M, ns)).astype(np.int32)
    gx = oracle.group_points(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx) - new_xyz.transpose(0, 2, 1)[..., None]
    want_cm = np.concatenate([gx, oracle.group_points(feats, idx)], axis=1) if c else gx     # (B,3+C,M,ns): reference order
    x, q, i = _dev(xyz), _dev(new_xyz), _dev(idx)
    f_cm = _dev(feats) if c else None
    out_cm = torch.empty(B, 3 + c, M, ns, device="cuda")
    pc.group_concat_wrapper(B, c, N, M, ns, x, q, f_cm, i, out_cm)
    np.testing.assert_array_equal(out_cm.cpu().numpy(), want_cm)
    kp = (c + 3 + 3) // 4 * 4
    f_pm = _dev(feats.transpose(0, 2, 1)) if c else None
    out_pm = torch.full((B * M * ns, kp), 7.0, device="cuda")
    pc.group_concat_pm_wrapper(B, c, N, M, ns, x, q, f_pm, i, out_pm)
    got = out_pm.cpu().numpy().reshape(B, M, ns, kp)
    want_pm = want_cm.transpose(0, 2, 3, 1)  # (B,M,ns,3+C), xyz first
    np.testing.assert_array_equal(got[..., :c], want_pm[..., 3:])   # point-major rows carry the features first ...
    np.testing.assert_array_equal(got[..., c:c + 3], want_pm[..., :3])  # ... then the re-centred xyz
    assert (got[..., c + 3:] == 0).all()


def test_bias_relu_and_pool():
    from epnet_b200 import pointnet2_cuda as pc
    x = torch.randn(2, 5, 40, 16, device="cuda")
    b = torch.randn(5, device="cuda")
    want = torch.relu(x + b[None, :, None, None])
    pooled = torch.empty(2, 9, 40, device="cuda").fill_(-1)
    pc.bias_relu_maxpool_wrapper(2, 5, 40, 16, x, b, pooled.data_ptr() + 4 * 3 * 40, 9 * 40)
    torch.testing.assert_close(pooled[:, 3:8], want.max(dim=3).values, rtol=0, atol=0)
    assert (pooled[:, :3] == -1).all() and (pooled[:, 8:] == -1).all()
    y = x.clone()
    pc.bias_relu_wrapper(2, 5, 40 * 16, y, b)
    torch.testing.assert_close(y, want, rtol=0, atol=0)


def test_three_nn_weights_and_interpolate_concat():
    from epnet_b200 import pointnet2_cuda as pc
    rng = np.random.RandomState(1)
    B, n, m, c2, c1 = 2, 333, 50, 16, 8
    unknown, known = cloud(8, B, n, "gauss"), cloud(9, B, m, "gauss")
    d2_w, idx_w = oracle.three_nn(unknown, known)
    u, k = _dev(unknown), _dev(known)
    d2 = torch.empty(B, n, 3, device="cuda"); idx = torch.empty(B, n, 3, dtype=torch.int32, device="cuda"); w = torch.empty(B, n, 3, device="cuda")
    pc.three_nn_weights_wrapper(B, n, m, u, k, d2, idx, w)
    np.testing.assert_array_equal(idx.cpu().numpy(), idx_w)
    np.testing.assert_array_equal(d2.cpu().numpy(), d2_w)
    recip = 1.0 / (torch.sqrt(torch.from_numpy(d2_w)) + 1e-8)
    w_ref = recip / recip.sum(dim=2, keepdim=True)                    # pointnet2_modules.py:157-159
    torch.testing.assert_close(w.cpu(), w_ref, rtol=1e-6, atol=1e-7)
    kf = rng.randn(B, c2, m).astype(np.float32)
    skip = rng.randn(B, c1, n).astype(np.float32)
    interp = oracle.three_interpolate(kf, idx_w, w.cpu().numpy())
    want = np.concatenate([interp, skip], axis=1)                     # (B, c2+c1, n)
    out_cm = torch.empty(B, c2 + c1, n, device="cuda")
    pc.three_interpolate_concat_wrapper(B, c2, m, n, c1, _dev(kf), idx, d2, _dev(skip), out_cm)
    np.testing.assert_allclose(out_cm.cpu().numpy(), want, rtol=1e-5, atol=1e-6)
    out_pm = torch.empty(B * n, c2 + c1, device="cuda")
    pc.three_interpolate_concat_pm_wrapper(B, c2, m, n, c1, _dev(kf.transpose(0, 2, 1)), idx, w, _dev(skip.transpose(0, 2, 1)), out_pm)
    np.testing.assert_array_equal(out_pm.cpu().numpy().reshape(B, n, c2 + c1).transpose(0, 2, 1), want)


@pytest.mark.parametrize("align", [False, True])
def test_grid_gather_point_major_variants(align):
    from epnet_b200 import pointnet2_cuda as pc
    rng = np.random.RandomState(2)
    B, C, H, W, n = 2, 12, 24, 80, 300
    fmap = rng.randn(B, C, H, W).astype(np.float32)
    xy = (rng.rand(B, n, 2).astype(np.float32) * 2.2 - 1.1)
    want = oracle.grid_gather_bilinear(fmap, xy, align).transpose(0, 2, 1).reshape(B * n, C)
    out = torch.empty(B * n, C, device="cuda")
    pc.grid_gather_pm_wrapper(B, C, H, W, n, _dev(fmap), _dev(xy), align, out)
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    out2 = torch.empty(B * n, C, device="cuda")
    pc.grid_gather_nhwc_pm_wrapper(B, C, H, W, n, _dev(fmap.transpose(0, 2, 3, 1)), _dev(xy), align, out2)
    np.testing.assert_array_equal(out2.cpu().numpy(), want)


def test_deconv_as_gemm_plus_shuffle_equals_conv_transpose():
    from epnet_b200 import pointnet2_cuda as pc
    from epnet_b200.gemm import PackedLinear
    torch.backends.cudnn.allow_tf32 = False
    B, ci, co, h, w, k = 2, 32, 16, 6, 10, 4
    de = torch.nn.ConvTranspose2d(ci, co, kernel_size=k, stride=k).cuda()
    x = torch.randn(B, ci, h, w, device="cuda")
    with torch.no_grad():
        want = de(x).permute(0, 2, 3, 1)                               # NHWC
        lin = PackedLinear(de.weight.permute(2, 3, 1, 0).reshape(k * k * co, ci), None)
        y = lin(x.permute(0, 2, 3, 1).reshape(-1, ci).contiguous(), relu=False)
        cat = torch.zeros(B, h * k, w * k, 24, device="cuda")
        pc.deconv_shuffle_nhwc_wrapper(B, h, w, k, co, y, cat, 4)
    got = cat[..., 4:20] + de.bias
    assert (cat[..., :4] == 0).all() and (cat[..., 20:] == 0).all()
    err = (got - want).abs().max().item()
    assert err <= 5e-6 * want.abs().max().item(), err


@pytest.mark.parametrize("rows,rc,c", [(8192, 24, 96), (2048, 64, 256), (130, 256, 1024), (32768, 8, 32), (5, 128, 512)])
def test_attention_scale_pm(rows, rc, c):
    """IA_Layer tail in one kernel == the op-by-op torch expression (lib/net/pointnet2_msg.py:88-95), written into a column
    slice of a wider buffer."""
    from epnet_b200 import pointnet2_cuda as pc
    g = torch.Generator(device="cpu").manual_seed(rows + c)
    r1 = torch.randn(rows, rc, generator=g).cuda()
    r2 = torch.randn(rows, rc, generator=g).cuda()
    w3 = (torch.randn(rc, generator=g) / rc ** 0.5).cuda()
    b3 = torch.randn(1, generator=g).cuda()
    x = torch.randn(rows, c, generator=g).cuda().relu()
    cat = torch.full((rows, 2 * c + 4), -3.0, device="cuda")
    pc.attention_scale_pm_wrapper(r1, r2, w3, b3, x, cat[:, c:2 * c])
    torch.cuda.synchronize()
    att = torch.sigmoid(torch.tanh(r1.double() + r2.double()) @ w3.double() + b3.double())
    want = x.double() * att[:, None]
    err = (cat[:, c:2 * c].double() - want).abs().max().item()
    assert err <= 2e-6 * max(want.abs().max().item(), 1.0)
    assert (cat[:, :c] == -3.0).all() and (cat[:, 2 * c:] == -3.0).all()
