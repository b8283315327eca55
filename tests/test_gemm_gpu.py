"""tcgen05 3xTF32 GEMM (csrc/gemm_tf32x3.cu) against a float64 reference: fp32-grade accuracy (the tolerance a plain
TF32 GEMM fails by two orders of magnitude), bias/ReLU/max-pool epilogues, ragged L / K / N."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, w, b, relu, pool):
    y = x.double() @ w.double().t()
    if b is not None:
        y = y + b.double()
    if relu:
        y = y.clamp_min(0)
    if pool > 1:
        y = y.view(-1, pool, y.shape[-1]).max(dim=1).values
    return y


@pytest.mark.parametrize("L,K,N,relu,pool", [
    (128, 32, 16, False, 1), (256, 32, 16, True, 1), (1000, 99, 64, True, 1), (4096, 259, 196, True, 1), (512, 1536, 512, True, 1),
    (4096, 6, 32, True, 16), (2048, 99, 128, True, 32), (384, 515, 384, False, 1), (96, 64, 96, True, 2), (33, 35, 17, False, 1),
    # narrow-tile persistent kernel: more 128-row tiles than resident CTAs (every CTA walks several tiles, barrier phases wrap),
    # ragged last tile, two column tiles, pooled epilogue; and a wide-tile launch under both tile policies
    (89600 + 77, 70, 48, True, 1), (65536, 33, 128, True, 16), (40000, 200, 100, False, 1), (3000, 300, 512, True, 1),
])
def test_gemm_matches_fp64(L, K, N, relu, pool):
    from epnet_b200.gemm import PackedLinear
    g = torch.Generator(device="cpu").manual_seed(L * 7 + K)
    x = torch.randn(L, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    lin = PackedLinear(w, b)
    y = lin(x, relu=relu, pool=pool)
    torch.cuda.synchronize()
    want = _ref(x, w, b, relu, pool)
    assert y.shape == want.shape
    err = (y.double() - want).abs().max().item()
    scale = want.abs().max().item()
    # 3xTF32 on the tensor core: hi*hi + hi*lo + lo*hi with fp32 accumulation -> a few 1e-6 of the output scale, inside
    # BASELINE.json's 1e-5; the exact figure per shape is printed for DESIGN.md
    print("gemm L=%d K=%d N=%d: max abs err %.3e, scale %.3e, rel %.2e" % (L, K, N, err, scale, err / scale))
    assert err <= 4e-6 * scale, "max abs err %.3e (scale %.3e)" % (err, scale)
    # the same product in plain TF32 is ~1e-3: make sure the test would notice a missing correction term
    torch.backends.cuda.matmul.allow_tf32 = True
    tf32 = (x @ w.t() + b)
    torch.backends.cuda.matmul.allow_tf32 = False
    if K >= 64 and pool == 1 and not relu:
        assert (tf32.double() - want).abs().max().item() > 20 * err


def test_tile_policies_agree():
    """latency tiles (narrow, more CTAs) and throughput tiles (widest) are two schedules of the same arithmetic per output"""
    from epnet_b200.gemm import PackedLinear, tile_policy
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn(600, 777, generator=g).cuda()
    lin = PackedLinear((torch.randn(512, 777, generator=g) / 28).cuda(), torch.randn(512, generator=g).cuda())
    with tile_policy("latency"):
        bn_l, a = lin.for_rows(600)[0], lin(x)
    with tile_policy("throughput"):
        bn_t, b = lin.for_rows(600)[0], lin(x)
    assert bn_l < bn_t
    want = _ref(x, lin._w, lin.bias, True, 1)
    for y in (a, b):
        assert (y.double() - want).abs().max().item() <= 4e-6 * want.abs().max().item()


def test_gemm_strided_input_and_output():
    from epnet_b200.gemm import PackedLinear
    x_full = torch.randn(640, 72, device="cuda")
    w = torch.randn(40, 67, device="cuda") * 0.1
    lin = PackedLinear(w, None)
    out_full = torch.zeros(640, 100, device="cuda")
    lin(x_full[:, :67], relu=False, out=out_full[:, 8:48])
    want = x_full[:, :67].double() @ w.double().t()
    assert (out_full[:, 8:48].double() - want).abs().max().item() < 1e-5
    assert out_full[:, :8].abs().max().item() == 0 and out_full[:, 48:].abs().max().item() == 0


@pytest.mark.parametrize("B,H,W,cin,cout,stride,relu", [(2, 12, 20, 3, 16, 1, True), (1, 9, 13, 8, 24, 2, False), (2, 48, 160, 64, 128, 1, True),
                                                         (2, 48, 160, 128, 128, 2, False), (1, 24, 80, 256, 512, 1, True)])
def test_conv3x3_nhwc_matches_fp64(B, H, W, cin, cout, stride, relu):
    from epnet_b200.gemm import PackedConv3x3
    g = torch.Generator(device="cpu").manual_seed(H * W + cin)
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).cuda()
    b = torch.randn(cout, generator=g).cuda()
    conv = PackedConv3x3(w, b, stride=stride)
    x_nhwc = torch.zeros(B, H, W, conv.cin_p, device="cuda")
    x_nhwc[..., :cin] = x.permute(0, 2, 3, 1)
    y = conv(x_nhwc, relu=relu)
    torch.cuda.synchronize()
    want = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=1)
    if relu:
        want = want.clamp_min(0)
    want = want.permute(0, 2, 3, 1)
    assert y.shape == want.shape
    err = (y.double() - want).abs().max().item()
    scale = want.abs().max().item()
    print("conv %dx%d cin=%d cout=%d s%d: max abs err %.3e, scale %.3e, rel %.2e" % (H, W, cin, cout, stride, err, scale, err / scale))
    assert err <= 1e-5 * scale  # K = 9*Cin up to 4608: the tensor core's truncating accumulate shows (~2e-9 per k)


@pytest.mark.parametrize("B,h,w,cin,cpad,k,co", [(2, 6, 10, 64, 64, 2, 16), (1, 5, 7, 24, 32, 4, 16), (2, 3, 5, 512, 512, 16, 16), (1, 7, 3, 40, 40, 3, 8)])
def test_deconv_nhwc_matches_fp64(B, h, w, cin, cpad, k, co):
    """ConvTranspose2d(kernel == stride) as one GEMM with a scattering epilogue, written into a slice of a wider NHWC buffer."""
    from epnet_b200.gemm import PackedDeconv
    g = torch.Generator(device="cpu").manual_seed(h * w + cin + k)
    x = torch.randn(B, cin, h, w, generator=g).cuda()
    wt = (torch.randn(cin, co, k, k, generator=g) / cin ** 0.5).cuda()
    b = torch.randn(co, generator=g).cuda()
    x_nhwc = torch.zeros(B, h, w, cpad, device="cuda")
    x_nhwc[..., :cin] = x.permute(0, 2, 3, 1)
    cat = torch.full((B, h * k, w * k, co + 24), 7.0, device="cuda")
    PackedDeconv(wt, b)(x_nhwc, cat[..., 8:8 + co])
    torch.cuda.synchronize()
    want = torch.nn.functional.conv_transpose2d(x.double(), wt.double(), b.double(), stride=k).permute(0, 2, 3, 1)
    err = (cat[..., 8:8 + co].double() - want).abs().max().item()
    assert err <= 4e-6 * want.abs().max().item()
    assert (cat[..., :8] == 7.0).all() and (cat[..., 8 + co:] == 7.0).all()  # neighbours of the slice untouched


@pytest.mark.parametrize("B,pts,K,N", [(2, 16384, 160, 128), (3, 1000, 70, 48), (1, 130, 300, 512)])
def test_gemm_channel_major_output(B, pts, K, N):
    """the epilogue can write (B, N, pts) -- the reference's (B, C, N) feature layout -- directly"""
    from epnet_b200.gemm import PackedLinear
    g = torch.Generator(device="cpu").manual_seed(pts + K)
    x = torch.randn(B * pts, K, generator=g).cuda()
    lin = PackedLinear((torch.randn(N, K, generator=g) / K ** 0.5).cuda(), torch.randn(N, generator=g).cuda())
    out = torch.empty(B, N, pts, device="cuda")
    lin(x, relu=True, out_cm=out)
    torch.cuda.synchronize()
    want = _ref(x, lin._w, lin.bias, True, 1).view(B, pts, N).transpose(1, 2)
    assert (out.double() - want).abs().max().item() <= 4e-6 * want.abs().max().item()


@pytest.mark.parametrize("B,n,m,ns,c,N", [(2, 4096, 1024, 16, 96, 64), (2, 16384, 4096, 16, 0, 16), (1, 1000, 130, 32, 30, 48),
                                          (2, 256, 64, 32, 512, 512), (3, 500, 77, 8, 5, 33)])
def test_grouped_first_layer_equals_group_then_gemm(B, n, m, ns, c, N):
    """QueryAndGroup fused into the GEMM operand == group_concat_pm followed by the plain GEMM, bit for bit (same values enter the
    same MMAs), and both match the float64 reference; feature widths that are not multiples of 4 and an xyz-only level included."""
    from epnet_b200 import pointnet2_cuda as pc
    from epnet_b200.gemm import PackedLinear, grouped_first_layer
    g = torch.Generator(device="cpu").manual_seed(n + c)
    xyz = torch.randn(B, n, 3, generator=g).cuda()
    new_xyz = xyz[:, :m].contiguous()
    feats = None if c == 0 else torch.randn(B, n, c, generator=g).cuda()  # c % 4 != 0: rows are not 16-byte aligned
    idx = torch.randint(0, n, (B, m, ns), generator=g).int().cuda()
    lin = PackedLinear((torch.randn(N, c + 3, generator=g) / (c + 3) ** 0.5).cuda(), torch.randn(N, generator=g).cuda())
    kp = (c + 3 + 3) // 4 * 4
    rows = torch.empty(B * m * ns, kp, device="cuda")
    pc.group_concat_pm_wrapper(B, c, n, m, ns, xyz, new_xyz, feats, idx, rows)
    two_step = lin(rows, relu=True)
    fused = grouped_first_layer(lin, xyz, new_xyz, feats, idx, relu=True)
    torch.cuda.synchronize()
    want = _ref(rows[:, :c + 3], lin._w, lin.bias, True, 1)
    assert (two_step.double() - want).abs().max().item() <= 4e-6 * want.abs().max().item()
    assert fused is not None
    if lin.for_rows(B * m * ns)[0] > 64:  # the two-step product ran on a wide (FP16-split) tile, the fused one on 64-column TF32-split tiles
        assert (fused.double() - want).abs().max().item() <= 4e-6 * want.abs().max().item()
    else:
        assert torch.equal(fused, two_step)


@pytest.mark.parametrize("f16_wide", [True, False])
def test_wide_tiles_both_splits(monkeypatch, f16_wide):
    """tiles wider than 64 columns run the FP16 two-term split by default and the TF32 split with gemm.F16_WIDE off: both are
    fp32-grade, and they are different kernels (results differ in the last bits)"""
    from epnet_b200 import gemm
    monkeypatch.setattr(gemm, "F16_WIDE", f16_wide)
    g = torch.Generator(device="cpu").manual_seed(17)
    x = (torch.randn(3000, 700, generator=g) * 3).cuda()
    lin = gemm.PackedLinear((torch.randn(300, 700, generator=g) / 26).cuda(), torch.randn(300, generator=g).cuda())
    conv = gemm.PackedConv3x3((torch.randn(256, 128, 3, 3, generator=g) / 34).cuda(), torch.randn(256, generator=g).cuda())
    xi = torch.randn(1, 24, 40, 128, generator=g).cuda()
    with gemm.tile_policy("throughput"):  # widest tiles whatever the row count
        assert lin.for_rows(3000)[0] > 64 and lin.wide_f16(lin.for_rows(3000)[0]) == f16_wide
        y = lin(x, relu=False)
        yc = conv(xi, relu=True)
    torch.cuda.synchronize()
    want = _ref(x, lin._w, lin.bias, False, 1)
    assert (y.double() - want).abs().max().item() <= 4e-6 * want.abs().max().item()
    wantc = torch.nn.functional.conv2d(xi.permute(0, 3, 1, 2).double(), conv.lin._w.view(256, 3, 3, 128).permute(0, 3, 1, 2).double(),
                                       conv.lin.bias.double(), padding=1).clamp_min(0).permute(0, 2, 3, 1)
    assert (yc.double() - wantc).abs().max().item() <= 1e-5 * wantc.abs().max().item()
