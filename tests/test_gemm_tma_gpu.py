"""The TMA-fed FP16-split kernel (csrc/gemm_tf32x3.cu, gemm_f16x3_tma_kernel): activations arrive as two FP16 planes written by the
producing layer's epilogue; 3x3 convolutions (stride 1/2, zero padding by TMA out-of-bounds fill, ragged tiles) and plain GEMMs
against float64 references of the same layer (torch conv2d / matmul on CPU-exact operands).  Replaces the cuDNN fp32 convolutions
of the reference's image stream (lib/net/pointnet2_msg.py:17-33); tolerance = the fp32-grade bound of the other GEMM tests."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 4e-6  # of the output scale (22-bit operand split; tests/test_gemm_gpu.py uses the same bound)


def _planes(x):
    from epnet_b200.gemm import Planes
    h1 = x.half()
    h2 = ((x - h1.float()) * 2048.0).half()
    return Planes(h1.contiguous(), h2.contiguous())


def _rel(a, b):
    return (a.double() - b).abs().max().item() / b.abs().max().item()


@pytest.mark.parametrize("cin,cout,stride,hw", [(64, 64, 2, (48, 80)), (64, 128, 1, (24, 48)), (128, 128, 2, (40, 64)), (128, 256, 1, (16, 32)),
                                                (256, 512, 1, (8, 16)), (64, 64, 1, (13, 21)), (64, 96, 2, (19, 35)), (512, 512, 2, (48, 160))])
def test_conv_from_planes_matches_float64(cin, cout, stride, hw):
    from epnet_b200.gemm import PackedConv3x3
    g = torch.Generator().manual_seed(cin * 7 + cout + stride)
    B, (H, W) = 2, hw
    x = torch.randn(B, H, W, cin, generator=g).cuda()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    conv = PackedConv3x3(w, bias, stride=stride)
    assert conv.planes_capable()
    px = _planes(x)
    y, py = conv(px, relu=True, planes_out=True)
    torch.cuda.synchronize()
    xe = px.float().double()  # the value the planes carry exactly (x to 22 bits)
    want = F.relu(F.conv2d(xe.permute(0, 3, 1, 2), w.double(), bias.double(), stride=stride, padding=1)).permute(0, 2, 3, 1)
    assert y.shape == want.shape
    tol = TOL if 9 * cin <= 2304 else 6e-6  # the error of an fp32 accumulation grows slowly with K (4608 here: 4.6e-6 measured)
    assert _rel(y, want) <= tol, _rel(y, want)
    # the planes output carries the fp32 result to 22 bits
    assert (py.float() - y).abs().max().item() <= 2.0 ** -21 * y.abs().max().item()
    # and agrees with the layer evaluated from the fp32 tensor (SIMT-producer kernels)
    y32 = conv(x, relu=True)
    assert _rel(y32, want) <= 2 * TOL  # `want` was formed from the 22-bit-rounded input the planes carry
    only_planes = conv(px, relu=True, planes_out=True, f32_out=False)
    torch.cuda.synchronize()
    assert only_planes[0] is None and torch.equal(only_planes[1].h1, py.h1) and torch.equal(only_planes[1].h2, py.h2)


def test_first_conv_ffma_kernel_matches_float64():
    """3 -> 64 channels, stride 1: the dedicated fp32 FFMA kernel (csrc/first_conv.cu), fp32 and planes outputs, borders included"""
    from epnet_b200.gemm import PackedConv3x3
    g = torch.Generator().manual_seed(11)
    for (H, W) in ((32, 64), (17, 36), (384, 1280)):
        x = torch.zeros(2, H, W, 4)
        x[..., :3] = torch.randn(2, H, W, 3, generator=g)
        x[..., 3] = 123.0  # the padded channel is ignored
        x = x.cuda()
        w = (torch.randn(64, 3, 3, 3, generator=g) / 5).cuda()
        bias = torch.randn(64, generator=g).cuda()
        conv = PackedConv3x3(w, bias, stride=1)
        assert conv.w_c3 is not None
        y, p = conv(x, relu=True, planes_out=True)
        _, p_only = conv(x, relu=True, planes_out=True, f32_out=False)
        y_only = conv(x, relu=True)
        torch.cuda.synchronize()
        want = F.relu(F.conv2d(x[..., :3].double().permute(0, 3, 1, 2), w.double(), bias.double(), padding=1)).permute(0, 2, 3, 1)
        assert _rel(y, want) <= 1e-6, _rel(y, want)  # plain fp32 accumulation of 27 products
        assert torch.equal(y, y_only)
        assert (p.float() - y).abs().max().item() <= 2.0 ** -21 * y.abs().max().item()
        assert torch.equal(p.h1, p_only.h1) and torch.equal(p.h2, p_only.h2)
    # a width that is not a multiple of 4 stays on the tensor-core path
    x = torch.randn(1, 9, 38, 4).cuda()
    y = conv(x, relu=True)
    want = F.relu(F.conv2d(x[..., :3].double().permute(0, 3, 1, 2), w.double(), bias.double(), padding=1)).permute(0, 2, 3, 1)
    assert _rel(y, want) <= TOL


def test_first_conv_writes_planes_and_chain_matches():
    from epnet_b200.gemm import PackedConv3x3
    g = torch.Generator().manual_seed(5)
    x = torch.zeros(2, 32, 64, 4)
    x[..., :3] = torch.randn(2, 32, 64, 3, generator=g)
    x = x.cuda()
    w0 = (torch.randn(64, 3, 3, 3, generator=g) / 5).cuda()
    w1 = (torch.randn(64, 64, 3, 3, generator=g) / 24).cuda()
    c0, c1 = PackedConv3x3(w0, torch.zeros(64).cuda(), stride=1), PackedConv3x3(w1, None, stride=2)
    y0, p0 = c0(x, relu=True, planes_out=True)
    y0_ref = c0(x, relu=True)
    torch.cuda.synchronize()
    assert torch.equal(y0, y0_ref)  # the same kernel, one more output
    assert (p0.float() - y0).abs().max().item() <= 2.0 ** -21 * y0.abs().max().item()
    y1 = c1(p0, relu=False)
    want = F.conv2d(y0.double().permute(0, 3, 1, 2), w1.double(), None, stride=2, padding=1).permute(0, 2, 3, 1)
    assert _rel(y1, want) <= 2 * TOL


@pytest.mark.parametrize("L,K,N,pool", [(4096, 256, 128, 1), (1000, 96, 64, 1), (8192, 512, 512, 1), (2048, 192, 256, 16), (130, 64, 16, 1)])
def test_plain_gemm_from_planes(L, K, N, pool):
    from epnet_b200.gemm import PackedLinear
    g = torch.Generator().manual_seed(L + K + N)
    ld = (K + 7) // 8 * 8
    xf = torch.zeros(L, ld)
    xf[:, :K] = torch.randn(L, K, generator=g)
    px = _planes(xf.cuda())
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    lin = PackedLinear(w, bias)
    y = lin.from_planes(px, relu=True, pool=pool)
    torch.cuda.synchronize()
    want = F.relu(px.float().double()[:, :K] @ w.double().t() + bias.double())
    if pool > 1:
        want = want.view(L // pool, pool, N).max(dim=1).values
    assert y.shape == want.shape and _rel(y, want) <= TOL, _rel(y, want)
