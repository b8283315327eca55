"""Sparse evaluation of the final image fusion (epnet_b200/sparse_tail.py) against the dense computation the
reference performs (lib/net/pointnet2_msg.py:237-246: 4 x ConvTranspose2d -> cat -> 1x1 conv + BN + ReLU over every pixel -> bilinear
gather): torch float64 on the same weights, including taps outside the canvas, on its border, and points sharing a pixel; and the
whole backbone with the sparse tail == the whole backbone with the dense tail."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dense_reference(deconvs, wf, bf, maps_nchw, xy):
    up = torch.cat([d.double()(m.double()) for d, m in zip(deconvs, maps_nchw)], dim=1)
    fused = F.relu(F.conv2d(up, wf.double()[:, :, None, None], bf.double()))
    # the sampling itself in fp32, as the reference runs it (ATen grid_sampler_2d): the pixel coordinate ((x + 1) * W - 1) / 2 carries
    # an fp32 rounding of ~W * 6e-8 pixels, i.e. up to 4e-5 in the bilinear weights at W = 1280 -- in ATen and in our kernels alike;
    # a float64 sampler would "disagree" with both by that much (measured: 5.8e-5 of the scale at W = 1280, 1.5e-6 at W = 48)
    return F.grid_sample(fused.float(), xy.unsqueeze(1), align_corners=False).squeeze(2).double()  # (B, Cf, N)


@pytest.mark.parametrize("hw,cins,n", [((96, 320), (8, 16, 24, 32), 1024), ((384, 1280), (64, 128, 256, 512), 16384), ((32, 48), (64, 32), 100)])
def test_sparse_tail_matches_dense_float64(hw, cins, n):
    from epnet_b200.sparse_tail import SparseImageTail
    torch.manual_seed(len(cins) * 100 + n)
    B, (H, W) = 2, hw
    ks = (2, 4, 8, 16)[:len(cins)]
    deconvs = [nn.ConvTranspose2d(c, 16, kernel_size=k, stride=k).cuda() for c, k in zip(cins, ks)]
    wf = (torch.randn(32, 16 * len(cins)) / (16 * len(cins)) ** 0.5).cuda()
    bf = torch.randn(32).cuda() * 0.1
    biases = torch.cat([d.bias.detach() for d in deconvs])
    tail = SparseImageTail(deconvs, wf, bf + wf @ biases, B, n, (H, W), torch.device("cuda"))
    maps = [torch.randn(B, H // k, W // k, c, device="cuda") for c, k in zip(cins, ks)]
    xy = torch.rand(B, n, 2, device="cuda") * 2.2 - 1.1          # ~9 % of the points have taps outside the canvas
    xy[0, :8] = torch.tensor([[-1.0, -1.0], [1.0, 1.0], [-1.0, 1.0], [0.0, 0.0], [0.0, 0.0], [1.5, 0.0], [-0.999, 0.3], [0.37, -1.2]], device="cuda")
    out = torch.full((B * n, 36), 7.0, device="cuda")
    tail(maps, xy.contiguous(), out)
    torch.cuda.synchronize()
    want = _dense_reference(deconvs, wf, bf, [m.permute(0, 3, 1, 2) for m in maps], xy)
    got = out[:, :32].view(B, n, 32).permute(0, 2, 1).double()
    assert torch.all(out[:, 32:] == 7.0)  # columns beyond Cf untouched
    err = (got - want).abs().max().item() / want.abs().max().item()
    print("sparse tail vs dense float64: max abs err / scale = %.2e" % err)
    assert err <= 1e-5


def test_backbone_with_sparse_tail_equals_dense_tail():
    from epnet_b200 import BackboneConfig, Pointnet2MSG, scenes
    torch.manual_seed(0)
    model = Pointnet2MSG(config=BackboneConfig()).cuda().eval()
    g = torch.Generator().manual_seed(1)
    for mod in model.modules():
        if isinstance(mod, (nn.BatchNorm1d, nn.BatchNorm2d)):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
    dev = torch.device("cuda")
    sparse = model.make_runner(2, 16384, dev, sparse_tail=True)
    dense = model.make_runner(2, 16384, dev, sparse_tail=False)
    assert sparse.sparse_tail is not None and dense.sparse_tail is None
    for seed in (1000, 1500):
        d = {k: v.cuda() for k, v in scenes.batch(seed, 2, 16384).items()}
        d["xy"][0, :64] = d["xy"][0, 64:128]      # points sharing pixels
        d["xy"][1, :16, 0] = 1279.0               # right border of the canvas: two taps fall outside
        xyz_s, f_s = [t.clone() for t in sparse(d["points"], d["image"], d["xy"])]
        xyz_d, f_d = [t.clone() for t in dense(d["points"], d["image"], d["xy"])]
        torch.cuda.synchronize()
        assert torch.equal(xyz_s, xyz_d)
        err = (f_s - f_d).abs().max().item() / f_d.abs().max().item()
        print("backbone sparse vs dense tail: max abs err / scale = %.2e" % err)
        assert err <= 1e-5
    assert sparse.kernel_launches_per_replay < dense.kernel_launches_per_replay + 8
