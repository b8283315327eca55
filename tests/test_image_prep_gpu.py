"""Device-side image preparation (input side): bit-identical to the oracle's restatement of the reference's
host-side float64 preparation (lib/datasets/kitti_dataset.py:44-55 + lib/net/train_functions.py:37), ragged sizes, strided rows,
and the backbone fed with the decoded uint8 image == fed with the host-prepared fp32 tensor."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", [(375, 1242), (370, 1224), (384, 1280), (1, 1)])
def test_uint8_prep_is_bit_identical_to_the_reference_arithmetic(hw):
    from epnet_b200 import image_prep
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, size=(2,) + hw + (3,)).astype(np.uint8)
    img[0, 0, 0] = (0, 128, 255)
    want = oracle.image_prep(img)
    dev = torch.from_numpy(img).cuda()
    nchw = image_prep.normalise_pad(dev)
    nhwc4 = torch.full((2, 384, 1280, 4), 7.0, device="cuda")
    image_prep.normalise_pad(dev, nhwc4=nhwc4)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(nchw.cpu().numpy(), want)
    np.testing.assert_array_equal(nhwc4[..., :3].permute(0, 3, 1, 2).cpu().numpy(), want)
    assert float(nhwc4[..., 3].abs().max()) == 0.0


def test_ragged_sizes_and_strided_rows():
    from epnet_b200 import image_prep
    rng = np.random.RandomState(1)
    canvas = rng.randint(0, 256, size=(3, 376, 1248, 3)).astype(np.uint8)  # allocation larger than any decoded image
    sizes = np.array([[375, 1242], [370, 1224], [376, 1241]], dtype=np.int32)
    want = oracle.image_prep([canvas[i, :h, :w] for i, (h, w) in enumerate(sizes)])
    dev = torch.from_numpy(canvas).cuda()
    out = image_prep.normalise_pad(dev, torch.from_numpy(sizes).cuda())
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    view = dev[:, :300, :1000]  # a strided view: rows 1248*3 bytes apart
    out = image_prep.normalise_pad(view)
    np.testing.assert_array_equal(out.cpu().numpy(), oracle.image_prep(canvas[:, :300, :1000]))


def test_nchw_to_nhwc4():
    from epnet_b200 import image_prep
    x = torch.randn(2, 3, 384, 1280, device="cuda")
    y = image_prep.nchw_to_nhwc4(x)
    assert torch.equal(y[..., :3], x.permute(0, 2, 3, 1)) and float(y[..., 3].abs().max()) == 0.0


def test_bad_arguments():
    from epnet_b200 import image_prep
    with pytest.raises(ValueError):
        image_prep.normalise_pad(torch.zeros(1, 400, 1280, 3, dtype=torch.uint8, device="cuda"))  # taller than the canvas
    with pytest.raises(ValueError):
        image_prep.normalise_pad(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))  # host tensor: no CPU path


@pytest.mark.parametrize("pipeline", [1, 3])
def test_backbone_from_decoded_image_equals_backbone_from_prepared_tensor(pipeline):
    from epnet_b200 import BackboneConfig, Pointnet2MSG, scenes
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    model = Pointnet2MSG(config=BackboneConfig()).cuda().eval()
    runner = model.make_runner(2, 16384, torch.device("cuda"), pipeline=pipeline)
    host = scenes.batch(1000, 2, 16384, with_u8=True)
    np.testing.assert_array_equal(host["image"].numpy(), oracle.image_prep(host["image_u8"].numpy()))  # scenes.py == oracle
    for src in ("device", "pinned"):
        b = {k: (v.cuda() if src == "device" else v.pin_memory()) for k, v in host.items()}
        def run(image):  # a pipelined slot runs on its own stream: wait for it before reading the runner-owned result
            out = runner(b["points"], image, b["xy"])
            if pipeline > 1:
                runner.join()
            torch.cuda.synchronize()
            return [t.clone() for t in out]

        xyz_a, f_a = run(b["image"])
        xyz_b, f_b = run(b["image_u8"])
        assert torch.equal(xyz_a, xyz_b) and torch.equal(f_a, f_b)  # same canvas bit for bit -> same features bit for bit
    # the module path takes the decoded image too
    model.auto_fast_inference = False
    with torch.no_grad():
        d = {k: v.cuda() for k, v in host.items()}
        _, f_m1 = model(d["points"], d["image"], d["xy"].clone())
        _, f_m2 = model(d["points"], d["image_u8"], d["xy"].clone())
    assert torch.equal(f_m1, f_m2)


@pytest.mark.parametrize("pipeline", [1, 3])
def test_backbone_from_ragged_decoded_frames(pipeline):
    """KITTI frames differ in size by a few pixels (kitti_dataset.py:44-55 pads each to the 384x1280 canvas on its own): a batch of
    frames of different sizes in one uint8 buffer + `sizes` must give the features of the per-frame host preparation, bit for bit,
    through the runner and through a pipelined slot."""
    from epnet_b200 import BackboneConfig, Pointnet2MSG, scenes
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    model = Pointnet2MSG(config=BackboneConfig()).cuda().eval()
    runner = model.make_runner(2, 16384, torch.device("cuda"), pipeline=pipeline)
    host = scenes.batch(1300, 2, 16384, with_u8=True)
    frames = host["image_u8"].numpy()
    sizes = np.array([[375, 1242], [370, 1224]], dtype=np.int32)
    buf = np.full_like(frames, 255)  # whatever lies outside a frame's own size must not reach the canvas
    for i, (h, w) in enumerate(sizes):
        buf[i, :h, :w] = frames[i, :h, :w]
    prepared = torch.from_numpy(oracle.image_prep([frames[i, :h, :w] for i, (h, w) in enumerate(sizes)]))

    def run(image, **kw):
        out = runner(host["points"].cuda(), image, host["xy"].cuda(), **kw)
        if pipeline > 1:
            runner.join()
        torch.cuda.synchronize()
        return [t.clone() for t in out]

    xyz_a, f_a = run(prepared.cuda())
    xyz_b, f_b = run(torch.from_numpy(buf).cuda(), sizes=torch.from_numpy(sizes).cuda())
    assert torch.equal(xyz_a, xyz_b) and torch.equal(f_a, f_b)
    xyz_c, f_c = run(torch.from_numpy(buf).pin_memory(), sizes=torch.from_numpy(sizes))  # host buffers: staged by the runner
    assert torch.equal(f_a, f_c)
    _, f_d = run(torch.from_numpy(buf).cuda())  # without sizes the 255 padding is part of the frames: a different image
    assert not torch.equal(f_a, f_d)
