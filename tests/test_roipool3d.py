"""3D RoI point pooling: oracle sanity on CPU; on the GPU the B200 kernel vs the reference's own
roipool3d kernels (bit-exact: same selection, same copies) and vs the oracle."""
import numpy as np
import pytest
import torch

import oracle


def _case(seed, B=2, N=4096, M=24, C=5):
    from epnet_b200 import scenes
    rng = np.random.RandomState(seed)
    pts = np.stack([scenes.lidar_scene(seed + i, N).numpy() for i in range(B)])
    feats = rng.randn(B, N, C).astype(np.float32)
    boxes = np.zeros((B, M, 7), dtype=np.float32)
    for b in range(B):
        centres = pts[b][rng.randint(0, N, size=M)]
        boxes[b, :, 0] = centres[:, 0] + rng.randn(M) * 0.3
        boxes[b, :, 1] = 1.8 + rng.randn(M) * 0.1           # bottom centre y
        boxes[b, :, 2] = centres[:, 2] + rng.randn(M) * 0.3
        boxes[b, :, 3:6] = np.array([1.6, 1.7, 4.0]) + rng.rand(M, 3) * 1.5
        boxes[b, :, 6] = rng.uniform(-np.pi, np.pi, size=M)
    boxes[0, 0, :3] = [500.0, 1.8, 500.0]                    # far away: empty
    boxes[0, 1, 3:6] = [4.0, 60.0, 60.0]                     # huge: more than 512 points, truncation in index order
    boxes[0, 1, 0], boxes[0, 1, 2], boxes[0, 1, 6] = 0.0, 20.0, 0.3
    return pts, feats, boxes.astype(np.float32)


def test_oracle_roipool3d_properties():
    pts, feats, boxes = _case(7)
    out, flag = oracle.roipool3d(pts, feats, boxes, sampled=64)
    assert flag[0, 0] == 1 and not out[0, 0].any()
    for b in range(2):
        for j in range(boxes.shape[1]):
            if flag[b, j]:
                continue
            rows = out[b, j]
            # every pooled row is one of the cloud's (xyz | feature) rows, first-come order, cyclic repetition
            full = np.concatenate([pts[b], feats[b]], axis=1)
            idx = [int(np.nonzero((full == r).all(axis=1))[0][0]) for r in rows]
            first = idx[:len(dict.fromkeys(idx))]
            assert first == sorted(first)
            cnt = len(first)
            assert idx == [first[k % cnt] for k in range(64)] or cnt == 64


@pytest.mark.gpu
@pytest.mark.parametrize("seed,N,M,C,sampled", [(1, 4096, 24, 5, 512), (2, 16384, 100, 128, 512), (3, 1000, 7, 0, 64), (4, 16384, 64, 1, 512)])
def test_roipool3d_matches_reference_kernel_and_oracle(seed, N, M, C, sampled):
    from epnet_b200 import roipool3d_utils
    from oracle import ref_cuda
    pts, feats, boxes = _case(seed, 2, N, M, max(C, 1))
    feats = feats[:, :, :C]
    tp, tf, tb = torch.from_numpy(pts).cuda(), torch.from_numpy(np.ascontiguousarray(feats)).cuda(), torch.from_numpy(boxes).cuda()
    got, got_flag = roipool3d_utils.roipool3d_gpu(tp, tf, tb, 1.0, sampled_pt_num=sampled)
    torch.cuda.synchronize()
    enlarged = roipool3d_utils.enlarge_box3d(tb.view(-1, 7), 1.0).view(2, M, 7).contiguous()
    assert got.shape == (2, M, sampled, 3 + C)
    if ref_cuda.available() and C > 0:
        ref, ref_flag = ref_cuda.roipool3d(tp, enlarged, tf, sampled)
        assert torch.equal(got_flag, ref_flag)
        assert torch.equal(got, ref)  # same libdevice cosf/sinf, same arithmetic: identical selections and copies
    want, want_flag = oracle.roipool3d(pts, feats, enlarged.cpu().numpy(), sampled)
    np.testing.assert_array_equal(got_flag.cpu().numpy(), want_flag)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
