"""Proposal selection (epnet_b200/proposal_select.py) against a literal restatement of the reference's per-scene loop
(/root/reference/lib/rpn/proposal_layer.py:33-143).  On the CPU the NMS is injected (the C oracle behind nms_batched's
interface), which exercises every tensor op of the batched, sync-free selection; on the GPU the real nms_batched runs and the
loop restatement uses the reference-signature nms_gpu / nms_normal_gpu."""
import numpy as np
import pytest
import torch

import oracle
from epnet_b200 import iou3d_utils, proposal_select


def scene_proposals(seed, B, N, far=True, near=True):
    """(scores (B,N), proposals (B,N,7)): clustered car-sized boxes, z spread over the two distance bands and beyond."""
    rng = np.random.RandomState(seed)
    objects = 12
    out = np.zeros((B, N, 7), dtype=np.float32)
    for b in range(B):
        lo, hi = (0.5 if near else 41.0), (85.0 if far else 39.0)
        cx, cz = rng.uniform(-30, 30, objects), rng.uniform(lo, hi, objects)
        which = rng.randint(0, objects, N)
        out[b, :, 0] = cx[which] + rng.randn(N) * 0.4
        out[b, :, 1] = 1.7 + rng.randn(N) * 0.1
        out[b, :, 2] = cz[which] + rng.randn(N) * 0.4
        out[b, :, 3] = 1.5 + rng.rand(N) * 0.2
        out[b, :, 4] = 1.6 + rng.randn(N) * 0.1
        out[b, :, 5] = 3.9 + rng.randn(N) * 0.3
        out[b, :, 6] = rng.uniform(-np.pi, np.pi, objects)[which] + rng.randn(N) * 0.1
    scores = rng.permutation(B * N).reshape(B, N).astype(np.float32) / (B * N)    # distinct: the sort order is unambiguous
    return torch.from_numpy(scores), torch.from_numpy(out)


def reference_loop(scores, proposals, pre, post, thresh, distance_based, nms_single):
    """proposal_layer.py:33-143 restated line by line; nms_single(boxes_bev (n,5), scores (n), thresh) -> kept indices."""
    B = scores.shape[0]
    sorted_idxs = torch.sort(scores, dim=1, descending=True)[1]
    ret_bbox3d, ret_scores = scores.new_zeros((B, post, 7)), scores.new_zeros((B, post))
    for k in range(B):
        so, po = scores[k][sorted_idxs[k]], proposals[k][sorted_idxs[k]]
        if distance_based:
            pre_list = [0, int(pre * 0.7), pre - int(pre * 0.7)]
            post_list = [0, int(post * 0.7), post - int(post * 0.7)]
            dist = po[:, 2]
            first_mask = (dist > 0.0) & (dist <= 40.0)
            s_list, p_list = [], []
            for i, (a, b) in enumerate(((0.0, 40.0), (40.0, 80.0)), start=1):
                m = (dist > a) & (dist <= b)
                if m.sum() != 0:
                    cs, cp = so[m][:pre_list[i]], po[m][:pre_list[i]]
                else:
                    if i != 2:      # the reference asserts here; the batched version yields nothing for an empty near band
                        continue
                    cs, cp = so[first_mask][pre_list[1]:][:pre_list[2]], po[first_mask][pre_list[1]:][:pre_list[2]]
                if cs.numel() == 0:
                    continue
                keep = nms_single(iou3d_utils.boxes3d_to_bev_torch(cp), cs, thresh)[:post_list[i]]
                s_list.append(cs[keep])
                p_list.append(cp[keep])
            ss = torch.cat(s_list) if s_list else so[:0]
            ps = torch.cat(p_list) if p_list else po[:0]
        else:
            cs, cp = so[:pre], po[:pre]
            keep = nms_single(iou3d_utils.boxes3d_to_bev_torch(cp), cs, thresh)[:post]
            ss, ps = cs[keep], cp[keep]
        ret_bbox3d[k, :ps.shape[0]] = ps
        ret_scores[k, :ss.shape[0]] = ss
    return ret_bbox3d, ret_scores


def oracle_nms_batched(rotated_default=True):
    """The C oracle behind epnet_b200.iou3d_utils.nms_batched's interface (CPU tensors)."""
    def nms(boxes, thresh, max_out=0, counts=None, rotated=rotated_default, workspace=None):
        S, N = boxes.shape[0], boxes.shape[1]
        keep = torch.full((S, N), -1, dtype=torch.int64)
        num = torch.zeros((S,), dtype=torch.int32)
        for s in range(S):
            n = N if counts is None else int(counts[s])
            k = oracle.nms_bev(boxes[s, :n].numpy(), thresh, rotated)
            if max_out > 0:
                k = k[:max_out]
            keep[s, :len(k)] = torch.from_numpy(k)
            num[s] = len(k)
        return keep, num
    return nms


def _select_on_oracle(*args):
    from backend_swap import swapped
    with swapped(nms_batched=oracle_nms_batched()):
        return proposal_select.select_proposals(*args)


def oracle_nms_single(rotated):
    def nms(boxes_bev, scores, thresh):
        order = scores.sort(0, descending=True)[1]
        return order[torch.from_numpy(oracle.nms_bev(boxes_bev[order].numpy(), thresh, rotated))]
    return nms


@pytest.mark.parametrize("distance_based,nms_type", [(True, "rotate"), (True, "normal"), (False, "rotate")])
def test_selection_equals_reference_loop_cpu(distance_based, nms_type):
    scores, props = scene_proposals(1, 3, 1500)
    rotated = nms_type == "rotate"
    for pre, post, thresh in ((900, 100, 0.7), (1200, 64, 0.85), (5000, 300, 0.5)):
        got = _select_on_oracle(scores, props, pre, post, thresh, distance_based, nms_type)
        want = reference_loop(scores, props, pre, post, thresh, distance_based, oracle_nms_single(rotated))
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        assert got[0].shape == (3, post, 7) and (got[1] > 0).any()


def test_empty_far_band_is_served_from_the_near_band_cpu():
    scores, props = scene_proposals(2, 2, 1200, far=False)             # nothing beyond 40 m
    got = _select_on_oracle(scores, props, 600, 90, 0.8, True, "rotate")
    want = reference_loop(scores, props, 600, 90, 0.8, True, oracle_nms_single(True))
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    # mixed batch: scene 0 has both bands, scene 1 only the near one
    s2, p2 = scene_proposals(3, 2, 1200)
    s2[1], p2[1] = scores[1], props[1]
    got = _select_on_oracle(s2, p2, 600, 90, 0.8, True, "normal")
    want = reference_loop(s2, p2, 600, 90, 0.8, True, oracle_nms_single(False))
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])


def test_edge_cases_cpu():
    scores, props = scene_proposals(4, 2, 300)
    props[1, :, 2] = 200.0                                              # scene 1: everything outside both bands
    got = _select_on_oracle(scores, props, 100, 50, 0.7, True, "rotate")
    want = reference_loop(scores, props, 100, 50, 0.7, True, oracle_nms_single(True))
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    assert not got[0][1].any() and not got[1][1].any()
    # fewer proposals than the quotas
    got = _select_on_oracle(scores[:, :40], props[:, :40], 9000, 300, 0.7, False, "rotate")
    want = reference_loop(scores[:, :40], props[:, :40], 9000, 300, 0.7, False, oracle_nms_single(True))
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    with pytest.raises(NotImplementedError):
        proposal_select.select_proposals(scores, props, 100, 50, 0.7, True, "soft")
    with pytest.raises(ValueError):
        proposal_select.select_proposals(scores, props[:, :, :5], 100, 50, 0.7)


@pytest.mark.gpu
@pytest.mark.parametrize("distance_based,nms_type", [(True, "rotate"), (True, "normal"), (False, "rotate")])
def test_selection_equals_reference_loop_gpu(distance_based, nms_type):
    scores, props = scene_proposals(5, 2, 16384)
    scores, props = scores.cuda(), props.cuda()
    single = iou3d_utils.nms_gpu if nms_type == "rotate" else iou3d_utils.nms_normal_gpu
    for pre, post, thresh in ((9000, 300, 0.7), (12000, 2048, 0.85)):    # lib/config.py:188-190,203-205
        got = proposal_select.select_proposals(scores, props, pre, post, thresh, distance_based, nms_type)
        want = reference_loop(scores, props, pre, post, thresh, distance_based, single)
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        assert (got[1] > 0).sum() > post // 4
