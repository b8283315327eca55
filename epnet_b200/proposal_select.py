"""Proposal selection around the batched NMS -- the part of /root/reference/lib/rpn/proposal_layer.py that follows box decoding
(`ProposalLayer.forward` :33-56, `distance_based_proposal` :58-119, `score_based_proposal` :121-143), restated without the
per-scene Python loop and without any host synchronisation.

The reference, per scene: sort by score, split the sorted proposals into the distance bands (0, 40] and (40, 80] m (z of the box),
keep the first 70 % / 30 % of RPN_PRE_NMS_TOP_N of each band (an empty far band is replaced by the NEXT near-band proposals,
:96-104), run NMS per band (one cudaMalloc, one blocking D2H copy of the mask and one host loop each), keep the first
70 % / 30 % of RPN_POST_NMS_TOP_N, concatenate, zero-pad.  That is 2 x B NMS calls with 2 x B device->host round trips.

Here every step is a batched tensor op: ranks inside a band come from a cumulative sum of the band mask, the selected proposals
are scattered into a padded (B, pre_n, 7) block with per-scene counts, ONE `nms_batched` launch per band handles all scenes and
stops at the band's post-NMS quota, and the survivors are scattered to their final rows.  Outputs are identical to the
reference's (same proposals, same order, zero padding) -- tests/test_zz_proposal_select.py checks this against a literal
restatement of the reference loop.

Box decoding (`decode_bbox_target`) and the network heads are out of scope (only the NMS row and its caller's
selection logic are "next")."""
import torch

from . import iou3d_utils

NMS_RANGES = (0.0, 40.0, 80.0)   # proposal_layer.py:65


def _band_quota(total):
    first = int(total * 0.7)     # proposal_layer.py:67,69
    return first, total - first


def _select_band(in_band, quota, skip):
    """in_band (B,N) bool over score-sorted proposals -> (src (B,quota) indices into the sorted order, count (B,)): the
    band members with rank in [skip, skip+quota), in order.  skip is (B,) int64."""
    B, N = in_band.shape
    rank = torch.cumsum(in_band.to(torch.int64), dim=1) - 1 - skip.view(B, 1)
    take = in_band & (rank >= 0) & (rank < quota)
    count = take.sum(dim=1)
    # scatter position j of the sorted order to slot rank[j]; everything else goes to a dump slot
    slot = torch.where(take, rank, torch.full_like(rank, quota))
    src = torch.zeros((B, quota + 1), dtype=torch.int64, device=in_band.device)
    src.scatter_(1, slot, torch.arange(N, device=in_band.device).expand(B, N))
    return src[:, :quota], count


def select_proposals(scores, proposals, pre_nms_top_n, post_nms_top_n, nms_thresh, distance_based=True, nms_type="rotate"):
    """scores (B,N), proposals (B,N,7) [x, y, z, h, w, l, ry] -> ret_bbox3d (B, post_nms_top_n, 7), ret_scores (B, post_nms_top_n),
    zero-padded, as `ProposalLayer.forward` returns them (proposal_layer.py:39-56).  nms_type: 'rotate' | 'normal'
    (cfg.RPN.NMS_TYPE, lib/config.py:90).  The NMS is iou3d_utils.nms_batched (libepnet_b200.so); the CPU tests replace that
    module attribute with the oracle (tests/backend_swap.py)."""
    if nms_type not in ("rotate", "normal"):
        raise NotImplementedError(nms_type)           # proposal_layer.py:108-109
    if scores.dim() != 2 or proposals.shape != scores.shape + (7,):
        raise ValueError("scores must be (B,N) and proposals (B,N,7)")
    nms = iou3d_utils.nms_batched  # looked up at call time
    B, N = scores.shape
    dev = scores.device
    order = torch.sort(scores, dim=1, descending=True)[1]
    scores_ordered = torch.gather(scores, 1, order)
    proposals_ordered = torch.gather(proposals, 1, order.unsqueeze(-1).expand(B, N, 7))

    zero = torch.zeros((B,), dtype=torch.int64, device=dev)
    if distance_based:
        dist = proposals_ordered[:, :, 2]
        near = (dist > NMS_RANGES[0]) & (dist <= NMS_RANGES[1])
        far = (dist > NMS_RANGES[1]) & (dist <= NMS_RANGES[2])
        pre = _band_quota(pre_nms_top_n)
        post = _band_quota(post_nms_top_n)
        far_empty = far.sum(dim=1) == 0
        # an empty far band is served from the near band, after the near band's own quota (proposal_layer.py:96-104)
        far_or_near = torch.where(far_empty.view(B, 1), near, far)
        bands = [(near, pre[0], zero, post[0]), (far_or_near, pre[1], torch.where(far_empty, torch.full_like(zero, pre[0]), zero), post[1])]
    else:
        everything = torch.ones((B, N), dtype=torch.bool, device=dev)
        bands = [(everything, min(pre_nms_top_n, N), zero, post_nms_top_n)]

    ret_bbox3d = scores.new_zeros((B, post_nms_top_n + 1, 7))     # last row: dump slot for padding entries
    ret_scores = scores.new_zeros((B, post_nms_top_n + 1))
    filled = zero
    for in_band, pre_n, skip, post_n in bands:
        if pre_n <= 0 or post_n <= 0:
            continue
        src, count = _select_band(in_band, pre_n, skip)
        cur_scores = torch.gather(scores_ordered, 1, src)
        cur_props = torch.gather(proposals_ordered, 1, src.unsqueeze(-1).expand(B, pre_n, 7))
        boxes_bev = iou3d_utils.boxes3d_to_bev_torch(cur_props.reshape(-1, 7)).view(B, pre_n, 5)
        keep, num = nms(boxes_bev, nms_thresh, max_out=post_n, counts=count.to(torch.int32), rotated=(nms_type == "rotate"))
        keep, num = keep[:, :post_n], num.to(torch.int64).clamp(max=post_n)
        k = keep.shape[1]
        valid = torch.arange(k, device=dev).view(1, k) < num.view(B, 1)
        safe = torch.where(valid, keep, torch.zeros_like(keep))
        dest = torch.where(valid, filled.view(B, 1) + torch.arange(k, device=dev).view(1, k), torch.full_like(keep, post_nms_top_n))
        ret_scores.scatter_(1, dest, torch.gather(cur_scores, 1, safe))
        ret_bbox3d.scatter_(1, dest.unsqueeze(-1).expand(B, k, 7), torch.gather(cur_props, 1, safe.unsqueeze(-1).expand(B, k, 7)))
        filled = filled + num
    ret_bbox3d[:, post_nms_top_n] = 0
    ret_scores[:, post_nms_top_n] = 0
    return ret_bbox3d[:, :post_nms_top_n].contiguous(), ret_scores[:, :post_nms_top_n].contiguous()
