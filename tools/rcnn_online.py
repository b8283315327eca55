"""BASELINE.json configs[3]: full EPNet rcnn_online inference (RPN + proposal layer + roipool3d + RCNN + final rotated NMS), batch 2,
synthetic KITTI-shaped scenes, random-init weights, yaml config (TEST.RPN_POST_NMS_TOP_N = 100).  The network and the post-processing
are the REFERENCE'S OWN, UNMODIFIED Python (baseline/_ref: lib/net/point_rcnn.py, rpn.py, rcnn_net.py, lib/rpn/proposal_layer.py,
lib/utils/*, tools/eval_rcnn.py:548-676 restated for the post-processing) in three arms:

  reference : on the reference's own CUDA kernels (pointnet2_cuda / iou3d_cuda / roipool3d_cuda from oracle/_ref) + ATen grid_sample
  dropin    : the same Python, UNCHANGED, on epnet_b200.install() (the product's three extension modules + grid_sample)
  native    : dropin + the two B200-native replacements of reference Python: backbone_net = epnet_b200.Pointnet2MSG (captured
              inference runner) and ProposalLayer.forward = decode + epnet_b200.proposal_select.select_proposals (batched device NMS)

    python tools/rcnn_online.py [--steps 20] [--warmup 3] [--arms reference,dropin,native]

Prints one JSON line per arm (scenes/s, device-timed) plus the parity of dropin / native against reference on the same inputs."""
import argparse
import contextlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_env  # noqa: E402

B, N = 2, 16384


def build(ref):
    cfg = ref.cfg
    cfg.RCNN.ENABLED = True          # tools/eval_rcnn.py:961-964 (rcnn_online)
    cfg.RPN.ENABLED = True
    cfg.RPN.FIXED = False
    with contextlib.redirect_stdout(sys.stderr):
        torch.manual_seed(0)
        model = ref.point_rcnn.PointRCNN(num_classes=2, use_xyz=True, mode="TEST").cuda().eval()
    return model


def post_process(ref, ret, batch_size):
    """tools/eval_rcnn.py:551-676 without the file output: decode, score, threshold, rotated NMS per scene -> list of kept boxes"""
    from lib.utils.bbox_transform import decode_bbox_target
    import lib.utils.kitti_utils as kitti_utils
    cfg = ref.cfg
    mean_size = torch.from_numpy(cfg.CLS_MEAN_SIZE[0]).cuda()
    roi_boxes3d = ret["rois"]
    rcnn_cls = ret["rcnn_cls"].view(batch_size, -1, ret["rcnn_cls"].shape[1])
    rcnn_reg = ret["rcnn_reg"].view(batch_size, -1, ret["rcnn_reg"].shape[1])
    pred = decode_bbox_target(roi_boxes3d.view(-1, 7), rcnn_reg.view(-1, rcnn_reg.shape[-1]), anchor_size=mean_size,
                              loc_scope=cfg.RCNN.LOC_SCOPE, loc_bin_size=cfg.RCNN.LOC_BIN_SIZE, num_head_bin=cfg.RCNN.NUM_HEAD_BIN,
                              get_xz_fine=True, get_y_by_bin=cfg.RCNN.LOC_Y_BY_BIN, loc_y_scope=cfg.RCNN.LOC_Y_SCOPE,
                              loc_y_bin_size=cfg.RCNN.LOC_Y_BIN_SIZE, get_ry_fine=True).view(batch_size, -1, 7)
    raw_scores = rcnn_cls
    norm_scores = torch.sigmoid(raw_scores)
    inds = norm_scores > cfg.RCNN.SCORE_THRESH
    out = []
    for k in range(batch_size):
        cur = inds[k].view(-1)
        if cur.sum() == 0:
            out.append(pred[k, :0])
            continue
        boxes, scores = pred[k, cur], raw_scores[k, cur].view(-1)
        keep = ref.iou3d_utils.nms_gpu(kitti_utils.boxes3d_to_bev_torch(boxes), scores, cfg.RCNN.NMS_THRESH).view(-1)
        out.append(boxes[keep])
    return pred, out


def make_native(ref, model):
    """the two B200-native replacements of reference Python (everything else stays the reference's)"""
    import epnet_b200
    from epnet_b200 import BackboneConfig, Pointnet2MSG, proposal_select
    from lib.utils.bbox_transform import decode_bbox_target
    cfg = ref.cfg
    ours = Pointnet2MSG(config=BackboneConfig.from_cfg(cfg)).cuda().eval()
    ours.load_state_dict(model.rpn.backbone_net.state_dict(), strict=True)
    model.rpn.backbone_net = ours
    layer = model.rpn.proposal_layer

    def forward(rpn_scores, rpn_reg, xyz):  # lib/rpn/proposal_layer.py:15-56 with the per-scene loop replaced
        batch_size = xyz.shape[0]
        proposals = decode_bbox_target(xyz.view(-1, 3), rpn_reg.view(-1, rpn_reg.shape[-1]), anchor_size=layer.MEAN_SIZE,
                                       loc_scope=cfg.RPN.LOC_SCOPE, loc_bin_size=cfg.RPN.LOC_BIN_SIZE, num_head_bin=cfg.RPN.NUM_HEAD_BIN,
                                       get_xz_fine=cfg.RPN.LOC_XZ_FINE, get_y_by_bin=False, get_ry_fine=False)
        proposals[:, 1] += proposals[:, 3] / 2
        proposals = proposals.view(batch_size, -1, 7)
        return proposal_select.select_proposals(rpn_scores, proposals, cfg[layer.mode].RPN_PRE_NMS_TOP_N, cfg[layer.mode].RPN_POST_NMS_TOP_N,
                                                cfg[layer.mode].RPN_NMS_THRESH, cfg.TEST.RPN_DISTANCE_BASED_PROPOSE, cfg.RPN.NMS_TYPE)

    layer.forward = forward
    return model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--arms", default="reference,dropin,native")
    args = ap.parse_args()
    ref = ref_env.import_reference("reference", with_rcnn=True)
    scenes = ref_env.load_by_path("_epnet_scenes", os.path.join(ROOT, "epnet_b200", "scenes.py"))
    pool = []
    for i in range(6):
        d = scenes.batch(5000 + 10 * i, B, N)
        pool.append({"pts_input": d["points"].cuda(), "img": d["image"].cuda(), "pts_origin_xy": d["xy"].cuda()})
    model = build(ref)
    state = {k: v.clone() for k, v in model.state_dict().items()}
    results, outputs = {}, {}
    for arm in args.arms.split(","):
        if arm != "reference":
            ref.use("product")
        else:
            ref.use("reference")
        m = model
        if arm == "native":
            m = build(ref)
            m.load_state_dict(state, strict=True)
            m = make_native(ref, m)

        def step(i):
            d = pool[i % len(pool)]
            inp = {"pts_input": d["pts_input"], "img": d["img"], "pts_origin_xy": d["pts_origin_xy"].clone()}  # xy is normalised in place
            with torch.no_grad():
                ret = m(inp)
                return ret, post_process(ref, ret, B)

        for i in range(args.warmup):
            out = step(i)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(args.steps):
            out = step(i)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / args.steps
        ret, (pred, kept) = step(0)
        torch.cuda.synchronize()
        outputs[arm] = {"rois": ret["rois"].clone(), "rcnn_cls": ret["rcnn_cls"].clone(), "rcnn_reg": ret["rcnn_reg"].clone(),
                        "pred": pred.clone(), "kept": [k.clone() for k in kept]}
        results[arm] = {"arm": arm, "metric": "EPNet rcnn_online inference scenes/s (RPN + proposals + roipool3d + RCNN + final NMS)", "value": round(B * 1e3 / ms, 2),
                        "unit": "scenes/s", "ms_per_step": round(ms, 3), "batch": B, "steps": args.steps, "rois_per_scene": int(ret["rois"].shape[1]),
                        "final_boxes": [int(k.shape[0]) for k in kept]}
        print(json.dumps(results[arm]), flush=True)
    ref.use("reference")
    if "reference" in outputs:
        base = outputs["reference"]
        for arm, o in outputs.items():
            if arm == "reference":
                continue
            par = {"arm": arm, "vs": "reference", "rois_equal": bool(torch.equal(o["rois"], base["rois"])),
                   "rois_max_abs_diff": float((o["rois"] - base["rois"]).abs().max()),
                   "rcnn_cls_max_abs_diff": float((o["rcnn_cls"] - base["rcnn_cls"]).abs().max()),
                   "rcnn_reg_max_abs_diff": float((o["rcnn_reg"] - base["rcnn_reg"]).abs().max()),
                   "final_box_counts": [int(k.shape[0]) for k in o["kept"]], "reference_final_box_counts": [int(k.shape[0]) for k in base["kept"]],
                   "speedup": round(results[arm]["value"] / results["reference"]["value"], 2)}
            print(json.dumps(par), flush=True)


if __name__ == "__main__":
    main()
