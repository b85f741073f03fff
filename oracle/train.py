"""Oracle: training-time target assignment of YOLOLayer and FCOSLayer (FCOS2).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates models/detlayers/yolov3.py:71-132 and models/detlayers/fcos2.py:72-143 with the same torch CPU
operators and the same per-image / per-GT loops, so the target tensors are bit-identical to the reference's
(pinned in tests/test_oracle_golden.py against tensors captured from the unmodified reference).
"""
import torch

from .atss import anchor_centers, unclamped_cxcywh
from .iou import bboxes_iou


def yolo_targets(p_xywh, gts, img_hw, stride, anchors_all, indices, ignore_thre, n_cls, grid_hw):
    """p_xywh (B, nA*nH*nW, 4) decoded boxes; gts: list of (boxes (n,4), cats (n,)).
    Returns dict(gt_mask, conf_loss_mask, tgt_xywh, tgt_conf, tgt_cls, weighted, valid_gt_num)."""
    n_b = p_xywh.shape[0]
    n_a, (n_h, n_w) = len(indices), grid_hw
    anchors_all = torch.tensor(anchors_all, dtype=torch.float32)
    anchors = anchors_all[indices, :]
    anch_00wh_all = torch.zeros(len(anchors_all), 4)
    anch_00wh_all[:, 2:4] = anchors_all
    gt_mask = torch.zeros(n_b, n_a, n_h, n_w, dtype=torch.bool)
    conf_loss_mask = torch.ones(n_b, n_a, n_h, n_w, dtype=torch.bool)
    weighted = torch.zeros(n_b, n_a, n_h, n_w)
    tgt_xywh = torch.zeros(n_b, n_a, n_h, n_w, 4)
    tgt_conf = torch.zeros(n_b, n_a, n_h, n_w, 1)
    tgt_cls = torch.zeros(n_b, n_a, n_h, n_w, n_cls)
    valid_gt_num = 0
    for b, (gt_bboxes, gt_cls_idx) in enumerate(gts):
        num_gt = gt_bboxes.shape[0]
        if num_gt == 0:                                                       # yolov3.py:83-85
            continue
        gt_00wh = torch.zeros(num_gt, 4)
        gt_00wh[:, 2:4] = gt_bboxes[:, 2:4]
        best_n_all = torch.argmax(bboxes_iou(gt_00wh, anch_00wh_all), dim=1)  # :91-95
        best_n = best_n_all % n_a
        valid_mask = torch.zeros(num_gt, dtype=torch.bool)
        for ind in indices:
            valid_mask = valid_mask | (best_n_all == ind)                     # :98-100
        if valid_mask.sum() == 0:
            continue
        valid_gt_num += int(valid_mask.sum())
        iou_with_gt, _ = bboxes_iou(p_xywh[b], gt_bboxes).max(dim=1)          # :106-107
        conf_loss_mask[b] = (iou_with_gt < ignore_thre).view(n_a, n_h, n_w)   # :109
        g = gt_bboxes[valid_mask, :]
        grid_tx, grid_ty = g[:, 0] / stride, g[:, 1] / stride                 # :113-114
        ti, tj = grid_tx.long().clamp(max=n_w - 1), grid_ty.long().clamp(max=n_h - 1)
        tn = best_n[valid_mask]
        conf_loss_mask[b, tn, tj, ti] = 1
        gt_mask[b, tn, tj, ti] = 1
        tgt_xywh[b, tn, tj, ti, 0] = grid_tx - grid_tx.floor()
        tgt_xywh[b, tn, tj, ti, 1] = grid_ty - grid_ty.floor()
        tgt_xywh[b, tn, tj, ti, 2] = torch.log(g[:, 2] / anchors[tn, 0] + 1e-8)
        tgt_xywh[b, tn, tj, ti, 3] = torch.log(g[:, 3] / anchors[tn, 1] + 1e-8)
        tgt_conf[b, tn, tj, ti] = 1
        if n_cls > 0:
            tgt_cls[b, tn, tj, ti, gt_cls_idx[valid_mask]] = 1
        weighted[b, tn, tj, ti] = 2 - g[:, 2] * g[:, 3] / (img_hw[0] * img_hw[1])   # :131-132
    return {'gt_mask': gt_mask, 'conf_loss_mask': conf_loss_mask, 'tgt_xywh': tgt_xywh, 'tgt_conf': tgt_conf,
            'tgt_cls': tgt_cls, 'weighted': weighted.unsqueeze(-1), 'valid_gt_num': valid_gt_num}


def _xywh_to_xyxy(bb, cr):
    cx, cy, w, h = bb                                                          # fcos2.py:408-414
    return cx - w * cr / 2, cy - h * cr / 2, cx + w * cr / 2, cy + h * cr / 2


def fcos2_targets(t_ltrb, gts, img_hw, stride, anch_min, anch_max, ignore_thre, n_cls, center_region=0.5):
    """t_ltrb (B,nH,nW,4) raw logits; gts: list of (boxes (n,4) cxcywh, cats (n,)).  fcos2.py:72-143."""
    n_b, n_h, n_w = t_ltrb.shape[:3]
    p_xywh = unclamped_cxcywh(t_ltrb, stride)                                  # :72
    gy, gx = anchor_centers(img_hw, stride)                                    # :76-78
    pos_all = torch.zeros(n_b, n_h, n_w, dtype=torch.bool)
    ign_all = torch.zeros(n_b, n_h, n_w, dtype=torch.bool)
    t_conf = torch.zeros(n_b, n_h, n_w, 1)
    t_box = torch.zeros(n_b, n_h, n_w, 4)
    t_cls = torch.zeros(n_b, n_h, n_w, n_cls)
    for b, (gt_xywh, gt_cls_idx) in enumerate(gts):
        if gt_xywh.shape[0] == 0:
            continue
        areas = gt_xywh[:, 2] * gt_xywh[:, 3]
        order = torch.argsort(areas, descending=True, stable=True)             # :98-100 (ties: by index)
        gt_xywh, gt_cls_idx = gt_xywh[order, :], gt_cls_idx[order]
        iou_with_gt, _ = torch.max(bboxes_iou(p_xywh[b].view(-1, 4), gt_xywh), dim=1)   # :104-105
        ign_all[b] = (iou_with_gt > ignore_thre).view(n_h, n_w)
        for bb, cidx in zip(gt_xywh, gt_cls_idx):
            tx1, ty1, tx2, ty2 = _xywh_to_xyxy(bb, 1)
            tgt = torch.stack([gx - tx1, gy - ty1, tx2 - gx, ty2 - gy], dim=-1)          # :113-116
            cx1, cy1, cx2, cy2 = _xywh_to_xyxy(bb, center_region)
            center_mask = (gx > cx1) & (gx < cx2) & (gy > cy1) & (gy < cy2)              # :123-124
            mx, _ = torch.max(tgt, dim=-1)
            pos_mask = center_mask & (anch_min < mx) & (mx < anch_max)                   # :126-129
            if not pos_mask.any():
                continue
            t_box[b, pos_mask, :] = tgt[pos_mask, :]
            t_conf[b, pos_mask] = 1
            hi, wi = pos_mask.nonzero(as_tuple=True)
            t_cls[b, hi, wi, cidx] = 1
            pos_all[b] = pos_all[b] | pos_mask
    return {'PositiveMask': pos_all, 'IgnoredMask': ign_all, 'TargetConf': t_conf, 'TargetLTRB': t_box, 'TargetCls': t_cls}
