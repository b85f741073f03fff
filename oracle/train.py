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


def retina_anchors(img_hw, stride, anchor_wh):
    """(nA, nH, nW, 4) cxcywh anchor boxes -- retinanet.py:54-60, :85-90."""
    img_h, img_w = img_hw
    n_a = anchor_wh.shape[0]
    n_h, n_w = int(img_h / stride), int(img_w / stride)
    a_cx = torch.arange(stride / 2, img_w, stride).view(1, 1, n_w, 1)
    a_cy = torch.arange(stride / 2, img_h, stride).view(1, n_h, 1, 1)
    a_wh = anchor_wh.view(n_a, 1, 1, 2)
    return torch.cat([a_cx.expand(n_a, n_h, n_w, 1), a_cy.expand(n_a, n_h, n_w, 1), a_wh.expand(n_a, n_h, n_w, 2)], dim=-1)


def retina_targets_and_loss(t_xywh, cls_logits, gts, img_hw, stride, anchor_wh, pos_thres, neg_thres, angle_loss=None):
    """RetinaLayer's training branch for 'cxcywh' boxes and ONE class (the reference's squeeze(-1), :125, supports
    nothing else) -- retinanet.py:84-160.  t_xywh (B,nA,nH,nW,4), cls_logits (B,nA,nH,nW,1).
    angle_loss: None for 'cxcywh'; 'Periodic_L1' | 'Periodic_L2' | 'Periodic_smoothL1' (models/losses.py:18-98) for
    'cxcywhd' boxes: t_xywh then has 5 columns, GT boxes 5, and the angle part of :133-136, :151-154 is added.
    Returns (per-image list of dict(M_pos, M_neg, gt_idx, tgt_xywh, tgt_cls, cls_penalty_mask) or None, loss, pos)."""
    import math
    import torch.nn.functional as tnf
    n_b, n_a, n_h, n_w = t_xywh.shape[:4]
    anch = retina_anchors(img_hw, stride, anchor_wh)
    loss_xywh, loss_cls, total_pos, per_image = 0, 0, 0, []
    for b, (gt_bbs, cats) in enumerate(gts):
        if gt_bbs.shape[0] == 0:                                               # :97-101
            loss_cls = loss_cls + tnf.binary_cross_entropy_with_logits(cls_logits[b], torch.zeros(n_a, n_h, n_w, 1), reduction='sum')
            per_image.append(None)
            continue
        ious = bboxes_iou(anch.view(-1, 4), gt_bbs[:, :4])                     # :106
        iou_with_gt, gt_idx = ious.max(dim=1)
        iou_with_gt, gt_idx = iou_with_gt.view(n_a, n_h, n_w), gt_idx.view(n_a, n_h, n_w)
        m_pos, m_neg = iou_with_gt > pos_thres, iou_with_gt < neg_thres        # :110-111
        total_pos += int(m_pos.sum())
        g = gt_bbs[gt_idx, :]
        tgt_xywh = torch.zeros(n_a, n_h, n_w, 4)
        tgt_xywh[..., 0:2] = (g[..., 0:2] - anch[..., 0:2]) / anch[..., 2:4]
        tgt_xywh[..., 2:4] = torch.log(g[..., 2:4] / anch[..., 2:4] + 1e-8)
        tgt_cls = torch.zeros(n_a, n_h, n_w, 1)
        tgt_cls[m_pos, cats[gt_idx[m_pos]]] = 1
        logit = cls_logits[b].detach().squeeze(-1)
        need_higher = m_pos & (logit < math.log(0.95 / (1 - 0.95)))
        need_lower = m_neg & (logit > math.log(0.01 / (1 - 0.01)))
        penalty = need_higher | need_lower
        if int(m_pos.sum()) > 0:                                               # fvcore.nn.smooth_l1_loss, beta 0.1
            n = torch.abs(t_xywh[b][m_pos][:, 0:4] - tgt_xywh[m_pos, :])
            im_loss = torch.where(n < 0.1, 0.5 * n ** 2 / 0.1, n - 0.05).sum()
            if angle_loss is not None:
                tgt_angle = g[..., 4] / 180 * math.pi                           # :136
                p_angle = torch.sigmoid(t_xywh[b][m_pos][:, 4]) * 2 * math.pi - math.pi     # :152
                d = torch.remainder(p_angle - tgt_angle[m_pos] - math.pi / 2, math.pi) - math.pi / 2
                if angle_loss == 'Periodic_L1':
                    im_loss = im_loss + torch.abs(d).sum()
                elif angle_loss == 'Periodic_L2':
                    im_loss = im_loss + (d ** 2).sum()
                else:
                    a = torch.abs(d)
                    im_loss = im_loss + torch.where(a < 0.4, 0.5 * a ** 2 / 0.4, a - 0.5 * 0.4).sum()
            loss_xywh = loss_xywh + im_loss
        loss_cls = loss_cls + tnf.binary_cross_entropy_with_logits(cls_logits[b, penalty], tgt_cls[penalty], reduction='sum')
        per_image.append({'M_pos': m_pos, 'M_neg': m_neg, 'gt_idx': gt_idx, 'tgt_xywh': tgt_xywh, 'tgt_cls': tgt_cls,
                          'cls_penalty_mask': penalty})
    return per_image, (loss_xywh + loss_cls) / n_b, total_pos


def rapid_targets(p_xywha, conf_logits, gts, img_hw, stride, anchors_all, indices, n_cls, grid_hw, ignore_thre=0.6):
    """RAPiDLayer's target assignment -- rapid.py:84-167, with the oracle's exact rotated IoU (oracle/iou.py: iou_rot)
    in place of the pycocotools raster (rotated IoU VALUES are "parity unpinned"; the control flow is what this pins).
    p_xywha (B, nA*nH*nW, 5) decoded boxes (degrees), conf_logits (B,nA,nH,nW,1), gts: list of (boxes (n,5), cats (n,)).
    Returns dict(PositiveMask, IgnoredMask, TargetXYWH, TargetAngle, TargetConf, TargetCls, weighted)."""
    import math
    from .iou import iou_rot
    n_b = p_xywha.shape[0]
    n_a, (n_h, n_w) = len(indices), grid_hw
    anchors_all = torch.tensor(anchors_all, dtype=torch.float32)
    anchors = anchors_all[indices, :]
    anch_00wha_all = torch.zeros(len(anchors_all), 5)
    anch_00wha_all[:, 2:4] = anchors_all
    idx_t = torch.tensor(indices)
    pos = torch.zeros(n_b, n_a, n_h, n_w, dtype=torch.bool)
    ign = torch.zeros(n_b, n_a, n_h, n_w, dtype=torch.bool)
    weighted = torch.zeros(n_b, n_a, n_h, n_w)
    t_xywh = torch.zeros(n_b, n_a, n_h, n_w, 4)
    t_angle = torch.zeros(n_b, n_a, n_h, n_w)
    t_conf = torch.zeros(n_b, n_a, n_h, n_w, 1)
    t_cls = torch.zeros(n_b, n_a, n_h, n_w, max(n_cls, 1))
    p5 = p_xywha.view(n_b, n_a, n_h, n_w, 5)
    for b, (gt_bboxes, gt_cats) in enumerate(gts):
        if gt_bboxes.shape[0] == 0:                                            # rapid.py:106-107
            continue
        selected = (conf_logits[b] > -math.log(1 / 0.005 - 1)).squeeze(-1)     # :115-116
        p_sel = p5[b][selected]
        if 0 < len(p_sel) < 1000:                                              # :120
            ious = iou_rot(p_sel.reshape(-1, 5), gt_bboxes)
            ign[b, selected] = ious.max(dim=1).values > ignore_thre            # :123-126
        for gt_bb, gt_c in zip(gt_bboxes, gt_cats):                            # :129
            g0 = gt_bb.clone()
            g0[0:2] = 0
            g0[4] = 0
            anch_idx_all = int(torch.argmax(iou_rot(g0.unsqueeze(0), anch_00wha_all), dim=1))   # :134-136
            if not bool((idx_t == anch_idx_all).any()):
                continue
            ta = anch_idx_all % n_a
            ti, tj = int(gt_bb[0] / stride), int(gt_bb[1] / stride)            # :148-149
            pos[b, ta, tj, ti] = True
            t_xywh[b, ta, tj, ti, 0] = (gt_bb[0] / stride) % 1
            t_xywh[b, ta, tj, ti, 1] = (gt_bb[1] / stride) % 1
            t_xywh[b, ta, tj, ti, 2] = torch.log(gt_bb[2] / anchors[ta, 0] + 1e-8)
            t_xywh[b, ta, tj, ti, 3] = torch.log(gt_bb[3] / anchors[ta, 1] + 1e-8)
            t_angle[b, ta, tj, ti] = gt_bb[4] / 180 * math.pi
            t_conf[b, ta, tj, ti] = 1
            if n_cls > 0:
                t_cls[b, ta, tj, ti, gt_c] = 1
            weighted[b, ta, tj, ti] = 2 - gt_bb[2] * gt_bb[3] / (img_hw[0] * img_hw[1])
    return {'PositiveMask': pos, 'IgnoredMask': ign, 'TargetXYWH': t_xywh, 'TargetAngle': t_angle, 'TargetConf': t_conf,
            'TargetCls': t_cls, 'weighted': weighted}


def uv5_targets_and_loss(t_bbox, conf_logits, cls_logits, p_bbox, gts, stride, anchors_all, indices, n_cls, conf_target, negative_thres):
    """Training branch of DetectLayer, models/detlayers/uv5.py:115-224, with the reference's per-image / per-GT loops.
    t_bbox (B,nA,nH,nW,4) logits, p_bbox (B, nA*nH*nW, 4) decoded boxes, gts: list of (boxes (n,4), cats (n,)).
    Returns dict(TargetConf, IgnoredMask | None, loss, valid_gt_num)."""
    import torch.nn.functional as tnf
    bce = tnf.binary_cross_entropy_with_logits
    n_b, n_a, n_h, n_w = t_bbox.shape[:4]
    anchors_all = torch.tensor(anchors_all, dtype=torch.float32)
    ind = torch.tensor(indices).long()
    anchors = anchors_all[ind, :]
    anch_00wh_all = torch.zeros(len(anchors_all), 4)
    anch_00wh_all[:, 2:4] = anchors_all
    tgt_conf = torch.zeros(n_b, n_a, n_h, n_w, 1)
    ignored = torch.zeros(n_b, n_a, n_h, n_w, dtype=torch.bool) if conf_target == 'zero-one' else None
    loss_xy = loss_wh = loss_cls = 0
    valid = 0
    for b, (gt_bboxes, gt_cls_idx) in enumerate(gts):
        if gt_bboxes.shape[0] == 0:                                            # :131-133
            continue
        for gt_bb, gt_c in zip(gt_bboxes, gt_cls_idx):
            g00 = gt_bb.clone()
            g00[0:2] = 0
            a_all = torch.argmax(bboxes_iou(g00, anch_00wh_all), dim=1).squeeze().item()   # :141-143
            if not (ind == a_all).any():
                continue
            ta = a_all % n_a
            ti, tj = (gt_bb[0] / stride).long(), (gt_bb[1] / stride).long()
            valid += 1
            tb = t_bbox[b, ta, tj, ti]
            loss_xy = loss_xy + bce(tb[:2], ((gt_bb[:2] / stride) % 1 + 0.5) / 2, reduction='sum')        # :161-163
            loss_wh = loss_wh + bce(tb[2:4], torch.sqrt(gt_bb[2:4] / anchors[ta, :]) / 2, reduction='sum')  # :164-166
            if n_cls > 0:
                tc = cls_logits[b, ta, tj, ti]
                tgt = torch.zeros_like(tc)
                tgt[..., gt_c] = 1
                loss_cls = loss_cls + bce(tc, tgt)                              # :178 (mean over the classes)
            if conf_target == 'zero-one':
                tgt_conf[b, ta, tj, ti] = 1
        iou_with_gt, _ = bboxes_iou(p_bbox[b], gt_bboxes).max(dim=1)           # :186-190
        if conf_target == 'IoU':
            tgt_conf[b] = iou_with_gt.view(n_a, n_h, n_w, 1)
        else:
            ignored[b] = (iou_with_gt > negative_thres).view(n_a, n_h, n_w)
    if conf_target == 'IoU':
        loss_conf = bce(conf_logits, tgt_conf, reduction='sum')
    else:
        pos = tgt_conf.squeeze(-1).bool()
        pen = pos | (~ignored)
        loss_conf = bce(conf_logits[pen], tgt_conf[pen], reduction='sum')
    loss = (loss_xy + loss_wh + loss_conf + loss_cls) / n_b
    return {'TargetConf': tgt_conf, 'IgnoredMask': ignored, 'loss': loss, 'valid_gt_num': valid}
