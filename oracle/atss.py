"""Oracle: ATSS anchor-to-GT assignment.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates FCOS_ATSS_Layer's training-time target construction,
models/detlayers/fcos2.py:253-341 with _get_atss_threshold :385-405, for ONE
pyramid level (the reference runs it once per level).  Same torch CPU operators,
so results are bit-identical to the reference when GT areas are tie-free (the
reference's area argsort, :301, is unstable; the oracle breaks ties by index).
"""
import torch

from .iou import bboxes_iou


def anchor_centers(img_hw, stride):
    """Cell-centre meshgrid (gy, gx), each (nH, nW) -- fcos2.py:256-259, :267-271."""
    img_h, img_w = img_hw
    n_h, n_w = img_h // stride, img_w // stride
    xs = torch.linspace(0, img_w, steps=n_w + 1)[:-1] + 0.5 * stride
    ys = torch.linspace(0, img_h, steps=n_h + 1)[:-1] + 0.5 * stride
    gy, gx = torch.meshgrid(ys, xs, indexing='ij')
    return gy.contiguous(), gx.contiguous()


def all_level_anchors(img_hw, strides, anchor_sides):
    """Square anchors (cx,cy,side,side) of every level, row-major -- fcos2.py:265-278."""
    out = []
    for s, side in zip(strides, anchor_sides):
        gy, gx = anchor_centers(img_hw, s)
        wh = torch.ones(gx.numel(), 2) * side
        out.append(torch.cat([gx.reshape(-1, 1), gy.reshape(-1, 1), wh], dim=1))
    return out


def atss_threshold(gt_box, anchors_per_level, k):
    """mean + unbiased std of the IoUs of the k nearest anchors per level -- fcos2.py:385-405."""
    cx, cy = gt_box[0], gt_box[1]
    cand = []
    for anchors in anchors_per_level:
        d2 = (cx - anchors[:, 0]).pow(2) + (cy - anchors[:, 1]).pow(2)       # :396
        _, near = torch.topk(d2, k, largest=False, sorted=False)              # :397
        cand.append(anchors[near, :])
    cand = torch.cat(cand, dim=0)
    ious = bboxes_iou(gt_box.view(1, 4), cand, xyxy=False).squeeze()         # :401
    return ious.mean() + ious.std()                                          # :403-404


def atss_threshold_index_ties(gt_box, anchors_per_level, k):
    """atss_threshold with the DECLARED tie policy of the kernels: anchors at equal distance are taken in ascending
    index order (stable sort).  torch.topk(sorted=False) leaves the choice among equidistant anchors open (it happens
    when a GT centre sits exactly on a cell boundary), and anchors mirrored across the diagonal have different IoUs
    with a non-square GT, so the reference's own result is implementation-defined there."""
    cx, cy = gt_box[0], gt_box[1]
    cand = []
    for anchors in anchors_per_level:
        d2 = (cx - anchors[:, 0]).pow(2) + (cy - anchors[:, 1]).pow(2)
        near = torch.sort(d2, stable=True).indices[:k]
        cand.append(anchors[near, :])
    cand = torch.cat(cand, dim=0)
    ious = bboxes_iou(gt_box.view(1, 4), cand, xyxy=False).squeeze()
    return ious.mean() + ious.std()


def unclamped_cxcywh(t_ltrb, stride):
    """exp-ltrb -> cxcywh WITHOUT clamping, the boxes the ignore mask uses -- fcos2.py:42, :253, :444-450."""
    n_h, n_w = t_ltrb.shape[-3], t_ltrb.shape[-2]
    ltrb = torch.exp(t_ltrb.detach()) * stride
    rows = torch.arange(n_h, dtype=torch.float32).view(n_h, 1) * stride + stride / 2
    cols = torch.arange(n_w, dtype=torch.float32).view(1, n_w) * stride + stride / 2
    return torch.stack([cols + (ltrb[..., 2] - ltrb[..., 0]) / 2,
                        rows + (ltrb[..., 3] - ltrb[..., 1]) / 2,
                        ltrb[..., 0] + ltrb[..., 2],
                        ltrb[..., 1] + ltrb[..., 3]], dim=-1)


def assign_level(level_i, t_ltrb, gts, img_hw, strides, anchor_sides, k, ignore_thre, n_cls):
    """Targets of one level.

    t_ltrb: (B,nH,nW,4) raw regression logits of this level.
    gts: list of B pairs (boxes (nGT,4) cxcywh float32, cats (nGT,) int64).
    Returns dict of PositiveMask, IgnoredMask (B,nH,nW) bool, TargetLTRB (B,nH,nW,4),
    TargetConf (B,nH,nW,1), TargetCls (B,nH,nW,C) float32  -- fcos2.py:279-341.
    """
    stride, side = strides[level_i], anchor_sides[level_i]
    n_b, n_h, n_w, _ = t_ltrb.shape
    pred = unclamped_cxcywh(t_ltrb, stride)
    gy, gx = anchor_centers(img_hw, stride)
    pyramid = all_level_anchors(img_hw, strides, anchor_sides)
    level_anchors = torch.cat([gx.reshape(-1, 1), gy.reshape(-1, 1),
                               torch.ones(n_h * n_w, 2) * side], dim=1)      # :325-328
    pos_all = torch.zeros(n_b, n_h, n_w, dtype=torch.bool)
    ign_all = torch.zeros(n_b, n_h, n_w, dtype=torch.bool)
    t_conf = torch.zeros(n_b, n_h, n_w, 1)
    t_box = torch.zeros(n_b, n_h, n_w, 4)
    t_cls = torch.zeros(n_b, n_h, n_w, n_cls)
    for b, (boxes, cats) in enumerate(gts):
        if boxes.shape[0] == 0:
            continue                                                          # :295-296
        order = torch.argsort(boxes[:, 2] * boxes[:, 3], descending=True, stable=True)  # :299-301
        boxes, cats = boxes[order], cats[order]
        overlap = bboxes_iou(pred[b].reshape(-1, 4), boxes, xyxy=False)       # :306
        ign_all[b] = (overlap.max(dim=1).values > ignore_thre).view(n_h, n_w)  # :307-308
        for box, c in zip(boxes, cats):                                       # big -> small, :312
            x1, y1 = box[0] - box[2] * 1 / 2, box[1] - box[3] * 1 / 2         # :408-414, cr=1
            x2, y2 = box[0] + box[2] * 1 / 2, box[1] + box[3] * 1 / 2
            ltrb = torch.stack([gx - x1, gy - y1, x2 - gx, y2 - gy], dim=-1)  # :316-318
            inside = (ltrb > 0).all(dim=-1)                                   # :321
            thr = atss_threshold(box, pyramid, k)                             # :323
            iou = bboxes_iou(level_anchors, box.view(1, 4), xyxy=False).squeeze(1)  # :329
            pos = (iou > thr).view(n_h, n_w) & inside                         # :330-331
            if not pos.any():
                continue
            t_box[b, pos, :] = ltrb[pos, :]                                   # :335 (later = smaller GT wins)
            t_conf[b, pos] = 1                                                # :337
            hh, ww = pos.nonzero(as_tuple=True)
            t_cls[b, hh, ww, c] = 1                                           # :339-340 (multi-hot accumulates)
            pos_all[b] |= pos                                                 # :342
    return {'PositiveMask': pos_all, 'IgnoredMask': ign_all, 'TargetLTRB': t_box,
            'TargetConf': t_conf, 'TargetCls': t_cls}
