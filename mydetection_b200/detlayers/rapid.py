import numpy as np
import torch
import torch.nn.functional as tnf

from .. import ops
from ._base import decode_level, last_writer, pack_labels, stage_raw


class RAPiDLayer(torch.nn.Module):
    '''RAPiD rotated-box layer (reference: models/detlayers/rapid.py:11-215).
    Test mode: decode to (cx,cy,w,h,degrees).  Training mode: the two rotated-IoU steps of the target assignment
    run on the device with exact polygon clipping (mydet_iou_rot_pairwise) instead of the reference's pycocotools
    raster on the host -- the ignore mask of the confident predictions (:112-127) and the GT-to-anchor matching
    (:129-141); targets are scattered with index ops and the loss is the reference's (:169-199).  Rotated IoU VALUES
    are "parity unpinned" (DESIGN.md section 3); the control flow is pinned against the reference driven by an exact
    polygon-clipping stub (tests/golden/train.npz).'''
    def __init__(self, level_i: int, cfg: dict):
        super().__init__()
        anchors_all = cfg['model.rapid.anchors']
        self.anchor_indices = list(cfg['model.rapid.anchor_indices'][level_i])
        self.anchors = torch.tensor([anchors_all[i] for i in self.anchor_indices], dtype=torch.float32)
        self.anch_00wha_all = torch.zeros(len(anchors_all), 5)
        self.anch_00wha_all[:, 2:4] = torch.tensor(anchors_all, dtype=torch.float32)
        self.num_anchors = len(self.anchor_indices)
        self.stride = cfg['model.fpn.out_strides'][level_i]
        self.n_cls = cfg['general.num_class']
        self.ignore_thre = 0.6
        assert cfg.get('model.angle.pred_range', 360) == 360
        self.wh_sl1_beta = cfg.get('model.rapid.wh_smooth_l1_beta', 1)
        self.loss_str = ''

    def forward(self, raw: dict, img_size, labels=None):
        assert isinstance(raw, dict)
        t = raw['bbox']
        assert t.shape[1] == self.num_anchors and t.shape[-1] == 5
        preds = decode_level(ops.KIND_RAPID, raw, self.stride, img_size, self.anchors.tolist())
        if labels is None:
            return preds, None
        assert isinstance(labels, list) and len(labels) == t.shape[0]
        keys = ('bbox', 'conf', 'class') if self.n_cls > 0 else ('bbox', 'conf')
        staged = stage_raw(raw, keys, detach=False)
        t_xywha, conf_logits = staged['bbox'], staged['conf']
        n_b, n_a, n_h, n_w = t_xywha.shape[:4]
        dev = t_xywha.device
        p_radian = torch.sigmoid(t_xywha[..., 4]) * 2 * np.pi - np.pi                       # :48
        for l in labels:
            assert l._bb_format == 'cxcywhd' and tuple(l.img_hw) == tuple(img_size)         # :100-101
        gt_box, gt_cls, counts = pack_labels(labels, 5, dev)
        n_g = gt_box.shape[1]
        live = torch.arange(n_g, device=dev)[None, :] < counts[:, None]

        # ignore mask (:112-127): rotated IoU of the confident predictions (conf > 0.005, fewer than 1000 of them)
        # with the image's GT; > thr is not penalised
        ignored = torch.zeros(n_b, n_a, n_h, n_w, dtype=torch.bool, device=dev)
        p_xywha = preds['bbox'].view(n_b, n_a, n_h, n_w, 5)
        selected = conf_logits.detach().squeeze(-1) > float(-np.log(1 / 0.005 - 1))
        n_sel = selected.view(n_b, -1).sum(dim=1).tolist()
        n_gt = counts.tolist()
        for b in range(n_b):
            if n_gt[b] > 0 and 0 < n_sel[b] < 1000:
                ious = ops.iou_rot(p_xywha[b][selected[b]], gt_box[b, :n_gt[b]])
                ignored[b][selected[b]] = ious.max(dim=1).values > self.ignore_thre

        # GT -> anchor (:129-141): rotated IoU of (0,0,w,h,0) with ALL anchors; this level owns a GT iff the winner
        # is one of its anchors
        gt_00wh0 = gt_box.reshape(-1, 5).clone()
        gt_00wh0[:, 0:2] = 0
        gt_00wh0[:, 4] = 0
        anch_idx_all = ops.iou_rot(gt_00wh0, self.anch_00wha_all.to(dev)).argmax(dim=1).view(n_b, n_g)
        valid = live & torch.isin(anch_idx_all, torch.tensor(self.anchor_indices, device=dev))
        bi, gi = valid.nonzero(as_tuple=True)
        g = gt_box[bi, gi]
        self.valid_gts = [row.clone() for row in g.cpu()]

        positive = torch.zeros(n_b, n_a, n_h, n_w, dtype=torch.bool, device=dev)
        weighted = torch.zeros(n_b, n_a, n_h, n_w, device=dev)
        tgt_xywh = torch.zeros(n_b, n_a, n_h, n_w, 4, device=dev)
        tgt_angle = torch.zeros(n_b, n_a, n_h, n_w, device=dev)
        tgt_conf = torch.zeros(n_b, n_a, n_h, n_w, 1, device=dev)
        tgt_cls = torch.zeros(n_b, n_a, n_h, n_w, max(self.n_cls, 1), device=dev)
        if bi.numel():
            ta = anch_idx_all[bi, gi] % n_a                                                  # :147
            # the reference indexes without a clamp (a GT centre outside the image raises there)
            ti = (g[:, 0] / self.stride).long().clamp(0, n_w - 1)
            tj = (g[:, 1] / self.stride).long().clamp(0, n_h - 1)
            anchors = self.anchors.to(dev)
            positive[bi, ta, tj, ti] = True
            if self.n_cls > 0:
                tgt_cls[bi, ta, tj, ti, gt_cls[bi, gi]] = 1                                  # classes accumulate (:163)
            # GTs are visited in order: of several GTs aimed at one cell the last one owns the other targets
            keep = last_writer(((bi * n_a + ta) * n_h + tj) * n_w + ti, n_b * n_a * n_h * n_w)
            bi, gi, ta, tj, ti, g = bi[keep], gi[keep], ta[keep], tj[keep], ti[keep], g[keep]
            tgt_xywh[bi, ta, tj, ti, 0] = (g[:, 0] / self.stride) % 1
            tgt_xywh[bi, ta, tj, ti, 1] = (g[:, 1] / self.stride) % 1
            tgt_xywh[bi, ta, tj, ti, 2] = torch.log(g[:, 2] / anchors[ta, 0] + 1e-8)
            tgt_xywh[bi, ta, tj, ti, 3] = torch.log(g[:, 3] / anchors[ta, 1] + 1e-8)
            tgt_angle[bi, ta, tj, ti] = g[:, 4] / 180 * np.pi
            tgt_conf[bi, ta, tj, ti] = 1
            weighted[bi, ta, tj, ti] = 2 - g[:, 2] * g[:, 3] / (img_size[0] * img_size[1])
        self.targets = {'PositiveMask': positive, 'IgnoredMask': ignored, 'TargetXYWH': tgt_xywh, 'TargetAngle': tgt_angle,
                        'TargetConf': tgt_conf, 'TargetCls': tgt_cls, 'weighted': weighted}

        # loss (:169-199): plain torch functional ops, outside the kernel path
        bce_logits = tnf.binary_cross_entropy_with_logits
        w = weighted.unsqueeze(-1)[positive]
        loss_xy = bce_logits(t_xywha[..., 0:2][positive], tgt_xywh[..., 0:2][positive], weight=w, reduction='sum')
        err = torch.abs(t_xywha[..., 2:4][positive] - tgt_xywh[..., 2:4][positive])
        beta = self.wh_sl1_beta
        loss_wh = (torch.cat([w, w], dim=1) * torch.where(err <= beta, 0.5 * err.pow(2) / beta, err - 0.5 * beta)).sum()
        loss_angle = tnf.mse_loss(p_radian[positive], tgt_angle[positive], reduction='sum')  # :35 overrides the cfg loss
        penalty = positive | (~ignored)
        loss_conf = bce_logits(conf_logits[penalty], tgt_conf[penalty], reduction='sum')
        loss_cls = bce_logits(staged['class'][positive], tgt_cls[positive], reduction='sum') if self.n_cls > 0 else 0
        loss = (loss_xy + loss_wh + loss_angle + loss_conf + loss_cls) / n_b
        pos_num = int(positive.sum())
        ngt = pos_num + 1e-16
        ignored_num = int((ignored & (~positive)).sum())
        self.loss_str = (f'level_{n_h}x{n_w} pos/ignore: {int(ngt)}/{ignored_num}, loss: xy/gt {loss_xy / ngt:.3f}, '
                         f'wh/gt {loss_wh / ngt:.3f}, angle/gt {loss_angle / ngt:.3f}, conf {loss_conf:.3f}, class {loss_cls:.3f}')
        self._assigned_num = pos_num
        return preds, loss
