import numpy as np
import torch
import torch.nn.functional as tnf

from .. import ops
from ._base import decode_level, pack_labels, periodic_angle_loss, stage_raw


class RetinaLayer(torch.nn.Module):
    '''RetinaNet anchor-delta layer (reference: models/detlayers/retinanet.py:12-160).
    Test mode: decode.  Training mode: the anchor-to-GT matching `bboxes_iou(anchors, gt).max(dim=1)`
    (:106-107) runs on the device for the whole batch without its matrix (mydet_iou_aabb_rowmax, one shared anchor
    set); targets and the loss follow from it with elementwise ops (:108-152).  Like the reference, training mode
    returns (None, loss).
    Rotated boxes ('cxcywhd'): the matching uses the first four box parameters (:106), the angle target is the matched
    GT's angle in radians (:133-136) and the periodic angle loss is added at the positives (:151-154).  The reference
    itself cannot construct this variant at HEAD -- its __init__ imports `.losses` from the detlayers package (:36; the
    module lives one level up) and reads a config key no config defines ('model.angle.loss_name') -- so the loss name is taken
    from 'model.angle.loss_name' or, failing that, 'model.angle.loss_angle' (the key the configs do have), and the branch
    is pinned to the reference's unmodified forward() run on a layer whose three rotated attributes were set by hand
    (tests/golden/train_retina_rot.npz).'''
    def __init__(self, level_i: int, cfg: dict):
        super().__init__()
        stride = cfg['model.fpn.out_strides'][level_i]
        base_size = cfg['model.retina.anchor.base'] * stride
        anchors = [(base_size * sc * rt[0], base_size * sc * rt[1])
                   for sc in cfg['model.retina.anchor.scales'] for rt in cfg['model.retina.anchor.ratios']]
        self.anchor_wh = torch.Tensor(anchors)
        self.num_anchors = len(anchors)
        self.positive_thres = cfg.get('model.retina.anchor.positive_threshold', 0.5)
        self.negative_thres = cfg.get('model.retina.anchor.negative_threshold', 0.5)
        self.stride = stride
        self.n_cls = cfg['general.num_class']
        self.pred_bbox_format = cfg['general.pred_bbox_format']
        self.n_bbparam = cfg['general.bbox_param']
        if self.pred_bbox_format == 'cxcywhd':
            self.loss_angle = periodic_angle_loss(cfg.get('model.angle.loss_name', cfg.get('model.angle.loss_angle', 'Periodic_L1')))
        self.loss_str = ''

    def anchor_boxes(self, img_size, n_h, n_w, device):
        '''(nA, nH, nW, 4) cxcywh anchors of this level (:54-60, :85-90).'''
        img_h, img_w = img_size
        a_cx = torch.arange(self.stride / 2, img_w, self.stride, device=device).view(1, 1, n_w, 1)
        a_cy = torch.arange(self.stride / 2, img_h, self.stride, device=device).view(1, n_h, 1, 1)
        a_wh = self.anchor_wh.to(device).view(self.num_anchors, 1, 1, 2)
        n_a = self.num_anchors
        return torch.cat([a_cx.expand(n_a, n_h, n_w, 1), a_cy.expand(n_a, n_h, n_w, 1), a_wh.expand(n_a, n_h, n_w, 2)], dim=-1)

    def forward(self, raw: dict, img_size, labels=None):
        img_h, img_w = img_size
        n_a = self.num_anchors
        n_h, n_w = int(img_h / self.stride), int(img_w / self.stride)
        n_b = raw['bbox'].shape[0]
        assert raw['bbox'].shape == (n_b, n_a, n_h, n_w, self.n_bbparam)
        assert raw['class'].shape == (n_b, n_a, n_h, n_w, self.n_cls)
        if labels is None:
            preds = decode_level(ops.KIND_RETINA, raw, self.stride, img_size, self.anchor_wh.tolist(), keys=('bbox', 'class'))
            return preds, None
        if self.pred_bbox_format not in ('cxcywh', 'cxcywhd'):
            raise NotImplementedError()
        rotated = self.pred_bbox_format == 'cxcywhd'
        assert isinstance(labels, list) and len(labels) == n_b
        staged = stage_raw(raw, ('bbox', 'class'), detach=False)
        t_xywh, cls_logits = staged['bbox'], staged['class']
        dev = t_xywh.device
        anch = self.anchor_boxes(img_size, n_h, n_w, dev)
        gt_full, gt_cls, counts = pack_labels(labels, 5 if rotated else 4, dev)
        gt_box = gt_full[..., :4].contiguous()
        iou_with_gt, gt_idx = ops.iou_rowmax(anch.reshape(-1, 4), gt_box, counts)              # :106-107, whole batch
        iou_with_gt, gt_idx = iou_with_gt.view(n_b, n_a, n_h, n_w), gt_idx.view(n_b, n_a, n_h, n_w)
        has_gt = (counts > 0).view(n_b, 1, 1, 1)
        m_pos = (iou_with_gt > self.positive_thres) & has_gt                                   # :110
        m_neg = (iou_with_gt < self.negative_thres) & has_gt                                   # :111
        bi = torch.arange(n_b, device=dev).view(n_b, 1, 1, 1).expand_as(gt_idx)
        gsel = gt_idx.clamp(min=0)
        g = gt_box[bi, gsel]                                                                   # (B,nA,nH,nW,4) matched GT
        tgt_xywh = torch.cat([(g[..., 0:2] - anch[..., 0:2]) / anch[..., 2:4],
                              torch.log(g[..., 2:4] / anch[..., 2:4] + 1e-8)], dim=-1)         # :118-120
        tgt_cls = torch.zeros(n_b, n_a, n_h, n_w, self.n_cls, device=dev)
        pb, pa, ph, pw = m_pos.nonzero(as_tuple=True)
        tgt_cls[pb, pa, ph, pw, gt_cls[pb, gsel[pb, pa, ph, pw]]] = 1                          # :122-123
        # only predictions that are not good enough yet are penalised (:125-131); the reference's squeeze(-1) makes
        # this branch single-class only, and so is this one
        assert self.n_cls == 1, 'RetinaLayer training: the reference supports num_class == 1 only (retinanet.py:125)'
        logit = cls_logits.detach().squeeze(-1)
        need_higher = m_pos & (logit < float(np.log(0.95 / (1 - 0.95))))
        need_lower = m_neg & (logit > float(np.log(0.01 / (1 - 0.01))))
        penalty = need_higher | need_lower | ~has_gt                                           # images without GT: everything (:97-101)
        self.targets = {'M_pos': m_pos, 'M_neg': m_neg, 'gt_idx': gt_idx, 'tgt_xywh': tgt_xywh, 'tgt_cls': tgt_cls,
                        'cls_penalty_mask': penalty}
        # loss (:137-152): smooth-L1 (fvcore's, beta 0.1) at the positives, BCE on the penalised cells
        err = torch.abs(t_xywh[m_pos][:, 0:4] - tgt_xywh[m_pos])
        loss_xywh = torch.where(err < 0.1, 0.5 * err.pow(2) / 0.1, err - 0.05).sum()
        if rotated:
            tgt_angle = gt_full[bi, gsel][..., 4] / 180 * np.pi                               # :133-136, radians
            self.targets['tgt_angle'] = tgt_angle
            if bool(m_pos.any()):
                p_angle = torch.sigmoid(t_xywh[m_pos][:, 4]) * 2 * np.pi - np.pi              # :152
                loss_xywh = loss_xywh + self.loss_angle(p_angle, tgt_angle[m_pos])             # :153-154
        loss_cls = tnf.binary_cross_entropy_with_logits(cls_logits[penalty], tgt_cls[penalty], reduction='sum')
        loss = (loss_xywh + loss_cls) / n_b
        total_pos, total = int(m_pos.sum()), int((counts > 0).sum()) * n_a * n_h * n_w
        self.loss_str = f'level_{n_h}x{n_w} pos {total_pos}/{total}: xywh {loss_xywh:.3f}, class {loss_cls:.3f}'
        return None, loss
