import torch
import torch.nn.functional as tnf

from .. import ops
from ._base import decode_level, pack_labels, stage_raw


class DetectLayer(torch.nn.Module):
    '''Ultralytics / YOLOv5 layer (reference: models/detlayers/uv5.py:10-224).
    Test mode: decode.  Training mode ('best' sample selection, the only one the reference implements): the GT-to-anchor
    matching (:139-143) and the row-max IoU of every decoded prediction with the image's GT (:186-190: the 'IoU'
    confidence target, or the ignore mask of the 'zero-one' target) run on the device without their matrices
    (mydet_iou_aabb_rowmax) -- the reference moves every prediction to the CPU for them; the per-GT losses are gathered
    with index ops and summed as in the reference's GT loop (:157-179).'''
    def __init__(self, level_i: int, cfg: dict):
        super().__init__()
        anchors_all = cfg['model.detect.anchors']
        self.indices = list(cfg['model.detect.anchor_indices'][level_i])
        self.anchors = torch.tensor([anchors_all[i] for i in self.indices], dtype=torch.float32)
        self.anch_00wh_all = torch.zeros(len(anchors_all), 4)
        self.anch_00wh_all[:, 2:4] = torch.tensor(anchors_all, dtype=torch.float32)
        self.num_anchors = len(self.indices)
        self.stride = cfg['model.fpn.out_strides'][level_i]
        self.n_cls = cfg['general.num_class']
        self.sample_selection = cfg.get('model.detect.sample_selection', 'best')
        self.conf_target = cfg.get('model.detect.confidence_target', 'zero-one')
        self.negative_thres = cfg.get('model.detect.negative_threshold', 0.7)
        self.loss_bbox = cfg.get('model.detect.loss_bbox', 'smooth_L1')
        self.bbox_format = cfg['general.pred_bbox_format']
        self.loss_str = ''

    def forward(self, raw: dict, img_size, labels=None):
        assert isinstance(raw, dict)
        t = raw['bbox']
        assert t.shape[1] == self.num_anchors and t.shape[-1] == 4
        if self.bbox_format != 'cxcywh':
            raise NotImplementedError()
        preds = decode_level(ops.KIND_UV5, raw, self.stride, img_size, self.anchors.tolist())
        if labels is None:
            return preds, None
        if self.sample_selection != 'best' or self.loss_bbox != 'smooth_L1' or self.conf_target not in ('IoU', 'zero-one'):
            raise NotImplementedError()                                          # as the reference (:92-93, :152-154, :168-169)
        assert isinstance(labels, list) and len(labels) == t.shape[0]
        staged = stage_raw(raw, ('bbox', 'conf', 'class'), detach=False)
        t_bbox, conf_logits, cls_logits = staged['bbox'], staged['conf'], staged['class']
        n_b, n_a, n_h, n_w = t_bbox.shape[:4]
        dev = t_bbox.device
        bce = tnf.binary_cross_entropy_with_logits
        gt_box, gt_cls, counts = pack_labels(labels, 4, dev)
        live = torch.arange(gt_box.shape[1], device=dev)[None, :] < counts[:, None]
        has_gt = (counts > 0).view(n_b, 1, 1, 1)

        # GT -> anchor over ALL anchors; this level owns the GT iff the winner is one of its anchors (:139-146)
        gt_00wh = gt_box.clone()
        gt_00wh[..., 0:2] = 0
        anch = self.anch_00wh_all.to(dev)[None].expand(n_b, -1, -1).contiguous()
        _, best_all = ops.iou_rowmax(gt_00wh, anch)
        valid = live & torch.isin(best_all, torch.tensor(self.indices, device=dev))
        bi, gi = valid.nonzero(as_tuple=True)
        valid_gt_num = int(bi.numel())
        tgt_conf = torch.zeros(n_b, n_a, n_h, n_w, 1, device=dev)
        loss_xy = loss_wh = loss_cls = torch.zeros((), device=dev)
        if valid_gt_num:
            g = gt_box[bi, gi]
            ta = best_all[bi, gi] % n_a
            ti, tj = (g[:, 0] / self.stride).long(), (g[:, 1] / self.stride).long()      # :148-149
            tb = t_bbox[bi, ta, tj, ti]
            loss_xy = bce(tb[:, 0:2], ((g[:, 0:2] / self.stride) % 1 + 0.5) / 2, reduction='sum')             # :159-161
            loss_wh = bce(tb[:, 2:4], torch.sqrt(g[:, 2:4] / self.anchors.to(dev)[ta]) / 2, reduction='sum')  # :162-164
            if self.n_cls > 0:
                tc = cls_logits[bi, ta, tj, ti]
                onehot = torch.zeros_like(tc)
                onehot[torch.arange(valid_gt_num, device=dev), gt_cls[bi, gi]] = 1
                loss_cls = bce(tc, onehot, reduction='none').mean(dim=1).sum()             # :172-178: a MEAN per GT
            if self.conf_target == 'zero-one':
                tgt_conf[bi, ta, tj, ti] = 1                                               # :181-182
        # confidence target / ignore mask from max_GT IoU(prediction, GT), images with GT only (:131-133, :185-196)
        iou_with_gt, _ = ops.iou_rowmax(preds['bbox'], gt_box, counts, want_arg=False)
        iou_with_gt = iou_with_gt.view(n_b, n_a, n_h, n_w)
        ignored = None
        if self.conf_target == 'IoU':
            tgt_conf = torch.where(has_gt, iou_with_gt, torch.zeros_like(iou_with_gt)).unsqueeze(-1)
            loss_conf = bce(conf_logits, tgt_conf, reduction='sum')
            ignored_num = 0
        else:
            ignored = (iou_with_gt > self.negative_thres) & has_gt
            pos = tgt_conf.squeeze(-1).bool()
            penalty = pos | ~ignored
            loss_conf = bce(conf_logits[penalty], tgt_conf[penalty], reduction='sum')
            ignored_num = int((ignored & ~pos).sum())
        self.targets = {'TargetConf': tgt_conf, 'IgnoredMask': ignored}
        loss = (loss_xy + loss_wh + loss_conf + loss_cls) / n_b
        ngt = valid_gt_num + 1e-16
        self.loss_str = (f'yolo_{n_h}x{n_w} pos/ignore: {int(ngt)}/{ignored_num}: xy/gt {loss_xy / ngt:.3f}, '
                         f'wh/gt {loss_wh / ngt:.3f}, conf {loss_conf:.3f}, class {loss_cls:.3f}')
        self._assigned_num = valid_gt_num
        return preds, loss
