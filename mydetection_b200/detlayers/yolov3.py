import torch
import torch.nn.functional as tnf

from .. import ops
from ._base import decode_level, last_writer, pack_labels, stage_raw


class YOLOLayer(torch.nn.Module):
    '''YOLOv3 detection layer (reference: models/detlayers/yolov3.py:9-163).
    Test mode: decode.  Training mode: both IoU steps of the target assignment run on the device without
    their matrices (mydet_iou_aabb_rowmax) -- GT-to-anchor matching (:91-95) and the ignore mask
    max_GT IoU(pred, GT) < thr (:106-109) -- where the reference moves every prediction to the CPU; the
    handful of per-GT targets are scattered with index ops and the loss is the reference's (:137-156).'''
    def __init__(self, level_i: int, cfg: dict):
        super().__init__()
        anchors_all = cfg['model.yolo.anchors']
        self.indices = list(cfg['model.yolo.anchor_indices'][level_i])
        self.anchors = torch.tensor([anchors_all[i] for i in self.indices], dtype=torch.float32)
        self.anch_00wh_all = torch.zeros(len(anchors_all), 4)
        self.anch_00wh_all[:, 2:4] = torch.tensor(anchors_all, dtype=torch.float32)
        self.ignore_thre = cfg.get('model.yolo.anchor.negative_threshold', 0.7)
        self.num_anchors = len(self.indices)
        self.stride = cfg['model.fpn.out_strides'][level_i]
        self.n_cls = cfg['general.num_class']
        self.loss_str = ''

    def forward(self, raw: dict, img_size, labels=None):
        assert isinstance(raw, dict)
        t = raw['bbox']
        assert t.shape[1] == self.num_anchors and t.shape[-1] == 4
        preds = decode_level(ops.KIND_YOLO, raw, self.stride, img_size, self.anchors.tolist())
        if labels is None:
            return preds, None
        assert isinstance(labels, list) and len(labels) == t.shape[0]
        staged = stage_raw(raw, ('bbox', 'conf', 'class'), detach=False)
        t_xywh, conf_logits, cls_logits = staged['bbox'], staged['conf'], staged['class']
        n_b, n_a, n_h, n_w = t_xywh.shape[:4]
        dev = t_xywh.device
        gt_box, gt_cls, counts = pack_labels(labels, 4, dev)
        n_g = gt_box.shape[1]
        live = torch.arange(n_g, device=dev)[None, :] < counts[:, None]                    # (B,G) real GT rows

        # GT -> anchor: arg-max IoU of (0,0,w,h) with ALL anchors; this level owns a GT iff the winner is
        # one of its anchors (:91-101)
        gt_00wh = gt_box.clone()
        gt_00wh[..., 0:2] = 0
        anch = self.anch_00wh_all.to(dev)[None].expand(n_b, -1, -1).contiguous()
        _, best_n_all = ops.iou_rowmax(gt_00wh, anch)
        best_n = best_n_all % self.num_anchors
        valid = live & torch.isin(best_n_all, torch.tensor(self.indices, device=dev))
        has_valid = valid.any(dim=1)
        valid_gt_num = int(valid.sum())

        # ignore mask: predictions that already overlap some GT by >= thr are not penalised -- only for images
        # that have a GT owned by this level (the reference `continue`s before it otherwise, :102-109)
        iou_with_gt, _ = ops.iou_rowmax(preds['bbox'], gt_box, counts, want_arg=False)
        conf_loss_mask = (iou_with_gt < self.ignore_thre).view(n_b, n_a, n_h, n_w) | ~has_valid.view(n_b, 1, 1, 1)

        # targets of the owned GTs (:111-132)
        gt_mask = torch.zeros(n_b, n_a, n_h, n_w, dtype=torch.bool, device=dev)
        weighted = torch.zeros(n_b, n_a, n_h, n_w, device=dev)
        tgt_xywh = torch.zeros(n_b, n_a, n_h, n_w, 4, device=dev)
        tgt_conf = torch.zeros(n_b, n_a, n_h, n_w, 1, device=dev)
        tgt_cls = torch.zeros(n_b, n_a, n_h, n_w, self.n_cls, device=dev)
        bi, gi = valid.nonzero(as_tuple=True)
        if bi.numel():
            g = gt_box[bi, gi]
            grid_tx, grid_ty = g[:, 0] / self.stride, g[:, 1] / self.stride
            ti, tj = grid_tx.long().clamp(max=n_w - 1), grid_ty.long().clamp(max=n_h - 1)
            tn = best_n[bi, gi]
            anchors = self.anchors.to(dev)
            conf_loss_mask[bi, tn, tj, ti] = True
            gt_mask[bi, tn, tj, ti] = True
            if self.n_cls > 0:
                tgt_cls[bi, tn, tj, ti, gt_cls[bi, gi]] = 1                                  # classes accumulate
            # several GTs aimed at one cell: the last one (GT order) owns the other targets, as on the CPU
            keep = last_writer(((bi * n_a + tn) * n_h + tj) * n_w + ti, n_b * n_a * n_h * n_w)
            bi, tn, tj, ti, g, grid_tx, grid_ty = bi[keep], tn[keep], tj[keep], ti[keep], g[keep], grid_tx[keep], grid_ty[keep]
            tgt_xywh[bi, tn, tj, ti, 0] = grid_tx - grid_tx.floor()
            tgt_xywh[bi, tn, tj, ti, 1] = grid_ty - grid_ty.floor()
            tgt_xywh[bi, tn, tj, ti, 2] = torch.log(g[:, 2] / anchors[tn, 0] + 1e-8)
            tgt_xywh[bi, tn, tj, ti, 3] = torch.log(g[:, 3] / anchors[tn, 1] + 1e-8)
            tgt_conf[bi, tn, tj, ti] = 1
            img_area = img_size[0] * img_size[1]
            weighted[bi, tn, tj, ti] = 2 - g[:, 2] * g[:, 3] / img_area
        weighted = weighted.unsqueeze(-1)
        self.targets = {'gt_mask': gt_mask, 'conf_loss_mask': conf_loss_mask, 'tgt_xywh': tgt_xywh, 'tgt_cls': tgt_cls,
                        'weighted': weighted}

        # loss (:137-156): plain torch functional ops, outside the kernel path
        bce_logits = tnf.binary_cross_entropy_with_logits
        loss_xy = bce_logits(t_xywh[..., 0:2][gt_mask], tgt_xywh[..., 0:2][gt_mask], weight=weighted[gt_mask], reduction='sum')
        loss_wh = (t_xywh[..., 2:4][gt_mask] - tgt_xywh[..., 2:4][gt_mask]).pow(2)
        loss_wh = 0.5 * (weighted[gt_mask] * loss_wh).sum()
        loss_conf = bce_logits(conf_logits[conf_loss_mask], tgt_conf[conf_loss_mask], reduction='sum')
        loss_cls = bce_logits(cls_logits[gt_mask], tgt_cls[gt_mask], reduction='sum') if self.n_cls > 0 else 0
        loss = (loss_xy + loss_wh + loss_conf + loss_cls) / n_b
        ngt = valid_gt_num + 1e-16
        self.loss_str = (f'yolo_{n_h}x{n_w} total {int(ngt)} objects: xy/gt {loss_xy / ngt:.3f}, wh/gt {loss_wh / ngt:.3f}, '
                         f'conf {loss_conf:.3f}, class {loss_cls:.3f}')
        self._assigned_num = valid_gt_num
        return preds, loss
