import torch
import torch.nn.functional as tnf

from .. import ops
from ._base import decode_level, pack_labels, stage_raw


# Thresholds of the forward pass in progress: at most ONE entry, written by level 0 and read by the later levels of
# the same pass.  The entry holds a strong reference to the labels list it was computed for (so its id cannot be
# recycled) and is only honoured when that very object comes back with the same geometry / layer configuration.
_THR_CACHE = {}


def _check_shapes(raw, img_size, stride, n_cls):
    img_h, img_w = img_size
    n_h, n_w = int(img_h / stride), int(img_w / stride)
    n_b = raw['bbox'].shape[0]
    assert raw['bbox'].shape == (n_b, n_h, n_w, 4)
    assert raw['conf'].shape == (n_b, n_h, n_w, 1)
    assert raw['class'].shape == (n_b, n_h, n_w, n_cls)
    return n_b, n_h, n_w


def _fcos2_loss(layer, staged, tg, n_b, n_h, n_w):
    """Loss of both FCOS2 layers on top of kernel-built targets: smooth-L1 on log-ltrb, BCE on conf where
    positive or not ignored, BCE on classes at positives (reference fcos2.py:157-189 == :351-382); plain torch
    functional ops, outside the kernel path."""
    t_ltrb, conf_logits, cls_logits = staged['bbox'], staged['conf'], staged['class']
    pos, ign = tg['PositiveMask'], tg['IgnoredMask']
    p_ltrb, t_ltrb_tgt = t_ltrb[pos], torch.log(tg['TargetLTRB'][pos] / layer.stride)
    err = torch.abs(p_ltrb - t_ltrb_tgt)
    beta = 0.2
    loss_bbox = torch.where(err <= beta, 0.5 * err.pow(2) / beta, err - 0.5 * beta).sum()
    penalty = pos | (~ign)
    loss_conf = tnf.binary_cross_entropy_with_logits(conf_logits[penalty], tg['TargetConf'][penalty], reduction='sum')
    loss_cls = tnf.binary_cross_entropy_with_logits(cls_logits[pos], tg['TargetCls'][pos], reduction='sum')
    pos_num, ignored_num = int(pos.sum()), int((ign & (~pos)).sum())
    total = n_b * n_h * n_w
    layer.loss_str = (f'level_{n_h}x{n_w}, pos {pos_num}/{total}, ignored {ignored_num}/{total}: '
                      f'bbox/gt {loss_bbox:.3f}, conf {loss_conf:.3f}, class/gt {loss_cls:.3f}')
    return loss_bbox + loss_conf + loss_cls


class FCOSLayer(torch.nn.Module):
    '''FCOS2 layer (conf + class heads) (reference: models/detlayers/fcos2.py:11-190).
    Test mode: decode.  Training mode: the target maps come from libmydet (mydet_fcos_assign: ignore mask =
    row-max IoU of the un-clamped predictions with the GT, positives by central region and ltrb range); the
    loss on top of them is the reference's (:157-180), divided by the batch size (:181).'''
    def __init__(self, level_i: int, cfg: dict):
        super().__init__()
        self.anch_min = cfg['model.fcos.anchors'][level_i]
        self.anch_max = cfg['model.fcos.anchors'][level_i + 1]
        self.stride = cfg['model.fpn.out_strides'][level_i]
        self.n_cls = cfg['general.num_class']
        self.center_region = 0.5
        self.ignore_thre = cfg['model.fcos2.ignored_threshold']
        self.bb_format = cfg['general.pred_bbox_format']
        self.loss_str = ''

    def assign(self, t_ltrb, img_size, labels):
        gt_box, gt_cls, counts = pack_labels(labels, 4, t_ltrb.device)
        return ops.fcos_assign(t_ltrb, self.stride, img_size, gt_box, gt_cls, counts, self.center_region,
                               self.anch_min, self.anch_max, self.ignore_thre, self.n_cls)

    def forward(self, raw, img_size, labels=None):
        assert isinstance(raw, dict)
        n_b, n_h, n_w = _check_shapes(raw, img_size, self.stride, self.n_cls)
        preds = decode_level(ops.KIND_FCOS, raw, self.stride, img_size)
        if labels is None:
            return preds, None
        assert isinstance(labels, list) and len(labels) == n_b and self.n_cls > 0
        staged = stage_raw(raw, ('bbox', 'conf', 'class'), detach=False)
        tg = self.assign(staged['bbox'].detach(), img_size, labels)
        loss = _fcos2_loss(self, staged, tg, n_b, n_h, n_w) / n_b                  # :181
        return preds, loss


class FCOS_ATSS_Layer(torch.nn.Module):
    '''FCOS2 layer with ATSS sample selection (reference: models/detlayers/fcos2.py:192-382).
    Test mode: decode.  Training mode: the ATSS target maps come from libmydet's kernels
    (mydet_atss_assign); the loss on top of them is the reference's (:351-371).'''
    def __init__(self, level_i: int, cfg: dict):
        super().__init__()
        self.level_i = level_i
        self.strides_all = cfg['model.fpn.out_strides']
        self.stride = cfg['model.fpn.out_strides'][level_i]
        self.n_cls = cfg['general.num_class']
        self.anchors_all = cfg['model.atss.anchors']
        self.anchor = self.anchors_all[level_i]
        self.topk = cfg['model.atss.topk_per_level']
        self.ignore_thre = cfg['model.fcos2.ignored_threshold']
        self.loss_str = ''

    def assign(self, t_ltrb, img_size, labels):
        '''ATSS targets of this level for a list of per-image GT (objects with .bboxes / .cats).'''
        n_b = len(labels)
        gt_box, gt_cls, counts = pack_labels(labels, 4, t_ltrb.device)
        max_gt = gt_box.shape[1]
        # the adaptive threshold of a GT is level independent: the first level's call computes it, the
        # other levels of the same forward pass (same labels list) reuse it
        key = (tuple(img_size), n_b, max_gt, str(t_ltrb.device), int(self.topk),
               tuple(self.strides_all), tuple(float(a) for a in self.anchors_all))
        thr = None
        if self.level_i > 0:                          # level 0 always recomputes: a pass that stopped early leaves nothing behind
            ent = _THR_CACHE.get('pass')
            if ent is not None and ent[0] is labels and ent[1] == key:
                thr = ent[2]
        out = ops.atss_assign(t_ltrb, self.level_i, self.strides_all, self.anchors_all, img_size,
                              gt_box, gt_cls, counts, self.topk, self.ignore_thre,
                              self.n_cls, thr=thr)
        _THR_CACHE.clear()
        if self.level_i + 1 < len(self.strides_all):
            _THR_CACHE['pass'] = (labels, key, out['thr'])
        return out

    def forward(self, raw, img_size, labels=None):
        assert isinstance(raw, dict)
        n_b, n_h, n_w = _check_shapes(raw, img_size, self.stride, self.n_cls)
        preds = decode_level(ops.KIND_FCOS, raw, self.stride, img_size)
        if labels is None:
            return preds, None
        assert isinstance(labels, list) and self.n_cls > 0
        staged = stage_raw(raw, ('bbox', 'conf', 'class'), detach=False)
        tg = self.assign(staged['bbox'].detach(), img_size, labels)
        loss = _fcos2_loss(self, staged, tg, n_b, n_h, n_w)
        return preds, loss
