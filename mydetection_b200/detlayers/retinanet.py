import torch

from .. import ops
from ._base import decode_level, no_training


class RetinaLayer(torch.nn.Module):
    '''RetinaNet anchor-delta layer, test-mode decode (reference: models/detlayers/retinanet.py:12-82).'''
    def __init__(self, level_i: int, cfg: dict):
        super().__init__()
        stride = cfg['model.fpn.out_strides'][level_i]
        base_size = cfg['model.retina.anchor.base'] * stride
        anchors = [(base_size * sc * rt[0], base_size * sc * rt[1])
                   for sc in cfg['model.retina.anchor.scales'] for rt in cfg['model.retina.anchor.ratios']]
        self.anchor_wh = torch.Tensor(anchors)
        self.num_anchors = len(anchors)
        self.stride = stride
        self.n_cls = cfg['general.num_class']
        self.pred_bbox_format = cfg['general.pred_bbox_format']
        self.n_bbparam = cfg['general.bbox_param']
        self.loss_str = ''

    def forward(self, raw: dict, img_size, labels=None):
        img_h, img_w = img_size
        n_h, n_w = int(img_h / self.stride), int(img_w / self.stride)
        n_b = raw['bbox'].shape[0]
        assert raw['bbox'].shape == (n_b, self.num_anchors, n_h, n_w, self.n_bbparam)
        assert raw['class'].shape == (n_b, self.num_anchors, n_h, n_w, self.n_cls)
        if labels is not None:
            no_training('RetinaLayer')
        preds = decode_level(ops.KIND_RETINA, raw, self.stride, img_size, self.anchor_wh.tolist(),
                             keys=('bbox', 'class'))
        return preds, None
