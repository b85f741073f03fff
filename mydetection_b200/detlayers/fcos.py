import torch

from .. import ops
from ._base import decode_level, no_training


class FCOSLayer(torch.nn.Module):
    '''Original FCOS layer (centerness head under raw['center']), test-mode decode
    (reference: models/detlayers/fcos.py:10-68).'''
    def __init__(self, level_i: int, cfg: dict):
        super().__init__()
        self.anch_min = cfg['model.fcos.anchors'][level_i]
        self.anch_max = cfg['model.fcos.anchors'][level_i + 1]
        self.stride = cfg['model.fpn.out_strides'][level_i]
        self.n_cls = cfg['general.num_class']
        self.loss_str = ''

    def forward(self, raw, img_size, labels=None):
        assert isinstance(raw, dict)
        img_h, img_w = img_size
        n_h, n_w = int(img_h / self.stride), int(img_w / self.stride)
        n_b = raw['bbox'].shape[0]
        assert raw['bbox'].shape == (n_b, n_h, n_w, 4)
        assert raw['center'].shape == (n_b, n_h, n_w, 1)
        assert raw['class'].shape == (n_b, n_h, n_w, self.n_cls)
        if labels is not None:
            no_training('FCOSLayer (v1)')
        preds = decode_level(ops.KIND_FCOS, raw, self.stride, img_size, conf_key='center',
                             keys=('bbox', 'center', 'class'))
        return preds, None
