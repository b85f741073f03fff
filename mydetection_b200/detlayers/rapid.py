import torch

from .. import ops
from ._base import decode_level, no_training


class RAPiDLayer(torch.nn.Module):
    '''RAPiD rotated-box layer, test-mode decode to (cx,cy,w,h,degrees)
    (reference: models/detlayers/rapid.py:11-82).'''
    def __init__(self, level_i: int, cfg: dict):
        super().__init__()
        anchors_all = cfg['model.rapid.anchors']
        self.anchor_indices = list(cfg['model.rapid.anchor_indices'][level_i])
        self.anchors = torch.tensor([anchors_all[i] for i in self.anchor_indices], dtype=torch.float32)
        self.num_anchors = len(self.anchor_indices)
        self.stride = cfg['model.fpn.out_strides'][level_i]
        self.n_cls = cfg['general.num_class']
        assert cfg.get('model.angle.pred_range', 360) == 360
        self.loss_str = ''

    def forward(self, raw: dict, img_size, labels=None):
        assert isinstance(raw, dict)
        t = raw['bbox']
        assert t.shape[1] == self.num_anchors and t.shape[-1] == 5
        if labels is not None:
            no_training('RAPiDLayer')
        return decode_level(ops.KIND_RAPID, raw, self.stride, img_size, self.anchors.tolist()), None
