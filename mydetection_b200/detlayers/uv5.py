import torch

from .. import ops
from ._base import decode_level, no_training


class DetectLayer(torch.nn.Module):
    '''Ultralytics / YOLOv5 layer, test-mode decode (reference: models/detlayers/uv5.py:10-91).'''
    def __init__(self, level_i: int, cfg: dict):
        super().__init__()
        anchors_all = cfg['model.detect.anchors']
        self.indices = list(cfg['model.detect.anchor_indices'][level_i])
        self.anchors = torch.tensor([anchors_all[i] for i in self.indices], dtype=torch.float32)
        self.num_anchors = len(self.indices)
        self.stride = cfg['model.fpn.out_strides'][level_i]
        self.n_cls = cfg['general.num_class']
        self.bbox_format = cfg['general.pred_bbox_format']
        self.loss_str = ''

    def forward(self, raw: dict, img_size, labels=None):
        assert isinstance(raw, dict)
        t = raw['bbox']
        assert t.shape[1] == self.num_anchors and t.shape[-1] == 4
        if self.bbox_format != 'cxcywh':
            raise NotImplementedError()
        if labels is not None:
            no_training('DetectLayer')
        return decode_level(ops.KIND_UV5, raw, self.stride, img_size, self.anchors.tolist()), None
