"""Host-side mirror of the reference's models/detlayers/*: same class names, constructor
`(level_i, cfg)` and `forward(raw, img_size, labels=None) -> (preds, loss)`;
the decode arithmetic runs in libmydet's CUDA kernels.  `get_det_layer` mirrors
models/registry.py:119-146."""
from .yolov3 import YOLOLayer
from .fcos import FCOSLayer as FCOSLayerV1
from .fcos2 import FCOSLayer, FCOS_ATSS_Layer
from .rapid import RAPiDLayer
from .retinanet import RetinaLayer
from .uv5 import DetectLayer


def get_det_layer(cfg: dict):
    '''Get final detection layer class (models/registry.py:119-146).'''
    name = cfg['model.pred_layer']
    table = {'YOLO': YOLOLayer, 'Ultralytics': DetectLayer, 'RetinaNet': RetinaLayer, 'FCOS': FCOSLayerV1,
             'FCOS2': FCOSLayer, 'FCOS2_ATSS': FCOS_ATSS_Layer, 'RAPiD': RAPiDLayer}
    if name not in table:
        raise NotImplementedError()
    return table[name]
