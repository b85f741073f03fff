"""One-off differential fuzz (CPU, needs the reference checkout): oracle/train.py against tensors captured from the
UNMODIFIED YOLOLayer / FCOSLayer (FCOS2) forward(raw, img_size, labels) -- 20 seeds x 3 levels, 3 image shapes, 0-30 GT
per image (images without GT included), ignore thresholds 0.2-0.7."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for sub in ('tests/golden', '', 'tests'):
    sys.path.insert(0, os.path.join(ROOT, sub))
import numpy as np                      # noqa: E402
import torch                            # noqa: E402
import make_golden as m                 # noqa: E402

m.import_reference()
torch.set_grad_enabled(False)
from models.detlayers.yolov3 import YOLOLayer            # noqa: E402
from models.detlayers.fcos2 import FCOSLayer             # noqa: E402
from utils.structures import ImageObjects                # noqa: E402
from oracle import decode as od, train as ot             # noqa: E402

bad = total = 0
for seed in range(20):
    gen = torch.Generator().manual_seed(7000 + seed)
    img_hw = [(256, 320), (320, 320), (384, 256)][seed % 3]
    n_cls, thr = 5, [0.2, 0.3, 0.5, 0.7][seed % 4]
    labels, gts = [], []
    for b in range(3):
        n = 0 if (b == 1 and seed % 2) else int(torch.randint(1, 30, (1,), generator=gen))
        while True:
            bx, ct = m._random_gt(gen, n, img_hw, n_cls, 8.0, 330.0) if n else (torch.zeros(0, 4), torch.zeros(0, dtype=torch.int64))
            break
        labels.append(ImageObjects(bx, ct, bb_format='cxcywh', img_hw=img_hw))
        gts.append((bx, ct))
    cfg = {'model.yolo.anchors': m.YOLO_ANCHORS, 'model.yolo.anchor_indices': m.IDX3, 'model.yolo.anchor.negative_threshold': thr,
           'model.fpn.out_strides': [8, 16, 32], 'general.num_class': n_cls}
    for li, s in enumerate((8, 16, 32)):
        store, raw = m.head_views(gen, 3, 3, img_hw[0] // s, img_hw[1] // s, 4, n_cls, conf_mu=-1.0)
        layer = YOLOLayer(li, cfg)
        names = ['gt_mask', 'conf_loss_mask', 'tgt_xywh', 'tgt_cls', 'weighted']
        _, g = m._grab_forward(layer, names, raw, img_hw, labels)
        anchors = torch.tensor(m.YOLO_ANCHORS, dtype=torch.float32)[m.IDX3[li]]
        box, _, _ = od.decode_yolo(raw, anchors, s, n_cls)
        tg = ot.yolo_targets(box, gts, img_hw, s, m.YOLO_ANCHORS, m.IDX3[li], thr, n_cls, raw['bbox'].shape[2:4])
        for k in names:
            if k not in g:          # no GT in the whole batch: the reference returns before building the targets
                continue
            total += 1
            if not torch.equal(tg[k], g[k]):
                bad += 1
                print('MISMATCH yolo seed', seed, 'level', li, k, int((tg[k] != g[k]).sum()))
        total += 1
        if tg['valid_gt_num'] != int(layer._assigned_num):
            bad += 1
            print('MISMATCH yolo assigned', seed, li, tg['valid_gt_num'], int(layer._assigned_num))
    fa = [0, 64, 128, 256, 512, 100000000]
    fcfg = {'model.fcos.anchors': fa, 'model.fpn.out_strides': [8, 16, 32, 64, 128], 'general.num_class': n_cls,
            'model.fcos2.ignored_threshold': thr, 'general.pred_bbox_format': 'cxcywh'}
    for li, s in zip((0, 1, 2), (8, 16, 32)):
        store, raw = m.head_views(gen, 3, 1, img_hw[0] // s, img_hw[1] // s, 4, n_cls, separate=True)
        names = ['PositiveMask', 'IgnoredMask', 'TargetConf', 'TargetLTRB', 'TargetCls']
        _, g = m._grab_forward(FCOSLayer(li, fcfg), names, raw, img_hw, labels)
        tg = ot.fcos2_targets(store['bbox_nchw'].permute(0, 2, 3, 1), gts, img_hw, s, fa[li], fa[li + 1], thr, n_cls)
        for k in names:
            total += 1
            if not torch.equal(tg[k], g[k]):
                bad += 1
                print('MISMATCH fcos2 seed', seed, 'level', li, k, int((tg[k] != g[k]).sum()))
print('compared', total, 'bad', bad)
