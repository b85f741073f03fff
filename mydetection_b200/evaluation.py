"""First step beyond the hot path: the detection-to-GT matching IoU of the
CEPDOF evaluator, utils/evaluation/cepdof.py:67-99 (`computeIoU`) and :210-243 (its own numpy `iou_rle`),
on the rotated pairwise-IoU kernel of the path (`mydet_iou_rot_pairwise`).

As everywhere in this package the IoU is the exact polygon intersection, not the pycocotools raster the
reference uses (parity unpinned for rotated IoU values); `img_size` is accepted and unused.
"""
import numpy as np
import torch

from . import ops
from .bbox_ops import _cuda_device


def iou_rle(boxes1, boxes2, img_size=2048):
    """IoU between rotated boxes given as lists of [cx, cy, w, h, degree] -> np.array[M, N] float64
    (same signature and return type as utils/evaluation/cepdof.py:210-243)."""
    assert isinstance(boxes1, list) and isinstance(boxes2, list)
    b1 = np.array(boxes1, dtype=np.float64).reshape(-1, 5)
    b2 = np.array(boxes2, dtype=np.float64).reshape(-1, 5)
    if b1.shape[0] == 0 or b2.shape[0] == 0:
        return np.zeros((b1.shape[0], b2.shape[0]))
    dev = _cuda_device()
    t1 = torch.from_numpy(b1).to(dev, torch.float32)
    t2 = torch.from_numpy(b2).to(dev, torch.float32)
    return ops.iou_rot(t1, t2).cpu().numpy()


def compute_iou(dts, gts, max_dets=100, img_size=2048):
    """CEPDOFeval.computeIoU for one (image, category): detections sorted by descending score with a STABLE
    sort (`np.argsort(..., kind='mergesort')`, cepdof.py:78), capped at max_dets (:80-81), IoU of every kept
    detection with every GT (:92-98).  dts / gts are lists of dicts with 'bbox' (and 'score' for dts).
    Returns (ious np.array[D, G], order): `order` are the indices of the kept detections."""
    if len(gts) == 0 and len(dts) == 0:
        return [], np.zeros(0, dtype=np.int64)
    order = np.argsort([-d['score'] for d in dts], kind='mergesort')
    if len(order) > max_dets:
        order = order[:max_dets]
    d = [dts[i]['bbox'] for i in order]
    g = [x['bbox'] for x in gts]
    return iou_rle(d, g, img_size=img_size), order
