"""The evaluator side of the path.  `install_cepdof()` (called by dropin.install() when the reference's evaluator is
importable) replaces `utils.evaluation.cepdof.iou_rle` and `CEPDOFeval.computeIoU`: the first computeIoU call of an
evaluation computes the matching IoU of EVERY (image, category) pair in one launch (mydet_iou_rot_segments) and the
other calls -- COCOeval.evaluate makes one per pair -- read their matrix from that result.

First step beyond the hot path: the detection-to-GT matching IoU of the
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
    dev = _cuda_device()                      # float64 boxes and corners, as the evaluator's numpy code (cepdof.py:232-236)
    flat, _ = ops.iou_rot_segments(torch.from_numpy(b1).to(dev), torch.from_numpy(b2).to(dev), [(0, b1.shape[0], 0, b2.shape[0])])
    return flat.cpu().numpy().reshape(b1.shape[0], b2.shape[0])


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


def compute_iou_all(dts_by_key, gts_by_key, keys, max_dets=100):
    """CEPDOFeval.computeIoU for many (image, category) keys in ONE launch.  dts_by_key / gts_by_key map a key to the
    list of dt / gt dicts; returns {key: ious} with exactly what computeIoU returns per key: [] when both lists are
    empty, else np.array[D, G] float64 (D capped at max_dets after the stable score sort)."""
    res, segs, a_rows, b_rows, live = {}, [], [], [], []
    for key in keys:
        dts, gts = dts_by_key.get(key, []), gts_by_key.get(key, [])
        if len(dts) == 0 and len(gts) == 0:
            res[key] = []
            continue
        order = np.argsort([-d['score'] for d in dts], kind='mergesort')[:max_dets]
        segs.append((len(a_rows), len(order), len(b_rows), len(gts)))
        a_rows.extend(dts[i]['bbox'] for i in order)
        b_rows.extend(g['bbox'] for g in gts)
        live.append(key)
    if not live:
        return res
    dev = _cuda_device()
    a = torch.tensor(np.array(a_rows, dtype=np.float64).reshape(-1, 5), dtype=torch.float64)     # float64 end to end,
    b = torch.tensor(np.array(b_rows, dtype=np.float64).reshape(-1, 5), dtype=torch.float64)     # like the evaluator's numpy
    if a.shape[0] == 0 or b.shape[0] == 0:
        flat, out0 = np.zeros(0), [0] * len(segs)
    else:
        flat_t, out0 = ops.iou_rot_segments(a.to(dev), b.to(dev), segs)
        flat = flat_t.cpu().numpy()
    for key, (_, na, _, nb), o in zip(live, segs, out0):
        res[key] = flat[o:o + na * nb].reshape(na, nb).copy() if na * nb else np.zeros((na, nb))
    return res


def _compute_iou_method(self, imgId, catId):
    """Replacement of CEPDOFeval.computeIoU (utils/evaluation/cepdof.py:67-99): same arguments, same return value; the
    first call of an evaluation serves all (imgId, catId) pairs of self.params from one launch."""
    p = self.params
    cache = self.__dict__.get('_mydet_ious')
    stamp = (id(self._gts), id(self._dts), bool(p.useCats), p.maxDets[-1], len(p.imgIds), len(p.catIds))   # _prepare() makes new dicts per evaluate()
    if cache is None or cache[0] != stamp:
        if p.useCats:
            keys = [(i, c) for i in p.imgIds for c in p.catIds]
            dts, gts = self._dts, self._gts
        else:
            keys = [(i, -1) for i in p.imgIds]
            dts = {(i, -1): [d for c in p.catIds for d in self._dts[i, c]] for i in p.imgIds}
            gts = {(i, -1): [g for c in p.catIds for g in self._gts[i, c]] for i in p.imgIds}
        cache = (stamp, compute_iou_all(dts, gts, keys, max_dets=p.maxDets[-1]))
        self._mydet_ious = cache
    key = (imgId, catId if p.useCats else -1)
    if key not in cache[1]:                       # a pair outside params: the single-pair path
        gt = self._gts[imgId, catId] if p.useCats else [g for c in p.catIds for g in self._gts[imgId, c]]
        dt = self._dts[imgId, catId] if p.useCats else [d for c in p.catIds for d in self._dts[imgId, c]]
        return compute_iou(dt, gt, max_dets=p.maxDets[-1])[0]
    return cache[1][key]


def install_cepdof(module=None):
    """Patch the reference's evaluator module (utils.evaluation.cepdof) in place: its numpy `iou_rle` (:210-243) and
    `CEPDOFeval.computeIoU` (:67-99) run on the rotated-IoU kernels.  Returns the module, or None when it cannot be
    imported (it needs pycocotools' COCOeval base class)."""
    if module is None:
        import importlib
        try:
            module = importlib.import_module('utils.evaluation.cepdof')
        except Exception:
            return None
    module.iou_rle = iou_rle
    module.CEPDOFeval.computeIoU = _compute_iou_method
    return module
