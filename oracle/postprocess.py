"""Oracle: threshold -> top-k -> per-class NMS.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

All functions return INDICES into the arrays they were given, in the order the
reference emits the surviving boxes (class ascending, then score descending,
utils/structures.py:158-171), so that callers can compare kept indices exactly.

Tie policy (the reference leaves it unspecified because
torch.topk / argsort are unstable): equal scores are ordered by ascending input
index everywhere.  For inputs without exact score ties inside the top-(K+1) this
coincides with whatever the reference does.
"""
import ctypes

import numpy as np
import torch

from .iou import lib, _f32

TOPK_CAP = 512  # utils/structures.py:99-101


def argsort_desc_stable(scores):
    a, p = _f32(scores)
    order = np.empty(a.shape[0], dtype=np.int64)
    lib().oracle_argsort_desc_stable(p, a.shape[0], order.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return torch.from_numpy(order)


def nms_aabb(xyxy, scores, thr):
    """Single-class NMS == torchvision.ops.nms on CPU (restated in oracle/nms.c)."""
    n = xyxy.shape[0]
    if n == 0:
        return torch.zeros(0, dtype=torch.int64)
    ab, pb = _f32(xyxy)
    asc, ps = _f32(scores)
    keep = np.empty(n, dtype=np.int64)
    k = lib().oracle_nms_aabb(pb, ps, n, float(thr), keep.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return torch.from_numpy(keep[:k].copy())


def to_corners(boxes, bb_format):
    """Box -> x1y1x2y2 exactly as ImageObjects.non_max_suppression does
    (utils/structures.py:124-143).  For 'cxcywhd' the ANGLE IS DROPPED."""
    if bb_format == 'x1y1x2y2':
        return boxes.clone()
    if bb_format not in ('cxcywh', 'cxcywhd'):
        raise NotImplementedError()
    out = boxes[:, 0:4].clone()
    out[:, 0] = boxes[:, 0] - boxes[:, 2] / 2
    out[:, 1] = boxes[:, 1] - boxes[:, 3] / 2
    out[:, 2] = boxes[:, 0] + boxes[:, 2] / 2
    out[:, 3] = boxes[:, 1] + boxes[:, 3] / 2
    return out


def class_nms(boxes, scores, cats, nms_thres, bb_format='cxcywh'):
    """ImageObjects.non_max_suppression -- utils/structures.py:111-173."""
    if boxes.shape[0] == 0:
        return torch.zeros(0, dtype=torch.int64)
    corners = to_corners(boxes, bb_format)
    picked = []
    for c in cats.unique():                                            # ascending, :157
        members = torch.nonzero(cats == c).flatten()                   # :159
        kept = nms_aabb(corners[members], scores[members], nms_thres)  # :161-162
        picked.append(members[kept])
    return torch.cat(picked)


def post_process(boxes, cats, scores, conf_thres, nms_thres, bb_format='cxcywh', topk=TOPK_CAP):
    """ImageObjects.post_process -- utils/structures.py:92-106.

    topk=None disables the cap (the dense-scene stress configuration).
    """
    alive = torch.nonzero(scores >= conf_thres).flatten()              # :98 (float32 compare)
    if topk is not None and alive.numel() > topk:                      # :99
        order = argsort_desc_stable(scores[alive])[:topk]              # :100 torch.topk, sorted desc
        alive = alive[order]
    sub = class_nms(boxes[alive], scores[alive], cats[alive], nms_thres, bb_format)   # :105
    return alive[sub]


def top_boundary_is_tie_free(scores, conf_thres, topk=TOPK_CAP):
    """Pre-condition of SURVEY.md section 8d: no exact tie between the K-th and (K+1)-th
    surviving score, so that the reference's unstable torch.topk is unambiguous."""
    s = scores[scores >= conf_thres]
    if topk is None or s.numel() <= topk:
        return True
    top = torch.sort(s, descending=True).values[:topk + 1]
    return bool(top[topk - 1] != top[topk])
