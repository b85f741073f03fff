"""Host-side mirror of the reference's utils/bbox_ops.py for the hot path: same function names,
argument meaning and error behaviour; the arithmetic runs in libmydet's CUDA kernels.

Inputs may live on the CPU (the reference's callers mostly pass CPU tensors): they are staged to
the current CUDA device, and the result is returned on the input's device like the reference does.
Without a CUDA device these functions raise -- there is no CPU implementation here.
"""
from math import pi

import os

import torch

from . import _lib, ops


def _cuda_device():
    if not torch.cuda.is_available():
        raise _lib.MydetError('mydetection_b200 needs a CUDA device (B200); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def _stage(t):
    """float32 copy of t on the GPU (no copy if it is already there)."""
    return t.detach().to(device=t.device if t.is_cuda else _cuda_device(), dtype=torch.float32)


def bboxes_iou(bboxes_a, bboxes_b, xyxy=False):
    """Pairwise IoU (N,4) x (K,4) -> (N,K) float32; utils/bbox_ops.py:6-49 (bit-exact)."""
    if bboxes_a.dim() == 1:
        bboxes_a = bboxes_a.unsqueeze(0)
    assert bboxes_a.dim() == bboxes_b.dim() == 2
    if bboxes_a.shape[1] != 4 or bboxes_b.shape[1] != 4:
        raise IndexError()
    out = ops.iou_aabb(_stage(bboxes_a), _stage(bboxes_b), xyxy=bool(xyxy))
    return out.to(bboxes_a.device)


def iou_rle(boxes1, boxes2, bb_format='cxcywhd', **kwargs):
    """IoU between rotated boxes (N,5) x (M,5) -> (N,M) float64; utils/bbox_ops.py:52-100.

    The reference rasterises the polygons with pycocotools on an `img_hw` canvas (default 2048 x 2048, :84-85).
    Default here: the EXACT intersection area of the same polygons (`img_hw` accepted and unused).  `raster=True`
    (or MYDET_IOU_RASTER=1 in the environment) selects the raster route instead -- pycocotools' polygon rule restated
    (oracle/raster.c, pinned to hand-derived run-length encodings; the library itself is not available to check
    against) and evaluated on the device, on the `img_hw` canvas exactly as the reference reads that keyword.
    """
    assert type(boxes1) == type(boxes2)
    assert bb_format == 'cxcywhd'
    if not (torch.is_tensor(boxes1) and torch.is_tensor(boxes2)):
        boxes1 = torch.from_numpy(boxes1).float()
        boxes2 = torch.from_numpy(boxes2).float()
    assert boxes1.device == boxes2.device
    device = boxes1.device
    if boxes1.dim() == 1:
        boxes1 = boxes1.unsqueeze(0)
    if boxes2.dim() == 1:
        boxes2 = boxes2.unsqueeze(0)
    assert boxes1.shape[1] == boxes2.shape[1] == 5
    if kwargs.get('raster', os.environ.get('MYDET_IOU_RASTER') == '1'):
        ious = ops.iou_raster(_stage(boxes1), _stage(boxes2), kwargs.get('img_hw', 2048))     # :84-85
    else:
        ious = ops.iou_rot(_stage(boxes1), _stage(boxes2))
    if kwargs.get('return_numpy', False):
        return ious.cpu().numpy()
    return ious.to(device=device)


def xywha2vertex(box, is_degree, stack=True):
    """(batch,5) (x,y,w,h,radians) -> corners tl,tr,br,bl: (batch,4,2), or (batch,8) if not stack;
    utils/bbox_ops.py:137-172."""
    assert is_degree == False and box.dim() == 2 and box.shape[1] >= 5  # noqa: E712
    src = _stage(box).contiguous()
    out = torch.empty(src.shape[0], 4, 2, dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        rc = _lib.lib().mydet_xywha2vertex(ops._ptr(src), src.shape[0], src.shape[1], ops._ptr(out), ops._stream())
    _lib.check(rc, 'mydet_xywha2vertex')
    out = out.to(box.device)
    return out if stack else out.reshape(-1, 8)


def nms_rotbb(boxes, scores, nms_thres=0.45, bb_format='cxcywhd', img_size=2048, majority=None):
    """Single-class NMS for rotated boxes, rows (x,y,w,h,degrees); utils/bbox_ops.py:250-306.

    Returns kept indices (int64) in descending score order.  A box is dropped iff its IoU with an
    already kept box is >= nms_thres; with `majority`, kept boxes with fewer votes are dropped too."""
    if bb_format != 'cxcywhd':
        raise NotImplementedError()
    assert (boxes.dim() == 2) and (boxes.shape[1] == 5)
    device = boxes.device
    if boxes.shape[0] == 0:
        return torch.zeros(0, dtype=torch.int64, device=device)
    b, s = _stage(boxes)[None], _stage(scores)[None]
    res = ops.nms_rot(b, s, nms_thres, ge=True, want_votes=majority is not None)
    n = int(res[1][0])
    keep = res[0][0, :n]
    if majority is not None:
        keep = keep[res[2][0, :n] >= majority]          # votes_valid filter, :304-306
    return keep.to(device)


def cxcywh_to_x1y1x2y2(cxcywh):
    """utils/bbox_ops.py:309-316; columns beyond the fourth are copied through."""
    assert cxcywh.shape[-1] >= 4
    src = _stage(cxcywh).contiguous()
    out = torch.empty_like(src)
    rows = src.numel() // src.shape[-1]
    with torch.cuda.device(src.device):
        rc = _lib.lib().mydet_cxcywh_to_x1y1x2y2(ops._ptr(src), rows, src.shape[-1], ops._ptr(out), ops._stream())
    _lib.check(rc, 'mydet_cxcywh_to_x1y1x2y2')
    return out.to(cxcywh.device)
