"""Oracle: IoU primitives of utils/bbox_ops.py.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import ctypes
import math

import numpy as np
import torch

from . import build as _build

_LIB = None


def lib():
    """The compiled C oracle (oracle/nms.c, oracle/rotiou.c)."""
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(_build.build())
        i64, f32p, i64p, f64p = (ctypes.c_int64, ctypes.POINTER(ctypes.c_float),
                                 ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_double))
        L.oracle_nms_aabb.restype = i64
        L.oracle_nms_aabb.argtypes = [f32p, f32p, i64, ctypes.c_double, i64p]
        L.oracle_argsort_desc_stable.restype = None
        L.oracle_argsort_desc_stable.argtypes = [f32p, i64, i64p]
        L.oracle_rot_iou_pairwise.restype = None
        L.oracle_rot_iou_pairwise.argtypes = [f32p, i64, f32p, i64, f64p]
        L.oracle_raster_rle.restype = ctypes.c_int64
        L.oracle_raster_rle.argtypes = [f64p, i64, i64, i64, ctypes.POINTER(ctypes.c_uint32), i64]
        L.oracle_raster_iou_pairwise.restype = None
        L.oracle_raster_iou_pairwise.argtypes = [f64p, i64, f64p, i64, i64, i64, f64p]
        L.oracle_quad_iou_pairwise.restype = None
        L.oracle_quad_iou_pairwise.argtypes = [f64p, i64, f64p, i64, f64p]
        L.oracle_nms_rot.restype = i64
        L.oracle_nms_rot.argtypes = [f32p, f32p, i64, ctypes.c_double, i64, i64p]
        _LIB = L
    return _LIB


def _f32(t):
    a = np.ascontiguousarray(t.detach().cpu().numpy() if torch.is_tensor(t) else t, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def bboxes_iou(a, b, xyxy=False):
    """Pairwise axis-aligned IoU (N,4)x(K,4)->(N,K) float32 -- utils/bbox_ops.py:6-49.

    Same torch CPU operators in the same order, so results are bit-identical to
    the reference: corners c -/+ wh/2, tl=max, br=min, en=prod(tl<br),
    area_i=prod(br-tl)*en, iou=area_i/(area_a+area_b-area_i).
    """
    if a.dim() == 1:
        a = a.unsqueeze(0)                                             # :25-26
    if a.shape[1] != 4 or b.shape[1] != 4:
        raise IndexError()                                             # :28-29
    if xyxy:
        lo_a, hi_a, lo_b, hi_b = a[:, :2], a[:, 2:], b[:, :2], b[:, 2:]
        area_a = torch.prod(hi_a - lo_a, 1)                            # :36
        area_b = torch.prod(hi_b - lo_b, 1)
    else:
        lo_a, hi_a = a[:, :2] - a[:, 2:] / 2, a[:, :2] + a[:, 2:] / 2  # :39-43
        lo_b, hi_b = b[:, :2] - b[:, 2:] / 2, b[:, :2] + b[:, 2:] / 2
        area_a = torch.prod(a[:, 2:], 1)                               # :45
        area_b = torch.prod(b[:, 2:], 1)
    tl = torch.max(lo_a[:, None, :], lo_b)
    br = torch.min(hi_a[:, None, :], hi_b)
    en = (tl < br).to(tl.dtype).prod(dim=2)                            # :47
    inter = torch.prod(br - tl, 2) * en                                # :48
    return inter / (area_a[:, None] + area_b - inter)                  # :49


def cxcywh_to_x1y1x2y2(t):
    """utils/bbox_ops.py:309-316."""
    out = t.clone()
    out[..., 0] = t[..., 0] - t[..., 2] / 2
    out[..., 1] = t[..., 1] - t[..., 3] / 2
    out[..., 2] = t[..., 0] + t[..., 2] / 2
    out[..., 3] = t[..., 1] + t[..., 3] / 2
    return out


def xywha2vertex(box_rad):
    """Corners (N,4,2) tl,tr,br,bl from (cx,cy,w,h,radians) -- utils/bbox_ops.py:137-172."""
    c, w, h, rad = box_rad[:, 0:2], box_rad[:, 2], box_rad[:, 3], box_rad[:, 4]
    verti = torch.stack([(h / 2) * torch.sin(rad), -(h / 2) * torch.cos(rad)], dim=1)
    hori = torch.stack([(w / 2) * torch.cos(rad), (w / 2) * torch.sin(rad)], dim=1)
    return torch.stack([c + verti - hori, c + verti + hori, c - verti + hori, c - verti - hori], dim=1)


def iou_rot(b1, b2):
    """Exact rotated IoU matrix (N,M) float64 for (cx,cy,w,h,degrees) boxes.

    Stands in for iou_rle (utils/bbox_ops.py:52-100); PARITY UNPINNED, see
    oracle/rotiou.c for why and for the algorithm.
    """
    if b1.dim() == 1:
        b1 = b1.unsqueeze(0)
    if b2.dim() == 1:
        b2 = b2.unsqueeze(0)
    a1, p1 = _f32(b1)
    a2, p2 = _f32(b2)
    out = np.empty((a1.shape[0], a2.shape[0]), dtype=np.float64)
    lib().oracle_rot_iou_pairwise(p1, a1.shape[0], p2, a2.shape[0],
                                  out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return torch.from_numpy(out)


def iou_rle_raster(b1, b2, canvas=2048):
    """Rasterised rotated IoU (N,M) float64 in the manner of the reference's iou_rle (utils/bbox_ops.py:84-96):
    corners from xywha2vertex in float32, then pycocotools-style polygon rasterisation on a canvas x canvas grid
    (the reference always uses 2048: its callers pass `img_size=` but iou_rle reads kwargs['img_hw']).
    PARITY UNPINNED: oracle/raster.c restates the published maskApi algorithm and could not be checked against the
    real library.  Only used to REPORT the raster-vs-exact gap."""
    def corners(b):
        if b.dim() == 1:
            b = b.unsqueeze(0)
        rad = b.clone().float()
        rad[:, 4] = deg2rad_f32(rad[:, 4])                                # :88-89
        v = xywha2vertex(rad)                                             # (N,4,2) tl,tr,br,bl
        return np.ascontiguousarray(v.reshape(-1, 8).double().numpy())
    c1, c2 = corners(b1), corners(b2)
    out = np.empty((c1.shape[0], c2.shape[0]), dtype=np.float64)
    f64p = ctypes.POINTER(ctypes.c_double)
    lib().oracle_raster_iou_pairwise(c1.ctypes.data_as(f64p), c1.shape[0], c2.ctypes.data_as(f64p), c2.shape[0],
                                     int(canvas), int(canvas), out.ctypes.data_as(f64p))
    return torch.from_numpy(out)


def nms_rot(boxes, scores, nms_thres=0.45, majority=None):
    """Rotated greedy NMS with nms_rotbb's control flow -- utils/bbox_ops.py:250-306.

    Returns kept indices (int64) into the input, in descending score order.
    """
    if boxes.shape[0] == 0:
        return torch.zeros(0, dtype=torch.int64)
    ab, pb = _f32(boxes)
    asc, ps = _f32(scores)
    keep = np.empty(ab.shape[0], dtype=np.int64)
    k = lib().oracle_nms_rot(pb, ps, ab.shape[0], float(nms_thres),
                             int(majority) if majority else 0,
                             keep.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return torch.from_numpy(keep[:k].copy())


def deg2rad_f32(deg):
    """degrees -> radians the way iou_rle does it (bbox_ops.py:88-89), float32."""
    return deg * math.pi / 180


def iou_rot_f64(boxes1, boxes2):
    """The EVALUATOR's rotated IoU (utils/evaluation/cepdof.py:210-243): boxes are float64 (lists of Python floats), the
    corners are built with numpy in float64 (:181-207), and the two polygons go to pycocotools -- here, as everywhere,
    to the exact clipping of oracle/rotiou.c.  Returns np.array[M, N] float64."""
    import ctypes
    import numpy as np
    b1 = np.array(boxes1, dtype=np.float64).reshape(-1, 5)
    b2 = np.array(boxes2, dtype=np.float64).reshape(-1, 5)
    if b1.shape[0] == 0 or b2.shape[0] == 0:
        return np.zeros((b1.shape[0], b2.shape[0]))

    def vertices(box):
        rad = box[:, 4] * np.pi / 180
        center, w, h = box[:, 0:2], box[:, 2], box[:, 3]
        verti = np.stack([(h / 2) * np.sin(rad), -(h / 2) * np.cos(rad)], axis=1)
        hori = np.stack([(w / 2) * np.cos(rad), (w / 2) * np.sin(rad)], axis=1)
        return np.ascontiguousarray(np.concatenate([center + verti - hori, center + verti + hori, center - verti + hori,
                                                    center - verti - hori], axis=1))
    a, b = vertices(b1), vertices(b2)
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.float64)
    f64p = ctypes.POINTER(ctypes.c_double)
    lib().oracle_quad_iou_pairwise(a.ctypes.data_as(f64p), a.shape[0], b.ctypes.data_as(f64p), b.shape[0], out.ctypes.data_as(f64p))
    return out


def raster_rle(xy, h, w):
    """Run lengths (alternating 0-runs / 1-runs, column major) of the polygon xy = [x0, y0, x1, y1, ...] on an h x w
    canvas, by oracle/raster.c (the restated rleFrPoly)."""
    import ctypes
    import numpy as np
    pts = np.ascontiguousarray(np.asarray(xy, dtype=np.float64))
    out = np.zeros(4096, dtype=np.uint32)
    m = lib().oracle_raster_rle(pts.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), pts.size // 2, int(h), int(w),
                                out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), out.size)
    return out[:m].tolist()
