"""Tensor-level wrappers over the C ABI (include/mydet.h).

PyTorch is used here only for device memory, the current stream and dtype/shape checks; every
result is produced by the hand-written CUDA kernels of libmydet.so.  CPU tensors are rejected --
there is no fallback path.
"""
import collections
import ctypes

import torch

from . import _lib
from ._lib import (KIND_YOLO, KIND_FCOS, KIND_RAPID, KIND_RETINA, KIND_UV5,  # noqa: F401
                   BOX_CXCYWH, BOX_X1Y1X2Y2, SMALL_K)

BOX_FORMATS = {'cxcywh': BOX_CXCYWH, 'cxcywhd': BOX_CXCYWH, 'x1y1x2y2': BOX_X1Y1X2Y2}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(t, dtype, what):
    if not torch.is_tensor(t):
        raise TypeError(f'{what}: expected a tensor')
    if not t.is_cuda:
        raise _lib.MydetError(f'{what}: expected a CUDA tensor (libmydet has no CPU path)')
    if t.dtype != dtype:
        raise TypeError(f'{what}: expected dtype {dtype}, got {t.dtype}')
    return t


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# Persistent workspaces of the large-N NMS (include/mydet.h, "persistent workspaces"): zeroed once, handed to the library
# with the clean flag, returned clean -- the n x ceil(n/64) suppression matrix is then never cleared wholesale again.
# Keyed by everything that fixes the buffer's layout AND by the stream (two streams must not share a buffer); a few
# entries, least recently used out first.
_PERSISTENT = collections.OrderedDict()
_PERSISTENT_MAX = 6


def _persistent_workspace(key, nbytes, device):
    ws = _PERSISTENT.pop(key, None)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)
    _PERSISTENT[key] = ws
    while len(_PERSISTENT) > _PERSISTENT_MAX:
        _PERSISTENT.popitem(last=False)
    return ws


def release_workspaces():
    """Drop the cached persistent workspaces (they hold device memory between calls)."""
    _PERSISTENT.clear()


# --------------------------------------------------------------------------------------- levels
def make_level(raw, stride, anchors_wh=None, conf_key='conf'):
    """Describe one raw dict of head views (models/rpns.py:29-41, :175-189) as a mydet_level_t.

    raw['bbox'] is (B,nA,nH,nW,P) or (B,nH,nW,P); raw[conf_key] and raw['class'] follow the same
    layout with a last dim of 1 / C.  No copy is made: pointers and element strides are passed.
    Returns (Level, keepalive tensors).
    """
    bbox = _dev(raw['bbox'], torch.float32, "raw['bbox']")
    multi = bbox.dim() == 5
    if bbox.dim() not in (4, 5):
        raise ValueError("raw['bbox'] must be (B,nA,nH,nW,P) or (B,nH,nW,P)")
    lv = _lib.Level()
    sb = bbox.stride()
    if multi:
        n_a, n_h, n_w = bbox.shape[1:4]
        lv.bbox_stride[:] = [sb[0], sb[1], sb[2], sb[3], sb[4]]
    else:
        n_a, (n_h, n_w) = 1, bbox.shape[1:3]
        lv.bbox_stride[:] = [sb[0], 0, sb[1], sb[2], sb[3]]
    lv.bbox = bbox.data_ptr()
    keep = [bbox]
    conf = raw.get(conf_key)
    if conf is not None:
        conf = _dev(conf, torch.float32, f"raw['{conf_key}']")
        sc = conf.stride()
        lv.conf_stride[:] = [sc[0], sc[1], sc[2], sc[3]] if multi else [sc[0], 0, sc[1], sc[2]]
        lv.conf = conf.data_ptr()
        keep.append(conf)
    cls = raw.get('class')
    if cls is not None and cls.shape[-1] > 0:
        cls = _dev(cls, torch.float32, "raw['class']")
        sk = cls.stride()
        lv.cls_stride[:] = [sk[0], sk[1], sk[2], sk[3], sk[4]] if multi else [sk[0], 0, sk[1], sk[2], sk[3]]
        lv.cls = cls.data_ptr()
        keep.append(cls)
    lv.n_anchor, lv.n_h, lv.n_w = int(n_a), int(n_h), int(n_w)
    lv.stride = float(stride)
    if anchors_wh is not None:
        aw = [float(a[0]) for a in anchors_wh]
        ah = [float(a[1]) for a in anchors_wh]
        if len(aw) != n_a or n_a > _lib.MAX_ANCHORS:
            raise ValueError(f'{len(aw)} anchors given for {n_a} anchor planes (max {_lib.MAX_ANCHORS})')
        lv.anchor_w[:n_a] = aw
        lv.anchor_h[:n_a] = ah
    return lv, keep


class LevelSet:
    """A host array of mydet_level_t for a fixed set of head tensors (reusable across calls)."""

    def __init__(self, raws, strides, anchors=None, conf_key='conf'):
        if not 1 <= len(raws) <= _lib.MAX_LEVELS:
            raise ValueError(f'1..{_lib.MAX_LEVELS} levels supported')
        self.array = (_lib.Level * len(raws))()
        self.keep = []
        self.n_total = 0
        for i, raw in enumerate(raws):
            lv, keep = make_level(raw, strides[i], None if anchors is None else anchors[i], conf_key)
            self.array[i] = lv
            self.keep += keep
            self.n_total += lv.n_anchor * lv.n_h * lv.n_w
        self.n_levels = len(raws)
        first = raws[0]['bbox']
        self.batch = int(first.shape[0])
        self.n_param = int(first.shape[-1])
        self.device = first.device
        cls = raws[0].get('class')
        self.n_cls = 0 if cls is None else int(cls.shape[-1])


# --------------------------------------------------------------------------------------- decode
def decode_dense(kind, levels: LevelSet, img_hw):
    """All levels -> level-concatenated (bbox (B,N,P) f32, class_idx (B,N) i64, score (B,N) f32)."""
    B, N, P = levels.batch, levels.n_total, levels.n_param
    dev = levels.device
    box = torch.empty(B, N, P, dtype=torch.float32, device=dev)
    cls = torch.empty(B, N, dtype=torch.int64, device=dev)
    score = torch.empty(B, N, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().mydet_decode_dense(kind, levels.array, levels.n_levels, B, levels.n_cls, P,
                                           float(img_hw[0]), float(img_hw[1]), _ptr(box), _ptr(cls), _ptr(score),
                                           N, _stream())
    _lib.check(rc, 'mydet_decode_dense')
    return box, cls, score


def decode_compact(kind, levels: LevelSet, img_hw, conf_thres, capacity=None):
    """Decode fused with `score >= conf_thres` and per-image compaction.

    Returns dict(box (B,cap,P), score (B,cap), cls (B,cap) i32, idx (B,cap) i32, count (B) i32)."""
    B, N, P = levels.batch, levels.n_total, levels.n_param
    cap = int(capacity or N)
    dev = levels.device
    out = {'box': torch.empty(B, cap, P, dtype=torch.float32, device=dev),
           'score': torch.empty(B, cap, dtype=torch.float32, device=dev),
           'cls': torch.empty(B, cap, dtype=torch.int32, device=dev),
           'idx': torch.empty(B, cap, dtype=torch.int32, device=dev),
           'count': torch.empty(B, dtype=torch.int32, device=dev)}
    with torch.cuda.device(dev):
        rc = _lib.lib().mydet_decode_compact(kind, levels.array, levels.n_levels, B, levels.n_cls, P,
                                             float(img_hw[0]), float(img_hw[1]), float(conf_thres),
                                             _ptr(out['box']), _ptr(out['score']), _ptr(out['cls']), _ptr(out['idx']),
                                             _ptr(out['count']), cap, 0, _stream())
    _lib.check(rc, 'mydet_decode_compact')
    return out


# --------------------------------------------------------------------------------------- post-process
def _alloc_dets(B, cap, P, dev):
    return {'box': torch.empty(B, cap, P, dtype=torch.float32, device=dev),
            'score': torch.empty(B, cap, dtype=torch.float32, device=dev),
            'cls': torch.empty(B, cap, dtype=torch.int64, device=dev),
            'idx': torch.empty(B, cap, dtype=torch.int32, device=dev),
            'count': torch.empty(B, dtype=torch.int32, device=dev),
            'status': torch.zeros(B, dtype=torch.int32, device=dev)}


def postprocess(boxes, scores, cls, conf_thres, nms_thres, topk=512, box_format='cxcywh', counts=None,
                src_idx=None, out_cap=None, out=None, consume=False, force_scan=False):
    """Batched threshold -> top-k -> class-aware NMS.  boxes (B,n,P), scores (B,n), cls (B,n) i32|i64.

    topk=None removes the cap.  Returns dict(box, score, cls (i64), idx (i32), count, status);
    rows [0, count[b]) of image b are valid and ordered class asc / score desc.
    consume=True zeroes `counts` once read (include/mydet.h: the state a decode_compact with state_clean needs);
    force_scan=True skips the sampled front end of the select (tests: the result must not change)."""
    boxes = _dev(boxes, torch.float32, 'boxes')
    scores = _dev(scores, torch.float32, 'scores')
    if cls.dtype not in (torch.int32, torch.int64):
        raise TypeError('cls must be int32 or int64')
    if not cls.is_cuda:
        raise _lib.MydetError('cls: expected a CUDA tensor')
    if boxes.dim() != 3 or scores.dim() != 2 or cls.dim() != 2:
        raise ValueError('expected boxes (B,n,P), scores (B,n), cls (B,n)')
    boxes, scores, cls = boxes.contiguous(), scores.contiguous(), cls.contiguous()
    B, n, P = boxes.shape
    k = int(topk) if topk else 0
    eff = min(k, n) if k > 0 else n
    cap = int(out_cap or max(eff, 1))
    dev = boxes.device
    if out is None:
        out = _alloc_dets(B, cap, P, dev)
    L = _lib.lib()
    wbytes = L.mydet_postprocess_workspace_bytes(B, n, k)
    ws = _workspace(wbytes, dev)
    with torch.cuda.device(dev):
        rc = L.mydet_postprocess(_ptr(boxes), _ptr(scores), _ptr(cls), 1 if cls.dtype == torch.int64 else 0,
                                 _ptr(src_idx), _ptr(counts), B, n, n, P, BOX_FORMATS[box_format],
                                 float(conf_thres), k, float(nms_thres), _ptr(out['box']), _ptr(out['score']),
                                 _ptr(out['cls']), _ptr(out['idx']), _ptr(out['count']), _ptr(out['status']), cap,
                                 _ptr(ws), ws.numel(), (1 if consume else 0) | (2 if force_scan else 0), _stream())
    _lib.check(rc, 'mydet_postprocess')
    return out


def detect(kind, levels: LevelSet, img_hw, conf_thres, nms_thres, topk=512, out=None, workspace=None):
    """decode + threshold + top-k + NMS for the whole batch in one C call."""
    B, N, P = levels.batch, levels.n_total, levels.n_param
    k = int(topk) if topk else 0
    cap = min(k, N) if k > 0 else N
    dev = levels.device
    if out is None:
        out = _alloc_dets(B, max(cap, 1), P, dev)
    L = _lib.lib()
    if workspace is None:
        workspace = _workspace(L.mydet_detect_workspace_bytes(B, N, P, k), dev)
    with torch.cuda.device(dev):
        rc = L.mydet_detect(kind, levels.array, levels.n_levels, B, levels.n_cls, P, float(img_hw[0]),
                            float(img_hw[1]), float(conf_thres), k, float(nms_thres), _ptr(out['box']),
                            _ptr(out['score']), _ptr(out['cls']), _ptr(out['idx']), _ptr(out['count']),
                            _ptr(out['status']), out['box'].shape[1], _ptr(workspace), workspace.numel(), _stream())
    _lib.check(rc, 'mydet_detect')
    return out


def detect_workspace(levels: LevelSet, topk=512, zeroed=False):
    """zeroed=True: a persistent workspace in its clean state, for mydet_detect_ws(..., workspace_clean=1)."""
    k = int(topk) if topk else 0
    nbytes = _lib.lib().mydet_detect_workspace_bytes(levels.batch, levels.n_total, levels.n_param, k)
    if zeroed:
        return torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=levels.device)
    return _workspace(nbytes, levels.device)


# --------------------------------------------------------------------------------------- rotated NMS / IoU
_SIDE_STREAMS = {}


def _side_streams(dev, k):
    """k cached side streams of a device (fork / join helpers of the chunked large-NMS calls)."""
    pool = _SIDE_STREAMS.setdefault(dev.index if dev.index is not None else torch.cuda.current_device(), [])
    while len(pool) < k:
        pool.append(torch.cuda.Stream(dev))
    return pool[:k]


def nms_rot(boxes, scores, thr, ge=True, counts=None, want_votes=False, chunks=None):
    """Batched single-class rotated NMS.  boxes (B,n,5) degrees, scores (B,n).
    Returns (keep (B,n) i64, keep_count (B) i32[, votes (B,n) i32]): kept indices per image in
    descending score; votes[b, p] = 1 + number of dropped boxes whose best overlap was keep[b, p].

    Large batches are issued as `chunks` (default 2) independent C calls on forked streams that join the
    current stream again: the per-image stages of the pipeline (shared-memory sort, greedy sweep: one CTA
    per image) of one chunk then run beside the grid-filling mask kernel of the other."""
    boxes = _dev(boxes, torch.float32, 'boxes').contiguous()
    scores = _dev(scores, torch.float32, 'scores').contiguous()
    if boxes.dim() != 3 or boxes.shape[-1] != 5:
        raise ValueError('boxes must be (B,n,5)')
    B, n, _ = boxes.shape
    dev = boxes.device
    keep = torch.empty(B, max(n, 1), dtype=torch.int64, device=dev)
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    votes = torch.empty(B, max(n, 1), dtype=torch.int32, device=dev) if want_votes else None
    L = _lib.lib()
    if B == 0:
        return (keep, cnt, votes) if want_votes else (keep, cnt)
    if chunks is None:
        chunks = 2 if (B >= 4 and n >= 2048) else 1
    chunks = max(1, min(int(chunks), B))
    bounds = [(B * c // chunks, B * (c + 1) // chunks) for c in range(chunks)]
    def ws_key(lo, hi, ci, stream):
        return ('nms_rot', dev.index, hi - lo, n, ci, stream.cuda_stream)

    def call(lo, hi, ci, stream):
        # one persistent workspace per (geometry, chunk, stream): clean on entry, left clean by the library
        key = ws_key(lo, hi, ci, stream)
        ws = _persistent_workspace(key, L.mydet_nms_rot_workspace_bytes(hi - lo, n), dev)
        sub = lambda t: _ptr(t[lo:hi]) if t is not None else _ptr(None)
        rc = L.mydet_nms_rot_ws(sub(boxes), sub(scores), sub(counts), hi - lo, n, n, float(thr), 1 if ge else 0,
                                sub(keep), sub(cnt), sub(votes), _ptr(ws), ws.numel(), 1, ctypes.c_void_p(stream.cuda_stream))
        if rc:
            _PERSISTENT.pop(key, None)             # its state is unknown now
        _lib.check(rc, 'mydet_nms_rot_ws')

    with torch.cuda.device(dev):
        main = torch.cuda.current_stream()
        if chunks == 1:
            call(0, B, 0, main)
        else:
            sides = _side_streams(dev, chunks)
            for ci, ((lo, hi), side) in enumerate(zip(bounds, sides)):       # a new buffer is zeroed on `main`, BEFORE the fork
                _persistent_workspace(ws_key(lo, hi, ci, side), L.mydet_nms_rot_workspace_bytes(hi - lo, n), dev)
            fork = torch.cuda.Event()
            fork.record(main)
            for ci, ((lo, hi), side) in enumerate(zip(bounds, sides)):
                side.wait_event(fork)
                call(lo, hi, ci, side)
                main.wait_stream(side)
    return (keep, cnt, votes) if want_votes else (keep, cnt)


def iou_aabb(a, b, xyxy=False):
    a = _dev(a, torch.float32, 'bboxes_a').contiguous()
    b = _dev(b, torch.float32, 'bboxes_b').contiguous()
    out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.lib().mydet_iou_aabb_pairwise(_ptr(a), a.shape[0], _ptr(b), b.shape[0], 1 if xyxy else 0,
                                                _ptr(out), _stream())
    _lib.check(rc, 'mydet_iou_aabb_pairwise')
    return out


def iou_rowmax(a, gt, gt_count=None, xyxy=False, want_arg=True):
    """`bboxes_iou(a[b], gt[b]).max(dim=1)` for every image, without the matrix (mydet_iou_aabb_rowmax).
    a: (B,n,P>=4) or (n,P) shared by all images; gt: (B,G,4); gt_count: (B,) i32 or None.
    Returns (max (B,n) f32, arg (B,n) i64 | None); images without GT give -1 / -1."""
    a = _dev(a, torch.float32, 'a')
    gt = _dev(gt, torch.float32, 'gt').contiguous()
    if gt.dim() != 3 or gt.shape[-1] != 4:
        raise ValueError('gt must be (B,G,4)')
    B, G, _ = gt.shape
    shared = a.dim() == 2
    if a.shape[-1] < 4 or a.dim() not in (2, 3) or (not shared and a.shape[0] != B):
        raise IndexError('a must be (B,n,P>=4) or (n,P>=4)')
    a = a.contiguous()
    n, pitch = a.shape[-2], a.shape[-1]
    dev = a.device
    out_max = torch.empty(B, n, dtype=torch.float32, device=dev)
    out_arg = torch.empty(B, n, dtype=torch.int64, device=dev) if want_arg else None
    if gt_count is not None:
        gt_count = _dev(gt_count, torch.int32, 'gt_count').contiguous()
    if G == 0:                                  # keep the 16-byte alignment check away from an empty tensor
        out_max.fill_(-1.0)
        if want_arg:
            out_arg.fill_(-1)
        return out_max, out_arg
    with torch.cuda.device(dev):
        rc = _lib.lib().mydet_iou_aabb_rowmax(_ptr(a), 0 if shared else n * pitch, pitch, n, _ptr(gt), _ptr(gt_count), G, B,
                                              1 if xyxy else 0, _ptr(out_max), _ptr(out_arg), _stream())
    _lib.check(rc, 'mydet_iou_aabb_rowmax')
    return out_max, out_arg


def iou_rot(a, b):
    a = _dev(a, torch.float32, 'boxes1').contiguous()
    b = _dev(b, torch.float32, 'boxes2').contiguous()
    out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.lib().mydet_iou_rot_pairwise(_ptr(a), a.shape[0], _ptr(b), b.shape[0], _ptr(out), _stream())
    _lib.check(rc, 'mydet_iou_rot_pairwise')
    return out


def iou_raster(a, b, canvas_hw=(2048, 2048)):
    """Rasterised rotated IoU (mydet_iou_raster_pairwise): the reference's pycocotools route, restated.  a (N,5), b (K,5)
    degrees -> (N,K) float64 on a canvas_hw canvas."""
    a = _dev(a, torch.float32, 'boxes1').contiguous()
    b = _dev(b, torch.float32, 'boxes2').contiguous()
    h, w = (int(canvas_hw), int(canvas_hw)) if isinstance(canvas_hw, int) else (int(canvas_hw[0]), int(canvas_hw[1]))
    out = torch.zeros(a.shape[0], b.shape[0], dtype=torch.float64, device=a.device)
    L = _lib.lib()
    ws = _workspace(L.mydet_iou_raster_workspace_bytes(a.shape[0], b.shape[0], w), a.device)
    with torch.cuda.device(a.device):
        rc = L.mydet_iou_raster_pairwise(_ptr(a), a.shape[0], _ptr(b), b.shape[0], h, w, _ptr(out), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, 'mydet_iou_raster_pairwise')
    return out


def iou_rot_segments(a, b, segments):
    """Many small rotated-IoU matrices in one launch (mydet_iou_rot_segments).  a (Na,5), b (Nb,5) degrees;
    segments: int64 (S,4) rows {a0, na, b0, nb} (host or device).  Returns (flat f64 tensor, out0 list): matrix s is
    flat[out0[s] : out0[s] + na*nb].view(na, nb)."""
    f64 = a.dtype == torch.float64            # float64 boxes: corners in float64 too (the evaluator's numpy arithmetic)
    a = _dev(a, torch.float64 if f64 else torch.float32, 'a').contiguous()
    b = _dev(b, torch.float64 if f64 else torch.float32, 'b').contiguous()
    seg = torch.as_tensor(segments, dtype=torch.int64).reshape(-1, 4).cpu()
    sizes = seg[:, 1] * seg[:, 3]
    out0 = torch.cumsum(sizes, 0) - sizes
    total = int(sizes.sum())
    seg5 = torch.cat([seg, out0[:, None]], dim=1).contiguous().to(a.device)
    out = torch.empty(max(total, 1), dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.lib().mydet_iou_rot_segments(_ptr(a), _ptr(b), 1 if f64 else 0, _ptr(seg5), seg5.shape[0], total, _ptr(out), _stream())
    _lib.check(rc, 'mydet_iou_rot_segments')
    return out[:total], out0.tolist()


# --------------------------------------------------------------------------------------- ATSS
def fcos_assign(t_ltrb, stride, img_hw, gt_box, gt_cls, gt_count, center_region, anch_min, anch_max, ignore_thres, n_cls):
    """FCOSLayer's targets of one level (mydet_fcos_assign).  Same tensors as atss_assign (no 'thr')."""
    t = _dev(t_ltrb, torch.float32, 't_ltrb')
    gt_box = _dev(gt_box, torch.float32, 'gt_box').contiguous()
    gt_cls = _dev(gt_cls, torch.int64, 'gt_cls').contiguous()
    gt_count = _dev(gt_count, torch.int32, 'gt_count').contiguous()
    B, n_h, n_w, _ = t.shape
    G = gt_box.shape[1]
    dev = t.device
    pos = torch.empty(B, n_h, n_w, dtype=torch.uint8, device=dev)
    ign = torch.empty(B, n_h, n_w, dtype=torch.uint8, device=dev)
    t_box = torch.empty(B, n_h, n_w, 4, dtype=torch.float32, device=dev)
    t_conf = torch.empty(B, n_h, n_w, 1, dtype=torch.float32, device=dev)
    t_cls = torch.empty(B, n_h, n_w, n_cls, dtype=torch.float32, device=dev)
    L = _lib.lib()
    ws = _workspace(L.mydet_atss_workspace_bytes(B, G), dev)
    st = (ctypes.c_int64 * 4)(*t.stride())
    with torch.cuda.device(dev):
        rc = L.mydet_fcos_assign(_ptr(t), st, B, int(stride), int(img_hw[0]), int(img_hw[1]), _ptr(gt_box), _ptr(gt_cls),
                                 _ptr(gt_count), G, float(center_region), float(anch_min), float(anch_max),
                                 float(ignore_thres), int(n_cls), _ptr(pos), _ptr(ign), _ptr(t_box), _ptr(t_conf),
                                 _ptr(t_cls), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, 'mydet_fcos_assign')
    return {'PositiveMask': pos.view(torch.bool), 'IgnoredMask': ign.view(torch.bool), 'TargetLTRB': t_box, 'TargetConf': t_conf,
            'TargetCls': t_cls}


def atss_assign(t_ltrb, level, strides, anchor_sides, img_hw, gt_box, gt_cls, gt_count, topk, ignore_thres, n_cls,
                thr=None):
    """Targets of one level.  t_ltrb (B,nH,nW,4) view; gt_box (B,G,4) f32, gt_cls (B,G) i64,
    gt_count (B) i32.  Returns dict of PositiveMask/IgnoredMask (bool), TargetLTRB, TargetConf,
    TargetCls and 'thr' (B,G).  Pass the 'thr' of an earlier level's call (same GT) as `thr` to skip the
    per-GT nearest-anchor search, which does not depend on the level."""
    t = _dev(t_ltrb, torch.float32, 't_ltrb')
    gt_box = _dev(gt_box, torch.float32, 'gt_box').contiguous()
    gt_cls = _dev(gt_cls, torch.int64, 'gt_cls').contiguous()
    gt_count = _dev(gt_count, torch.int32, 'gt_count').contiguous()
    B, n_h, n_w, _ = t.shape
    G = gt_box.shape[1]
    dev = t.device
    pos = torch.empty(B, n_h, n_w, dtype=torch.uint8, device=dev)
    ign = torch.empty(B, n_h, n_w, dtype=torch.uint8, device=dev)
    t_box = torch.empty(B, n_h, n_w, 4, dtype=torch.float32, device=dev)
    t_conf = torch.empty(B, n_h, n_w, 1, dtype=torch.float32, device=dev)
    t_cls = torch.empty(B, n_h, n_w, n_cls, dtype=torch.float32, device=dev)
    thr_is_input = thr is not None
    if thr is None:
        thr = torch.full((B, max(G, 1)), float('nan'), dtype=torch.float32, device=dev)
    else:
        thr = _dev(thr, torch.float32, 'thr').contiguous()
    L = _lib.lib()
    ws = _workspace(L.mydet_atss_workspace_bytes(B, G), dev)
    n_l = len(strides)
    st = (ctypes.c_int64 * 4)(*t.stride())
    c_strides = (ctypes.c_int32 * n_l)(*[int(s) for s in strides])
    c_sides = (ctypes.c_float * n_l)(*[float(s) for s in anchor_sides])
    with torch.cuda.device(dev):
        rc = L.mydet_atss_assign(_ptr(t), st, B, int(level), n_l, c_strides, c_sides, int(img_hw[0]), int(img_hw[1]),
                                 _ptr(gt_box), _ptr(gt_cls), _ptr(gt_count), G, int(topk), float(ignore_thres),
                                 int(n_cls), _ptr(pos), _ptr(ign), _ptr(t_box), _ptr(t_conf), _ptr(t_cls), _ptr(thr),
                                 1 if thr_is_input else 0, _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, 'mydet_atss_assign')
    return {'PositiveMask': pos.view(torch.bool), 'IgnoredMask': ign.view(torch.bool), 'TargetLTRB': t_box, 'TargetConf': t_conf,
            'TargetCls': t_cls, 'thr': thr}


def atss_assign_levels(t_ltrbs, strides, anchor_sides, img_hw, gt_box, gt_cls, gt_count, topk, ignore_thres, n_cls):
    """Targets of ALL levels in one call (mydet_atss_assign_levels): t_ltrbs = list of (B,nH_l,nW_l,4) views, finest level
    first.  Returns a list of dicts like atss_assign's (the 'thr' entry is shared); bit-identical to calling
    atss_assign level by level, in 3 launches and 6 allocations instead of 15 and ~35."""
    ts = [_dev(t, torch.float32, 't_ltrb') for t in t_ltrbs]
    gt_box = _dev(gt_box, torch.float32, 'gt_box').contiguous()
    gt_cls = _dev(gt_cls, torch.int64, 'gt_cls').contiguous()
    gt_count = _dev(gt_count, torch.int32, 'gt_count').contiguous()
    n_l = len(ts)
    if n_l != len(strides) or n_l != len(anchor_sides):
        raise ValueError('one t_ltrb view, stride and anchor side per level')
    B, G, dev = ts[0].shape[0], gt_box.shape[1], ts[0].device
    cells = [t.shape[1] * t.shape[2] for t in ts]
    total = sum(cells)
    # one allocation per kind; level l owns the contiguous block [B * sum(cells[:l]), B * sum(cells[:l+1]))
    pos = torch.empty(B * total, dtype=torch.uint8, device=dev)
    ign = torch.empty(B * total, dtype=torch.uint8, device=dev)
    t_box = torch.empty(B * total * 4, dtype=torch.float32, device=dev)
    t_conf = torch.empty(B * total, dtype=torch.float32, device=dev)
    t_cls = torch.empty(B * total * n_cls, dtype=torch.float32, device=dev)
    thr = torch.full((B, max(G, 1)), float('nan'), dtype=torch.float32, device=dev)
    L = _lib.lib()
    ws = _workspace(L.mydet_atss_workspace_bytes(B, G), dev)
    lv = (_lib.AtssLevel * n_l)()
    out, off = [], 0
    for i, t in enumerate(ts):
        n_h, n_w = t.shape[1], t.shape[2]
        lo, hi = B * off, B * (off + cells[i])
        views = {'PositiveMask': pos[lo:hi].view(B, n_h, n_w), 'IgnoredMask': ign[lo:hi].view(B, n_h, n_w),
                 'TargetLTRB': t_box[lo * 4:hi * 4].view(B, n_h, n_w, 4), 'TargetConf': t_conf[lo:hi].view(B, n_h, n_w, 1),
                 'TargetCls': t_cls[lo * n_cls:hi * n_cls].view(B, n_h, n_w, n_cls)}
        lv[i].t_ltrb = t.data_ptr()
        lv[i].t_stride[:] = list(t.stride())
        lv[i].positive, lv[i].ignored = views['PositiveMask'].data_ptr(), views['IgnoredMask'].data_ptr()
        lv[i].target_ltrb, lv[i].target_conf = views['TargetLTRB'].data_ptr(), views['TargetConf'].data_ptr()
        lv[i].target_cls = views['TargetCls'].data_ptr()
        views['PositiveMask'] = views['PositiveMask'].view(torch.bool)
        views['IgnoredMask'] = views['IgnoredMask'].view(torch.bool)
        views['thr'] = thr
        out.append(views)
        off += cells[i]
    c_strides = (ctypes.c_int32 * n_l)(*[int(s) for s in strides])
    c_sides = (ctypes.c_float * n_l)(*[float(s) for s in anchor_sides])
    with torch.cuda.device(dev):
        rc = L.mydet_atss_assign_levels(lv, n_l, c_strides, c_sides, B, int(img_hw[0]), int(img_hw[1]), _ptr(gt_box), _ptr(gt_cls),
                                        _ptr(gt_count), G, int(topk), float(ignore_thres), int(n_cls), _ptr(thr), _ptr(ws),
                                        ws.numel(), _stream())
    _lib.check(rc, 'mydet_atss_assign_levels')
    return out
