"""Oracle: per-level head decode.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every function takes the *raw* dict a reference det layer receives (permuted
views of NCHW conv outputs, models/rpns.py:29-41, :175-189) and returns the
triple the reference layer returns in test mode:
    bbox  (B, nA*nH*nW, P) float32
    cls   (B, nA*nH*nW)    int64
    score (B, nA*nH*nW)    float32
The float arithmetic uses the same torch CPU operators, in the same order and
with the same scalar types, as the reference, so that the outputs are
bit-identical to the reference on the same machine (checked in
tests/test_oracle_golden.py against fixtures produced by the real reference).
"""
import math

import torch


def _grid(n_h, n_w):
    """Cell indices as float32 column / row vectors (yolov3.py:45-46, fcos2.py:436-437)."""
    rows = torch.arange(n_h, dtype=torch.float32).view(n_h, 1)
    cols = torch.arange(n_w, dtype=torch.float32).view(1, n_w)
    return rows, cols


def _class_score(cls_logits):
    """max_c sigmoid(logit_c) and the FIRST index attaining it (torch.max, yolov3.py:57-58)."""
    probs = torch.sigmoid(cls_logits)
    return torch.max(probs, dim=-1)


def _flatten(box, cls_idx, score):
    n_b, p = box.shape[0], box.shape[-1]
    return (box.reshape(n_b, -1, p).contiguous(),
            cls_idx.reshape(n_b, -1).to(torch.int64).contiguous(),
            score.reshape(n_b, -1).contiguous())


def decode_yolo(raw, anchors_wh, stride, n_cls):
    """YOLOv3 decode -- models/detlayers/yolov3.py:41-69.

    raw['bbox'] (B,nA,nH,nW,4), raw['conf'] (B,nA,nH,nW,1), raw['class'] (B,nA,nH,nW,C).
    anchors_wh: (nA,2) float32 tensor, this level's anchors in pixels.
    """
    # :43 -- the reference works on a CONTIGUOUS copy, so the per-channel slices below are
    # strided (stride 4) and torch takes its scalar sigmoid/exp path, whose last bit can differ
    # from the vectorised path; keep the same layout to stay bit-identical.
    t = raw['bbox'].detach().clone().contiguous()
    n_b, n_a, n_h, n_w, _ = t.shape
    rows, cols = _grid(n_h, n_w)
    box = torch.empty(n_b, n_a, n_h, n_w, 4, dtype=torch.float32)
    box[..., 0] = (torch.sigmoid(t[..., 0]) + cols) * stride           # :47
    box[..., 1] = (torch.sigmoid(t[..., 1]) + rows) * stride           # :48
    box[..., 2:4] = torch.exp(t[..., 2:4]) * anchors_wh.view(1, n_a, 1, 1, 2)  # :50-51
    p_conf = torch.sigmoid(raw['conf'].detach())[..., 0]               # :54
    if n_cls > 0:
        best, idx = _class_score(raw['class'].detach())               # :57-58
        score = p_conf * best                                          # :59
    else:
        idx = torch.zeros(n_b, n_a, n_h, n_w, dtype=torch.int64)       # :61
        score = p_conf                                                 # :62
    return _flatten(box, idx, score)


def _fcos_boxes(t_ltrb, stride, img_hw):
    """exp-ltrb -> clamped x1y1x2y2 -> cxcywh: fcos2.py:41-56 with helpers :417-458."""
    img_h, img_w = img_hw
    n_h, n_w = t_ltrb.shape[-3], t_ltrb.shape[-2]
    ltrb = torch.exp(t_ltrb) * stride                                  # :42
    rows, cols = _grid(n_h, n_w)
    cy = rows * stride + stride / 2                                    # :441
    cx = cols * stride + stride / 2                                    # :442
    x1 = (cx - ltrb[..., 0]).clamp_(min=0, max=img_w)                  # :454, :51
    y1 = (cy - ltrb[..., 1]).clamp_(min=0, max=img_h)
    x2 = (cx + ltrb[..., 2]).clamp_(min=0, max=img_w)
    y2 = (cy + ltrb[..., 3]).clamp_(min=0, max=img_h)
    return torch.stack([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], dim=-1)  # :420-423


def decode_fcos(raw, stride, img_hw, conf_key='conf'):
    """FCOS / FCOS2 / FCOS2_ATSS decode -- fcos2.py:40-69, :222-251; v1 fcos.py:41-68
    (v1 reads the centerness head: conf_key='center').

    raw['bbox'] (B,nH,nW,4), raw[conf_key] (B,nH,nW,1), raw['class'] (B,nH,nW,C), C>0.
    """
    box = _fcos_boxes(raw['bbox'].detach(), stride, img_hw)
    p_conf = torch.sigmoid(raw[conf_key].detach())[..., 0]             # :58
    best, idx = _class_score(raw['class'].detach())                    # :60-61
    score = torch.sqrt(p_conf * best)                                  # :62
    return _flatten(box, idx, score)


def decode_rapid(raw, anchors_wh, stride, n_cls):
    """RAPiD rotated decode -- models/detlayers/rapid.py:48-82.  bbox is (cx,cy,w,h,degrees)."""
    t = raw['bbox'].detach()
    n_b, n_a, n_h, n_w, _ = t.shape
    rows, cols = _grid(n_h, n_w)
    radian = torch.sigmoid(t[..., 4]) * 2 * math.pi - math.pi          # :49
    box = torch.empty(n_b, n_a, n_h, n_w, 5, dtype=torch.float32)
    box[..., 0] = (torch.sigmoid(t[..., 0]) + cols) * stride           # :58
    box[..., 1] = (torch.sigmoid(t[..., 1]) + rows) * stride           # :59
    box[..., 2:4] = torch.exp(t[..., 2:4]) * anchors_wh.view(1, n_a, 1, 1, 2)  # :61-62
    box[..., 4] = radian / math.pi * 180                               # :63
    p_conf = torch.sigmoid(raw['conf'].detach())[..., 0]               # :67
    if n_cls > 0:
        best, idx = _class_score(raw['class'].detach())               # :70-71
        score = torch.sqrt(p_conf * best)                              # :72
    else:
        idx = torch.zeros(n_b, n_a, n_h, n_w, dtype=torch.int64)       # :75
        score = p_conf
    return _flatten(box, idx, score)


def decode_retina(raw, anchors_wh, stride, img_hw, with_angle=False):
    """RetinaNet anchor-delta decode -- models/detlayers/retinanet.py:55-82 (no 'conf' head)."""
    img_h, img_w = img_hw
    t = raw['bbox'].detach()
    n_b, n_a, n_h, n_w, n_p = t.shape
    a_cx = torch.arange(stride / 2, img_w, stride).view(1, 1, 1, n_w)  # :57
    a_cy = torch.arange(stride / 2, img_h, stride).view(1, 1, n_h, 1)  # :58
    a_wh = anchors_wh.view(1, n_a, 1, 1, 2)
    box = torch.empty(n_b, n_a, n_h, n_w, n_p, dtype=torch.float32)
    box[..., 0] = a_cx + t[..., 0] * a_wh[..., 0]                      # :67
    box[..., 1] = a_cy + t[..., 1] * a_wh[..., 1]                      # :68
    box[..., 2:4] = torch.exp(t[..., 2:4]) * a_wh                      # :69
    box[..., 0:4].clamp_(min=1, max=max(img_hw))                       # :70
    if with_angle:
        box[..., 4] = torch.sigmoid(t[..., 4]) * 360 - 180             # :72
    best, idx = _class_score(raw['class'].detach())                    # :74-75
    return _flatten(box, idx, best)


def decode_uv5(raw, anchors_wh, stride):
    """Ultralytics/YOLOv5 decode -- models/detlayers/uv5.py:60-91 (requires C>0)."""
    t = raw['bbox'].detach()
    n_b, n_a, n_h, n_w, _ = t.shape
    rows, cols = _grid(n_h, n_w)
    s = torch.sigmoid(t)                                               # :68
    box = torch.empty(n_b, n_a, n_h, n_w, 4, dtype=torch.float32)
    box[..., 0] = (s[..., 0] * 2 - 0.5 + cols) * stride                # :70
    box[..., 1] = (s[..., 1] * 2 - 0.5 + rows) * stride
    box[..., 2:4] = (s[..., 2:4] * 2) ** 2 * anchors_wh.view(1, n_a, 1, 1, 2)  # :73
    p_conf = torch.sigmoid(raw['conf'].detach())[..., 0]               # :81
    best, idx = _class_score(raw['class'].detach())                    # :85
    return _flatten(box, idx, p_conf * best)                           # :86


def merge_levels(per_level):
    """Level concatenation along dim 1, strides ascending -- models/general.py:74-76."""
    return tuple(torch.cat([lvl[k] for lvl in per_level], dim=1) for k in range(3))
