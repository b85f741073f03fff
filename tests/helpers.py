"""Shared helpers for the parity tests: rebuild the views the reference heads emit from
stored NCHW tensors (models/rpns.py:29-41, :175-189) and the constants of the golden cases."""
import numpy as np
import torch

YOLO_ANCHORS = [[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198], [373, 326]]
RAPID_ANCHORS = [[18.7807, 33.4659], [28.8912, 61.7536], [48.6849, 68.3897], [45.0668, 101.4673],
                 [63.0952, 113.5382], [81.3909, 134.4554], [91.7364, 144.9949], [137.5189, 178.4791],
                 [194.4429, 250.7985]]


def T(a):
    return torch.from_numpy(np.asarray(a))


def level_anchors(table, level_i):
    return torch.tensor(table, dtype=torch.float32)[3 * level_i:3 * level_i + 3]


def yolo_views(nchw, n_a, n_p, n_c):
    """YOLOHead.forward views of a (B, nA*(P+1+C), nH, nW) tensor."""
    n_b, _, n_h, n_w = nchw.shape
    v = nchw.view(n_b, n_a, n_p + 1 + n_c, n_h, n_w)
    return {'bbox': v[:, :, 0:n_p].permute(0, 1, 3, 4, 2),
            'conf': v[:, :, n_p:n_p + 1].permute(0, 1, 3, 4, 2),
            'class': v[:, :, n_p + 1:].permute(0, 1, 3, 4, 2)}


def efdet_views(bbox_nchw, cls_nchw):
    """EfDetHead.forward views (nA == 1, enable_conf): bbox (B,4,nH,nW), cls (B,1+C,nH,nW)."""
    c = cls_nchw.permute(0, 2, 3, 1)
    return {'bbox': bbox_nchw.permute(0, 2, 3, 1), 'conf': c[..., 0:1], 'class': c[..., 1:]}


def anchor_views(bbox_nchw, cls_nchw, n_a):
    """Multi-anchor EfDetHead views without a conf head (RetinaNet)."""
    n_b, _, n_h, n_w = bbox_nchw.shape
    return {'bbox': bbox_nchw.view(n_b, n_a, -1, n_h, n_w).permute(0, 1, 3, 4, 2),
            'class': cls_nchw.view(n_b, n_a, -1, n_h, n_w).permute(0, 1, 3, 4, 2)}
