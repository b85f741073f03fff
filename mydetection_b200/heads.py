"""The raw dicts the reference heads hand to the det layers, rebuilt as zero-copy views of NCHW
conv outputs (models/rpns.py:29-41 YOLOHead, :175-189 EfDetHead).  Layout only, no arithmetic."""


def yolo_head_views(nchw, n_anchor, n_param, n_cls):
    """(B, nA*(P+1+C), nH, nW) -> {'bbox' (B,nA,nH,nW,P), 'conf' (...,1), 'class' (...,C)}."""
    n_b, _, n_h, n_w = nchw.shape
    v = nchw.view(n_b, n_anchor, n_param + 1 + n_cls, n_h, n_w)
    return {'bbox': v[:, :, 0:n_param].permute(0, 1, 3, 4, 2),
            'conf': v[:, :, n_param:n_param + 1].permute(0, 1, 3, 4, 2),
            'class': v[:, :, n_param + 1:].permute(0, 1, 3, 4, 2)}


def efdet_head_views(bbox_nchw, cls_nchw):
    """bbox (B,4,nH,nW) + cls (B,1+C,nH,nW) -> {'bbox' (B,nH,nW,4), 'conf' (B,nH,nW,1), 'class' (B,nH,nW,C)}."""
    c = cls_nchw.permute(0, 2, 3, 1)
    return {'bbox': bbox_nchw.permute(0, 2, 3, 1), 'conf': c[..., 0:1], 'class': c[..., 1:]}
