"""Generate tests/golden/*.npz by running the UNMODIFIED reference on seeded inputs.

Runs only in the build container (needs /root/reference); the fixtures it writes
are committed, and nothing at test/bench time reads /root/reference.

    python tests/golden/make_golden.py

Third-party imports the reference needs but this image lacks are stubbed in
sys.modules: matplotlib, fvcore, pycocotools.  The
pycocotools stub is NOT the real rasteriser: its ``iou`` is an independent
pure-Python exact polygon-clipping IoU, used only to pin nms_rotbb's CONTROL FLOW
(utils/bbox_ops.py:276-306) -- rotated IoU values remain "parity unpinned".
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get('MYDET_REFERENCE', '/root/reference')
OUT = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------- stubs
def _poly_area(p):
    return 0.5 * sum(p[i][0] * p[(i + 1) % len(p)][1] - p[(i + 1) % len(p)][0] * p[i][1]
                     for i in range(len(p)))


def _clip(subject, clipper):
    """Sutherland-Hodgman, pure Python, independent of oracle/rotiou.c."""
    sgn = 1.0 if _poly_area(clipper) >= 0 else -1.0
    out = list(subject)
    for i in range(len(clipper)):
        a, b = clipper[i], clipper[(i + 1) % len(clipper)]
        ex, ey = b[0] - a[0], b[1] - a[1]
        src, out = out, []
        if not src:
            break
        for k in range(len(src)):
            p, q = src[k], src[(k + 1) % len(src)]
            dp = sgn * (ex * (p[1] - a[1]) - ey * (p[0] - a[0]))
            dq = sgn * (ex * (q[1] - a[1]) - ey * (q[0] - a[0]))
            if dp >= 0:
                out.append(p)
            if (dp >= 0) != (dq >= 0):
                t = dp / (dp - dq)
                out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
    return out


def _poly_iou(v1, v2):
    p1 = [(v1[2 * i], v1[2 * i + 1]) for i in range(4)]
    p2 = [(v2[2 * i], v2[2 * i + 1]) for i in range(4)]
    inter_poly = _clip(p1, p2)
    inter = abs(_poly_area(inter_poly)) if len(inter_poly) >= 3 else 0.0
    union = abs(_poly_area(p1)) + abs(_poly_area(p2)) - inter
    return inter / union if union > 0 else 0.0


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mask = mod('pycocotools.mask',
               frPyObjects=lambda polys, h, w: polys,
               iou=lambda d, g, crowd: np.array([[_poly_iou(a, b) for b in g] for a in d], dtype=np.float64))
    mod('pycocotools', mask=mask)
    mod('pycocotools.coco', COCO=object)
    mod('pycocotools.cocoeval', COCOeval=object)
    plt = mod('matplotlib.pyplot')
    mod('matplotlib', pyplot=plt)
    def smooth_l1_loss(input, target, beta, reduction='none'):
        # fvcore.nn.smooth_l1_loss restated (fvcore is not in this image); only RetinaLayer's training loss uses it
        n = torch.abs(input - target)
        loss = torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta) if beta >= 1e-5 else n
        return loss.sum() if reduction == 'sum' else (loss.mean() if reduction == 'mean' else loss)
    nn = mod('fvcore.nn', smooth_l1_loss=smooth_l1_loss, sigmoid_focal_loss=None)
    mod('fvcore', nn=nn)


def import_reference():
    install_stubs()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings('ignore')


# ----------------------------------------------------------------------------- inputs
def head_views(gen, n_b, n_a, n_h, n_w, n_p, n_c, conf_mu=0.0, separate=False):
    """NCHW head output + the permuted views the reference heads hand to the det layers
    (models/rpns.py:29-41 for the YOLO head, :175-189 for the EfficientDet head)."""
    if separate:  # EfDetHead: bbox tensor and (conf+class) tensor, nA == 1
        bb = torch.randn(n_b, n_p, n_h, n_w, generator=gen) * 0.5
        cc = torch.randn(n_b, 1 + n_c, n_h, n_w, generator=gen) * 1.5
        cc[:, 0] += conf_mu
        cc[:, 1:] -= 1.0
        raw = {'bbox': bb.permute(0, 2, 3, 1), 'conf': cc.permute(0, 2, 3, 1)[..., 0:1],
               'class': cc.permute(0, 2, 3, 1)[..., 1:]}
        return {'bbox_nchw': bb, 'cls_nchw': cc}, raw
    ch = n_p + 1 + n_c
    t = torch.randn(n_b, n_a * ch, n_h, n_w, generator=gen)
    v = t.view(n_b, n_a, ch, n_h, n_w)
    v[:, :, :n_p] *= 0.5
    v[:, :, n_p] = v[:, :, n_p] * 1.5 + conf_mu
    v[:, :, n_p + 1:] = v[:, :, n_p + 1:] * 1.5 - 1.0
    raw = {'bbox': v[:, :, 0:n_p].permute(0, 1, 3, 4, 2),
           'conf': v[:, :, n_p:n_p + 1].permute(0, 1, 3, 4, 2),
           'class': v[:, :, n_p + 1:].permute(0, 1, 3, 4, 2)}
    return {'nchw': t}, raw


def np_(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def save(name, **arrays):
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **np_(arrays))
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} kB')


YOLO_ANCHORS = [[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198], [373, 326]]
RAPID_ANCHORS = [[18.7807, 33.4659], [28.8912, 61.7536], [48.6849, 68.3897], [45.0668, 101.4673],
                 [63.0952, 113.5382], [81.3909, 134.4554], [91.7364, 144.9949], [137.5189, 178.4791],
                 [194.4429, 250.7985]]
IDX3 = [[0, 1, 2], [3, 4, 5], [6, 7, 8]]


# ----------------------------------------------------------------------------- cases
def gen_decode():
    from models.detlayers.yolov3 import YOLOLayer
    from models.detlayers.fcos2 import FCOSLayer as FCOS2Layer, FCOS_ATSS_Layer
    from models.detlayers.fcos import FCOSLayer as FCOS1Layer
    from models.detlayers.rapid import RAPiDLayer
    from models.detlayers.retinanet import RetinaLayer
    from models.detlayers.uv5 import DetectLayer

    gen = torch.Generator().manual_seed(1001)
    img_hw = (96, 128)
    strides = [8, 16, 32]
    out = {}
    # --- YOLOv3, 3 anchors, 5 classes; level 2 is 3x4 (odd plane, exercises the scalar path)
    cfg = {'model.yolo.anchors': YOLO_ANCHORS, 'model.yolo.anchor_indices': IDX3,
           'model.yolo.anchor.negative_threshold': 0.7, 'model.fpn.out_strides': strides,
           'general.num_class': 5}
    for li, s in enumerate(strides):
        store, raw = head_views(gen, 2, 3, img_hw[0] // s, img_hw[1] // s, 4, 5)
        p, _ = YOLOLayer(li, cfg)(raw, img_hw, None)
        out.update({f'yolo{li}_in': store['nchw'], f'yolo{li}_bbox': p['bbox'],
                    f'yolo{li}_cls': p['class_idx'], f'yolo{li}_score': p['score']})
    # --- YOLOv3 with zero classes (score = sigmoid(conf))
    cfg0 = dict(cfg, **{'general.num_class': 0})
    store, raw = head_views(gen, 2, 3, 6, 8, 4, 0)
    p, _ = YOLOLayer(1, cfg0)(raw, img_hw, None)
    out.update({'yolo_c0_in': store['nchw'], 'yolo_c0_bbox': p['bbox'], 'yolo_c0_cls': p['class_idx'],
                'yolo_c0_score': p['score']})
    # --- RAPiD, zero classes (configs/rapid.json) and 3 classes
    for tag, nc in (('rapid_c0', 0), ('rapid_c3', 3)):
        cfgr = {'model.rapid.anchors': RAPID_ANCHORS, 'model.rapid.anchor_indices': IDX3,
                'model.fpn.out_strides': strides, 'general.num_class': nc,
                'model.rapid.wh_smooth_l1_beta': 1, 'model.angle.loss_angle': 'Periodic_L1',
                'model.angle.pred_range': 360}
        for li, s in enumerate(strides):
            store, raw = head_views(gen, 2, 3, img_hw[0] // s, img_hw[1] // s, 5, nc)
            raw['bbox'][..., 4].mul_(4.0)  # angle logits spread over (-3, 3)-ish
            p, _ = RAPiDLayer(li, cfgr)(raw, img_hw, None)
            out.update({f'{tag}_{li}_in': store['nchw'], f'{tag}_{li}_bbox': p['bbox'],
                        f'{tag}_{li}_cls': p['class_idx'], f'{tag}_{li}_score': p['score']})
    # --- FCOS2 / FCOS2_ATSS / FCOS v1 on an EfficientDet-style head, 5 levels, 6 classes
    img2 = (256, 384)
    strides5 = [8, 16, 32, 64, 128]
    cfgf = {'model.fcos.anchors': [0, 64, 128, 256, 512, 100000000], 'model.fpn.out_strides': strides5,
            'general.num_class': 6, 'model.fcos2.ignored_threshold': 0.7,
            'general.pred_bbox_format': 'cxcywh', 'model.atss.anchors': [24, 48, 96, 192, 384],
            'model.atss.topk_per_level': 9}
    for li, s in enumerate(strides5):
        store, raw = head_views(gen, 2, 1, img2[0] // s, img2[1] // s, 4, 6, separate=True)
        raw['bbox'].mul_(3.0)  # make some boxes hit the image-border clamp
        p, _ = FCOS2Layer(li, cfgf)(raw, img2, None)
        p_atss, _ = FCOS_ATSS_Layer(li, cfgf)(raw, img2, None)
        assert all(torch.equal(p[k], p_atss[k]) for k in p)
        raw1 = {'bbox': raw['bbox'], 'center': raw['conf'], 'class': raw['class']}
        p1, _ = FCOS1Layer(li, cfgf)(raw1, img2, None)
        assert all(torch.equal(p[k], p1[k]) for k in p)
        out.update({f'fcos{li}_bbox_in': store['bbox_nchw'], f'fcos{li}_cls_in': store['cls_nchw'],
                    f'fcos{li}_bbox': p['bbox'], f'fcos{li}_cls': p['class_idx'], f'fcos{li}_score': p['score']})
    # --- RetinaNet (9 anchors), plain and rotated
    for tag, npar, fmt in (('retina', 4, 'cxcywh'), ('retina_rot', 5, 'cxcywhd')):
        cfgt = {'model.fpn.out_strides': strides, 'model.retina.anchor.base': 4,
                'model.retina.anchor.scales': [1, 1.26, 1.5874],
                'model.retina.anchor.ratios': [[1, 1], [1.4, 0.7], [0.7, 1.4]],
                'model.retina.anchor.positive_threshold': 0.5, 'model.retina.anchor.negative_threshold': 0.5,
                'general.num_class': 4, 'general.pred_bbox_format': fmt, 'general.bbox_param': npar,
                'model.angle.loss_name': 'Periodic_L1'}
        li, s = 1, 16
        n_h, n_w = img_hw[0] // s, img_hw[1] // s
        bb = torch.randn(2, 9 * npar, n_h, n_w, generator=gen) * 0.5
        cc = torch.randn(2, 9 * 4, n_h, n_w, generator=gen) * 1.5
        raw = {'bbox': bb.view(2, 9, npar, n_h, n_w).permute(0, 1, 3, 4, 2),
               'class': cc.view(2, 9, 4, n_h, n_w).permute(0, 1, 3, 4, 2)}
        # the 'cxcywhd' constructor branch is broken at HEAD (retinanet.py:36 imports a module
        # that does not exist), so build the plain layer and switch the two attributes the
        # decode branch reads (:71, :53)
        layer = RetinaLayer(li, dict(cfgt, **{'general.pred_bbox_format': 'cxcywh'}))
        layer.pred_bbox_format, layer.n_bbparam = fmt, npar
        p, _ = layer(raw, img_hw, None)
        out.update({f'{tag}_bbox_in': bb, f'{tag}_cls_in': cc, f'{tag}_anchors': layer.anchor_wh,
                    f'{tag}_bbox': p['bbox'], f'{tag}_cls': p['class_idx'], f'{tag}_score': p['score']})
    # --- Ultralytics / YOLOv5
    cfgu = {'model.detect.anchors': YOLO_ANCHORS, 'model.detect.anchor_indices': IDX3,
            'model.fpn.out_strides': strides, 'general.num_class': 5,
            'model.detect.sample_selection': 'best', 'model.detect.confidence_target': 'zero-one',
            'model.detect.loss_bbox': 'smooth_L1', 'general.pred_bbox_format': 'cxcywh'}
    store, raw = head_views(gen, 2, 3, 12, 16, 4, 5)
    p, _ = DetectLayer(0, cfgu)(raw, img_hw, None)
    out.update({'uv5_in': store['nchw'], 'uv5_bbox': p['bbox'], 'uv5_cls': p['class_idx'], 'uv5_score': p['score']})
    save('decode', **out)


def gen_postprocess():
    from utils.structures import ImageObjects
    gen = torch.Generator().manual_seed(1002)
    out = {}

    def boxes_(n, p, span=400.0):
        b = torch.empty(n, p)
        b[:, 0:2] = torch.rand(n, 2, generator=gen) * span
        b[:, 2:4] = torch.rand(n, 2, generator=gen) * 80 + 4
        if p == 5:
            b[:, 4] = torch.rand(n, generator=gen) * 360 - 180
        return b

    def run(tag, n, p, n_cls, fmt, conf, nms, direct=False):
        bxs = boxes_(n, p)
        sc = torch.rand(n, generator=gen)
        cats = torch.randint(0, max(n_cls, 1), (n,), generator=gen)
        keys = torch.cat([bxs, sc[:, None], cats[:, None].float()], dim=1)
        obj = ImageObjects(bxs.clone(), cats.clone(), scores=sc.clone(), bb_format=fmt, img_hw=(512, 512))
        res = obj.nms(nms) if direct else obj.post_process(conf, nms)
        # recover kept indices: rows are unique with probability 1
        got = torch.cat([res.bboxes, res.scores[:, None], res.cats[:, None].float()], dim=1)
        idx = torch.tensor([int(torch.nonzero((keys == g).all(dim=1))[0, 0]) for g in got], dtype=torch.int64)
        assert idx.unique().numel() == idx.numel()
        out.update({f'{tag}_boxes': bxs, f'{tag}_scores': sc, f'{tag}_cats': cats, f'{tag}_keep': idx,
                    f'{tag}_params': np.array([conf, nms], dtype=np.float64)})

    run('pp_small', 300, 4, 5, 'cxcywh', 0.3, 0.45)           # below the 512 cap
    run('pp_cap', 2000, 4, 7, 'cxcywh', 0.05, 0.5)            # top-512 kicks in
    run('pp_rot', 1500, 5, 1, 'cxcywhd', 0.01, 0.45)          # angle ignored, single class
    run('pp_empty', 50, 4, 3, 'cxcywh', 2.0, 0.5)             # nothing survives
    run('nms_direct', 1200, 4, 3, 'cxcywh', 0.0, 0.3, direct=True)   # un-capped ImageObjects.nms
    # adversarial: duplicates, zero-area and contained boxes, exact-threshold IoU, all one class
    b = torch.tensor([[10., 10, 20, 20], [10, 10, 20, 20], [10, 10, 0, 0], [12, 10, 20, 20],
                      [100, 100, 30, 30], [100, 100, 10, 10], [30, 10, 20, 20], [10, 30, 20, 20]])
    s = torch.tensor([0.9, 0.8, 0.7, 0.6, 0.5, 0.4, 0.3, 0.2])
    obj = ImageObjects(b.clone(), torch.zeros(8, dtype=torch.int64), scores=s.clone(), bb_format='cxcywh')
    res = obj.nms(1.0 / 9.0)   # IoU(100,100,30,30 vs 10x10 inside) == 1/9 exactly?  strict '>' keeps it
    out.update({'adv_boxes': b, 'adv_scores': s, 'adv_keep_scores': res.scores,
                'adv_thr': np.array([1.0 / 9.0])})
    save('postprocess', **out)


def gen_iou():
    from utils.bbox_ops import bboxes_iou, nms_rotbb, xywha2vertex, cxcywh_to_x1y1x2y2
    import json
    gen = torch.Generator().manual_seed(1003)
    a = torch.rand(37, 4, generator=gen) * 100 + 1
    b = torch.rand(11, 4, generator=gen) * 100 + 1
    ax = cxcywh_to_x1y1x2y2(a)
    bx = cxcywh_to_x1y1x2y2(b)
    out = {'a': a, 'b': b, 'iou_cxcywh': bboxes_iou(a, b, xyxy=False), 'iou_xyxy': bboxes_iou(ax, bx, xyxy=True),
           'a_xyxy': ax}
    # realistic GT from the reference's own debug fixture (datasets/debug/debug3.json, x1y1wh)
    anns = json.load(open(os.path.join(REF, 'datasets/debug/debug3.json')))['annotations']
    gt = torch.tensor([[x + w / 2, y + h / 2, w, h] for x, y, w, h in (an['bbox'] for an in anns)], dtype=torch.float32)
    out.update({'gt_debug3': gt, 'iou_gt_self': bboxes_iou(gt, gt)})
    # rotated: corners, and nms_rotbb control flow under the stub polygon IoU
    rb = torch.empty(120, 5)
    rb[:, 0:2] = torch.rand(120, 2, generator=gen) * 300 + 50
    rb[:, 2:4] = torch.rand(120, 2, generator=gen) * 90 + 10
    rb[:, 4] = torch.rand(120, generator=gen) * 360 - 180
    rs = torch.rand(120, generator=gen)
    rad = rb.clone()
    rad[:, 4] = rad[:, 4] * np.pi / 180
    out.update({'rot_boxes': rb, 'rot_scores': rs, 'rot_vertices': xywha2vertex(rad, is_degree=False),
                'rot_keep_045': nms_rotbb(rb, rs, 0.45), 'rot_keep_02': nms_rotbb(rb, rs, 0.2),
                'rot_keep_maj2': nms_rotbb(rb, rs, 0.3, majority=2)})
    rot_anns = json.load(open(os.path.join(REF, 'datasets/debug/rotbb_debug3.json')))['annotations']
    rgt = torch.tensor([an['bbox'] for an in rot_anns], dtype=torch.float32)
    from utils.bbox_ops import iou_rle
    out.update({'rot_gt_debug3': rgt, 'rot_gt_iou_stub': iou_rle(rgt[:40], rgt[:40])})
    save('iou', **out)


def gen_atss():
    from models.detlayers.fcos2 import FCOS_ATSS_Layer
    from utils.structures import ImageObjects
    gen = torch.Generator().manual_seed(1004)
    img_hw = (384, 512)   # the coarsest level must still hold >= k=9 anchors (torch.topk, fcos2.py:397)
    strides = [8, 16, 32, 64, 128]
    cfg = {'model.fpn.out_strides': strides, 'general.num_class': 6, 'model.fcos2.ignored_threshold': 0.7,
           'model.atss.anchors': [24, 48, 96, 192, 384], 'model.atss.topk_per_level': 9}
    n_b = 2
    labels = []
    gts = {}
    for b in range(n_b):
        n = 12 if b == 0 else 5
        bx = torch.empty(n, 4)
        bx[:, 0] = torch.rand(n, generator=gen) * (img_hw[1] - 40) + 20
        bx[:, 1] = torch.rand(n, generator=gen) * (img_hw[0] - 40) + 20
        bx[:, 2:4] = torch.rand(n, 2, generator=gen) * 120 + 12
        ct = torch.randint(0, 6, (n,), generator=gen)
        assert (bx[:, 2] * bx[:, 3]).unique().numel() == n
        labels.append(ImageObjects(bx, ct, bb_format='cxcywh', img_hw=img_hw))
        gts[f'gt{b}_boxes'] = bx
        gts[f'gt{b}_cats'] = ct
    out = dict(gts)
    for li, s in enumerate(strides):
        store, raw = head_views(gen, n_b, 1, img_hw[0] // s, img_hw[1] // s, 4, 6, separate=True)
        layer = FCOS_ATSS_Layer(li, cfg)
        grabbed = {}

        def prof(frame, event, arg):
            if event == 'return' and frame.f_code.co_name == 'forward' and 'PositiveMask' in frame.f_locals:
                for k in ('PositiveMask', 'IgnoredMask', 'TargetConf', 'TargetLTRB', 'TargetCls'):
                    grabbed[k] = frame.f_locals[k].clone()
        sys.setprofile(prof)
        try:
            layer(raw, img_hw, labels)
        finally:
            sys.setprofile(None)
        out.update({f'atss{li}_bbox_in': store['bbox_nchw'], f'atss{li}_cls_in': store['cls_nchw']})
        out.update({f'atss{li}_{k}': v for k, v in grabbed.items()})
    save('atss', **out)


def _grab_forward(layer, names, *args):
    """Run layer(*args) and return (result, {name: clone of the forward()'s local tensor})."""
    grabbed = {}

    def prof(frame, event, arg):
        if event == 'return' and frame.f_code.co_name == 'forward' and names[0] in frame.f_locals:
            for k in names:
                v = frame.f_locals.get(k)
                if torch.is_tensor(v):
                    grabbed[k] = v.detach().clone().cpu()
    sys.setprofile(prof)
    try:
        res = layer(*args)
    finally:
        sys.setprofile(None)
    return res, grabbed


def _random_gt(gen, n, img_hw, n_cls, lo=10.0, hi=200.0):
    bx = torch.empty(n, 4)
    bx[:, 0] = torch.rand(n, generator=gen) * (img_hw[1] - 40) + 20
    bx[:, 1] = torch.rand(n, generator=gen) * (img_hw[0] - 40) + 20
    bx[:, 2:4] = torch.exp(torch.rand(n, 2, generator=gen) * (np.log(hi) - np.log(lo)) + np.log(lo))
    ct = torch.randint(0, n_cls, (n,), generator=gen)
    assert (bx[:, 2] * bx[:, 3]).unique().numel() == n          # area ties would leave the GT order open
    return bx, ct


def gen_train():
    """Training branches: YOLOLayer and FCOSLayer (FCOS2) forward(raw, img_size, labels) of the
    unmodified reference -- target tensors captured from forward()'s locals, and the loss."""
    from models.detlayers.yolov3 import YOLOLayer
    from models.detlayers.fcos2 import FCOSLayer
    from utils.structures import ImageObjects
    gen = torch.Generator().manual_seed(1006)
    img_hw = (256, 320)
    n_cls = 5
    counts = [9, 0, 4]
    out, labels = {}, []
    for b, n in enumerate(counts):
        bx, ct = _random_gt(gen, n, img_hw, n_cls, 10.0, 330.0) if n else (torch.zeros(0, 4), torch.zeros(0, dtype=torch.int64))
        if b == 0:
            bx[-1, 2:4] = torch.tensor([300.0, 270.0])            # a GT for the coarsest YOLO level
            bx[-2, 2:4] = torch.tensor([150.0, 180.0])
        labels.append(ImageObjects(bx, ct, bb_format='cxcywh', img_hw=img_hw))
        out[f'gt{b}_boxes'], out[f'gt{b}_cats'] = bx, ct
    # ---- YOLO, three levels
    cfg = {'model.yolo.anchors': YOLO_ANCHORS, 'model.yolo.anchor_indices': IDX3, 'model.yolo.anchor.negative_threshold': 0.3,
           'model.fpn.out_strides': [8, 16, 32], 'general.num_class': n_cls}
    for li, s in enumerate(cfg['model.fpn.out_strides']):
        store, raw = head_views(gen, len(counts), 3, img_hw[0] // s, img_hw[1] // s, 4, n_cls, conf_mu=-1.0)
        layer = YOLOLayer(li, cfg)
        (_, loss), g = _grab_forward(layer, ['gt_mask', 'conf_loss_mask', 'tgt_xywh', 'tgt_cls', 'weighted'], raw, img_hw, labels)
        out[f'yolo{li}_in'] = store['nchw']
        out[f'yolo{li}_loss'] = loss
        out[f'yolo{li}_assigned'] = torch.tensor(int(layer._assigned_num))
        out.update({f'yolo{li}_{k}': v for k, v in g.items()})
    # ---- FCOS2 (FCOSLayer), three of the five levels
    strides = [8, 16, 32, 64, 128]
    cfg = {'model.fcos.anchors': [0, 64, 128, 256, 512, 100000000], 'model.fpn.out_strides': strides, 'general.num_class': n_cls,
           'model.fcos2.ignored_threshold': 0.2, 'general.pred_bbox_format': 'cxcywh'}
    n_pos = n_ign = 0
    for li in (0, 1, 2):
        s = strides[li]
        store, raw = head_views(gen, len(counts), 1, img_hw[0] // s, img_hw[1] // s, 4, n_cls, separate=True)
        layer = FCOSLayer(li, cfg)
        (_, loss), g = _grab_forward(layer, ['PositiveMask', 'IgnoredMask', 'TargetConf', 'TargetLTRB', 'TargetCls'], raw, img_hw, labels)
        out[f'fcos{li}_bbox_in'], out[f'fcos{li}_cls_in'] = store['bbox_nchw'], store['cls_nchw']
        out[f'fcos{li}_loss'] = loss
        out.update({f'fcos{li}_{k}': v for k, v in g.items()})
        n_pos += int(g['PositiveMask'].sum()); n_ign += int(g['IgnoredMask'].sum())
    assert n_pos > 0 and n_ign > 0
    # ---- RetinaNet, single class (the reference's training branch squeezes the class dim, retinanet.py:125)
    from models.detlayers.retinanet import RetinaLayer
    rcfg = {'model.fpn.out_strides': strides, 'model.retina.anchor.base': 4, 'model.retina.anchor.scales': [1, 1.26, 1.5874],
            'model.retina.anchor.ratios': [[1, 1], [1.4, 0.7], [0.7, 1.4]], 'model.retina.anchor.positive_threshold': 0.5,
            'model.retina.anchor.negative_threshold': 0.4, 'general.num_class': 1, 'general.pred_bbox_format': 'cxcywh',
            'general.bbox_param': 4}
    labels1 = [ImageObjects(l.bboxes, torch.zeros_like(l.cats), bb_format='cxcywh', img_hw=img_hw) for l in labels]
    for li in (1, 2, 3):
        s = strides[li]
        n_h, n_w = img_hw[0] // s, img_hw[1] // s
        bb = torch.randn(len(counts), 9 * 4, n_h, n_w, generator=gen) * 0.3
        cc = torch.randn(len(counts), 9 * 1, n_h, n_w, generator=gen) * 3.0
        raw = {'bbox': bb.view(len(counts), 9, 4, n_h, n_w).permute(0, 1, 3, 4, 2),
               'class': cc.view(len(counts), 9, 1, n_h, n_w).permute(0, 1, 3, 4, 2)}
        layer = RetinaLayer(li, rcfg)
        (_, loss), _ = _grab_forward(layer, ['M_pos'], raw, img_hw, labels1)
        out[f'retina{li}_bbox_in'], out[f'retina{li}_cls_in'] = bb, cc
        out[f'retina{li}_loss'] = loss
        out[f'retina{li}_loss_str'] = np.array(layer.loss_str)
    # ---- RAPiD (rotated boxes, no classes): the stub pycocotools (exact polygon clipper) drives iou_rle, so this pins
    # the CONTROL FLOW of the training branch; rotated IoU values stay "parity unpinned"
    from models.detlayers.rapid import RAPiDLayer
    pcfg = {'model.rapid.anchors': RAPID_ANCHORS, 'model.rapid.anchor_indices': IDX3, 'model.fpn.out_strides': [8, 16, 32],
            'general.num_class': 0, 'model.rapid.wh_smooth_l1_beta': 1, 'model.angle.loss_angle': 'Periodic_L1',
            'model.angle.pred_range': 360}
    rlabels = []
    for b, n in enumerate(counts):
        bx = torch.zeros(n, 5)
        if n:
            bx[:, :4] = labels[b].bboxes
            bx[:, 2:4] = torch.exp(torch.rand(n, 2, generator=gen) * (np.log(260.0) - np.log(15.0)) + np.log(15.0))
            bx[:, 4] = torch.rand(n, generator=gen) * 180 - 90
        rlabels.append(ImageObjects(bx, torch.zeros(n, dtype=torch.int64), bb_format='cxcywhd', img_hw=img_hw))
        out[f'rgt{b}_boxes'] = bx
    for li, s in enumerate(pcfg['model.fpn.out_strides']):
        store, raw = head_views(gen, len(counts), 3, img_hw[0] // s, img_hw[1] // s, 5, 0, conf_mu=-7.5)
        layer = RAPiDLayer(li, pcfg)
        layer.ignore_thre = 0.1        # instance attribute (0.6 in __init__): random predictions rarely reach 0.6
        (_, loss), g = _grab_forward(layer, ['PositiveMask', 'IgnoredMask', 'TargetXYWH', 'TargetAngle', 'TargetConf'], raw, img_hw, rlabels)
        out[f'rapid{li}_in'] = store['nchw']
        out[f'rapid{li}_loss'] = loss
        out.update({f'rapid{li}_{k}': v for k, v in g.items()})
        print('rapid', li, layer.loss_str)
    save('train', **out)


def gen_train_uv5():
    """Training branch of the Ultralytics / YOLOv5 layer (models/detlayers/uv5.py:92-224) of the unmodified reference:
    both confidence targets ('zero-one' with the ignore mask, 'IoU'), three levels, three images (one without GT)."""
    from models.detlayers.uv5 import DetectLayer
    from utils.structures import ImageObjects
    gen = torch.Generator().manual_seed(1011)
    img_hw, n_cls, counts = (256, 320), 5, [9, 0, 4]
    out, labels = {}, []
    for b, n in enumerate(counts):
        bx, ct = _random_gt(gen, n, img_hw, n_cls, 10.0, 330.0) if n else (torch.zeros(0, 4), torch.zeros(0, dtype=torch.int64))
        if b == 0:
            bx[-1, 2:4] = torch.tensor([300.0, 270.0])
            bx[-2, 2:4] = torch.tensor([150.0, 180.0])
            bx[1, 0:2] = bx[0, 0:2] + 1.0                          # two GTs in one cell: both contribute to the loss
            bx[1, 2:4] = bx[0, 2:4] * 1.05
        labels.append(ImageObjects(bx, ct, bb_format='cxcywh', img_hw=img_hw))
        out[f'gt{b}_boxes'], out[f'gt{b}_cats'] = bx, ct
    for mode in ('zero-one', 'IoU'):
        cfg = {'model.detect.anchors': YOLO_ANCHORS, 'model.detect.anchor_indices': IDX3, 'model.fpn.out_strides': [8, 16, 32],
               'general.num_class': n_cls, 'model.detect.sample_selection': 'best', 'model.detect.confidence_target': mode,
               'model.detect.negative_threshold': 0.3, 'model.detect.loss_bbox': 'smooth_L1', 'general.pred_bbox_format': 'cxcywh'}
        tag = 'zo' if mode == 'zero-one' else 'iou'
        for li, s in enumerate(cfg['model.fpn.out_strides']):
            store, raw = head_views(gen, len(counts), 3, img_hw[0] // s, img_hw[1] // s, 4, n_cls, conf_mu=-1.0)
            layer = DetectLayer(li, cfg)
            (_, loss), g = _grab_forward(layer, ['TargetConf', 'IgnoredMask'], raw, img_hw, labels)
            out[f'{tag}{li}_in'] = store['nchw']
            out[f'{tag}{li}_loss'] = loss
            out[f'{tag}{li}_assigned'] = torch.tensor(int(layer._assigned_num))
            out[f'{tag}{li}_loss_str'] = np.array(layer.loss_str)
            out.update({f'{tag}{li}_{k}': v for k, v in g.items()})
            print('uv5', mode, li, layer.loss_str)
    save('train_uv5', **out)


def gen_train_retina_rot():
    """RetinaLayer.forward(raw, img_size, labels) with ROTATED boxes (retinanet.py:84-160, angle part :133-136, :151-154).
    At HEAD the reference cannot construct this variant: __init__ does `from .losses import get_angle_loss` (:36) inside
    the detlayers package (the module is models/losses.py) and reads 'model.angle.loss_name', a key no config has.  The
    layer is therefore built with the 'cxcywh' config and the three attributes __init__ would have set are assigned by
    hand (pred_bbox_format, n_bbparam, loss_angle = models.losses.get_angle_loss(name, 'sum')); forward() itself runs
    unmodified."""
    from models.detlayers.retinanet import RetinaLayer
    from models.losses import get_angle_loss
    from utils.structures import ImageObjects
    gen = torch.Generator().manual_seed(1013)
    img_hw, counts, strides = (256, 320), [7, 0, 5], [8, 16, 32, 64, 128]
    rcfg = {'model.fpn.out_strides': strides, 'model.retina.anchor.base': 4, 'model.retina.anchor.scales': [1, 1.26, 1.5874],
            'model.retina.anchor.ratios': [[1, 1], [1.4, 0.7], [0.7, 1.4]], 'model.retina.anchor.positive_threshold': 0.5,
            'model.retina.anchor.negative_threshold': 0.4, 'general.num_class': 1, 'general.pred_bbox_format': 'cxcywh',
            'general.bbox_param': 4}
    out, labels = {}, []
    for b, n in enumerate(counts):
        bx = torch.zeros(n, 5)
        if n:
            bx[:, :4], _ = _random_gt(gen, n, img_hw, 1, 20.0, 200.0)
            bx[:, 4] = torch.rand(n, generator=gen) * 180 - 90
        labels.append(ImageObjects(bx, torch.zeros(n, dtype=torch.int64), bb_format='cxcywhd', img_hw=img_hw))
        out[f'gt{b}_boxes'] = bx
    for name in ('Periodic_L1', 'Periodic_smoothL1'):
        for li in (1, 2, 3):
            s = strides[li]
            n_h, n_w = img_hw[0] // s, img_hw[1] // s
            bb = torch.randn(len(counts), 9 * 5, n_h, n_w, generator=gen) * 0.3
            cc = torch.randn(len(counts), 9 * 1, n_h, n_w, generator=gen) * 3.0
            raw = {'bbox': bb.view(len(counts), 9, 5, n_h, n_w).permute(0, 1, 3, 4, 2),
                   'class': cc.view(len(counts), 9, 1, n_h, n_w).permute(0, 1, 3, 4, 2)}
            layer = RetinaLayer(li, rcfg)
            layer.pred_bbox_format, layer.n_bbparam, layer.loss_angle = 'cxcywhd', 5, get_angle_loss(name, reduction='sum')
            (_, loss), _ = _grab_forward(layer, ['M_pos'], raw, img_hw, labels)
            out[f'{name}{li}_bbox_in'], out[f'{name}{li}_cls_in'] = bb, cc
            out[f'{name}{li}_loss'] = loss
            out[f'{name}{li}_loss_str'] = np.array(layer.loss_str)
            print('retina rot', name, li, layer.loss_str)
    save('train_retina_rot', **out)


def gen_preprocess():
    """Detector._preprocess_pil + tvf.to_tensor + format_tensor_img (api/detection.py:158-162, :177-205) of the unmodified
    reference on small seeded uint8 images: every pre-processing name x every input format, up- and down-scaling."""
    import PIL.Image
    import torchvision.transforms.functional as tvf
    from api.detection import Detector
    import utils.image_ops as imgUtils
    rng = np.random.default_rng(77)
    det = Detector.__new__(Detector)          # no model needed for the pre-processing methods
    out, cases = {}, []
    specs = [((67, 101), 'pad_divisible', None, 32, 'RGB_1'),
             ((96, 64), 'pad_divisible', None, 32, 'BGR_255_norm'),          # already divisible: no padding at all
             ((150, 201), 'resize_pad_divisible', 96, 32, 'RGB_1_norm'),
             ((201, 150), 'resize_pad_divisible', 128, 128, 'BGR_255_norm'),
             ((45, 80), 'resize_pad_divisible', 160, 32, 'RGB_1'),           # up-scaling
             ((240, 135), 'resize_pad_square', 96, 32, 'RGB_1_norm'),
             ((77, 77), 'resize_pad_square', 64, 32, 'RGB_1'),
             ((50, 121), 'resize_pad_square', 192, 32, 'BGR_255_norm'),      # up-scaling + centred padding
             ((64, 64), 'resize_pad_square', 64, 32, 'RGB_1_norm')]          # identity resize
    for i, (hw, name, size, div, code) in enumerate(specs):
        img = rng.integers(0, 256, hw + (3,), dtype=np.uint8)
        if i % 2:                                                            # smooth content on every other case
            yy, xx = np.mgrid[0:hw[0], 0:hw[1]]
            img = np.stack([(yy * 3 + xx) % 256, (xx * 5) % 256, (yy * 2 + xx * 7) % 256], -1).astype(np.uint8)
        det.divisibe = div
        pil, pad_info = det._preprocess_pil(PIL.Image.fromarray(img), name, size)
        t = imgUtils.format_tensor_img(tvf.to_tensor(pil), code=code)
        out[f'pre{i}_img'] = img
        out[f'pre{i}_out'] = t
        out[f'pre{i}_pad'] = np.array(pad_info if pad_info is not None else [-1] * 6, dtype=np.int64)
        cases.append(f'{name}|{size}|{div}|{code}')
    out['cases'] = np.array(cases)
    save('preprocess', **out)


def gen_tracking():
    """the tracklet state machine of utils/structures.py:447-529 (KFTracklet) on top of
    utils/kalman_filter.py:77-142 (RotBBoxKalmanFilter), run UNMODIFIED for 12 tracklets over 8 frames: predict every
    frame, update when the tracklet has a measurement, likelihood of 6 candidate boxes under every tracklet.
    The reference writes `np.bool`, which NumPy >= 1.24 no longer has: the alias is restored for the run (the only
    accommodation; no reference file is touched)."""
    import_reference()
    if not hasattr(np, 'bool'):
        np.bool = bool
    from utils.structures import KFTracklet
    rng = np.random.RandomState(77)
    n_t, n_f, n_c = 12, 8, 6
    init = np.concatenate([rng.uniform(100, 900, (n_t, 2)), rng.uniform(20, 120, (n_t, 2)), rng.uniform(-200, 400, (n_t, 1))], axis=1)
    init_score = rng.uniform(0.2, 1.0, n_t)
    tracks = [KFTracklet(init[i].copy(), float(init_score[i]), object_id=i, img_hw=(1024, 1024)) for i in range(n_t)]
    meas = np.zeros((n_f, n_t, 5)); meas_score = np.zeros((n_f, n_t)); has = np.zeros((n_f, n_t), dtype=bool)
    cand = np.zeros((n_f, n_c, 5))
    out = {k: [] for k in ('pred', 'upd', 'x', 'P', 'score', 'feasible', 'lik')}
    truth = init.copy()
    for f in range(n_f):
        pred = np.stack([t.predict() for t in tracks])
        truth[:, :2] += rng.randn(n_t, 2) * 3
        truth[:, 4] += rng.randn(n_t) * 4
        has[f] = rng.rand(n_t) < 0.7
        meas[f] = truth + np.concatenate([rng.randn(n_t, 4) * 1.5, rng.randn(n_t, 1) * 3 + 180 * rng.randint(-1, 2, (n_t, 1))], axis=1)
        meas_score[f] = rng.uniform(0.1, 1.0, n_t)
        cand[f] = np.concatenate([truth[:n_c, :4] + rng.randn(n_c, 4) * 2, (truth[:n_c, 4:] % 180) + rng.randn(n_c, 1) * 2], axis=1)
        lik = np.stack([t.likelihood(cand[f]) for t in tracks])
        upd = np.zeros((n_t, 5))
        for i, t in enumerate(tracks):
            if has[f, i]:
                upd[i] = t.update(meas[f, i].copy(), float(meas_score[f, i]))
        out['pred'].append(pred); out['upd'].append(upd); out['lik'].append(lik)
        out['x'].append(np.stack([t.kf.x for t in tracks])); out['P'].append(np.stack([t.kf.P for t in tracks]))
        out['score'].append(np.array([t.score for t in tracks])); out['feasible'].append(np.array([t.is_feasible() for t in tracks]))
    save('tracking', init=init, init_score=init_score, meas=meas, meas_score=meas_score, has=has, cand=cand,
         **{k: np.stack(v) for k, v in out.items()})


def _build_reference_model(cfg):
    """OneStageBBox(cfg) of the unmodified reference with RANDOM weights: the two places that fetch pre-trained weights
    (models/registry.py:15 torch.load of weights/dark53_imgnet.pth, absent; external/efficientnet's download) are answered
    with the freshly initialised parameters instead."""
    from models.general import OneStageBBox
    from models.backbones import Darknet53
    import external.efficientnet.model as efn_model
    real_load, real_pre = torch.load, efn_model.load_pretrained_weights
    torch.load = lambda path, *a, **k: Darknet53(cfg).state_dict() if str(path).endswith('dark53_imgnet.pth') else real_load(path, *a, **k)
    efn_model.load_pretrained_weights = lambda *a, **k: None
    try:
        return OneStageBBox(cfg).eval()
    finally:
        torch.load, efn_model.load_pretrained_weights = real_load, real_pre


def gen_fullmodel(name='yolov3_80'):
    """A BASELINE config end to end through the UNMODIFIED reference: OneStageBBox(configs/<name>.json) with random
    weights (backbone + FPN + head), one synthetic image, forward -> det layers -> level concatenation
    (models/general.py:67-84) -> ImageObjects.post_process (api/detection.py:172).
    A random-init network in eval mode emits logits of ~1e-3 (exact score ties), so every head output tensor
    is standardised PER CHANNEL to zero mean and a standard deviation of 1.5 over the positions -- the same as rescaling
    the weights and shifting the bias of the head's last convolution, channel by channel; the weights stay random, the ranking becomes tie-free.  To keep the fixture small
    the image is 256 x 256 and the head outputs are rounded to float16-representable values BEFORE they enter the
    reference's det layers: 2 bytes per logit are stored and both sides see identical float32 inputs (the 18-channel RAPiD
    head is kept in float32: its score is sigmoid(conf) alone, and float16 logits would tie).  The image seed is the
    first one whose ranking keeps comfortable margins for a float32 GPU path."""
    import json
    from utils.structures import ImageObjects
    from utils.bbox_ops import bboxes_iou
    cfg = json.load(open(os.path.join(REF, 'configs', name + '.json')))
    torch.manual_seed(2024)
    model = _build_reference_model(cfg)
    img_hw = (256, 256)
    conf_thres, nms_thres = cfg['test.ap_conf_thres'], cfg['test.nms_thres']
    n_cls = cfg['general.num_class']

    def nchw_of(raw):
        """The NCHW tensors behind the permuted views of models/rpns.py:29-41 (YOLOHead) / :175-189 (EfDetHead)."""
        if raw['bbox'].dim() == 5:                                   # (B,nA,nH,nW,K) views of one (B, nA*(P+1+C), nH, nW) tensor
            t = torch.cat([raw['bbox'], raw['conf'], raw['class']], dim=-1).permute(0, 1, 4, 2, 3)
            return [t.reshape(t.shape[0], -1, t.shape[3], t.shape[4]).contiguous()]
        return [raw['bbox'].permute(0, 3, 1, 2).contiguous(),       # (B,4,nH,nW) and (B,1+C,nH,nW)
                torch.cat([raw['conf'], raw['class']], dim=-1).permute(0, 3, 1, 2).contiguous()]

    def views_of(tensors, n_param):
        if len(tensors) == 1:
            t = tensors[0]
            v = t.view(t.shape[0], 3, n_param + 1 + n_cls, t.shape[2], t.shape[3])
            return {'bbox': v[:, :, 0:n_param].permute(0, 1, 3, 4, 2), 'conf': v[:, :, n_param:n_param + 1].permute(0, 1, 3, 4, 2),
                    'class': v[:, :, n_param + 1:].permute(0, 1, 3, 4, 2)}
        bb, cc = tensors
        c = cc.permute(0, 2, 3, 1)
        return {'bbox': bb.permute(0, 2, 3, 1), 'conf': c[..., 0:1], 'class': c[..., 1:]}

    for seed in range(11, 80):
        gen = torch.Generator().manual_seed(seed)
        x = torch.rand(1, 3, *img_hw, generator=gen)
        raws = model.rpn(model.fpn(model.backbone(x)))
        out, dts_all = {}, []
        for i, raw in enumerate(raws):
            stored = []
            for j, t in enumerate(nchw_of(raw)):
                t = (t - t.mean(dim=(0, 2, 3), keepdim=True)) / t.std(dim=(0, 2, 3), keepdim=True) * 1.5
                t = t.half() if n_cls > 0 else t      # class-less RAPiD: score = sigmoid(conf) alone, float16 logits would tie
                out[f'head{i}_{j}'] = t
                stored.append(t.float())
            dts_all.append(model.det_layers[i](views_of(stored, raw['bbox'].shape[-1]), img_hw, None)[0])
        bbs = torch.cat([d['bbox'] for d in dts_all], dim=1)[0]         # models/general.py:74-76
        cls = torch.cat([d['class_idx'] for d in dts_all], dim=1)[0]
        sc = torch.cat([d['score'] for d in dts_all], dim=1)[0]
        srt = sc.sort(descending=True).values
        top = sc.argsort(descending=True)[:512]
        iou = bboxes_iou(bbs[top][:, :4], bbs[top][:, :4])
        same = cls[top][:, None] == cls[top][None, :]
        margins = {'seed': seed, 'score_512_513': float(srt[511] - srt[512]), 'score_to_conf_thres': float((sc - conf_thres).abs().min()),
                   'iou_to_nms_thres': float((iou[same] - nms_thres).abs().min()), 'distinct_scores': int(sc.unique().numel())}
        if os.environ.get('MYDET_GOLDEN_VERBOSE'):
            print(margins)
        if margins['score_512_513'] > 5e-5 and margins['score_to_conf_thres'] > 5e-5 and margins['iou_to_nms_thres'] > 1e-4 \
                and srt[:513].unique().numel() == 513:
            break
    else:
        raise RuntimeError('no seed with safe margins')
    res = ImageObjects(bboxes=bbs.clone(), cats=cls.clone(), scores=sc.clone(), bb_format=model.bb_format,
                       img_hw=img_hw).post_process(conf_thres, nms_thres)
    keys = torch.cat([bbs, sc[:, None], cls[:, None].float()], dim=1)
    got = torch.cat([res.bboxes, res.scores[:, None], res.cats[:, None].float()], dim=1)
    keep = torch.tensor([int(torch.nonzero((keys == g).all(dim=1))[0, 0]) for g in got], dtype=torch.int64)
    assert keep.unique().numel() == keep.numel()
    print('fullmodel', name, len(res), 'kept of', int((sc >= conf_thres).sum()), 'candidates;', margins)
    out.update({'keep': keep, 'kept_boxes': res.bboxes, 'kept_scores': res.scores, 'kept_cats': res.cats,
                'params': np.array([conf_thres, nms_thres, img_hw[0], img_hw[1]], dtype=np.float64)})
    save('fullmodel_' + name, **out)


def gen_fullmodel_atss():
    """BASELINE configs[3] through the UNMODIFIED reference in TRAINING mode: OneStageBBox(configs/d1_fcs2_atss.json) with
    random weights (see gen_fullmodel), one synthetic 384 x 384 image (the coarsest level must hold >= 9 anchors for
    torch.topk, fcos2.py:397) with 100 GT boxes, model(x, labels): the ATSS target
    tensors of all five FCOS_ATSS_Layer levels are captured from forward()'s locals.  Only the regression head output enters
    the assignment (the ignore mask uses the predicted boxes); it is standardised per channel to sigma 0.5 and stored as
    float16-representable values, as in gen_fullmodel."""
    import json
    from utils.structures import ImageObjects
    cfg = json.load(open(os.path.join(REF, 'configs', 'd1_fcs2_atss.json')))
    torch.manual_seed(2025)
    model = _build_reference_model(cfg)
    gen = torch.Generator().manual_seed(77)
    img_hw = (384, 384)
    n_gt, n_cls = 100, cfg['general.num_class']
    bx = torch.empty(n_gt, 4)
    bx[:, 0:2] = torch.rand(n_gt, 2, generator=gen) * 364 + 10
    bx[:, 2:4] = torch.exp(torch.rand(n_gt, 2, generator=gen) * (np.log(140.0) - np.log(8.0)) + np.log(8.0))
    ct = torch.randint(0, n_cls, (n_gt,), generator=gen)
    assert (bx[:, 2] * bx[:, 3]).unique().numel() == n_gt
    labels = [ImageObjects(bx, ct, bb_format='cxcywh', img_hw=img_hw)]
    x = torch.rand(1, 3, *img_hw, generator=gen)
    raws = model.rpn(model.fpn(model.backbone(x)))
    out = {'gt_boxes': bx, 'gt_cats': ct, 'params': np.array([img_hw[0], img_hw[1], cfg['model.atss.topk_per_level'],
                                                              cfg.get('model.fcos2.ignored_threshold', 0.7), n_cls], dtype=np.float64),
           'strides': np.array(cfg['model.fpn.out_strides']), 'anchors': np.array(cfg['model.atss.anchors'], dtype=np.float64)}
    names = ['PositiveMask', 'IgnoredMask', 'TargetConf', 'TargetLTRB', 'TargetCls']
    n_pos = n_ign = 0
    for i, raw in enumerate(raws):
        bb = raw['bbox'].permute(0, 3, 1, 2).contiguous()
        bb = ((bb - bb.mean(dim=(0, 2, 3), keepdim=True)) / bb.std(dim=(0, 2, 3), keepdim=True).clamp(min=1e-12) * 0.5).half()
        out[f'atss{i}_bbox'] = bb
        raw_in = {'bbox': bb.float().permute(0, 2, 3, 1), 'conf': raw['conf'], 'class': raw['class']}
        _, g = _grab_forward(model.det_layers[i], names, raw_in, img_hw, labels)
        out.update({f'atss{i}_{k}': g[k] for k in names})
        n_pos += int(g['PositiveMask'].sum()); n_ign += int(g['IgnoredMask'].sum())
    print('fullmodel atss: positives', n_pos, 'ignored', n_ign)
    assert n_pos > 50
    save('fullmodel_d1_fcs2_atss', **out)


if __name__ == '__main__':
    import_reference()
    torch.set_grad_enabled(False)
    gen_decode()
    gen_postprocess()
    gen_iou()
    gen_atss()
    gen_train()
    gen_preprocess()
    gen_tracking()
    gen_train_uv5()
    gen_train_retina_rot()
    for cfg_name in ('yolov3_80', 'rapid', 'd1_fcs2'):
        gen_fullmodel(cfg_name)
    gen_fullmodel_atss()
