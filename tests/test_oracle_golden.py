"""CPU: the oracle against fixtures produced by the real reference (tests/golden/make_golden.py)
and against the installed torchvision.  Bit-exact unless stated."""
import numpy as np
import pytest
import torch

from oracle import decode as od, postprocess as opp, iou as oi, atss as oa
from helpers import T, level_anchors, yolo_views, efdet_views, anchor_views, YOLO_ANCHORS, RAPID_ANCHORS


def same(got, g, key):
    b, c, s = got
    assert torch.equal(b, T(g[key + '_bbox'])), key
    assert torch.equal(c, T(g[key + '_cls'])), key
    assert torch.equal(s, T(g[key + '_score'])), key


def test_decode_yolo(golden):
    g = golden('decode')
    for li, s in enumerate((8, 16, 32)):
        raw = yolo_views(T(g[f'yolo{li}_in']), 3, 4, 5)
        same(od.decode_yolo(raw, level_anchors(YOLO_ANCHORS, li), s, 5), g, f'yolo{li}')
    raw = yolo_views(T(g['yolo_c0_in']), 3, 4, 0)
    same(od.decode_yolo(raw, level_anchors(YOLO_ANCHORS, 1), 16, 0), g, 'yolo_c0')


def test_decode_rapid(golden):
    g = golden('decode')
    for tag, nc in (('rapid_c0', 0), ('rapid_c3', 3)):
        for li, s in enumerate((8, 16, 32)):
            raw = yolo_views(T(g[f'{tag}_{li}_in']), 3, 5, nc)
            same(od.decode_rapid(raw, level_anchors(RAPID_ANCHORS, li), s, nc), g, f'{tag}_{li}')


def test_decode_fcos(golden):
    g = golden('decode')
    for li, s in enumerate((8, 16, 32, 64, 128)):
        raw = efdet_views(T(g[f'fcos{li}_bbox_in']), T(g[f'fcos{li}_cls_in']))
        same(od.decode_fcos(raw, s, (256, 384)), g, f'fcos{li}')


def test_decode_retina_uv5(golden):
    g = golden('decode')
    for tag, rot in (('retina', False), ('retina_rot', True)):
        raw = anchor_views(T(g[f'{tag}_bbox_in']), T(g[f'{tag}_cls_in']), 9)
        same(od.decode_retina(raw, T(g[f'{tag}_anchors']), 16, (96, 128), with_angle=rot), g, tag)
    raw = yolo_views(T(g['uv5_in']), 3, 4, 5)
    same(od.decode_uv5(raw, level_anchors(YOLO_ANCHORS, 0), 8), g, 'uv5')


@pytest.mark.parametrize('tag,fmt,cap', [('pp_small', 'cxcywh', 512), ('pp_cap', 'cxcywh', 512),
                                         ('pp_rot', 'cxcywhd', 512), ('pp_empty', 'cxcywh', 512)])
def test_post_process(golden, tag, fmt, cap):
    g = golden('postprocess')
    boxes, scores, cats = T(g[tag + '_boxes']), T(g[tag + '_scores']), T(g[tag + '_cats'])
    conf, nms = g[tag + '_params']
    assert opp.top_boundary_is_tie_free(scores, conf, cap)
    keep = opp.post_process(boxes, cats, scores, float(conf), float(nms), fmt, cap)
    assert torch.equal(keep, T(g[tag + '_keep']))


def test_nms_direct_and_adversarial(golden):
    g = golden('postprocess')
    boxes, scores, cats = T(g['nms_direct_boxes']), T(g['nms_direct_scores']), T(g['nms_direct_cats'])
    keep = opp.class_nms(boxes, scores, cats, float(g['nms_direct_params'][1]))
    assert torch.equal(keep, T(g['nms_direct_keep']))
    keep = opp.class_nms(T(g['adv_boxes']), T(g['adv_scores']), torch.zeros(8, dtype=torch.int64), float(g['adv_thr'][0]))
    assert torch.equal(T(g['adv_scores'])[keep], T(g['adv_keep_scores']))


def test_nms_restatement_matches_torchvision():
    tv = pytest.importorskip('torchvision')
    gen = torch.Generator().manual_seed(7)
    for n, thr in ((1, 0.5), (64, 0.5), (777, 0.3), (2000, 0.7)):
        xy = torch.rand(n, 2, generator=gen) * 200
        wh = torch.rand(n, 2, generator=gen) * 60
        box = torch.cat([xy, xy + wh], 1)
        sc = (torch.rand(n, generator=gen) * 50).round() / 50      # many exact score ties
        assert torch.equal(opp.nms_aabb(box, sc, thr), tv.ops.nms(box, sc, thr))
    # threshold semantics: float IoU against a DOUBLE threshold, strict '>'
    box = torch.tensor([[0., 0, 10, 10], [0, 0, 10, 4.5]])
    sc = torch.tensor([1.0, 0.5])
    for thr in (0.45, float(np.float32(0.45)), 0.4499999):
        assert torch.equal(opp.nms_aabb(box, sc, thr), tv.ops.nms(box, sc, thr))


def test_bboxes_iou(golden):
    g = golden('iou')
    a, b = T(g['a']), T(g['b'])
    assert torch.equal(oi.bboxes_iou(a, b), T(g['iou_cxcywh']))
    assert torch.equal(oi.cxcywh_to_x1y1x2y2(a), T(g['a_xyxy']))
    assert torch.equal(oi.bboxes_iou(oi.cxcywh_to_x1y1x2y2(a), oi.cxcywh_to_x1y1x2y2(b), xyxy=True), T(g['iou_xyxy']))
    gt = T(g['gt_debug3'])
    assert torch.equal(oi.bboxes_iou(gt, gt), T(g['iou_gt_self']))
    with pytest.raises(IndexError):
        oi.bboxes_iou(torch.zeros(3, 5), b)


def test_rotated(golden):
    g = golden('iou')
    rb, rs = T(g['rot_boxes']), T(g['rot_scores'])
    rad = rb.clone()
    rad[:, 4] = oi.deg2rad_f32(rad[:, 4])
    assert torch.equal(oi.xywha2vertex(rad), T(g['rot_vertices']))
    # control flow of nms_rotbb, pinned against the reference driven by an independent polygon IoU
    assert torch.equal(oi.nms_rot(rb, rs, 0.45), T(g['rot_keep_045']))
    assert torch.equal(oi.nms_rot(rb, rs, 0.2), T(g['rot_keep_02']))
    assert torch.equal(oi.nms_rot(rb, rs, 0.3, majority=2), T(g['rot_keep_maj2']))
    rgt = T(g['rot_gt_debug3'])[:40]
    np.testing.assert_allclose(oi.iou_rot(rgt, rgt).numpy(), g['rot_gt_iou_stub'], rtol=0, atol=2e-6)
    # geometric known answers
    sq = torch.tensor([[50., 50, 20, 20, 0]])
    assert abs(float(oi.iou_rot(sq, torch.tensor([[50., 50, 20, 20, 90]]))) - 1.0) < 1e-6
    assert abs(float(oi.iou_rot(sq, torch.tensor([[60., 50, 20, 20, 0]]))) - 1.0 / 3.0) < 1e-6
    octo = float(oi.iou_rot(sq, torch.tensor([[50., 50, 20, 20, 45]])))       # square vs 45-deg square
    inter = 400 * 2 * (2 ** 0.5 - 1)
    assert abs(octo - inter / (800 - inter)) < 1e-6
    assert float(oi.iou_rot(sq, torch.tensor([[500., 50, 20, 20, 10]]))) == 0.0


def test_atss(golden):
    g = golden('atss')
    gts = [(T(g[f'gt{b}_boxes']), T(g[f'gt{b}_cats'])) for b in range(2)]
    strides, sides = [8, 16, 32, 64, 128], [24, 48, 96, 192, 384]
    for li in range(5):
        t = T(g[f'atss{li}_bbox_in']).permute(0, 2, 3, 1)
        out = oa.assign_level(li, t, gts, (384, 512), strides, sides, 9, 0.7, 6)
        for k, v in out.items():
            assert torch.equal(v, T(g[f'atss{li}_{k}'])), (li, k)
    assert sum(int(g[f'atss{li}_PositiveMask'].sum()) for li in range(5)) > 0


def test_training_targets(golden):
    """oracle/train.py against tensors captured from the unmodified reference's YOLOLayer / FCOSLayer
    forward(raw, img_size, labels) (tests/golden/train.npz): bit-exact."""
    from oracle import decode as od, train as ot
    from helpers import yolo_views, YOLO_ANCHORS
    g = golden('train')
    gts = [(T(g[f'gt{b}_boxes']), T(g[f'gt{b}_cats'])) for b in range(3)]
    idx3 = [[0, 1, 2], [3, 4, 5], [6, 7, 8]]
    for li, s in enumerate((8, 16, 32)):
        raw = yolo_views(T(g[f'yolo{li}_in']), 3, 4, 5)
        anchors = torch.tensor(YOLO_ANCHORS, dtype=torch.float32)[idx3[li]]
        box, _, _ = od.decode_yolo(raw, anchors, s, 5)
        n_h, n_w = raw['bbox'].shape[2:4]
        tg = ot.yolo_targets(box, gts, (256, 320), s, YOLO_ANCHORS, idx3[li], 0.3, 5, (n_h, n_w))
        for k in ('gt_mask', 'conf_loss_mask', 'tgt_xywh', 'tgt_cls', 'weighted'):
            assert torch.equal(tg[k], T(g[f'yolo{li}_{k}'])), (li, k)
        assert tg['valid_gt_num'] == int(g[f'yolo{li}_assigned'])
    fa = [0, 64, 128, 256, 512, 100000000]
    for li, s in zip((0, 1, 2), (8, 16, 32)):
        t = T(g[f'fcos{li}_bbox_in']).permute(0, 2, 3, 1)
        tg = ot.fcos2_targets(t, gts, (256, 320), s, fa[li], fa[li + 1], 0.2, 5)
        for k in ('PositiveMask', 'IgnoredMask', 'TargetConf', 'TargetLTRB', 'TargetCls'):
            assert torch.equal(tg[k], T(g[f'fcos{li}_{k}'])), (li, k)

    # RetinaNet (single class): loss and positive count of the reference run (its per-image targets are loop locals)
    scales, ratios = [1, 1.26, 1.5874], [[1, 1], [1.4, 0.7], [0.7, 1.4]]
    gts1 = [(b, torch.zeros_like(c)) for b, c in gts]
    for li, s in zip((1, 2, 3), (16, 32, 64)):
        wh = torch.Tensor([(4 * s * sc * rt[0], 4 * s * sc * rt[1]) for sc in scales for rt in ratios])
        bb, cc = T(g[f'retina{li}_bbox_in']), T(g[f'retina{li}_cls_in'])
        n_b, _, n_h, n_w = bb.shape
        t = bb.view(n_b, 9, 4, n_h, n_w).permute(0, 1, 3, 4, 2)
        c = cc.view(n_b, 9, 1, n_h, n_w).permute(0, 1, 3, 4, 2)
        _, loss, pos = ot.retina_targets_and_loss(t, c, gts1, (256, 320), s, wh, 0.5, 0.4)
        assert float(loss) == float(g[f'retina{li}_loss']), li
        assert f' pos {pos}/' in str(g[f'retina{li}_loss_str'])


def test_rapid_training_targets(golden):
    """oracle/train.py: rapid_targets against the RAPiDLayer tensors captured from the reference driven by the exact
    polygon-clipping pycocotools stub (control flow; rotated IoU values stay "parity unpinned")."""
    from oracle import decode as od, train as ot
    from helpers import yolo_views, RAPID_ANCHORS
    g = golden('train')
    gts = [(T(g[f'rgt{b}_boxes']), torch.zeros(len(g[f'rgt{b}_boxes']), dtype=torch.int64)) for b in range(3)]
    idx3 = [[0, 1, 2], [3, 4, 5], [6, 7, 8]]
    for li, s in enumerate((8, 16, 32)):
        raw = yolo_views(T(g[f'rapid{li}_in']), 3, 5, 0)
        anchors = torch.tensor(RAPID_ANCHORS, dtype=torch.float32)[idx3[li]]
        box, _, _ = od.decode_rapid(raw, anchors, s, 0)
        n_h, n_w = raw['bbox'].shape[2:4]
        tg = ot.rapid_targets(box, raw['conf'], gts, (256, 320), s, RAPID_ANCHORS, idx3[li], 0, (n_h, n_w), ignore_thre=0.1)
        for k in ('PositiveMask', 'IgnoredMask', 'TargetXYWH', 'TargetAngle', 'TargetConf'):
            assert torch.equal(tg[k], T(g[f'rapid{li}_{k}'])), (li, k)


def test_atss_threshold_tie_policy_variant():
    """atss_threshold_index_ties (the kernels' declared tie policy) equals the reference-faithful atss_threshold
    wherever the k-th nearest anchor is not tied, i.e. for GT centres off the cell boundaries."""
    from oracle import atss as oa
    strides, sides, img = [8, 16, 32, 64, 128], [24, 48, 96, 192, 384], (384, 640)
    anchors = oa.all_level_anchors(img, strides, sides)
    gen = torch.Generator().manual_seed(3)
    for _ in range(40):
        c = torch.rand(2, generator=gen) * torch.tensor([640.0, 384.0])
        wh = torch.rand(2, generator=gen) * 200 + 10
        gt = torch.cat([c, wh])
        a, b = float(oa.atss_threshold(gt, anchors, 9)), float(oa.atss_threshold_index_ties(gt, anchors, 9))
        assert abs(a - b) <= 1e-6 * max(1.0, abs(a))     # same anchors, the summation order of mean / std may differ


def test_raster_vs_exact_gap():
    """oracle/raster.c restates pycocotools' polygon rasterisation + RLE IoU (PARITY UNPINNED: the library is not in this
    image).  Known answers on integer axis-aligned rectangles must be exact; on random rotated boxes the raster IoU on
    the reference's 2048 x 2048 canvas must stay within the O(perimeter / area) band of the exact polygon IoU that the
    kernels compute.  The measured gap is what DESIGN.md section 3 reports -- nothing else is compared against the raster."""
    from oracle import iou as oi
    a = torch.tensor([[100.0, 100.0, 40.0, 20.0, 0.0]])
    b = torch.tensor([[110.0, 100.0, 40.0, 20.0, 0.0], [100.0, 100.0, 40.0, 20.0, 0.0], [300.0, 300.0, 10.0, 10.0, 0.0],
                      [100.0, 100.0, 20.0, 40.0, 90.0]])
    assert oi.iou_rle_raster(a, b).tolist() == [[0.6, 1.0, 0.0, 1.0]]
    gen = torch.Generator().manual_seed(1)
    for lo, hi, bound in ((6, 20, 0.08), (20, 60, 0.02), (60, 250, 0.005)):
        n = 200
        bx = torch.cat([torch.rand(n, 2, generator=gen) * (hi * 4) + 200, torch.rand(n, 2, generator=gen) * (hi - lo) + lo,
                        torch.rand(n, 1, generator=gen) * 180 - 90], 1)
        r, e = oi.iou_rle_raster(bx[:100], bx[100:]), oi.iou_rot(bx[:100], bx[100:])
        gap = (r - e).abs()
        assert float(gap.max()) < bound, (lo, hi, float(gap.max()))
        assert bool(((r > 0) == (e > 1e-3))[e > 0.02].all())          # same overlap structure


def test_preprocess_against_reference(golden):
    """oracle/preprocess.py against the unmodified reference's Detector._preprocess_pil + to_tensor + format_tensor_img
    (api/detection.py:158-162, :177-205): every output float bit-exact, pad_info equal."""
    from oracle import preprocess as op
    g = golden('preprocess')
    for i, case in enumerate(g['cases']):
        name, size, div, code = str(case).split('|')
        got, pad = op.preprocess(g[f'pre{i}_img'], name, None if size == 'None' else int(size), int(div), code)
        assert got.dtype == np.float32 and got.shape == g[f'pre{i}_out'].shape, case
        assert np.array_equal(got.view(np.int32), g[f'pre{i}_out'].view(np.int32)), case
        assert (list(pad) if pad is not None else [-1] * 6) == g[f'pre{i}_pad'].tolist(), case


def test_resize_restatement_matches_pillow():
    """The restated two-pass 8-bit resampler against the installed Pillow (through torchvision's resize, as the reference
    calls it): random sizes, up- and down-scaling, one or both axes unchanged."""
    import PIL.Image
    import torchvision.transforms.functional as tvf
    from oracle import preprocess as op
    rng = np.random.default_rng(5)
    for it in range(40):
        h, w, oh, ow = (int(v) for v in rng.integers(3, 200, 4))
        ow = w if it % 7 == 0 else ow
        oh = h if it % 11 == 0 else oh
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = np.array(tvf.resize(PIL.Image.fromarray(img), (oh, ow)))
        assert np.array_equal(op.resize_bilinear_u8(img, oh, ow), want), (h, w, oh, ow)


def fullmodel_levels(g, name):
    """Rebuild the raw dicts the reference heads emitted from the stored NCHW head outputs and decode them with the oracle.
    Returns (box (N,P), cls (N,), score (N,), bb_format)."""
    levels = []
    if name == 'd1_fcs2':
        for li, s in enumerate((8, 16, 32, 64, 128)):
            raw = efdet_views(T(g[f'head{li}_0']).float(), T(g[f'head{li}_1']).float())
            levels.append(od.decode_fcos(raw, s, (int(g['params'][2]), int(g['params'][3]))))
        fmt = 'cxcywh'
    elif name == 'rapid':
        for li, s in enumerate((8, 16, 32)):
            levels.append(od.decode_rapid(yolo_views(T(g[f'head{li}_0']).float(), 3, 5, 0), level_anchors(RAPID_ANCHORS, li), s, 0))
        fmt = 'cxcywhd'
    else:
        for li, s in enumerate((8, 16, 32)):
            levels.append(od.decode_yolo(yolo_views(T(g[f'head{li}_0']).float(), 3, 4, 80), level_anchors(YOLO_ANCHORS, li), s, 80))
        fmt = 'cxcywh'
    box, cls, score = (t[0] for t in od.merge_levels(levels))
    return box, cls, score, fmt


@pytest.mark.parametrize('name', ['yolov3_80', 'rapid', 'd1_fcs2'])
def test_fullmodel_end_to_end(golden, name):
    """BASELINE configs[0] / [2] / [1] geometry end to end: the oracle's decode -> level concatenation -> post_process
    against the UNMODIFIED reference (OneStageBBox(configs/<name>.json) with random weights: backbone + FPN + head, det
    layers, general.py's concatenation, ImageObjects.post_process) on head outputs of a real network forward
    (tests/golden/make_golden.py: gen_fullmodel).  Kept indices and their order, boxes, scores, classes: bit-exact."""
    g = golden('fullmodel_' + name)
    conf, nms = float(g['params'][0]), float(g['params'][1])
    box, cls, score, fmt = fullmodel_levels(g, name)
    keep = opp.post_process(box, cls, score, conf, nms, fmt, 512)
    assert keep.numel() > 300 and torch.equal(keep, T(g['keep']))
    assert torch.equal(box[keep], T(g['kept_boxes'])) and torch.equal(score[keep], T(g['kept_scores']))
    assert torch.equal(cls[keep], T(g['kept_cats']))


def test_fullmodel_atss_training_targets(golden):
    """BASELINE configs[3]: oracle/atss.py against the ATSS target tensors captured from the UNMODIFIED reference's
    OneStageBBox(d1_fcs2_atss) in training mode (random weights, one 384 x 384 image, 100 GT boxes, all five levels;
    tests/golden/make_golden.py: gen_fullmodel_atss).  Bit-exact."""
    g = golden('fullmodel_d1_fcs2_atss')
    img_h, img_w, topk, ign, n_cls = g['params']
    gts = [(T(g['gt_boxes']), T(g['gt_cats']))]
    strides, sides = [int(v) for v in g['strides']], [float(v) for v in g['anchors']]
    n_pos = 0
    for li in range(5):
        t = T(g[f'atss{li}_bbox']).float().permute(0, 2, 3, 1)
        out = oa.assign_level(li, t, gts, (int(img_h), int(img_w)), strides, sides, int(topk), float(ign), int(n_cls))
        for k, v in out.items():
            assert torch.equal(v, T(g[f'atss{li}_{k}'])), (li, k)
        n_pos += int(out['PositiveMask'].sum())
    assert n_pos > 500


def test_tracking_oracle_bit_exact_vs_reference(golden):
    """oracle/tracking.py against the unmodified KFTracklet / RotBBoxKalmanFilter run (tests/golden/tracking.npz)."""
    from oracle import tracking as ot
    g = golden('tracking')
    bank = ot.Bank(g['init'], g['init_score'])
    for f in range(g['pred'].shape[0]):
        pred = bank.predict()
        assert np.array_equal(pred, g['pred'][f])
        assert np.array_equal(bank.likelihood(g['cand'][f]), g['lik'][f])
        upd = bank.update(g['meas'][f], g['meas_score'][f], g['has'][f])
        assert np.array_equal(upd, g['upd'][f])
        assert np.array_equal(bank.x, g['x'][f]) and np.array_equal(bank.P, g['P'][f]) and np.array_equal(bank.score, g['score'][f])
        assert np.array_equal(bank.feasible((1024, 1024), np.where(g['has'][f][:, None], upd, pred)), g['feasible'][f])


def test_uv5_training_oracle_vs_reference(golden):
    """oracle/train.py: uv5_targets_and_loss against the unmodified DetectLayer.forward(raw, img_size, labels)
    (tests/golden/train_uv5.npz): target tensors bit-exact, the loss exactly equal, both confidence targets."""
    from oracle import decode as od, train as ot
    from helpers import yolo_views, YOLO_ANCHORS
    g = golden('train_uv5')
    gts = [(T(g[f'gt{b}_boxes']), T(g[f'gt{b}_cats'])) for b in range(3)]
    idx3 = [[0, 1, 2], [3, 4, 5], [6, 7, 8]]
    for tag, mode in (('zo', 'zero-one'), ('iou', 'IoU')):
        for li, s in enumerate((8, 16, 32)):
            raw = yolo_views(T(g[f'{tag}{li}_in']), 3, 4, 5)
            anchors = torch.tensor(YOLO_ANCHORS, dtype=torch.float32)[idx3[li]]
            box, _, _ = od.decode_uv5(raw, anchors, s)
            r = ot.uv5_targets_and_loss(raw['bbox'], raw['conf'], raw['class'], box, gts, s, YOLO_ANCHORS, idx3[li], 5, mode, 0.3)
            assert torch.equal(r['TargetConf'], T(g[f'{tag}{li}_TargetConf'])), (tag, li)
            if mode == 'zero-one':
                assert torch.equal(r['IgnoredMask'], T(g[f'{tag}{li}_IgnoredMask']))
            assert float(r['loss']) == float(g[f'{tag}{li}_loss']), (tag, li, float(r['loss']), float(g[f'{tag}{li}_loss']))
            assert r['valid_gt_num'] == int(g[f'{tag}{li}_assigned'])


def test_retina_rotated_training_oracle_vs_reference(golden):
    """RetinaLayer training with rotated boxes: oracle/train.py against the loss of the reference's UNMODIFIED forward()
    (tests/golden/train_retina_rot.npz; the reference's __init__ cannot build this variant at HEAD, the fixture sets
    its three attributes by hand -- see make_golden.py: gen_train_retina_rot)."""
    from oracle import train as ot
    g = golden('train_retina_rot')
    gts = [(T(g[f'gt{b}_boxes']), torch.zeros(len(g[f'gt{b}_boxes']), dtype=torch.int64)) for b in range(3)]
    scales, ratios = [1, 1.26, 1.5874], [[1, 1], [1.4, 0.7], [0.7, 1.4]]
    for name in ('Periodic_L1', 'Periodic_smoothL1'):
        for li, s in zip((1, 2, 3), (16, 32, 64)):
            wh = torch.Tensor([(4 * s * sc * rt[0], 4 * s * sc * rt[1]) for sc in scales for rt in ratios])
            bb, cc = T(g[f'{name}{li}_bbox_in']), T(g[f'{name}{li}_cls_in'])
            n_b, _, n_h, n_w = bb.shape
            t = bb.view(n_b, 9, 5, n_h, n_w).permute(0, 1, 3, 4, 2)
            c = cc.view(n_b, 9, 1, n_h, n_w).permute(0, 1, 3, 4, 2)
            _, loss, pos = ot.retina_targets_and_loss(t, c, gts, (256, 320), s, wh, 0.5, 0.4, angle_loss=name)
            ref = float(g[f'{name}{li}_loss'])
            assert abs(float(loss) - ref) <= 1e-6 * max(1.0, abs(ref)), (name, li, float(loss), ref)
            assert f'pos {pos}/' in str(g[f'{name}{li}_loss_str'])


def test_raster_known_answers():
    """oracle/raster.c (pycocotools' rleFrPoly + rleIou restated; the library is not in this image) against run-length
    encodings DERIVED BY HAND from the published algorithm -- X = (int)(5x + .5); walk every edge along its major axis;
    the mask toggles where the walk steps from sub-column 5c+2 to 5c+3, at row ceil((v + .5)/5 - .5) of the smaller
    sub-row v; sorted column-major positions, differences = run lengths (zeros first).
      A  rectangle (1,1)-(4,3), 5 x 6 canvas: top edge v=5 -> row 1 at columns 1,2,3; bottom edge v=15 -> row 3:
         positions 6,8,11,13,16,18 then 30                                       -> 6,2,3,2,3,2,12
      B  diamond (3,1),(5,3),(3,5),(1,3), 7 x 7: slopes +-1, the four edges cross columns 3,4 / 4,3 / 2,1 / 1,2 at rows
         1,2 / 3,4 / 4,3 / 2,1: positions 9,10,15,18,22,25,30,31 then 49          -> 9,1,5,3,4,3,5,1,18  (8 pixels = its area)
      C  0.8 x 4.8 box (2.3,0.6)-(3.1,5.4), 7 x 6: X = 12..16 crosses only 12|13 (column 2); v=3 -> row 1, v=27 -> row 5
                                                                                 -> 15,4,23
      D  0.6-wide box (2.7,1)-(3.3,4): X = 14..17 steps over no 5c+2|5c+3 line     -> 42  (empty mask)
      E  parallelogram (1,1),(2,1),(4,5),(3,5), 7 x 6 (slope 1/2, y-major edges): right edge crosses 12|13 at t=5
         (v=9 -> row 2) and 17|18 at t=15 (v=19 -> row 4); left edge 12|13 at v=19 -> row 4 and 7|8 at v=9 -> row 2;
         top v=5 -> (1,1); bottom v=25 -> (3,5): positions 8,9,16,18,25,26 then 42 -> 8,1,7,2,7,1,16  (4 pixels = its area)"""
    assert oi.raster_rle([1, 1, 4, 1, 4, 3, 1, 3], 5, 6) == [6, 2, 3, 2, 3, 2, 12]
    assert oi.raster_rle([3, 1, 5, 3, 3, 5, 1, 3], 7, 7) == [9, 1, 5, 3, 4, 3, 5, 1, 18]
    assert oi.raster_rle([2.3, 0.6, 3.1, 0.6, 3.1, 5.4, 2.3, 5.4], 7, 6) == [15, 4, 23]
    assert oi.raster_rle([2.7, 1, 3.3, 1, 3.3, 4, 2.7, 4], 7, 6) == [42]
    assert oi.raster_rle([1, 1, 2, 1, 4, 5, 3, 5], 7, 6) == [8, 1, 7, 2, 7, 1, 16]
    # rleIou on two of them: rectangle A (6 px) and rectangle (2,1)-(5,4) (9 px) share columns 2,3 x rows 1,2
    import ctypes
    c1 = np.array([[1, 1, 4, 1, 4, 3, 1, 3]], dtype=np.float64)
    c2 = np.array([[2, 1, 5, 1, 5, 4, 2, 4], [2.7, 1, 3.3, 1, 3.3, 4, 2.7, 4]], dtype=np.float64)
    out = np.empty((1, 2))
    f64p = ctypes.POINTER(ctypes.c_double)
    oi.lib().oracle_raster_iou_pairwise(c1.ctypes.data_as(f64p), 1, c2.ctypes.data_as(f64p), 2, 6, 7, out.ctypes.data_as(f64p))
    assert out.tolist() == [[4 / 11, 0.0]]
