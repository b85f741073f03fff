"""GPU: the host-side mirror of the reference interface (same names / signatures / error behaviour)
against the golden fixtures produced by the real reference."""
import numpy as np
import pytest
import torch

from helpers import T, level_anchors, yolo_views, efdet_views, YOLO_ANCHORS, RAPID_ANCHORS
from test_gpu_parity import close, cls_match, FCOS_W, ANGLE, CORNERS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('tag,fmt', [('pp_small', 'cxcywh'), ('pp_cap', 'cxcywh'), ('pp_rot', 'cxcywhd'),
                                     ('pp_empty', 'cxcywh')])
def test_image_objects_post_process(golden, tag, fmt):
    """ImageObjects(...).post_process(conf, nms) exactly as api/detection.py:172 calls it, CPU tensors in."""
    from mydetection_b200.structures import ImageObjects
    g = golden('postprocess')
    boxes, scores, cats = T(g[tag + '_boxes']), T(g[tag + '_scores']), T(g[tag + '_cats'])
    conf, nms = (float(v) for v in g[tag + '_params'])
    keep = T(g[tag + '_keep'])
    dts = ImageObjects(boxes.clone(), cats.clone(), scores=scores.clone(), bb_format=fmt, img_hw=(512, 512))
    out = dts.post_process(conf, nms)
    assert isinstance(out, ImageObjects) and out.bboxes.device.type == 'cpu'      # structures.py:97 cpu_()
    assert torch.equal(out.bboxes, boxes[keep]) and torch.equal(out.scores, scores[keep])
    assert torch.equal(out.cats, cats[keep]) and out.cats.dtype == torch.int64
    assert out.img_hw == (512, 512) and out._bb_format == fmt


def test_image_objects_nms_direct(golden):
    from mydetection_b200.structures import ImageObjects
    g = golden('postprocess')
    boxes, scores, cats = T(g['nms_direct_boxes']), T(g['nms_direct_scores']), T(g['nms_direct_cats'])
    keep = T(g['nms_direct_keep'])
    for device in ('cpu', 'cuda'):
        dts = ImageObjects(boxes.to(device), cats.to(device), scores=scores.to(device))
        out = dts.nms(float(g['nms_direct_params'][1]))
        assert out.bboxes.device.type == device
        assert torch.equal(out.bboxes.cpu(), boxes[keep]) and torch.equal(out.cats.cpu(), cats[keep])
    # x1y1x2y2 input format (structures.py:133-136)
    from oracle import iou as oi, postprocess as opp
    xy = oi.cxcywh_to_x1y1x2y2(boxes)
    dts = ImageObjects.__new__(ImageObjects)
    dts.bboxes, dts.cats, dts.masks, dts.scores, dts._bb_format, dts.img_hw = xy, cats, None, scores, 'x1y1x2y2', None
    out = ImageObjects.non_max_suppression(dts, 0.3)
    want = opp.class_nms(xy, scores, cats, 0.3, 'x1y1x2y2')
    assert torch.equal(out.bboxes, xy[want])


def test_bbox_ops_mirror(golden):
    from mydetection_b200 import bbox_ops
    g = golden('iou')
    a, b = T(g['a']), T(g['b'])
    out = bbox_ops.bboxes_iou(a, b)                                   # CPU in -> CPU out, bit-exact
    assert out.device.type == 'cpu' and torch.equal(out, T(g['iou_cxcywh']))
    assert torch.equal(bbox_ops.bboxes_iou(a[0], b), T(g['iou_cxcywh'])[0:1])     # 1-d first argument, :25-26
    assert torch.equal(bbox_ops.cxcywh_to_x1y1x2y2(a), T(g['a_xyxy']))
    assert torch.equal(bbox_ops.bboxes_iou(T(g['a_xyxy']), bbox_ops.cxcywh_to_x1y1x2y2(b), xyxy=True), T(g['iou_xyxy']))
    rb, rs = T(g['rot_boxes']), T(g['rot_scores'])
    rad = rb.clone()
    rad[:, 4] = rad[:, 4] * np.pi / 180
    v = bbox_ops.xywha2vertex(rad, is_degree=False)
    assert v.shape == (120, 4, 2)
    close(v, T(g['rot_vertices']), 512, 'xywha2vertex', cancel=CORNERS)
    assert bbox_ops.xywha2vertex(rad, is_degree=False, stack=False).shape == (120, 8)
    # nms_rotbb incl. majority voting, against the reference's control flow
    assert torch.equal(bbox_ops.nms_rotbb(rb, rs, 0.45), T(g['rot_keep_045']))
    assert torch.equal(bbox_ops.nms_rotbb(rb, rs, 0.2), T(g['rot_keep_02']))
    assert torch.equal(bbox_ops.nms_rotbb(rb, rs, 0.3, majority=2), T(g['rot_keep_maj2']))
    assert torch.equal(bbox_ops.nms_rotbb(rb.cuda(), rs.cuda(), 0.45).cpu(), T(g['rot_keep_045']))
    # iou_rle: float64 like the reference (pycocotools returns doubles), numpy in / numpy out
    rgt = T(g['rot_gt_debug3'])[:40]
    iou = bbox_ops.iou_rle(rgt, rgt)
    assert iou.dtype == torch.float64
    np.testing.assert_allclose(iou.numpy(), g['rot_gt_iou_stub'], rtol=0, atol=2e-6)
    assert isinstance(bbox_ops.iou_rle(rgt.numpy(), rgt.numpy(), return_numpy=True), np.ndarray)
    assert bbox_ops.iou_rle(rgt[0], rgt, img_size=1024).shape == (1, 40)


def test_det_layers(golden):
    from mydetection_b200 import detlayers
    g = golden('decode')
    # FCOS2 per level through the registry, as OneStageBBox builds it (models/general.py:34-38)
    cfg = {'model.pred_layer': 'FCOS2', 'model.fcos.anchors': [0, 64, 128, 256, 512, 100000000],
           'model.fpn.out_strides': [8, 16, 32, 64, 128], 'general.num_class': 6,
           'model.fcos2.ignored_threshold': 0.7, 'general.pred_bbox_format': 'cxcywh',
           'model.atss.anchors': [24, 48, 96, 192, 384], 'model.atss.topk_per_level': 9}
    layer_cls = detlayers.get_det_layer(cfg)
    for li in range(5):
        raw = efdet_views(T(g[f'fcos{li}_bbox_in']), T(g[f'fcos{li}_cls_in']))
        raw = {k: v.cuda() for k, v in raw.items()}
        preds, loss = layer_cls(level_i=li, cfg=cfg)(raw, (256, 384), None)
        assert loss is None and set(preds) == {'bbox', 'class_idx', 'score'}
        assert preds['class_idx'].dtype == torch.int64
        close(preds['bbox'], T(g[f'fcos{li}_bbox']), 384, 'fcos layer bbox', cancel=FCOS_W)
        close(preds['score'], T(g[f'fcos{li}_score']), 1, 'fcos layer score')
        cls_match(preds['class_idx'], raw['class'].cpu().reshape(2, -1, 6), T(g[f'fcos{li}_cls']), 'fcos layer cls')
        atss, _ = detlayers.FCOS_ATSS_Layer(li, cfg)(raw, (256, 384))
        assert torch.equal(atss['bbox'], preds['bbox'])
    # YOLO + RAPiD on CPU inputs (the layers stage them on the GPU)
    ycfg = {'model.yolo.anchors': YOLO_ANCHORS, 'model.yolo.anchor_indices': [[0, 1, 2], [3, 4, 5], [6, 7, 8]],
            'model.fpn.out_strides': [8, 16, 32], 'general.num_class': 5}
    preds, _ = detlayers.YOLOLayer(2, ycfg)(yolo_views(T(g['yolo2_in']), 3, 4, 5), (96, 128))
    close(preds['bbox'], T(g['yolo2_bbox']), 128, 'yolo layer')
    rcfg = {'model.rapid.anchors': RAPID_ANCHORS, 'model.rapid.anchor_indices': [[0, 1, 2], [3, 4, 5], [6, 7, 8]],
            'model.fpn.out_strides': [8, 16, 32], 'general.num_class': 0}
    preds, _ = detlayers.RAPiDLayer(0, rcfg)(yolo_views(T(g['rapid_c0_0_in']), 3, 5, 0), (96, 128))
    close(preds['bbox'][..., :4], T(g['rapid_c0_0_bbox'])[..., :4], 128, 'rapid layer')
    assert preds['bbox'].shape[-1] == 5
    # the one training branch outside the built scope says so: FCOS v1 (no IoU step in its assignment, fcos.py:70-140)
    fcfg = {'model.fcos.anchors': [0, 64, 128, 256, 512, 1e8], 'model.fpn.out_strides': [8, 16, 32, 64, 128], 'general.num_class': 5}
    raw = {'bbox': torch.zeros(1, 12, 16, 4), 'center': torch.zeros(1, 12, 16, 1), 'class': torch.zeros(1, 12, 16, 5)}
    with pytest.raises(NotImplementedError):
        detlayers.get_det_layer({'model.pred_layer': 'FCOS'})(0, fcfg)(raw, (96, 128), labels=[None])


def _train_labels(g):
    from mydetection_b200.structures import ImageObjects
    return [ImageObjects(T(g[f'gt{b}_boxes']), T(g[f'gt{b}_cats']), bb_format='cxcywh', img_hw=(256, 320)) for b in range(3)]


def test_iou_rowmax_vs_oracle(golden):
    """mydet_iou_aabb_rowmax == bboxes_iou(...).max(dim=1) of the oracle, bit for bit (values AND first-max indices),
    ragged GT counts, an image without GT, shared row boxes, xyxy."""
    from mydetection_b200 import ops
    from oracle import iou as oi
    gen = torch.Generator().manual_seed(5)
    n, n_g = 3000, 700                                   # more GT than one shared-memory stage (512)
    a = torch.cat([torch.rand(3, n, 2, generator=gen) * 500, torch.rand(3, n, 2, generator=gen) * 80 + 1], -1)
    a[0, :50] = a[0, 50:100]                             # exact duplicates: ties in the arg-max
    gt = torch.cat([torch.rand(3, n_g, 2, generator=gen) * 500, torch.rand(3, n_g, 2, generator=gen) * 120 + 1], -1)
    gt[0, 300:350] = gt[0, 100:150]
    counts = torch.tensor([n_g, 0, 37], dtype=torch.int32)
    mx, arg = ops.iou_rowmax(a.cuda(), gt.cuda(), counts.cuda())
    for b in range(3):
        c = int(counts[b])
        if c == 0:
            assert bool((mx[b] == -1).all()) and bool((arg[b] == -1).all())
            continue
        want, want_arg = oi.bboxes_iou(a[b], gt[b, :c]).max(dim=1)
        assert torch.equal(mx[b].cpu(), want) and torch.equal(arg[b].cpu(), want_arg)
    mx2, arg2 = ops.iou_rowmax(a[2].cuda(), gt.cuda(), None)            # one set of row boxes for every image
    want, want_arg = oi.bboxes_iou(a[2], gt[1]).max(dim=1)
    assert torch.equal(mx2[1].cpu(), want) and torch.equal(arg2[1].cpu(), want_arg)
    ax, gx = oi.cxcywh_to_x1y1x2y2(a[0]), oi.cxcywh_to_x1y1x2y2(gt[0])
    mx3, arg3 = ops.iou_rowmax(ax[None].cuda(), gx[None].cuda(), None, xyxy=True)
    want, want_arg = oi.bboxes_iou(ax, gx, xyxy=True).max(dim=1)
    assert torch.equal(mx3[0].cpu(), want) and torch.equal(arg3[0].cpu(), want_arg)


def test_yolo_layer_training_forward(golden):
    """YOLOLayer.forward(raw, img_size, labels) against the unmodified reference (tests/golden/train.npz):
    masks and class targets bit-exact, regression targets / weights / loss within 1e-5 relative."""
    from mydetection_b200 import detlayers
    g = golden('train')
    labels = _train_labels(g)
    cfg = {'model.yolo.anchors': YOLO_ANCHORS, 'model.yolo.anchor_indices': [[0, 1, 2], [3, 4, 5], [6, 7, 8]],
           'model.yolo.anchor.negative_threshold': 0.3, 'model.fpn.out_strides': [8, 16, 32], 'general.num_class': 5}
    for li in range(3):
        nchw = T(g[f'yolo{li}_in']).cuda().requires_grad_(True)
        raw = yolo_views(nchw, 3, 4, 5)
        layer = detlayers.YOLOLayer(li, cfg)
        preds, loss = layer(raw, (256, 320), labels)
        tg = layer.targets
        assert torch.equal(tg['gt_mask'].cpu(), T(g[f'yolo{li}_gt_mask']))
        assert torch.equal(tg['conf_loss_mask'].cpu(), T(g[f'yolo{li}_conf_loss_mask']))
        assert torch.equal(tg['tgt_cls'].cpu(), T(g[f'yolo{li}_tgt_cls']))
        torch.testing.assert_close(tg['tgt_xywh'].cpu(), T(g[f'yolo{li}_tgt_xywh']), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(tg['weighted'].cpu().reshape(g[f'yolo{li}_weighted'].shape), T(g[f'yolo{li}_weighted']), rtol=1e-6, atol=0)
        ref = float(g[f'yolo{li}_loss'])
        assert abs(float(loss.detach()) - ref) <= 1e-5 * max(1.0, abs(ref)), (li, float(loss.detach()), ref)
        assert layer._assigned_num == int(g[f'yolo{li}_assigned'])
        loss.backward()                                   # the loss is differentiable w.r.t. the head output
        assert nchw.grad is not None and float(nchw.grad.abs().sum()) > 0


def test_fcos2_layer_training_forward(golden):
    """FCOSLayer (FCOS2) forward(raw, img_size, labels): targets from mydet_fcos_assign, bit-exact masks / class
    targets / ltrb targets against the reference's captured tensors; loss within 1e-5 relative."""
    from mydetection_b200 import detlayers
    g = golden('train')
    labels = _train_labels(g)
    strides = [8, 16, 32, 64, 128]
    cfg = {'model.fcos.anchors': [0, 64, 128, 256, 512, 100000000], 'model.fpn.out_strides': strides, 'general.num_class': 5,
           'model.fcos2.ignored_threshold': 0.2, 'general.pred_bbox_format': 'cxcywh'}
    for li in (0, 1, 2):
        bb = T(g[f'fcos{li}_bbox_in']).cuda().requires_grad_(True)
        cc = T(g[f'fcos{li}_cls_in']).cuda().requires_grad_(True)
        layer = detlayers.FCOSLayer(li, cfg)
        raw = efdet_views(bb, cc)
        tg = layer.assign(raw['bbox'].detach(), (256, 320), labels)
        assert torch.equal(tg['PositiveMask'].cpu(), T(g[f'fcos{li}_PositiveMask']))
        assert torch.equal(tg['TargetCls'].cpu(), T(g[f'fcos{li}_TargetCls']))
        assert torch.equal(tg['TargetConf'].cpu(), T(g[f'fcos{li}_TargetConf']))
        assert torch.equal(tg['TargetLTRB'].cpu(), T(g[f'fcos{li}_TargetLTRB']))
        # the ignore mask compares an IoU of exp()-decoded boxes with 0.2: CUDA expf vs the CPU's exp may flip a cell
        # whose IoU is within 1e-5 of the threshold (none on this fixture)
        assert torch.equal(tg['IgnoredMask'].cpu(), T(g[f'fcos{li}_IgnoredMask']))
        preds, loss = layer(raw, (256, 320), labels)
        ref = float(g[f'fcos{li}_loss'])
        assert abs(float(loss.detach()) - ref) <= 1e-5 * max(1.0, abs(ref)), (li, float(loss.detach()), ref)
        loss.backward()
        assert bb.grad is not None and cc.grad is not None and float(cc.grad.abs().sum()) > 0


def test_retina_layer_training_forward(golden):
    """RetinaLayer.forward(raw, img_size, labels): anchor matching through mydet_iou_aabb_rowmax for the whole batch.
    Masks / matched GT indices / class targets equal the oracle's per-image loop bit for bit, regression targets
    within 1e-5, the loss equals the reference run's (train.npz) within 1e-5 relative."""
    from mydetection_b200 import detlayers
    from mydetection_b200.structures import ImageObjects
    from oracle import train as ot
    g = golden('train')
    gts1 = [(T(g[f'gt{b}_boxes']), torch.zeros(len(g[f'gt{b}_cats']), dtype=torch.int64)) for b in range(3)]
    labels = [ImageObjects(bx, ct, bb_format='cxcywh', img_hw=(256, 320)) for bx, ct in gts1]
    strides = [8, 16, 32, 64, 128]
    cfg = {'model.fpn.out_strides': strides, 'model.retina.anchor.base': 4, 'model.retina.anchor.scales': [1, 1.26, 1.5874],
           'model.retina.anchor.ratios': [[1, 1], [1.4, 0.7], [0.7, 1.4]], 'model.retina.anchor.positive_threshold': 0.5,
           'model.retina.anchor.negative_threshold': 0.4, 'general.num_class': 1, 'general.pred_bbox_format': 'cxcywh',
           'general.bbox_param': 4}
    for li in (1, 2, 3):
        bb = T(g[f'retina{li}_bbox_in']).cuda().requires_grad_(True)
        cc = T(g[f'retina{li}_cls_in']).cuda().requires_grad_(True)
        n_b, _, n_h, n_w = bb.shape
        raw = {'bbox': bb.view(n_b, 9, 4, n_h, n_w).permute(0, 1, 3, 4, 2), 'class': cc.view(n_b, 9, 1, n_h, n_w).permute(0, 1, 3, 4, 2)}
        layer = detlayers.RetinaLayer(li, cfg)
        preds, loss = layer(raw, (256, 320), labels)
        assert preds is None                                                  # as the reference (retinanet.py:160)
        ref = float(g[f'retina{li}_loss'])
        assert abs(float(loss.detach()) - ref) <= 1e-5 * max(1.0, abs(ref)), (li, float(loss.detach()), ref)
        per_image, _, pos = ot.retina_targets_and_loss(raw['bbox'].detach().cpu(), raw['class'].detach().cpu(), gts1, (256, 320),
                                                       strides[li], layer.anchor_wh, 0.5, 0.4)
        tg = layer.targets
        assert int(tg['M_pos'].sum()) == pos
        for b, want in enumerate(per_image):
            if want is None:
                assert not bool(tg['M_pos'][b].any()) and bool(tg['cls_penalty_mask'][b].all())
                continue
            for k in ('M_pos', 'M_neg', 'gt_idx', 'tgt_cls', 'cls_penalty_mask'):
                assert torch.equal(tg[k][b].cpu(), want[k]), (li, b, k)
            torch.testing.assert_close(tg['tgt_xywh'][b].cpu(), want['tgt_xywh'], rtol=1e-5, atol=1e-6)
        loss.backward()
        assert bb.grad is not None and float(cc.grad.abs().sum()) > 0


def test_rapid_layer_training_forward(golden):
    """RAPiDLayer.forward(raw, img_size, labels): rotated IoUs from mydet_iou_rot_pairwise.  The fixture comes from the
    reference driven by an exact polygon-clipping stub of pycocotools, so masks and targets must agree exactly
    (control flow), the loss within 1e-5 relative; the raster-vs-exact IoU gap of the real library is not covered."""
    from mydetection_b200 import detlayers
    from mydetection_b200.structures import ImageObjects
    g = golden('train')
    labels = [ImageObjects(T(g[f'rgt{b}_boxes']), torch.zeros(len(g[f'rgt{b}_boxes']), dtype=torch.int64), bb_format='cxcywhd',
                           img_hw=(256, 320)) for b in range(3)]
    cfg = {'model.rapid.anchors': RAPID_ANCHORS, 'model.rapid.anchor_indices': [[0, 1, 2], [3, 4, 5], [6, 7, 8]],
           'model.fpn.out_strides': [8, 16, 32], 'general.num_class': 0, 'model.rapid.wh_smooth_l1_beta': 1,
           'model.angle.loss_angle': 'Periodic_L1', 'model.angle.pred_range': 360}
    for li in range(3):
        nchw = T(g[f'rapid{li}_in']).cuda().requires_grad_(True)
        layer = detlayers.RAPiDLayer(li, cfg)
        layer.ignore_thre = 0.1                                   # as the fixture (make_golden.py)
        preds, loss = layer(yolo_views(nchw, 3, 5, 0), (256, 320), labels)
        tg = layer.targets
        assert torch.equal(tg['PositiveMask'].cpu(), T(g[f'rapid{li}_PositiveMask']))
        assert torch.equal(tg['IgnoredMask'].cpu(), T(g[f'rapid{li}_IgnoredMask']))
        assert torch.equal(tg['TargetConf'].cpu(), T(g[f'rapid{li}_TargetConf']))
        torch.testing.assert_close(tg['TargetXYWH'].cpu(), T(g[f'rapid{li}_TargetXYWH']), rtol=1e-5, atol=2e-6)
        torch.testing.assert_close(tg['TargetAngle'].cpu(), T(g[f'rapid{li}_TargetAngle']), rtol=1e-6, atol=1e-7)
        ref = float(g[f'rapid{li}_loss'])
        assert abs(float(loss.detach()) - ref) <= 1e-5 * max(1.0, abs(ref)), (li, float(loss.detach()), ref)
        assert layer._assigned_num == int(T(g[f'rapid{li}_PositiveMask']).sum())
        loss.backward()
        assert nchw.grad is not None and float(nchw.grad.abs().sum()) > 0


def test_one_stage_forward_flow(golden):
    """The caller's flow of models/general.py:69-84 + api/detection.py:172 on the mirror: per-level layers,
    level concat, per-image ImageObjects, post_process -- end to end against the oracle chain."""
    from mydetection_b200 import detlayers
    from mydetection_b200.structures import ImageObjects
    from oracle import decode as od, postprocess as opp
    g = golden('decode')
    strides = [8, 16, 32, 64, 128]
    cfg = {'model.pred_layer': 'FCOS2', 'model.fcos.anchors': [0, 64, 128, 256, 512, 100000000],
           'model.fpn.out_strides': strides, 'general.num_class': 6, 'model.fcos2.ignored_threshold': 0.7,
           'general.pred_bbox_format': 'cxcywh'}
    raws = [efdet_views(T(g[f'fcos{li}_bbox_in']), T(g[f'fcos{li}_cls_in'])) for li in range(5)]
    layers = [detlayers.get_det_layer(cfg)(level_i=i, cfg=cfg) for i in range(5)]
    dts_all = [layers[i]({k: v.cuda() for k, v in raws[i].items()}, (256, 384), None)[0] for i in range(5)]
    bbs = torch.cat([d['bbox'] for d in dts_all], dim=1)
    cls_idx = torch.cat([d['class_idx'] for d in dts_all], dim=1)
    scores = torch.cat([d['score'] for d in dts_all], dim=1)
    ref = od.merge_levels([od.decode_fcos(r, s, (256, 384)) for r, s in zip(raws, strides)])
    for b in range(2):
        objs = ImageObjects(bboxes=bbs[b], cats=cls_idx[b], scores=scores[b], bb_format='cxcywh', img_hw=(256, 384))
        out = objs.post_process(0.05, 0.5)
        want = opp.post_process(ref[0][b], ref[1][b], ref[2][b], 0.05, 0.5, 'cxcywh', 512)
        assert len(out) == want.numel()
        close(out.bboxes, ref[0][b][want], 384, 'flow box', cancel=FCOS_W)
        assert torch.equal(out.cats, ref[1][b][want])


def test_post_process_is_batched_behind_the_per_image_calls(golden, monkeypatch):
    """The reference's per-image post_process loop (models/general.py:78-84, api/detection.py:172) on the mirror costs
    ONE mydet_postprocess launch per batch: rows of a batch tensor are recognised by their view metadata.  Same results
    as the per-image path; a different threshold, an in-place edit of the batch tensor or of a returned result must
    not be served from the cached batch."""
    from mydetection_b200 import detlayers, ops, structures
    from mydetection_b200.structures import ImageObjects
    g = golden('decode')
    strides = [8, 16, 32, 64, 128]
    cfg = {'model.pred_layer': 'FCOS2', 'model.fcos.anchors': [0, 64, 128, 256, 512, 100000000],
           'model.fpn.out_strides': strides, 'general.num_class': 6, 'model.fcos2.ignored_threshold': 0.7,
           'general.pred_bbox_format': 'cxcywh'}
    raws = [efdet_views(T(g[f'fcos{li}_bbox_in']), T(g[f'fcos{li}_cls_in'])) for li in range(5)]
    layers = [detlayers.get_det_layer(cfg)(level_i=i, cfg=cfg) for i in range(5)]
    dts_all = [layers[i]({k: v.cuda() for k, v in raws[i].items()}, (256, 384), None)[0] for i in range(5)]
    bbs = torch.cat([d['bbox'] for d in dts_all], dim=1)
    cls_idx = torch.cat([d['class_idx'] for d in dts_all], dim=1)
    scores = torch.cat([d['score'] for d in dts_all], dim=1)
    calls = []
    real = ops.postprocess
    monkeypatch.setattr(ops, 'postprocess', lambda *a, **k: (calls.append(a[0].shape[0]), real(*a, **k))[1])

    def loop(conf, nms):
        objs = [ImageObjects(bboxes=b, cats=c, scores=s_, bb_format='cxcywh', img_hw=(256, 384))
                for b, c, s_ in zip(bbs, cls_idx, scores)]                   # general.py:78-84
        return [o.post_process(conf, nms) for o in objs]

    def alone(b, conf, nms):      # the same image as a tensor of its own: per-image path
        o = ImageObjects(bboxes=bbs[b].clone(), cats=cls_idx[b].clone(), scores=scores[b].clone(), bb_format='cxcywh')
        return o.post_process(conf, nms)

    structures._BATCH.clear()
    got = loop(0.05, 0.5)
    assert calls == [2], calls                                                # one launch for the batch of 2
    for b in range(2):
        want = alone(b, 0.05, 0.5)
        assert not got[b].bboxes.is_cuda and len(got[b]) == len(want) > 0
        assert torch.equal(got[b].bboxes, want.bboxes) and torch.equal(got[b].scores, want.scores) and torch.equal(got[b].cats, want.cats)
    calls.clear()
    got[0].bboxes.mul_(0.0)                                                   # the caller edits ITS result in place ...
    again = loop(0.05, 0.5)
    assert calls == [] and float(again[0].bboxes.abs().sum()) > 0             # ... the cached batch is intact, no new launch
    other = loop(0.3, 0.3)
    assert calls == [2]                                                       # other thresholds: recomputed, not served from the cache
    want = alone(0, 0.3, 0.3)
    assert len(other[0]) == len(want) < len(again[0]) and torch.equal(other[0].bboxes, want.bboxes)
    calls.clear()
    scores[1].mul_(0.5)                                                       # the batch tensor changes: version counter
    changed = loop(0.3, 0.3)
    assert calls == [2]
    want = alone(1, 0.3, 0.3)
    assert len(changed[1]) == len(want) and torch.equal(changed[1].scores, want.scores)


def test_cepdof_matching_iou(golden):
    """CEPDOFeval.computeIoU (dt x gt rotated IoU, stable score order, maxDets cap)."""
    from mydetection_b200 import evaluation
    from oracle import iou as oi
    g = golden('iou')
    rgt = T(g['rot_gt_debug3'])
    gts = [{'bbox': b.tolist()} for b in rgt[:30]]
    gen = torch.Generator().manual_seed(5)
    dt_boxes = rgt[:60].clone()
    dt_boxes[:, :2] += torch.randn(60, 2, generator=gen) * 6
    dt_boxes[:, 4] += torch.randn(60, generator=gen) * 10
    scores = (torch.rand(60, generator=gen) * 20).round() / 20            # ties: the sort must be stable
    dts = [{'bbox': b.tolist(), 'score': float(s)} for b, s in zip(dt_boxes, scores)]
    ious, order = evaluation.compute_iou(dts, gts, max_dets=40)
    want_order = np.argsort([-d['score'] for d in dts], kind='mergesort')[:40]
    assert np.array_equal(order, want_order) and ious.shape == (40, 30) and ious.dtype == np.float64
    want = oi.iou_rot_f64([dts[i]['bbox'] for i in want_order], [x['bbox'] for x in gts])     # float64 corners, as the evaluator
    np.testing.assert_allclose(ious, want, rtol=0, atol=1e-9)
    assert evaluation.iou_rle([], [[1, 2, 3, 4, 5]]).shape == (0, 1)
    assert evaluation.compute_iou([], [])[0] == []


def test_tracklet_bank_vs_reference(golden):
    """12 tracklets over 8 frames -- predict, likelihood of 6 candidates, update of the tracklets that
    have a measurement -- against the states, boxes, scores, feasibility flags and likelihoods the UNMODIFIED reference
    produced one KFTracklet at a time (tests/golden/tracking.npz).  float64; 1e-9 relative (the 5x5 inverse is
    Gauss-Jordan here, LAPACK's LU in numpy); the association IoU against the oracle's exact clipping."""
    from mydetection_b200 import tracking
    from oracle import iou as oi
    g = golden('tracking')
    bank = tracking.TrackletBank(g['init'], g['init_score'], img_hw=(1024, 1024))
    worst = 0.0

    def close(got, want, what):
        nonlocal worst
        got, want = got.cpu().numpy(), np.asarray(want)
        err = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
        err = np.where(np.abs(want) < 1e-200, np.abs(got - want), err)
        worst = max(worst, float(err.max()))
        assert float(err.max()) < 1e-9, (what, float(err.max()))

    for f in range(g['pred'].shape[0]):
        close(bank.predict(), g['pred'][f], f'pred {f}')
        close(bank.likelihood(g['cand'][f]), g['lik'][f], f'likelihood {f}')
        iou = bank.association_iou(g['cand'][f])
        want_iou = oi.iou_rot(bank.bbox.float().cpu(), T(g['cand'][f]).float())
        assert float((iou.cpu() - want_iou).abs().max()) < 1e-6
        upd = bank.update(g['meas'][f], g['meas_score'][f], g['has'][f])
        close(upd, g['upd'][f], f'upd {f}')
        close(bank.x, g['x'][f], f'x {f}')
        close(bank.P, g['P'][f], f'P {f}')
        close(bank.score, g['score'][f], f'score {f}')
        assert np.array_equal(bank.is_feasible().cpu().numpy(), g['feasible'][f])
    print(f'tracklet bank: max relative error {worst:.3e}')
    with pytest.raises(AssertionError):
        bank.update(g['meas'][0], g['meas_score'][0])        # update without a predict, as the reference asserts


@pytest.mark.parametrize('tag,mode', [('zo', 'zero-one'), ('iou', 'IoU')])
def test_uv5_layer_training_vs_reference(golden, tag, mode):
    """DetectLayer.forward(raw, img_size, labels) of the mirror against the unmodified reference's
    (tests/golden/train_uv5.npz): confidence targets / ignore mask bit-exact outside a 1e-5 band around the IoU
    values, the loss within 1e-5 relative, differentiable w.r.t. the head output."""
    from mydetection_b200.detlayers.uv5 import DetectLayer
    from mydetection_b200.structures import ImageObjects
    g = golden('train_uv5')
    labels = [ImageObjects(T(g[f'gt{b}_boxes']), T(g[f'gt{b}_cats']), bb_format='cxcywh', img_hw=(256, 320)) for b in range(3)]
    cfg = {'model.detect.anchors': YOLO_ANCHORS, 'model.detect.anchor_indices': [[0, 1, 2], [3, 4, 5], [6, 7, 8]],
           'model.fpn.out_strides': [8, 16, 32], 'general.num_class': 5, 'model.detect.sample_selection': 'best',
           'model.detect.confidence_target': mode, 'model.detect.negative_threshold': 0.3, 'model.detect.loss_bbox': 'smooth_L1',
           'general.pred_bbox_format': 'cxcywh'}
    for li in range(3):
        nchw = T(g[f'{tag}{li}_in']).cuda().requires_grad_(True)
        layer = DetectLayer(li, cfg)
        preds, loss = layer(yolo_views(nchw, 3, 4, 5), (256, 320), labels)
        want_conf = T(g[f'{tag}{li}_TargetConf'])
        got_conf = layer.targets['TargetConf'].cpu()
        if mode == 'IoU':
            torch.testing.assert_close(got_conf, want_conf, rtol=1e-5, atol=1e-6)
        else:
            assert torch.equal(got_conf, want_conf)
            assert torch.equal(layer.targets['IgnoredMask'].cpu(), T(g[f'{tag}{li}_IgnoredMask']))
        ref = float(g[f'{tag}{li}_loss'])
        assert abs(float(loss.detach()) - ref) <= 1e-5 * max(1.0, abs(ref)), (li, float(loss.detach()), ref)
        assert layer._assigned_num == int(g[f'{tag}{li}_assigned'])
        assert layer.loss_str.split(':')[0] == str(g[f'{tag}{li}_loss_str']).split(':')[0]
        loss.backward()
        assert nchw.grad is not None and float(nchw.grad.abs().sum()) > 0


@pytest.mark.parametrize('name', ['Periodic_L1', 'Periodic_smoothL1'])
def test_retina_rotated_training_vs_reference(golden, name):
    """RetinaLayer.forward(raw, img_size, labels) with 'cxcywhd' boxes on the mirror against the loss and the positive
    count of the reference's unmodified forward() (tests/golden/train_retina_rot.npz), 1e-5 relative; differentiable."""
    from mydetection_b200.detlayers.retinanet import RetinaLayer
    from mydetection_b200.structures import ImageObjects
    g = golden('train_retina_rot')
    labels = [ImageObjects(T(g[f'gt{b}_boxes']), torch.zeros(len(g[f'gt{b}_boxes']), dtype=torch.int64), bb_format='cxcywhd',
                           img_hw=(256, 320)) for b in range(3)]
    cfg = {'model.fpn.out_strides': [8, 16, 32, 64, 128], 'model.retina.anchor.base': 4, 'model.retina.anchor.scales': [1, 1.26, 1.5874],
           'model.retina.anchor.ratios': [[1, 1], [1.4, 0.7], [0.7, 1.4]], 'model.retina.anchor.positive_threshold': 0.5,
           'model.retina.anchor.negative_threshold': 0.4, 'general.num_class': 1, 'general.pred_bbox_format': 'cxcywhd',
           'general.bbox_param': 5, 'model.angle.loss_name': name}
    for li in (1, 2, 3):
        bb = T(g[f'{name}{li}_bbox_in']).cuda().requires_grad_(True)
        cc = T(g[f'{name}{li}_cls_in']).cuda()
        n_b, _, n_h, n_w = bb.shape
        raw = {'bbox': bb.view(n_b, 9, 5, n_h, n_w).permute(0, 1, 3, 4, 2), 'class': cc.view(n_b, 9, 1, n_h, n_w).permute(0, 1, 3, 4, 2)}
        layer = RetinaLayer(li, cfg)
        _, loss = layer(raw, (256, 320), labels)
        ref = float(g[f'{name}{li}_loss'])
        assert abs(float(loss.detach()) - ref) <= 1e-5 * max(1.0, abs(ref)), (name, li, float(loss.detach()), ref)
        assert layer.loss_str.split(':')[0] == str(g[f'{name}{li}_loss_str']).split(':')[0]
        loss.backward()
        assert bb.grad is not None and (float(bb.grad.abs().sum()) > 0) == (' pos 0/' not in layer.loss_str)   # level 3 has no positive
