"""GPU: the five BASELINE.json configurations at their FULL sizes.  A few images of each are compared
with the oracle directly; the whole batch is covered by size-independent properties (determinism,
compaction == dense threshold, ordered output, greedy-NMS invariants, idempotence)."""
import pytest
import torch

from test_gpu_parity import close, cls_match, FCOS_W, ANGLE, CORNERS

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def fcos_batch(gen, batch, img, strides, n_cls, conf_mu):
    from mydetection_b200.heads import efdet_head_views
    raws = []
    for s in strides:
        n = img // s
        bb = torch.randn(batch, 4, n, n, generator=gen) * 0.5
        cc = torch.randn(batch, 1 + n_cls, n, n, generator=gen) * 1.5
        cc[:, 0] += conf_mu
        cc[:, 1:] -= 2.0
        raws.append(efdet_head_views(bb, cc))
    return raws


def to_dev(raws):
    return [{k: v.to(DEV) for k, v in r.items()} for r in raws]


def check_output_invariants(out, thr, box_format='cxcywh'):
    """Ordered (class asc, score desc), and no kept same-class pair above the IoU threshold."""
    from mydetection_b200 import ops
    from oracle import iou as oi
    counts = out['count'].tolist()
    for b, n in enumerate(counts):
        cls, sc, bx = out['cls'][b, :n], out['score'][b, :n], out['box'][b, :n, :4]
        same = cls[1:] == cls[:-1]
        assert bool((cls[1:] >= cls[:-1]).all())
        assert bool((sc[1:][same] <= sc[:-1][same]).all())
        if n and b % 8 == 0:
            xy = ops.iou_aabb(bx.contiguous(), bx.contiguous())
            same_cls = cls[:, None] == cls[None, :]
            off = ~torch.eye(n, dtype=torch.bool, device=xy.device)
            assert float((xy * (same_cls & off)).max()) <= thr + 1e-6


# ---------------------------------------------------------------------------------------------- config 2
@pytest.mark.parametrize('conf_mu', [-4.0, 2.0])
def test_cfg2_d1_fcos2_batch64(conf_mu):
    """EfficientDet-D1 + FCOS2, batch 64 @640: 'trained-like' (mu=-4) and 'all-pass' (mu=+2) operating points."""
    from mydetection_b200 import ops, pipeline as pl
    from oracle import decode as od, postprocess as opp
    strides, img, n_cls = (8, 16, 32, 64, 128), 640, 80
    gen = torch.Generator().manual_seed(1002)
    raws = fcos_batch(gen, 64, img, strides, n_cls, conf_mu)
    pipe = pl.DetectionPipeline('FCOS2', strides, n_cls, (img, img), 0.005, 0.5, 512)
    bound = pipe.bind(to_dev(raws))
    out1 = {k: v.clone() for k, v in bound.launch().items()}
    out2 = bound.launch()
    torch.cuda.synchronize()
    for k in out1:                                                  # deterministic run to run
        assert torch.equal(out1[k], out2[k]), k
    assert int(out1['status'].abs().sum()) == 0
    # compaction == dense threshold, for every image
    box, cls, score = ops.decode_dense(ops.KIND_FCOS, bound.levels, (img, img))
    cand = bound.launch_decode()
    torch.cuda.synchronize()
    assert torch.equal(cand['count'].long(), (score >= 0.005).sum(dim=1))
    if conf_mu > 0:
        assert int(cand['count'].min()) == 8525                     # all-pass: every location is a candidate
    check_output_invariants(out1, 0.5)
    # oracle, four images end to end; every image: the GPU's own dense decode through the oracle's
    # post-process must reproduce the fused path exactly (stage-wise parity)
    for b in (0, 21, 42, 63):
        sub = [{k: v[b:b + 1] for k, v in r.items()} for r in raws]
        ref = od.merge_levels([od.decode_fcos(r, s, (img, img)) for r, s in zip(sub, strides)])
        close(box[b], ref[0][0], img, 'cfg2 box', cancel=FCOS_W)
        close(score[b], ref[2][0], 1, 'cfg2 score')
        cls_match(cls[b], torch.cat([r['class'].reshape(1, -1, n_cls) for r in sub], 1), ref[1][0], 'cfg2 cls')
    box_c, cls_c, score_c = box.cpu(), cls.cpu(), score.cpu()
    for b in range(0, 64, 7):
        want = opp.post_process(box_c[b], cls_c[b], score_c[b], 0.005, 0.5, 'cxcywh', 512)
        n = int(out1['count'][b])
        assert n == want.numel() and torch.equal(out1['idx'][b, :n].cpu().long(), want)
        assert torch.equal(out1['box'][b, :n].cpu(), box_c[b][want])
    # idempotence: NMS of the survivors keeps all of them
    again = ops.postprocess(out1['box'], out1['score'], out1['cls'], float('-inf'), 0.5, topk=512, counts=out1['count'])
    torch.cuda.synchronize()
    assert torch.equal(again['count'], out1['count'])


# ---------------------------------------------------------------------------------------------- config 1
def test_cfg1_yolov3_608():
    """YOLOv3-80 head geometry at 608x608 (22 743 candidates, 255 channels, 19x19 level on the scalar path)."""
    from mydetection_b200 import ops
    from mydetection_b200.heads import yolo_head_views
    from oracle import decode as od, postprocess as opp
    anchors = [[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198], [373, 326]]
    gen = torch.Generator().manual_seed(1001)
    raws, refs = [], []
    for li, s in enumerate((8, 16, 32)):
        n = 608 // s
        t = torch.randn(1, 255, n, n, generator=gen)
        t.view(1, 3, 85, n, n)[:, :, 5:] -= 2.0
        raw = yolo_head_views(t, 3, 4, 80)
        raws.append(raw)
        refs.append(od.decode_yolo(raw, torch.tensor(anchors[3 * li:3 * li + 3], dtype=torch.float32), s, 80))
    ref = od.merge_levels(refs)
    ls = ops.LevelSet(to_dev(raws), (8, 16, 32), [anchors[0:3], anchors[3:6], anchors[6:9]])
    assert ls.n_total == 22743
    box, cls, score = ops.decode_dense(ops.KIND_YOLO, ls, (608, 608))
    close(box, ref[0], 608, 'cfg1 box')
    close(score, ref[2], 1, 'cfg1 score')
    cls_match(cls, torch.cat([r['class'].reshape(1, -1, 80) for r in raws], 1), ref[1], 'cfg1 cls')
    out = ops.detect(ops.KIND_YOLO, ls, (608, 608), 0.005, 0.45, topk=512)
    torch.cuda.synchronize()
    want = opp.post_process(box[0].cpu(), cls[0].cpu(), score[0].cpu(), 0.005, 0.45, 'cxcywh', 512)
    n = int(out['count'][0])
    assert n == want.numel() and torch.equal(out['idx'][0, :n].cpu().long(), want)


# ---------------------------------------------------------------------------------------------- config 3
def test_cfg3_rapid_batch32_1024():
    """RAPiD @1024, batch 32: xywha decode, post_process as HEAD does it (AABB, angle ignored)
    and true rotated NMS (nms_rotbb semantics) on the 10 000 best boxes of an image."""
    from mydetection_b200 import ops, pipeline as pl
    from mydetection_b200.heads import yolo_head_views
    from oracle import decode as od, postprocess as opp, iou as oi
    anchors = [[18.7807, 33.4659], [28.8912, 61.7536], [48.6849, 68.3897], [45.0668, 101.4673], [63.0952, 113.5382],
               [81.3909, 134.4554], [91.7364, 144.9949], [137.5189, 178.4791], [194.4429, 250.7985]]
    gen = torch.Generator().manual_seed(1003)
    B = 32
    raws = []
    for s in (8, 16, 32):
        n = 1024 // s
        t = torch.randn(B, 18, n, n, generator=gen) * 0.5
        v = t.view(B, 3, 6, n, n)
        v[:, :, 4] = torch.rand(B, 3, n, n, generator=gen) * 6 - 3
        v[:, :, 5] = torch.randn(B, 3, n, n, generator=gen) * 1.5 - 1.5
        raws.append(yolo_head_views(t, 3, 5, 0))
    groups = [anchors[0:3], anchors[3:6], anchors[6:9]]
    ls = ops.LevelSet(to_dev(raws), (8, 16, 32), groups)
    assert ls.n_total == 64512
    box, cls, score = ops.decode_dense(ops.KIND_RAPID, ls, (1024, 1024))
    for b in (0, 31):
        sub = [{k: v[b:b + 1] for k, v in r.items()} for r in raws]
        ref = od.merge_levels([od.decode_rapid(r, torch.tensor(a, dtype=torch.float32), s, 0)
                               for r, a, s in zip(sub, groups, (8, 16, 32))])
        close(box[b, :, :4], ref[0][0, :, :4], 1024, 'cfg3 box')
        close(box[b, :, 4], ref[0][0, :, 4], 180, 'cfg3 angle', cancel=ANGLE)
        close(score[b], ref[2][0], 1, 'cfg3 score')
    assert int(cls.abs().sum()) == 0
    # post_process parity (threshold chosen so that ~10k candidates pass per image)
    thr = float(score[0].kthvalue(64512 - 10000).values)
    pipe = pl.DetectionPipeline('RAPiD', (8, 16, 32), 0, (1024, 1024), thr, 0.45, 512, anchors=groups)
    out = pipe(to_dev(raws))
    torch.cuda.synchronize()
    box_c, cls_c, score_c = box.cpu(), cls.cpu(), score.cpu()
    for b in (0, 13, 31):
        want = opp.post_process(box_c[b], cls_c[b], score_c[b], thr, 0.45, 'cxcywhd', 512)
        n = int(out['count'][b])
        assert n == want.numel() and torch.equal(out['idx'][b, :n].cpu().long(), want)
        assert torch.equal(out['box'][b, :n].cpu(), box_c[b][want])           # 5 columns, angle carried through
    # rotated NMS on the 10 000 best boxes per image, 4 images, vs the oracle
    top = score.topk(10000, dim=1).indices[:4]
    rb = torch.gather(box[:4], 1, top[..., None].expand(-1, -1, 5)).contiguous()
    rs = torch.gather(score[:4], 1, top).contiguous()
    keep, cnt = ops.nms_rot(rb, rs, 0.45)
    torch.cuda.synchronize()
    for b in range(4):
        want = oi.nms_rot(rb[b].cpu(), rs[b].cpu(), 0.45)
        assert int(cnt[b]) == want.numel() and torch.equal(keep[b, :int(cnt[b])].cpu(), want)


# ---------------------------------------------------------------------------------------------- config 4
def test_cfg4_atss_100gt():
    """D1 + FCOS2 + ATSS at 640x640 with 100 GT per image: assignment maps of every level vs the oracle
    (2 images), plus the pairwise anchor-GT IoU matrix (8 525 x 100) bit-exactly."""
    from mydetection_b200 import ops
    from oracle import atss as oa, iou as oi
    strides, sides, img = [8, 16, 32, 64, 128], [24, 48, 96, 192, 384], (640, 640)
    gen = torch.Generator().manual_seed(1004)
    B, G = 2, 100
    gt_box = torch.empty(B, G, 4)
    gt_box[..., 0:2] = torch.rand(B, G, 2, generator=gen) * 600 + 20
    gt_box[..., 2:4] = torch.rand(B, G, 2, generator=gen) * 200 + 16
    gt_cls = torch.randint(0, 80, (B, G), generator=gen)
    counts = torch.tensor([100, 63], dtype=torch.int32)
    gts = [(gt_box[b, :int(counts[b])], gt_cls[b, :int(counts[b])]) for b in range(B)]
    anchors = torch.cat(oa.all_level_anchors(img, strides, sides))
    assert anchors.shape[0] == 8525
    got = ops.iou_aabb(anchors.to(DEV), gt_box[0].to(DEV))
    assert torch.equal(got.cpu(), oi.bboxes_iou(anchors, gt_box[0]))
    flips = 0
    for li, s in enumerate(strides):
        n = 640 // s
        t = (torch.randn(B, 4, n, n, generator=gen) * 0.5).permute(0, 2, 3, 1)
        want = oa.assign_level(li, t, gts, img, strides, sides, 9, 0.7, 80)
        out = ops.atss_assign(t.to(DEV), li, strides, sides, img, gt_box.to(DEV), gt_cls.to(DEV), counts.to(DEV), 9, 0.7, 80)
        torch.cuda.synchronize()
        # the adaptive threshold is mean + std of 45 float32 IoUs; the reference's summation order is
        # torch's, so a cell may flip only if an IoU lies within 1e-6 of its GT's threshold
        for k in ('PositiveMask', 'IgnoredMask'):
            flips += int((out[k].cpu() != want[k]).sum())
        same = out['PositiveMask'].cpu() == want['PositiveMask']
        close(out['TargetLTRB'].cpu()[same], want['TargetLTRB'][same], 640, f'cfg4 ltrb L{li}')
    assert flips <= 2, flips


def test_atss_threshold_window_edges():
    """The k nearest anchors are searched in an 11x11 window around the GT centre (exhaustively when the centre is
    outside the image): GT centres on corners, edges, cell boundaries, half cells and outside the image must give
    the oracle's adaptive threshold (torch.topk over ALL anchors of every level)."""
    from mydetection_b200 import ops
    from oracle import atss as oa
    strides, sides, img = [8, 16, 32, 64, 128], [24, 48, 96, 192, 384], (384, 640)
    pts = [(0, 0), (640, 384), (0, 384), (640, 0), (320, 0), (0, 192), (4, 4), (8, 8), (12, 12), (636, 380), (64, 64), (128, 256),
           (63.999, 64.001), (320, 192), (-30, 100), (700, 200), (100, -5), (300, 500), (-1e-3, 50), (639.5, 383.5), (3.9, 380.1),
           (37.3, 211.7), (501.2, 17.9), (1.1, 2.3), (638.7, 1.4), (333.3, 383.9), (-7.7, 391.3)]
    gt = torch.tensor([[x, y, 90.0 + 7 * i, 60.0 + 5 * i] for i, (x, y) in enumerate(pts)], dtype=torch.float32)[None]
    cls = torch.zeros(1, len(pts), dtype=torch.int64)
    counts = torch.tensor([len(pts)], dtype=torch.int32)
    t = torch.zeros(1, 48, 80, 4)
    out = ops.atss_assign(t.to(DEV), 0, strides, sides, img, gt.to(DEV), cls.to(DEV), counts.to(DEV), 9, 0.7, 3)
    torch.cuda.synchronize()
    anchors = oa.all_level_anchors(img, strides, sides)
    n_ref = 0
    for i in range(len(pts)):
        got = float(out['thr'][0, i])
        # declared tie policy (ascending index among equidistant anchors): always
        want = float(oa.atss_threshold_index_ties(gt[0, i], anchors, 9))
        assert abs(got - want) <= 2e-6 * max(1.0, abs(want)), (pts[i], got, want)
        # the reference's torch.topk: wherever the k-th distance is not tied (centres off the cell boundaries)
        ref = float(oa.atss_threshold(gt[0, i], anchors, 9))
        on_boundary = any(pts[i][0] % (s / 2) == 0 or pts[i][1] % (s / 2) == 0 for s in strides)
        if not on_boundary:
            n_ref += 1
            assert abs(got - ref) <= 2e-6 * max(1.0, abs(ref)), (pts[i], got, ref)
    assert n_ref >= 8


# ---------------------------------------------------------------------------------------------- config 5
@pytest.mark.parametrize('img,batch', [(704, 6), (1024, 3), (1536, 2)])
def test_cfg5_dense_scene_uncapped(img, batch):
    """Dense scenes: single class, every location a candidate (10 164 / 21 504 / 48 384 per image), no
    top-k cap -> the tiled large-N path end to end, against the oracle on the GPU's own dense decode."""
    from mydetection_b200 import ops
    from mydetection_b200.heads import yolo_head_views
    from oracle import postprocess as opp
    gen = torch.Generator().manual_seed(1005 + img)
    raws = []
    for s in (8, 16, 32):
        n = img // s
        t = torch.randn(batch, 6, n, n, generator=gen) * 0.5
        t[:, 4] = torch.randn(batch, n, n, generator=gen) * 1.5 + 2.0
        raws.append({k: v[:, 0] for k, v in yolo_head_views(t, 1, 4, 1).items()})     # nA == 1 -> (B,nH,nW,C) views
    ls = ops.LevelSet(to_dev(raws), (8, 16, 32))
    assert ls.n_total == sum((img // s) ** 2 for s in (8, 16, 32))
    box, cls, score = ops.decode_dense(ops.KIND_FCOS, ls, (img, img))
    out = ops.detect(ops.KIND_FCOS, ls, (img, img), 0.005, 0.45, topk=None)
    torch.cuda.synchronize()
    assert int(out['status'].abs().sum()) == 0
    b = batch - 1
    want = opp.post_process(box[b].cpu(), cls[b].cpu(), score[b].cpu(), 0.005, 0.45, 'cxcywh', None)
    n = int(out['count'][b])
    assert n == want.numel() and torch.equal(out['idx'][b, :n].cpu().long(), want)
    for bb in range(batch):                                         # descending scores, single class
        k = int(out['count'][bb])
        s = out['score'][bb, :k]
        assert bool((s[1:] <= s[:-1]).all()) and k > 1000


# ---------------------------------------------------------------------------------------------- repeatability
def test_large_paths_repeatable():
    """compute-sanitizer is closed on this pool, so races are hunted the blunt way: the tiled large-N paths
    (sort, mask queues with shared-memory atomics, sweep with atomicOr) must give bit-identical results over
    many repetitions on the same input."""
    from mydetection_b200 import ops
    gen = torch.Generator().manual_seed(77)
    B, n = 4, 6000
    xy = torch.rand(B, n, 2, generator=gen) * 700 + 50
    wh = torch.rand(B, n, 2, generator=gen) * 90 + 8
    ang = torch.rand(B, n, 1, generator=gen) * 360 - 180
    rb = torch.cat([xy, wh, ang], 2).to(DEV)
    scores = torch.rand(B, n, generator=gen).to(DEV)
    cls = torch.randint(0, 3, (B, n), generator=gen).to(DEV)
    first = None
    for rep in range(12):
        keep, cnt, votes = ops.nms_rot(rb, scores, 0.4, want_votes=True)
        out = ops.postprocess(rb[..., :4].contiguous(), scores, cls, 0.05, 0.5, topk=None)
        torch.cuda.synchronize()
        snap = [cnt.clone(), out['count'].clone()]
        for b in range(B):
            snap += [keep[b, :int(cnt[b])].clone(), votes[b, :int(cnt[b])].clone(), out['idx'][b, :int(out['count'][b])].clone()]
        if first is None:
            first = snap
        else:
            assert all(torch.equal(a, c) for a, c in zip(first, snap)), f'repetition {rep} differs'
