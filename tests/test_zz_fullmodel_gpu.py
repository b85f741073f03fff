"""GPU parity against the UNMODIFIED reference run end to end on real network forwards (tests/golden/fullmodel_*.npz:
OneStageBBox of yolov3_80 / rapid / d1_fcs2 with random weights, one synthetic image, det layers, level concatenation,
post_process -- BASELINE configs[0], [2], [1] geometry at 256 x 256).

STATUS: green on a B200 in the round-1 driver run (GPUTEST_r01.json); hard tests since round 2.  The generator picks image
seeds whose rankings keep >= 5e-5 (score at the top-512 boundary and at the confidence threshold) and >= 1e-4 (IoU to the
NMS threshold) of margin, far above the float32 tolerances of DESIGN.md section 4, so kept indices must be exact.
"""
import numpy as np
import pytest
import torch

from helpers import T, YOLO_ANCHORS, RAPID_ANCHORS

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize('name', ['yolov3_80', 'rapid', 'd1_fcs2'])
def test_fullmodel_against_reference(golden, name):
    from mydetection_b200 import ops
    from mydetection_b200.heads import yolo_head_views, efdet_head_views
    g = golden('fullmodel_' + name)
    conf, nms, img_h, img_w = (float(v) for v in g['params'])
    dev = torch.device('cuda', 0)
    on = lambda raw: {k: v.to(dev) for k, v in raw.items()}
    if name == 'd1_fcs2':
        strides = (8, 16, 32, 64, 128)
        raws = [on(efdet_head_views(T(g[f'head{li}_0']).float().to(dev), T(g[f'head{li}_1']).float().to(dev))) for li in range(5)]
        kind, anchors = ops.KIND_FCOS, None
    else:
        strides = (8, 16, 32)
        table, n_p, n_c, kind = ((RAPID_ANCHORS, 5, 0, ops.KIND_RAPID) if name == 'rapid' else (YOLO_ANCHORS, 4, 80, ops.KIND_YOLO))
        raws = [on(yolo_head_views(T(g[f'head{li}_0']).float().to(dev), 3, n_p, n_c)) for li in range(3)]
        anchors = [table[0:3], table[3:6], table[6:9]]
    ls = ops.LevelSet(raws, strides, anchors)
    out = ops.detect(kind, ls, (img_h, img_w), conf, nms, topk=512)
    torch.cuda.synchronize()
    n = int(out['count'][0])
    keep = T(g['keep'])
    assert n == keep.numel() and torch.equal(out['idx'][0, :n].cpu().long(), keep)
    assert torch.equal(out['cls'][0, :n].cpu(), T(g['kept_cats']))
    assert torch.allclose(out['score'][0, :n].cpu(), T(g['kept_scores']), rtol=1e-5, atol=0)
    assert torch.allclose(out['box'][0, :n].cpu(), T(g['kept_boxes']), rtol=1e-5, atol=2 * float(np.spacing(np.float32(256))))


def test_fullmodel_atss_targets_against_reference(golden):
    """BASELINE configs[3]: mydet_atss_assign on the regression head of a real forward, 100 GT boxes, against the target
    tensors of the unmodified reference in training mode (fullmodel_d1_fcs2_atss.npz).  A cell may flip only where an IoU
    lies within 1e-6 of its adaptive threshold: at most 2 cells over the five levels."""
    from mydetection_b200 import ops
    g = golden('fullmodel_d1_fcs2_atss')
    img_h, img_w, topk, ign, n_cls = g['params']
    dev = torch.device('cuda', 0)
    gt_box, gt_cls = T(g['gt_boxes'])[None].to(dev), T(g['gt_cats'])[None].to(dev)
    cnt = torch.tensor([gt_box.shape[1]], dtype=torch.int32, device=dev)
    strides, sides = [int(v) for v in g['strides']], [float(v) for v in g['anchors']]
    flips, thr = 0, None
    for li in range(5):
        t = T(g[f'atss{li}_bbox']).float().to(dev).permute(0, 2, 3, 1)
        out = ops.atss_assign(t, li, strides, sides, (int(img_h), int(img_w)), gt_box, gt_cls, cnt, int(topk), float(ign),
                              int(n_cls), thr=thr)
        thr = out['thr']                                   # level independent: computed once, reused (include/mydet.h)
        torch.cuda.synchronize()
        pos, want_pos = out['PositiveMask'].cpu(), T(g[f'atss{li}_PositiveMask'])
        same = pos == want_pos
        flips += int((~same).sum()) + int((out['IgnoredMask'].cpu() != T(g[f'atss{li}_IgnoredMask'])).sum())
        assert torch.equal(out['TargetCls'].cpu()[same], T(g[f'atss{li}_TargetCls'])[same]), li
        assert torch.equal(out['TargetConf'].cpu()[same], T(g[f'atss{li}_TargetConf'])[same]), li
        assert torch.allclose(out['TargetLTRB'].cpu()[same], T(g[f'atss{li}_TargetLTRB'])[same], rtol=1e-5, atol=1e-4), li
    assert flips <= 2, flips
