"""GPU parity against the UNMODIFIED reference run end to end on real network forwards (tests/golden/fullmodel_*.npz:
OneStageBBox of yolov3_80 / rapid / d1_fcs2 with random weights, one synthetic image, det layers, level concatenation,
post_process -- BASELINE configs[0], [2], [1] geometry at 256 x 256).

STATUS: the fixtures were generated after round 1's GPU budget was spent; the kernels they exercise are the ones
tests/test_gpu_configs.py already verifies on synthetic logits, but THESE inputs have not been on a B200 yet, so the
tests are non-strict xfail until their first run (the file sorts after the verified tests).  The generator picks image
seeds whose rankings keep >= 5e-5 (score at the top-512 boundary and at the confidence threshold) and >= 1e-4 (IoU to the
NMS threshold) of margin, far above the float32 tolerances of DESIGN.md section 4, so kept indices must be exact.
"""
import numpy as np
import pytest
import torch

from helpers import T, YOLO_ANCHORS, RAPID_ANCHORS

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason='first GPU run pending (fixtures added after the round-1 GPU budget was spent)')]


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
