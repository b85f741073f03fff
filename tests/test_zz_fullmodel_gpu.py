"""GPU parity on BASELINE configs[0] against the UNMODIFIED reference run end to end (tests/golden/fullmodel.npz:
OneStageBBox(yolov3_80), random-init Darknet-53 + FPN + head, one synthetic image, post_process).

STATUS: the fixture was generated after round 1's GPU budget was spent; the kernels it exercises are the ones
tests/test_gpu_configs.py::test_cfg1_yolov3_608 already verifies on synthetic logits, but THIS input has not been on a
B200 yet, so the test is non-strict xfail until its first run (the file sorts after the verified tests).  The fixture's
ranking margins (score gap at the top-512 boundary 2e-4, nearest IoU to the NMS threshold 4e-4) are far above the float32
tolerances of DESIGN.md section 4, so kept indices must be exact.
"""
import numpy as np
import pytest
import torch

from helpers import T, YOLO_ANCHORS

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason='first GPU run pending (fixture added after the round-1 GPU budget was spent)')]


def test_fullmodel_yolov3_against_reference(golden):
    from mydetection_b200 import ops
    from mydetection_b200.heads import yolo_head_views
    g = golden('fullmodel')
    conf, nms, img_h, img_w = (float(v) for v in g['params'])
    dev = torch.device('cuda', 0)
    raws = [{k: v.to(dev) for k, v in yolo_head_views(T(g[f'head{li}_f16']).float(), 3, 4, 80).items()} for li in range(3)]
    ls = ops.LevelSet(raws, (8, 16, 32), [YOLO_ANCHORS[0:3], YOLO_ANCHORS[3:6], YOLO_ANCHORS[6:9]])
    out = ops.detect(ops.KIND_YOLO, ls, (img_h, img_w), conf, nms, topk=512)
    torch.cuda.synchronize()
    n = int(out['count'][0])
    keep = T(g['keep'])
    assert n == keep.numel() and torch.equal(out['idx'][0, :n].cpu().long(), keep)
    assert torch.equal(out['cls'][0, :n].cpu(), T(g['kept_cats']))
    assert torch.allclose(out['score'][0, :n].cpu(), T(g['kept_scores']), rtol=1e-5, atol=0)
    assert torch.allclose(out['box'][0, :n].cpu(), T(g['kept_boxes']), rtol=1e-5, atol=2 * np.spacing(np.float32(256)))
