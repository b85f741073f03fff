"""The UNMODIFIED reference running on top of the drop-in, on the B200.

`oracle/_ref/` is a byte-for-byte copy of the reference's api/, models/, utils/, configs/, external/ made by the
committed recipe oracle/fetch_ref.py (git-ignored, ships with the gpurun snapshot; /root/reference is never read here).
After `mydetection_b200.dropin.install()`:
  * `models/general.py::OneStageBBox.__init__` (:28-42) instantiates the mirror det layers through the reference's own
    `models/registry.py::get_det_layer`, and `OneStageBBox.forward` (:44-97) -- backbone, FPN, head, the per-level
    `det_layers[i](raw, img_size, labels)` loop, the three `torch.cat`s and the per-image `ImageObjects(...)` -- runs
    unchanged with CUDA kernels underneath;
  * `api/detection.py::Detector.detect_one / _predict_pil` (:113-175) runs unchanged, including `return_img=True`
    (`ImageObjects.draw_on_np`).
Two kinds of checks:
  1. against the committed fixtures (tests/golden/fullmodel_*.npz: kept boxes of the unmodified reference on the CPU):
     the fixture's head outputs are injected with a forward hook on `model.rpn`, everything after it is the unmodified flow;
  2. A/B on the same box: the reference alone (its own det layers + post_process with torchvision NMS) against the
     reference + drop-in on the head outputs of one real GPU forward (random weights, per-channel standardised logits).
"""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import T

pytestmark = [pytest.mark.gpu]
NAMES = ['yolov3_80', 'rapid', 'd1_fcs2']


@pytest.fixture
def ref():
    from oracle import refload
    if not refload.available():
        pytest.skip('oracle/_ref absent: run `python -m oracle.fetch_ref` where /root/reference exists (build() does)')
    saved_path = list(sys.path)
    refload.activate(fresh=True)
    yield refload
    refload.deactivate()
    sys.path[:] = saved_path


def _nchw_of(raw):
    """The NCHW tensors behind the permuted views of models/rpns.py:29-41 (YOLOHead) / :175-189 (EfDetHead)."""
    if raw['bbox'].dim() == 5:
        t = torch.cat([raw['bbox'], raw['conf'], raw['class']], dim=-1).permute(0, 1, 4, 2, 3)
        return [t.reshape(t.shape[0], -1, t.shape[3], t.shape[4]).contiguous()]
    return [raw['bbox'].permute(0, 3, 1, 2).contiguous(),
            torch.cat([raw['conf'], raw['class']], dim=-1).permute(0, 3, 1, 2).contiguous()]


def _views_of(tensors, n_param, n_cls):
    if len(tensors) == 1:
        t = tensors[0]
        v = t.view(t.shape[0], 3, n_param + 1 + n_cls, t.shape[2], t.shape[3])
        return {'bbox': v[:, :, 0:n_param].permute(0, 1, 3, 4, 2), 'conf': v[:, :, n_param:n_param + 1].permute(0, 1, 3, 4, 2),
                'class': v[:, :, n_param + 1:].permute(0, 1, 3, 4, 2)}
    bb, cc = tensors
    c = cc.permute(0, 2, 3, 1)
    return {'bbox': bb.permute(0, 2, 3, 1), 'conf': c[..., 0:1], 'class': c[..., 1:]}


def _fixture_raws(g, name, dev):
    n_p, n_c = (5, 0) if name == 'rapid' else (4, 80)
    raws, li = [], 0
    while f'head{li}_0' in g:
        ts = [T(g[f'head{li}_{j}']).float().to(dev) for j in range(2) if f'head{li}_{j}' in g]
        raws.append(_views_of(ts, n_p, n_c))
        li += 1
    return raws


def _same_objects(got, g, img=256.0):
    assert len(got) == len(g['keep'])
    assert torch.equal(got.cats.cpu(), T(g['kept_cats']))
    assert torch.allclose(got.scores.cpu(), T(g['kept_scores']), rtol=1e-5, atol=0)
    assert torch.allclose(got.bboxes.cpu(), T(g['kept_boxes']), rtol=1e-5, atol=2 * float(np.spacing(np.float32(img))))


@pytest.mark.parametrize('name', NAMES)
def test_unmodified_forward_and_detector_on_dropin(ref, golden, name):
    """models/general.py:44-97 and api/detection.py:113-175 of the reference, unedited, on the mirror layers; the head
    outputs are the fixture's (forward hook on model.rpn), so the result must be the reference's own CPU result."""
    from mydetection_b200 import dropin
    dropin.install()
    dev = torch.device('cuda', 0)
    model, cfg = ref.build_model(name, device=dev)
    import models.general as general
    import utils.structures as structures
    assert general.__file__.startswith(ref.ROOT)                         # the reference's own file ...
    assert general.ImageObjects is structures.ImageObjects and structures.__name__ == 'mydetection_b200.structures'
    assert all(type(l).__module__.startswith('mydetection_b200.detlayers') for l in model.det_layers)   # ... on the mirror
    g = golden('fullmodel_' + name)
    conf, nms, img_h, img_w = (float(v) for v in g['params'])
    raws = _fixture_raws(g, name, dev)
    model.rpn.register_forward_hook(lambda mod, inp, out: raws)
    x = torch.rand(1, 3, int(img_h), int(img_w), device=dev)
    with torch.no_grad():
        dts = model(x)                                                    # OneStageBBox.forward, unchanged
    assert isinstance(dts, list) and len(dts) == 1 and isinstance(dts[0], structures.ImageObjects)
    _same_objects(dts[0].post_process(conf, nms), g)                      # api/detection.py:172

    from api.detection import Detector
    import PIL.Image
    det = Detector(model_and_cfg=(model, cfg))
    assert det.on_cpu is False
    img = PIL.Image.fromarray((np.random.RandomState(0).rand(int(img_h), int(img_w), 3) * 255).astype(np.uint8))
    out = det.detect_one(pil_img=img, input_size=int(img_h), conf_thres=conf, nms_thres=nms)
    _same_objects(out, g)                  # resize to the same size + bboxes_to_original_ with unit scale: same boxes
    drawn = det.detect_one(pil_img=img, input_size=int(img_h), conf_thres=conf, nms_thres=nms, return_img=True)
    assert isinstance(drawn, np.ndarray) and drawn.shape == (int(img_h), int(img_w), 3)
    assert (drawn != np.array(img)).any()                                 # ImageObjects.draw_on_np drew the boxes
    js = out.to_json(img_id=7, eval_type='cxcywhd' if name == 'rapid' else 'x1y1wh')
    assert len(js) == len(out) and js[0]['image_id'] == 7


def _margins_ok(d, conf, nms, bboxes_iou):
    sc, cls, bbs = d.scores.cpu(), d.cats.cpu(), d.bboxes.cpu()
    srt = sc.sort(descending=True).values
    if srt[:513].unique().numel() != min(513, srt.numel()):
        return False
    top = sc.argsort(descending=True)[:512]
    iou = bboxes_iou(bbs[top][:, :4], bbs[top][:, :4])
    same = cls[top][:, None] == cls[top][None, :]
    return bool(srt[511] - srt[512] > 5e-5 and (sc - conf).abs().min() > 5e-5 and (iou[same] - nms).abs().min() > 1e-4)


@pytest.mark.parametrize('name', NAMES)
def test_reference_alone_vs_reference_on_dropin_same_box(ref, name):
    """A/B on the B200: OneStageBBox of the unmodified reference (its own det layers, ImageObjects.post_process with
    torchvision NMS on the CPU) against the same unmodified OneStageBBox after dropin.install(), on the head outputs of
    one real forward of the random-weight network on the GPU (batch 2 @320 / @384; logits standardised per channel to sigma 1.5
    to leave the exact-tie regime of a random-init network).  The image seed is the first whose rankings keep
    the margins the fixtures use (>= 5e-5 on scores, >= 1e-4 on IoUs), so kept sets must be identical."""
    dev = torch.device('cuda', 0)
    model, cfg = ref.build_model(name, device=dev)
    from utils.bbox_ops import bboxes_iou
    assert type(model.det_layers[0]).__module__.startswith('models.detlayers')        # the reference's own layers
    conf, nms = cfg['test.ap_conf_thres'], cfg['test.nms_thres']
    n_p, n_c = cfg['general.bbox_param'], cfg['general.num_class']
    captured = {}
    size = 384 if name == 'd1_fcs2' else 320          # d1_fcs2: general.input_divisibility = 128
    dither = torch.Generator().manual_seed(99)

    def standardise(mod, inp, out):
        # a random-init network repeats feature vectors at different positions (exact duplicates among the logits):
        # a 1e-3 dither on top of the standardisation makes the ranking of the reference itself well defined
        raws = []
        for raw in out:
            ts = []
            for t in _nchw_of(raw):
                t = (t - t.mean(dim=(0, 2, 3), keepdim=True)) / t.std(dim=(0, 2, 3), keepdim=True) * 1.5
                ts.append((t + 1e-3 * torch.rand(t.shape, generator=dither).to(t.device)).contiguous())
            raws.append(_views_of(ts, n_p, n_c))
        captured['raws'] = raws
        return raws
    model.rpn.register_forward_hook(standardise)
    for seed in range(200):
        x = torch.rand(2, 3, size, size, generator=torch.Generator().manual_seed(seed)).to(dev)
        with torch.no_grad():
            dts_ref = model(x)
        if all(_margins_ok(d, conf, nms, bboxes_iou) for d in dts_ref):
            break
    else:
        pytest.fail('no seed with safe margins')
    raws = captured['raws']
    dense_ref = [(d.bboxes.cpu().clone(), d.cats.cpu().clone(), d.scores.cpu().clone()) for d in dts_ref]
    res_ref = [d.post_process(conf, nms) for d in dts_ref]

    ref.purge()                                                           # fresh import of models.* on top of the drop-in
    from mydetection_b200 import dropin
    dropin.install()
    model2, _ = ref.build_model(name, device=dev)
    assert type(model2.det_layers[0]).__module__.startswith('mydetection_b200.detlayers')
    model2.rpn.register_forward_hook(lambda mod, inp, out: raws)
    with torch.no_grad():
        dts_new = model2(x)
    assert len(dts_new) == 2
    for b in range(2):
        bb, cc, ss = dense_ref[b]
        atol = 2 * float(np.spacing(np.float32(size)))
        assert torch.allclose(dts_new[b].bboxes.cpu(), bb, rtol=1e-5, atol=atol)
        assert torch.allclose(dts_new[b].scores.cpu(), ss, rtol=1e-5, atol=0)
        assert torch.equal(dts_new[b].cats.cpu(), cc)
        got, want = dts_new[b].post_process(conf, nms), res_ref[b]
        assert len(got) == len(want) and len(want) > 0
        assert torch.equal(got.cats, want.cats)
        assert torch.allclose(got.scores, want.scores, rtol=1e-5, atol=0)
        assert torch.allclose(got.bboxes, want.bboxes, rtol=1e-5, atol=atol)


def test_cepdof_evaluator_on_dropin(ref, monkeypatch):
    """utils/evaluation/cepdof.py of the reference, unedited, under dropin.install(): `CEPDOFeval(...).evaluate()` builds
    its IoU table through the patched computeIoU -- ONE mydet_iou_rot_segments launch for all (image, category) pairs
    -- and the table equals what the unpatched evaluator computes pair by pair (its raster replaced by the oracle's
    exact clipping, refload's stand-in): same keys, same shapes, same empties, values within 1e-9 (float64 throughout)."""
    import copy
    rng = np.random.RandomState(3)
    images = [{'id': i, 'height': 1024, 'width': 1024} for i in range(1, 13)]
    cats = [{'id': 1}, {'id': 2}, {'id': 5}]
    anns, dts = [], []
    for im in images:
        for c in cats:
            n_gt = int(rng.randint(0, 7)) if im['id'] != 4 else 0
            for _ in range(n_gt):
                box = [*rng.uniform(100, 900, 2), *rng.uniform(20, 160, 2), rng.uniform(-90, 90)]
                anns.append({'image_id': im['id'], 'category_id': c['id'], 'bbox': [float(v) for v in box]})
                for _ in range(int(rng.randint(0, 4))):       # detections near the GT, tied scores included
                    d = [box[0] + rng.randn() * 8, box[1] + rng.randn() * 8, box[2] * rng.uniform(0.8, 1.2),
                         box[3] * rng.uniform(0.8, 1.2), box[4] + rng.randn() * 12]
                    dts.append({'image_id': im['id'], 'category_id': c['id'], 'bbox': [float(v) for v in d],
                                'score': float(np.round(rng.uniform(0, 1), 1))})
    for _ in range(130):                                       # one crowded pair: more than maxDets detections
        dts.append({'image_id': 2, 'category_id': 1, 'bbox': [float(v) for v in (*rng.uniform(100, 900, 2), 60., 30., rng.uniform(-90, 90))],
                    'score': float(rng.uniform(0, 1))})
    gt_json = {'images': images, 'categories': cats, 'annotations': anns}

    import utils.evaluation.cepdof as cep_ref                  # the reference alone (stand-in raster = exact clipping)
    assert cep_ref.__file__.startswith(ref.ROOT)
    ev_ref = cep_ref.CEPDOFeval(copy.deepcopy(gt_json), copy.deepcopy(dts))
    ev_ref.evaluate()

    ref.activate(fresh=True)
    from mydetection_b200 import dropin, evaluation, ops
    dropin.install()
    import utils.evaluation.cepdof as cep
    assert cep.__file__.startswith(ref.ROOT) and cep.iou_rle is evaluation.iou_rle
    launches = []
    real = ops.iou_rot_segments
    monkeypatch.setattr(ops, 'iou_rot_segments', lambda *a, **k: (launches.append(1), real(*a, **k))[1])
    ev = cep.CEPDOFeval(copy.deepcopy(gt_json), copy.deepcopy(dts))
    ev.evaluate()
    assert launches == [1]
    assert set(ev.ious) == set(ev_ref.ious) and len(ev.ious) == 36
    worst = 0.0
    for key, want in ev_ref.ious.items():
        got = ev.ious[key]
        if isinstance(want, list):
            assert isinstance(got, list) and got == []
            continue
        assert got.shape == want.shape and got.dtype == np.float64, key
        if want.size:
            worst = max(worst, float(np.abs(got - want).max()))
    assert ev.ious[(2, 1)].shape[0] == 100                     # maxDets cap after the stable score sort
    assert worst < 1e-9, worst                                  # float64 boxes and corners on both sides
