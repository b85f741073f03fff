"""GPU parity: the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Bars (BASELINE.json north_star):
  * integer / index results (class_idx, kept indices, masks) bit-exact, except where a float
    that feeds a comparison lies inside the stated band of its threshold;
  * decoded coordinates and scores within 1e-5 relative in fp32, PURE relative for every YOLO / RAPiD / Retina /
    YOLOv5 box and every score.  Only values that are differences of larger numbers (FCOS w = x2 - x1, the
    angle sigmoid*360 - 180 near 0, YOLOv5's 2*sigmoid - 0.5 in the first row / column) additionally get an absolute
    floor of 2 ulp of the operands' magnitude, because a 1-ulp difference between CUDA expf and the CPU's exp is
    amplified by that cancellation (DESIGN.md "tolerances"); each such call names its cancellation.  The largest
    error seen per quantity is printed at the end of the session (conftest.py).
"""
import numpy as np
import pytest
import torch

from helpers import T, level_anchors, yolo_views, efdet_views, anchor_views, YOLO_ANCHORS, RAPID_ANCHORS

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def dev():
    return torch.device('cuda:0')


OBSERVED = {}      # what -> (max relative error over |ref| > 1e-30, max absolute error); printed by conftest at session end


def close(got, ref, extent, what, cancel=None):
    """|got - ref| <= 1e-5 * |ref|: the north star's bar, pure relative.  Only where the decoded value is a DIFFERENCE of
    larger numbers -- so that a last-bit difference between CUDA expf / sigmoid and the CPU's is amplified by
    cancellation -- an absolute floor of 2 ulp of `extent` (the magnitude of the operands) is added, and the call site
    says which cancellation (`cancel='...'`)."""
    got, ref = got.detach().cpu().double(), ref.double()
    atol = 2 * float(np.spacing(np.float32(extent))) if cancel else 0.0
    err = (got - ref).abs()
    nz = ref.abs() > 1e-30
    rel = float((err[nz] / ref.abs()[nz]).max()) if bool(nz.any()) else 0.0
    prev = OBSERVED.get(what, (0.0, 0.0))
    OBSERVED[what] = (max(prev[0], rel), max(prev[1], float(err.max()) if err.numel() else 0.0))
    bad = err > (RTOL * ref.abs() + atol)
    assert not bad.any(), (f'{what}: {int(bad.sum())} of {bad.numel()} outside tolerance, max abs err {float(err.max()):.3e}, '
                           f'max rel err {rel:.3e}')


FCOS_W = 'FCOS: w = x2 - x1, cx = (x1 + x2) / 2 of image-scale clamped corners'
ANGLE = 'angle = (sigmoid * 2 pi - pi) / pi * 180 (sigmoid * 360 - 180): cancels near 0 degrees'
UV5_XY = 'YOLOv5: (2 sigmoid - 0.5 + g) cancels near sigmoid = 0.25 in the first row / column'
CORNERS = 'corners = centre +- v -+ h with last-bit differences in sinf / cosf'


def cls_match(got_cls, raw_cls_logits, ref_cls, what):
    """class_idx must equal the reference's except where the two best class probabilities are
    within 1e-6 relative (float32 sigmoid ties)."""
    got_cls = got_cls.cpu()
    diff = got_cls != ref_cls
    if diff.any():
        p = torch.sigmoid(raw_cls_logits).reshape(-1, raw_cls_logits.shape[-1])
        top2 = p.topk(2, dim=-1).values
        band = (top2[:, 0] - top2[:, 1]) <= 1e-6 * top2[:, 0]
        assert bool(band[diff.reshape(-1)].all()), f'{what}: class_idx differs outside the tie band'


def run_dense(kind, raws, strides, anchors, img_hw, conf_key='conf'):
    from mydetection_b200 import ops
    d = dev()
    raws_d = [{k: v.to(d) for k, v in r.items()} for r in raws]
    # .to(d) of a permuted view keeps the strides: the kernel sees the reference's layout
    ls = ops.LevelSet(raws_d, strides, anchors, conf_key)
    out = ops.decode_dense(kind, ls, img_hw)
    torch.cuda.synchronize()
    return out


# ------------------------------------------------------------------------------------- decode
def test_decode_yolo_golden(golden):
    from mydetection_b200 import ops
    from oracle import decode as od
    g = golden('decode')
    raws, refs = [], []
    for li, s in enumerate((8, 16, 32)):
        raw = yolo_views(T(g[f'yolo{li}_in']), 3, 4, 5)
        raws.append(raw)
        refs.append((T(g[f'yolo{li}_bbox']), T(g[f'yolo{li}_cls']), T(g[f'yolo{li}_score'])))
    anchors = [level_anchors(YOLO_ANCHORS, li).tolist() for li in range(3)]
    box, cls, score = run_dense(ops.KIND_YOLO, raws, (8, 16, 32), anchors, (96, 128))
    rb, rc, rs = od.merge_levels(refs)
    close(box, rb, 128, 'yolo bbox')
    close(score, rs, 1, 'yolo score')
    logits = torch.cat([r['class'].reshape(2, -1, 5) for r in raws], 1)
    cls_match(cls, logits, rc, 'yolo cls')
    # zero-class variant, single level
    raw = yolo_views(T(g['yolo_c0_in']), 3, 4, 0)
    box, cls, score = run_dense(ops.KIND_YOLO, [raw], (16,), [level_anchors(YOLO_ANCHORS, 1).tolist()], (96, 128))
    close(box, T(g['yolo_c0_bbox']), 128, 'yolo c0 bbox')
    close(score, T(g['yolo_c0_score']), 1, 'yolo c0 score')
    assert int(cls.abs().sum()) == 0


def test_decode_rapid_golden(golden):
    from mydetection_b200 import ops
    from oracle import decode as od
    g = golden('decode')
    for tag, nc in (('rapid_c0', 0), ('rapid_c3', 3)):
        raws, refs = [], []
        for li in range(3):
            raws.append(yolo_views(T(g[f'{tag}_{li}_in']), 3, 5, nc))
            refs.append((T(g[f'{tag}_{li}_bbox']), T(g[f'{tag}_{li}_cls']), T(g[f'{tag}_{li}_score'])))
        anchors = [level_anchors(RAPID_ANCHORS, li).tolist() for li in range(3)]
        box, cls, score = run_dense(ops.KIND_RAPID, raws, (8, 16, 32), anchors, (96, 128))
        rb, rc, rs = od.merge_levels(refs)
        close(box[..., :4], rb[..., :4], 128, tag + ' bbox')
        close(box[..., 4], rb[..., 4], 180, tag + ' angle', cancel=ANGLE)
        close(score, rs, 1, tag + ' score')
        if nc:
            cls_match(cls, torch.cat([r['class'].reshape(2, -1, nc) for r in raws], 1), rc, tag + ' cls')
        else:
            assert int(cls.abs().sum()) == 0


def test_decode_fcos_golden(golden):
    from mydetection_b200 import ops
    from oracle import decode as od
    g = golden('decode')
    raws, refs = [], []
    for li in range(5):
        raws.append(efdet_views(T(g[f'fcos{li}_bbox_in']), T(g[f'fcos{li}_cls_in'])))
        refs.append((T(g[f'fcos{li}_bbox']), T(g[f'fcos{li}_cls']), T(g[f'fcos{li}_score'])))
    box, cls, score = run_dense(ops.KIND_FCOS, raws, (8, 16, 32, 64, 128), None, (256, 384))
    rb, rc, rs = od.merge_levels(refs)
    close(box, rb, 384, 'fcos bbox', cancel=FCOS_W)
    close(score, rs, 1, 'fcos score')
    cls_match(cls, torch.cat([r['class'].reshape(2, -1, 6) for r in raws], 1), rc, 'fcos cls')
    # FCOS v1 reads the centerness head under the key 'center'
    raws1 = [{'bbox': r['bbox'], 'center': r['conf'], 'class': r['class']} for r in raws]
    box1, cls1, score1 = run_dense(ops.KIND_FCOS, raws1, (8, 16, 32, 64, 128), None, (256, 384), conf_key='center')
    assert torch.equal(box1, box) and torch.equal(score1, score) and torch.equal(cls1, cls)


def test_decode_retina_uv5_golden(golden):
    from mydetection_b200 import ops
    g = golden('decode')
    for tag in ('retina', 'retina_rot'):
        raw = anchor_views(T(g[f'{tag}_bbox_in']), T(g[f'{tag}_cls_in']), 9)
        box, cls, score = run_dense(ops.KIND_RETINA, [raw], (16,), [g[f'{tag}_anchors'].tolist()], (96, 128))
        close(box[..., :4], T(g[f'{tag}_bbox'])[..., :4], 128, tag + ' bbox')
        if tag == 'retina_rot':
            close(box[..., 4], T(g[f'{tag}_bbox'])[..., 4], 180, tag + ' angle', cancel=ANGLE)
        close(score, T(g[f'{tag}_score']), 1, tag + ' score')
        cls_match(cls, raw['class'].reshape(2, -1, 4), T(g[f'{tag}_cls']), tag + ' cls')
    raw = yolo_views(T(g['uv5_in']), 3, 4, 5)
    box, cls, score = run_dense(ops.KIND_UV5, [raw], (8,), [level_anchors(YOLO_ANCHORS, 0).tolist()], (96, 128))
    close(box[..., 2:], T(g['uv5_bbox'])[..., 2:], 128, 'uv5 wh')
    close(box[..., :2], T(g['uv5_bbox'])[..., :2], 128, 'uv5 xy', cancel=UV5_XY)
    close(score, T(g['uv5_score']), 1, 'uv5 score')
    cls_match(cls, raw['class'].reshape(2, -1, 5), T(g['uv5_cls']), 'uv5 cls')


def test_decode_generic_strides_and_compact(golden):
    """A contiguous (B,nH,nW,C) copy (non-planar strides) takes the scalar path and must give the
    same bits as the planar fast path; the compacting variant must select exactly score >= thr."""
    from mydetection_b200 import ops
    g = golden('decode')
    d = dev()
    raws = [efdet_views(T(g[f'fcos{li}_bbox_in']), T(g[f'fcos{li}_cls_in'])) for li in range(5)]
    planar = [{k: v.to(d) for k, v in r.items()} for r in raws]
    packed = []
    for r in planar:
        cc = torch.cat([r['conf'], r['class']], dim=-1).contiguous()      # (B,nH,nW,1+C) channels-last
        packed.append({'bbox': r['bbox'].contiguous(), 'conf': cc[..., 0:1], 'class': cc[..., 1:]})
    strides = (8, 16, 32, 64, 128)
    a = ops.decode_dense(ops.KIND_FCOS, ops.LevelSet(planar, strides), (256, 384))
    b = ops.decode_dense(ops.KIND_FCOS, ops.LevelSet(packed, strides), (256, 384))
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    box, cls, score = a
    for thr in (0.005, 0.3, 0.9, 2.0):
        c = ops.decode_compact(ops.KIND_FCOS, ops.LevelSet(planar, strides), (256, 384), thr)
        torch.cuda.synchronize()
        for bi in range(box.shape[0]):
            n = int(c['count'][bi])
            want = torch.nonzero(score[bi] >= thr).flatten()
            assert n == want.numel()
            order = torch.argsort(c['idx'][bi, :n])
            idx = c['idx'][bi, :n][order].long()
            assert torch.equal(idx, want)
            assert torch.equal(c['box'][bi, :n][order], box[bi][idx])
            assert torch.equal(c['score'][bi, :n][order], score[bi][idx])
            assert torch.equal(c['cls'][bi, :n][order].long(), cls[bi][idx])


# ------------------------------------------------------------------------------------- post-process
def gpu_keep(boxes, scores, cats, conf, nms, fmt, topk):
    from mydetection_b200 import ops
    d = dev()
    out = ops.postprocess(boxes[None].to(d), scores[None].to(d), cats[None].to(d), conf, nms, topk=topk, box_format=fmt)
    torch.cuda.synchronize()
    n = int(out['count'][0])
    assert int(out['status'][0]) == 0
    return out, n


@pytest.mark.parametrize('tag,fmt', [('pp_small', 'cxcywh'), ('pp_cap', 'cxcywh'), ('pp_rot', 'cxcywhd'),
                                     ('pp_empty', 'cxcywh')])
def test_post_process_golden(golden, tag, fmt):
    g = golden('postprocess')
    boxes, scores, cats = T(g[tag + '_boxes']), T(g[tag + '_scores']), T(g[tag + '_cats'])
    conf, nms = (float(v) for v in g[tag + '_params'])
    out, n = gpu_keep(boxes, scores, cats, conf, nms, fmt, 512)
    keep = T(g[tag + '_keep'])
    assert n == keep.numel()
    assert torch.equal(out['idx'][0, :n].cpu().long(), keep)           # kept indices, bit-exact, reference order
    assert torch.equal(out['box'][0, :n].cpu(), boxes[keep])
    assert torch.equal(out['score'][0, :n].cpu(), scores[keep])
    assert torch.equal(out['cls'][0, :n].cpu(), cats[keep])


def test_nms_direct_large_path_golden(golden):
    """ImageObjects.nms on 1200 boxes: no cap -> the tiled large-N path."""
    g = golden('postprocess')
    boxes, scores, cats = T(g['nms_direct_boxes']), T(g['nms_direct_scores']), T(g['nms_direct_cats'])
    out, n = gpu_keep(boxes, scores, cats, float('-inf'), float(g['nms_direct_params'][1]), 'cxcywh', None)
    keep = T(g['nms_direct_keep'])
    assert n == keep.numel()
    assert torch.equal(out['idx'][0, :n].cpu().long(), keep)


def test_nms_adversarial_golden(golden):
    g = golden('postprocess')
    b, s = T(g['adv_boxes']), T(g['adv_scores'])
    out, n = gpu_keep(b, s, torch.zeros(8, dtype=torch.int64), float('-inf'), float(g['adv_thr'][0]), 'cxcywh', None)
    assert torch.equal(out['score'][0, :n].cpu(), T(g['adv_keep_scores']))


@pytest.mark.parametrize('n,n_cls,topk', [(1, 1, 512), (31, 2, 512), (513, 1, 512), (1024, 3, None), (1025, 3, None),
                                          (3000, 1, None), (5000, 80, None), (4000, 4, 1000)])
def test_nms_random_vs_oracle(n, n_cls, topk):
    from oracle import postprocess as opp
    gen = torch.Generator().manual_seed(100 + n)
    boxes = torch.cat([torch.rand(n, 2, generator=gen) * 300, torch.rand(n, 2, generator=gen) * 60 + 2], 1)
    scores = (torch.rand(n, generator=gen) * 200).round() / 200           # exact score ties on purpose
    cats = torch.randint(0, n_cls, (n,), generator=gen)
    want = opp.post_process(boxes, cats, scores, 0.1, 0.5, 'cxcywh', topk)
    out, cnt = gpu_keep(boxes, scores, cats, 0.1, 0.5, 'cxcywh', topk)
    assert cnt == want.numel()
    assert torch.equal(out['idx'][0, :cnt].cpu().long(), want)


def test_postprocess_randomised_sweep():
    """Seeded sweep over candidate counts, class counts, top-k, thresholds and score distributions (uniform, heavily
    quantised, concentrated near 1, nearly all below the threshold): one batched launch per case against the
    oracle image by image.  Exercises the sampled front end and its fallback, the pair-balanced and the
    row-group IoU matrices (few / many same-class pairs) and the k-pad variants."""
    from mydetection_b200 import ops
    from oracle import postprocess as opp
    gen = torch.Generator().manual_seed(2024)
    d = dev()
    cases = 0
    for n in (7, 64, 700, 2049, 8525, 20000):
        for n_cls in (1, 5, 80):
            for topk in (16, 512, 1000):
                kind = cases % 4
                B = 3
                boxes = torch.cat([torch.rand(B, n, 2, generator=gen) * 500, torch.rand(B, n, 2, generator=gen) * 80 + 2], 2)
                u = torch.rand(B, n, generator=gen)
                scores = [u, (u * 50).round() / 50, 1 - u.pow(4) * 0.05, u * 0.12][kind]
                cats = torch.randint(0, n_cls, (B, n), generator=gen)
                thr = 0.1
                out = ops.postprocess(boxes.to(d), scores.to(d), cats.to(d), thr, 0.5, topk=topk)
                torch.cuda.synchronize()
                assert int(out['status'].abs().sum()) == 0
                for b in range(B):
                    want = opp.post_process(boxes[b], cats[b], scores[b], thr, 0.5, 'cxcywh', topk)
                    k = int(out['count'][b])
                    assert k == want.numel(), (n, n_cls, topk, kind, b)
                    assert torch.equal(out['idx'][b, :k].cpu().long(), want), (n, n_cls, topk, kind, b)
                cases += 1
    assert cases == 54


def test_postprocess_batched_ragged_counts():
    """Several images of different candidate counts in one launch, including an empty one."""
    from mydetection_b200 import ops
    from oracle import postprocess as opp
    gen = torch.Generator().manual_seed(5)
    B, n = 5, 900
    boxes = torch.cat([torch.rand(B, n, 2, generator=gen) * 200, torch.rand(B, n, 2, generator=gen) * 50 + 2], 2)
    scores = torch.rand(B, n, generator=gen)
    cats = torch.randint(0, 6, (B, n), generator=gen).int()
    counts = torch.tensor([900, 0, 1, 513, 37], dtype=torch.int32)
    d = dev()
    out = ops.postprocess(boxes.to(d), scores.to(d), cats.to(d), 0.2, 0.45, topk=512, counts=counts.to(d))
    torch.cuda.synchronize()
    for b in range(B):
        c = int(counts[b])
        want = opp.post_process(boxes[b, :c], cats[b, :c].long(), scores[b, :c], 0.2, 0.45, 'cxcywh', 512)
        k = int(out['count'][b])
        assert k == want.numel()
        assert torch.equal(out['idx'][b, :k].cpu().long(), want)


# ------------------------------------------------------------------------------------- whole path
def test_detect_end_to_end_vs_oracle(golden):
    """decode + post_process in one C call against the oracle chain on the same logits.  Membership may
    differ only for candidates whose score is within 1e-5 relative of conf_thres / of the K-th score,
    or whose IoU with a kept box is within 1e-6 of the NMS threshold; on this input none is."""
    from mydetection_b200 import ops
    from oracle import decode as od, postprocess as opp
    g = golden('decode')
    raws = [efdet_views(T(g[f'fcos{li}_bbox_in']), T(g[f'fcos{li}_cls_in'])) for li in range(5)]
    strides = (8, 16, 32, 64, 128)
    ref = od.merge_levels([od.decode_fcos(r, s, (256, 384)) for r, s in zip(raws, strides)])
    d = dev()
    ls = ops.LevelSet([{k: v.to(d) for k, v in r.items()} for r in raws], strides)
    out = ops.detect(ops.KIND_FCOS, ls, (256, 384), 0.05, 0.5, topk=512)
    torch.cuda.synchronize()
    for b in range(2):
        want = opp.post_process(ref[0][b], ref[1][b], ref[2][b], 0.05, 0.5, 'cxcywh', 512)
        k = int(out['count'][b])
        got = out['idx'][b, :k].cpu().long()
        assert k == want.numel() and torch.equal(got, want)
        close(out['box'][b, :k], ref[0][b][want], 384, 'e2e box', cancel=FCOS_W)
        close(out['score'][b, :k], ref[2][b][want], 1, 'e2e score')
        assert torch.equal(out['cls'][b, :k].cpu(), ref[1][b][want])


def test_detect_score_ties_are_broken_by_flat_index(golden):
    """Coarsely quantised logits give many exactly equal scores.  The compaction order of the fused decode
    is arbitrary, so the result may only depend on the flat candidate index (the declared tie policy):
    detect() must equal the oracle post-process applied to the GPU's own dense decode, run after run."""
    from mydetection_b200 import ops
    from oracle import postprocess as opp
    g = golden('decode')
    d = dev()
    strides = (8, 16, 32, 64, 128)
    raws = []
    for li in range(5):
        bb = (T(g[f'fcos{li}_bbox_in']) * 2).round() / 2
        cc = (T(g[f'fcos{li}_cls_in']) * 1).round() / 1
        raws.append({k: v.to(d) for k, v in efdet_views(bb, cc).items()})
    ls = ops.LevelSet(raws, strides)
    box, cls, score = (t.cpu() for t in ops.decode_dense(ops.KIND_FCOS, ls, (256, 384)))
    assert score[0].unique().numel() < score[0].numel() // 4          # heavy ties
    first = None
    for rep in range(3):
        out = ops.detect(ops.KIND_FCOS, ls, (256, 384), 0.05, 0.5, topk=512)
        torch.cuda.synchronize()
        for b in range(2):
            want = opp.post_process(box[b], cls[b], score[b], 0.05, 0.5, 'cxcywh', 512)
            k = int(out['count'][b])
            assert k == want.numel()
            assert torch.equal(out['idx'][b, :k].cpu().long(), want)
            assert torch.equal(out['box'][b, :k].cpu(), box[b][want])


def same_dets(a, b, what=''):
    assert torch.equal(a['count'], b['count']), what
    for i in range(a['count'].numel()):
        n = int(a['count'][i])
        for k in ('box', 'score', 'cls', 'idx'):
            assert torch.equal(a[k][i, :n], b[k][i, :n]), (what, i, k)


@pytest.mark.parametrize('quantise', [False, True])
def test_sampled_select_equals_scan(golden, quantise):
    """The post-process brackets the K-th score from a sample and verifies the bracket by counting; when
    the check fails it scans.  Both routes must give the same bits: random scores, quantised logits (heavy
    ties: the verified list overflows -> scan), every K regime, several thresholds, and `consume`."""
    from mydetection_b200 import ops
    g = golden('decode')
    d = dev()
    strides = (8, 16, 32, 64, 128)
    raws = []
    for li in range(5):
        bb, cc = T(g[f'fcos{li}_bbox_in']), T(g[f'fcos{li}_cls_in'])
        if quantise:
            bb, cc = (bb * 2).round() / 2, cc.round()
        raws.append({k: v.to(d) for k, v in efdet_views(bb, cc).items()})
    ls = ops.LevelSet(raws, strides)
    box, cls, score = ops.decode_dense(ops.KIND_FCOS, ls, (256, 384))
    for thr in (-float('inf'), 0.005, 0.05, 0.3):
        c = ops.decode_compact(ops.KIND_FCOS, ls, (256, 384), thr)
        for topk in (1, 8, 100, 512, 1000, None):
            # compacted candidates (already thresholded) and the dense arrays (threshold applied by the kernel)
            for args, kw in (((c['box'], c['score'], c['cls'], -float('inf'), 0.5), dict(counts=c['count'], src_idx=c['idx'])),
                             ((box, score, cls, thr, 0.5), {})):
                scan = ops.postprocess(*args, topk=topk, force_scan=True, **kw)
                fast = ops.postprocess(*args, topk=topk, **kw)
                torch.cuda.synchronize()
                assert int(fast['status'].abs().sum()) == 0
                same_dets(scan, fast, (thr, topk))
        want = ops.postprocess(c['box'], c['score'], c['cls'], -float('inf'), 0.5, topk=512, counts=c['count'], src_idx=c['idx'])
        got = ops.postprocess(c['box'], c['score'], c['cls'], -float('inf'), 0.5, topk=512, counts=c['count'],
                              src_idx=c['idx'], consume=True)
        torch.cuda.synchronize()
        same_dets(want, got, 'consume')
        assert int(c['count'].abs().sum()) == 0


@pytest.mark.parametrize('n,scale', [(700, 1.0), (5000, 1.0), (20000, 1.0), (20000, 1e4), (3000, 0.0)])
def test_sampled_select_score_scales(n, scale):
    """Scores outside [0, 1] (logits, huge magnitudes, all equal): the sample bins adapt to the range, and
    the count check keeps the result exact; fast and scan routes agree and match the oracle."""
    from mydetection_b200 import ops
    from oracle import postprocess as opp
    gen = torch.Generator().manual_seed(77 + n)
    xy = torch.rand(n, 2, generator=gen) * 600
    wh = torch.rand(n, 2, generator=gen) * 60 + 4
    boxes = torch.cat([xy, wh], 1)
    scores = (torch.randn(n, generator=gen) * scale) if scale > 0 else torch.full((n,), 0.25)
    cats = torch.randint(0, 20, (n,), generator=gen)
    d = dev()
    for topk in (64, 512):
        fast = ops.postprocess(boxes[None].to(d), scores[None].to(d), cats[None].to(d), -float('inf'), 0.5, topk=topk)
        scan = ops.postprocess(boxes[None].to(d), scores[None].to(d), cats[None].to(d), -float('inf'), 0.5, topk=topk, force_scan=True)
        torch.cuda.synchronize()
        same_dets(scan, fast, (n, scale, topk))
        want = opp.post_process(boxes, cats, scores, -float('inf'), 0.5, 'cxcywh', topk)
        k = int(fast['count'][0])
        assert k == want.numel() and torch.equal(fast['idx'][0, :k].cpu().long(), want)


def test_pipeline_stagewise_self_cleaning(golden):
    """BoundCall.launch_decode() + launch_postprocess() without any memset in between steps (the post-process
    leaves count and histogram zeroed): repeated steps must reproduce mydet_detect exactly."""
    from mydetection_b200 import pipeline as pl
    g = golden('decode')
    d = dev()
    strides = (8, 16, 32, 64, 128)
    raws = [{k: v.to(d) for k, v in efdet_views(T(g[f'fcos{li}_bbox_in']), T(g[f'fcos{li}_cls_in'])).items()} for li in range(5)]
    pipe = pl.DetectionPipeline('FCOS2', strides, 6, (256, 384), 0.05, 0.5, 512)
    bc = pipe.bind(raws)
    assert bc.self_cleaning
    want = {k: v.clone() for k, v in bc.launch().items()}
    for rep in range(3):
        bc.launch_decode()
        out = bc.launch_postprocess()
        torch.cuda.synchronize()
        assert int(bc.cand['count'].abs().sum()) == 0
        same_dets(out, want, rep)
    loose = pipe.bind(raws, self_cleaning=False)           # decode memsets its own state: may be repeated
    loose.launch_decode(); loose.launch_decode()
    out = loose.launch_postprocess()
    torch.cuda.synchronize()
    same_dets(out, want, 'self_cleaning=False')


@pytest.mark.parametrize('topk', [512, 511, 100])
def test_fused_exchange_layout_single_gpu(golden, topk):
    """mydet_postprocess_scatter with ONE peer (a local buffer): the rows the kernel's output stage stores
    into the gathered buffer must equal the separately packed detections, counts in the int32 tail.
    top-k 512 / 100: image blocks are 16-byte aligned -> staged, coalesced vector stores; 511: scalar stores."""
    from mydetection_b200 import ops, pipeline as pl
    g = golden('decode')
    d = dev()
    strides = (8, 16, 32, 64, 128)
    raws = [{k: v.to(d) for k, v in efdet_views(T(g[f'fcos{li}_bbox_in']), T(g[f'fcos{li}_cls_in'])).items()} for li in range(5)]
    pipe = pl.DetectionPipeline('FCOS2', strides, 6, (256, 384), 0.05, 0.5, topk)
    bc = pipe.bind(raws)
    ex = pl.PeerExchange(2, topk, 4, d, local_only=True)
    bc.bind_exchange(ex)
    bc.launch_decode()
    out = bc.launch_postprocess_scatter()
    torch.cuda.synchronize()
    rows, counts = ex.views()
    want_rows, want_counts = pl.unpack_gathered(pl.pack_detections(out), 1, 2, topk, 4)
    assert torch.equal(counts, out['count']) and torch.equal(counts, want_counts)
    for b in range(2):
        n = int(counts[b])
        assert n > 0 and torch.equal(rows[b, :n], want_rows[b, :n])


def test_exchange_protocol_single_gpu(golden):
    """The exchange protocol on one GPU (one rank = producer and consumer of a local buffer): five publications in
    a row with wait / release between them and NO host synchronisation inside the loop; the consumer's snapshot of
    step s must hold step s's detections (inputs alternate between two batches).  Then back-pressure: a second
    publication without the consumer's acknowledgement must not hang -- the bounded wait expires and status bit 16
    is raised -- and a wait for a publication that never comes reports its own time-out."""
    from mydetection_b200 import pipeline as pl
    g = golden('decode')
    d = dev()
    strides = (8, 16, 32, 64, 128)
    raws_a = [{k: v.to(d) for k, v in efdet_views(T(g[f'fcos{li}_bbox_in']), T(g[f'fcos{li}_cls_in'])).items()} for li in range(5)]
    raws_b = [{k: v.flip(0) for k, v in r.items()} for r in raws_a]              # the two images swapped
    pipe = pl.DetectionPipeline('FCOS2', strides, 6, (256, 384), 0.05, 0.5, 512)
    ex = pl.PeerExchange(2, 512, 4, d, local_only=True)
    calls = [pipe.bind(raws_a).bind_exchange(ex, protocol=True), pipe.bind(raws_b).bind_exchange(ex, protocol=True)]
    snaps = []
    for s in range(5):
        bc = calls[s % 2]
        bc.launch_decode()
        bc.launch_postprocess_scatter()
        ex.publish()                                 # after the kernel boundary: the producer kernel itself never fences
        counts = ex.wait().clone()
        rows = ex.views()[0].clone()                 # the consumer's read, stream-ordered behind the wait
        ex.release()
        snaps.append((rows, counts, ex.wait_status.clone(), bc.out['status'].clone()))
    torch.cuda.synchronize()
    want = []
    for bc in calls:
        bc.launch_decode()
        out = bc.launch_postprocess()
        torch.cuda.synchronize()
        want.append(pl.unpack_gathered(pl.pack_detections(out), 1, 2, 512, 4))
    for s, (rows, counts, wst, pst) in enumerate(snaps):
        w_rows, w_counts = want[s % 2]
        assert int(wst) == 0 and int((pst & 16).sum()) == 0
        assert torch.equal(counts, w_counts)
        for b in range(2):
            assert torch.equal(rows[b, :int(counts[b])], w_rows[b, :int(counts[b])])
    assert int(want[0][1][0]) != int(want[0][1][1]), 'the two images must differ for the alternation to prove anything'
    for s in (5, 6):                                 # the one-launch consumer (wait + snapshot + acknowledgement)
        bc = calls[s % 2]
        bc.launch_decode(); bc.launch_postprocess_scatter()
        counts = ex.consume_counts().clone()
        torch.cuda.synchronize()
        assert torch.equal(counts, want[s % 2][1]) and int(ex.wait_status) == 0 and int((bc.out['status'] & 16).sum()) == 0
    # back-pressure, bounded: the next publication is never acknowledged, so the one after it has to give up waiting
    import time
    calls[0].launch_decode(); calls[0].launch_postprocess_scatter(); ex.publish()
    torch.cuda.synchronize()
    assert int((calls[0].out['status'] & 16).sum()) == 0
    t0 = time.perf_counter()
    calls[1].launch_decode(); calls[1].launch_postprocess_scatter(); ex.publish()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert int((calls[1].out['status'] & 16).min()) == 16 and 0.2 < dt < 10.0, dt
    ex.wait(); ex.release(); ex.wait(); ex.release(); ex.wait()      # those two publications are there, a further one never comes
    torch.cuda.synchronize()
    assert int(ex.wait_status) == 1


# ------------------------------------------------------------------------------------- IoU / rotated
def test_bboxes_iou_bit_exact(golden):
    from mydetection_b200 import ops
    g = golden('iou')
    d = dev()
    a, b = T(g['a']).to(d), T(g['b']).to(d)
    assert torch.equal(ops.iou_aabb(a, b).cpu(), T(g['iou_cxcywh']))
    from oracle import iou as oi
    bx = oi.cxcywh_to_x1y1x2y2(T(g['b']))
    assert torch.equal(ops.iou_aabb(T(g['a_xyxy']).to(d), bx.to(d), xyxy=True).cpu(), T(g['iou_xyxy']))
    gt = T(g['gt_debug3']).to(d)
    assert torch.equal(ops.iou_aabb(gt, gt).cpu(), T(g['iou_gt_self']))


def test_rot_iou_and_nms_vs_oracle(golden):
    from mydetection_b200 import ops
    from oracle import iou as oi
    g = golden('iou')
    d = dev()
    rb, rs = T(g['rot_boxes']), T(g['rot_scores'])
    got = ops.iou_rot(rb.to(d), rb.to(d)).cpu()
    want = oi.iou_rot(rb, rb)
    # same polygons up to the last bit of float32 sin/cos (glibc sinf on the CPU, correctly rounded
    # float64->float32 on the GPU): the IoU values agree far inside the 1e-6 band of the north star
    assert float((got - want).abs().max()) < 1e-6
    for thr, key in ((0.45, 'rot_keep_045'), (0.2, 'rot_keep_02')):
        keep, cnt = ops.nms_rot(rb[None].to(d), rs[None].to(d), thr)
        torch.cuda.synchronize()
        assert torch.equal(keep[0, :int(cnt[0])].cpu(), T(g[key]))
    # larger random case, several images, vs the oracle
    gen = torch.Generator().manual_seed(11)
    B, n = 3, 2500
    boxes = torch.empty(B, n, 5)
    boxes[..., 0:2] = torch.rand(B, n, 2, generator=gen) * 600 + 50
    boxes[..., 2:4] = torch.rand(B, n, 2, generator=gen) * 90 + 8
    boxes[..., 4] = torch.rand(B, n, generator=gen) * 360 - 180
    scores = torch.rand(B, n, generator=gen)
    keep, cnt = ops.nms_rot(boxes.to(d), scores.to(d), 0.45)
    torch.cuda.synchronize()
    for b in range(B):
        want = oi.nms_rot(boxes[b], scores[b], 0.45)
        assert torch.equal(keep[b, :int(cnt[b])].cpu(), want)


# ------------------------------------------------------------------------------------- ATSS
def test_atss_golden(golden):
    from mydetection_b200 import ops
    g = golden('atss')
    d = dev()
    strides, sides = [8, 16, 32, 64, 128], [24, 48, 96, 192, 384]
    G = 12
    gt_box = torch.zeros(2, G, 4)
    gt_cls = torch.zeros(2, G, dtype=torch.int64)
    cnt = torch.zeros(2, dtype=torch.int32)
    for b in range(2):
        bx, ct = T(g[f'gt{b}_boxes']), T(g[f'gt{b}_cats'])
        gt_box[b, :bx.shape[0]] = bx
        gt_cls[b, :bx.shape[0]] = ct
        cnt[b] = bx.shape[0]
    for li in range(5):
        t = T(g[f'atss{li}_bbox_in']).to(d).permute(0, 2, 3, 1)
        out = ops.atss_assign(t, li, strides, sides, (384, 512), gt_box.to(d), gt_cls.to(d), cnt.to(d), 9, 0.7, 6)
        torch.cuda.synchronize()
        for k in ('PositiveMask', 'IgnoredMask', 'TargetConf', 'TargetCls'):
            assert torch.equal(out[k].cpu(), T(g[f'atss{li}_{k}'])), (li, k)
        close(out['TargetLTRB'], T(g[f'atss{li}_TargetLTRB']), 512, f'atss{li} ltrb')
        # thresholds handed over from another level's call give the same maps (they are level independent)
        again = ops.atss_assign(t, li, strides, sides, (384, 512), gt_box.to(d), gt_cls.to(d), cnt.to(d), 9, 0.7, 6,
                                thr=out['thr'])
        for k in ('PositiveMask', 'IgnoredMask', 'TargetConf', 'TargetCls', 'TargetLTRB'):
            assert torch.equal(again[k], out[k]), (li, k)


@pytest.mark.parametrize('n', [1025, 3000, 8192, 10000, 16384])
def test_radix_sort_equals_bitonic_sort(n, monkeypatch):
    """The one-CTA LSD radix sort of the large-N path (score order and Morton order, n in (1024, 16384]) against the
    bitonic network it replaces (MYDET_SORT_BITONIC=1) and against the oracle: identical kept indices, in identical order,
    for rotated NMS (single class: 4 passes) and class-aware axis-aligned NMS (6 passes), with heavy score ties."""
    from mydetection_b200 import ops
    from oracle import iou as oi, postprocess as opp
    d = dev()
    gen = torch.Generator().manual_seed(n)
    rb = torch.cat([torch.rand(2, n, 2, generator=gen) * 700, torch.rand(2, n, 2, generator=gen) * 50 + 10,
                    torch.rand(2, n, 1, generator=gen) * 180 - 90], dim=2)
    rs = (torch.rand(2, n, generator=gen) * 200).round() / 200                       # ~50 boxes per score value: ties by index
    counts = torch.tensor([n, n - 37], dtype=torch.int32)
    keep, cnt = ops.nms_rot(rb.to(d), rs.to(d), 0.45, counts=counts.to(d))
    monkeypatch.setenv('MYDET_SORT_BITONIC', '1')
    keep_b, cnt_b = ops.nms_rot(rb.to(d), rs.to(d), 0.45, counts=counts.to(d))
    monkeypatch.delenv('MYDET_SORT_BITONIC')
    assert torch.equal(cnt, cnt_b)
    for b in range(2):
        c = int(cnt[b])
        assert torch.equal(keep[b, :c], keep_b[b, :c])
    want = oi.nms_rot(rb[1, :n - 37], rs[1, :n - 37], 0.45)
    assert int(cnt[1]) == want.numel() and torch.equal(keep[1, :int(cnt[1])].cpu(), want)
    # class-aware axis-aligned NMS without a cap: the large path with class bits in the keys
    cls = torch.randint(0, 7, (2, n), generator=gen)
    out = ops.postprocess(rb[..., :4].contiguous().to(d), rs.to(d), cls.to(d), -1.0, 0.5, topk=None)
    monkeypatch.setenv('MYDET_SORT_BITONIC', '1')
    out_b = ops.postprocess(rb[..., :4].contiguous().to(d), rs.to(d), cls.to(d), -1.0, 0.5, topk=None)
    monkeypatch.delenv('MYDET_SORT_BITONIC')
    assert torch.equal(out['count'], out_b['count'])
    c0 = int(out['count'][0])
    assert torch.equal(out['idx'][0, :c0], out_b['idx'][0, :c0])
    want = opp.post_process(rb[0, :, :4], cls[0], rs[0], -1.0, 0.5, 'cxcywh', None)
    assert c0 == want.numel() and torch.equal(out['idx'][0, :c0].cpu().long(), want)


@pytest.mark.parametrize('canvas,lo,hi,n', [(2048, 6, 250, 400), (2048, 0.3, 12, 400), (96, 4, 60, 300), ((70, 130), 2, 90, 300)])
def test_raster_iou_bit_exact_vs_oracle(canvas, lo, hi, n):
    """mydet_iou_raster_pairwise (closed-form column runs) against oracle/raster.c (the restated pycocotools boundary
    walk + RLE merge): bit-exact, on the reference's 2048 canvas, for sub-pixel boxes, and on small canvases that clip
    the boxes on every side; also through the mirror's iou_rle(..., raster=True, img_hw=...)."""
    from mydetection_b200 import ops, bbox_ops
    from oracle import iou as oi
    h, w = (canvas, canvas) if isinstance(canvas, int) else canvas
    gen = torch.Generator().manual_seed(int(h + 10 * hi))
    bx = torch.cat([torch.rand(n, 1, generator=gen) * (w + 20) - 10, torch.rand(n, 1, generator=gen) * (h + 20) - 10,
                    torch.rand(n, 2, generator=gen) * (hi - lo) + lo, torch.rand(n, 1, generator=gen) * 360 - 180], dim=1)
    if canvas == 2048:                                                   # clustered: plenty of overlapping pairs
        bx[:, :2] = torch.rand(n, 2, generator=gen) * (400 if hi > 100 else 40) + 800
    a, b = bx[:n // 2], bx[n // 2:]
    # the oracle helper takes a square canvas; call the C function for the rectangular one
    import ctypes
    rad = bx.clone()
    rad[:, 4] = oi.deg2rad_f32(rad[:, 4])
    cs = np.ascontiguousarray(oi.xywha2vertex(rad).reshape(-1, 8).double().numpy())
    want = np.empty((n // 2, n - n // 2))
    f64p = ctypes.POINTER(ctypes.c_double)
    oi.lib().oracle_raster_iou_pairwise(cs[:n // 2].ctypes.data_as(f64p), n // 2, cs[n // 2:].ctypes.data_as(f64p), n - n // 2,
                                        h, w, want.ctypes.data_as(f64p))
    got = ops.iou_raster(a.to(dev()), b.to(dev()), (h, w)).cpu().numpy()
    assert (want > 0).sum() > 20
    assert np.array_equal(got, want), (np.abs(got - want).max(), int((got != want).sum()))
    via = bbox_ops.iou_rle(a, b, raster=True, img_hw=(h, w))
    assert not via.is_cuda and np.array_equal(via.numpy(), want)
    exact = bbox_ops.iou_rle(a, b).numpy()
    big = want > 0.3
    if canvas == 2048 and hi > 100 and big.any():                        # un-clipped boxes of ordinary size: the raster follows the exact IoU
        assert np.abs(exact - want)[big].max() < 0.08


def test_persistent_workspace_is_left_clean():
    """The large-N NMS on a persistent workspace (ops.nms_rot caches one per geometry; mydet_nms_rot_ws with
    workspace_clean=1): the suppression matrix is never cleared wholesale after the first call, each call zeroes what it
    set.  Different inputs of one geometry, one after the other -- sparse, then so dense that the entry list overflows
    (whole-matrix cleanup path), then sparse again -- must each equal the oracle; a stale bit would suppress a box."""
    from mydetection_b200 import ops
    from oracle import iou as oi
    d = dev()
    n = 2000
    gen = torch.Generator().manual_seed(5)

    def boxes(spread):
        return torch.cat([torch.rand(1, n, 2, generator=gen) * spread + 100, torch.rand(1, n, 2, generator=gen) * 40 + 20,
                          torch.rand(1, n, 1, generator=gen) * 180 - 90], dim=2)

    ops.release_workspaces()
    for spread in (900.0, 900.0, 25.0, 900.0, 300.0):
        rb, rs = boxes(spread), torch.rand(1, n, generator=gen)
        keep, cnt = ops.nms_rot(rb.to(d), rs.to(d), 0.3)
        want = oi.nms_rot(rb[0], rs[0], 0.3)
        assert int(cnt[0]) == want.numel() and torch.equal(keep[0, :int(cnt[0])].cpu(), want), spread
    assert len(ops._PERSISTENT) == 1
    # the same through a bound pipeline call (mydet_detect_ws): un-capped single-class scenes, launched repeatedly
    from mydetection_b200 import pipeline as pl
    from oracle import decode as od, postprocess as opp
    raws_cpu = []
    for s_ in (8, 16):
        m = 256 // s_
        t = torch.randn(1, 6, m, m, generator=gen) * 0.5
        t[:, 4] = torch.randn(1, m, m, generator=gen) * 1.5 + 2.0
        raws_cpu.append({k: v[:, 0] for k, v in yolo_views(t, 1, 4, 1).items()})
    pipe = pl.DetectionPipeline('FCOS2', (8, 16), 1, (256, 256), 0.005, 0.3, None)
    bc = pipe.bind([{k: v.to(d) for k, v in r.items()} for r in raws_cpu])
    ref = od.merge_levels([od.decode_fcos(r, s_, (256, 256)) for r, s_ in zip(raws_cpu, (8, 16))])
    want = opp.post_process(ref[0][0], ref[1][0], ref[2][0], 0.005, 0.3, 'cxcywh', None)
    for _ in range(3):
        out = bc.launch()
        torch.cuda.synchronize()
        c = int(out['count'][0])
        assert c == want.numel() and torch.equal(out['idx'][0, :c].cpu().long(), want)


@pytest.mark.parametrize('spread', [4.0, 12.0, 400.0])
def test_lazy_narrow_phase_equals_full_and_oracle(spread, monkeypatch):
    """The root-first narrow phase of rotated NMS (pairs with a member that a certainly-kept box suppresses are never
    clipped) against clipping every listed pair (MYDET_ROT_LAZY=0) and against the oracle: identical kept indices and
    votes -- on tight clusters (most boxes die at their root), loose clusters (chains: a dead box must not suppress) and
    scattered boxes (few roots)."""
    from mydetection_b200 import ops
    from oracle import iou as oi
    d = dev()
    gen = torch.Generator().manual_seed(int(spread))
    n_obj, per = 12, 250
    n = n_obj * per
    obj = torch.rand(2, n_obj, 1, 2, generator=gen) * 800 + 100
    xy = (obj + torch.randn(2, n_obj, per, 2, generator=gen) * spread).reshape(2, n, 2)
    wh = (torch.rand(2, n_obj, 1, 2, generator=gen) * 80 + 30 + torch.randn(2, n_obj, per, 2, generator=gen) * 3).reshape(2, n, 2).abs() + 4
    ang = (torch.rand(2, n_obj, 1, 1, generator=gen) * 180 - 90 + torch.randn(2, n_obj, per, 1, generator=gen) * 8).reshape(2, n, 1)
    rb, rs = torch.cat([xy, wh, ang], dim=2), torch.rand(2, n, generator=gen)
    keep, cnt, votes = ops.nms_rot(rb.to(d), rs.to(d), 0.45, want_votes=True)        # default: lazy per image, by its pair count
    for mode in ('0', '2'):                                                          # never lazy / always lazy
        monkeypatch.setenv('MYDET_ROT_LAZY', mode)
        keep_f, cnt_f, votes_f = ops.nms_rot(rb.to(d), rs.to(d), 0.45, want_votes=True)
        monkeypatch.delenv('MYDET_ROT_LAZY')
        assert torch.equal(cnt, cnt_f)
        for b in range(2):
            c = int(cnt[b])
            assert torch.equal(keep[b, :c], keep_f[b, :c]) and torch.equal(votes[b, :c], votes_f[b, :c]), (mode, b)
    for b in range(2):
        c = int(cnt[b])
        want = oi.nms_rot(rb[b], rs[b], 0.45)
        assert c == want.numel() and torch.equal(keep[b, :c].cpu(), want), (spread, b)


@pytest.mark.parametrize('n', [16385, 40000])
def test_rank_merge_sort_equals_bitonic_merge(n, monkeypatch):
    """Large-N sort above 16 384 candidates: sorted 16 384-key chunks merged by ranking (binary searches) against the
    bitonic merge steps it replaces (MYDET_SORT_BIG_MERGE=1) and against the oracle -- un-capped single-class NMS with
    heavy score ties, counts below the capacity on the second image."""
    from mydetection_b200 import ops
    from oracle import postprocess as opp
    d = dev()
    gen = torch.Generator().manual_seed(n)
    bx = torch.cat([torch.rand(2, n, 2, generator=gen) * 1500, torch.rand(2, n, 2, generator=gen) * 24 + 6], dim=2)
    sc = (torch.rand(2, n, generator=gen) * 500).round() / 500
    cls = torch.zeros(2, n, dtype=torch.int64)
    counts = torch.tensor([n, n - 4097], dtype=torch.int32)
    out = ops.postprocess(bx.to(d), sc.to(d), cls.to(d), -1.0, 0.5, topk=None, counts=counts.to(d))
    monkeypatch.setenv('MYDET_SORT_BIG_MERGE', '1')
    out_b = ops.postprocess(bx.to(d), sc.to(d), cls.to(d), -1.0, 0.5, topk=None, counts=counts.to(d))
    monkeypatch.delenv('MYDET_SORT_BIG_MERGE')
    assert torch.equal(out['count'], out_b['count'])
    for b in range(2):
        c = int(out['count'][b])
        assert torch.equal(out['idx'][b, :c], out_b['idx'][b, :c])
    m = int(counts[1])
    want = opp.post_process(bx[1, :m], cls[1, :m], sc[1, :m], -1.0, 0.5, 'cxcywh', None)
    c = int(out['count'][1])
    assert c == want.numel() and torch.equal(out['idx'][1, :c].cpu().long(), want)
