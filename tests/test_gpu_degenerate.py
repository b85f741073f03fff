"""GPU: empty, single-element and degenerate inputs through the entry points added in round 2 (persistent-workspace
rotated NMS, raster IoU, segmented rotated IoU, tracklet bank, exchange publish) and through the pairwise IoU family.
The reference's behaviour on these is what its own tests pin (utils/bbox_ops.py:271-272 empty NMS input returns an empty
index tensor; structures.py:120-121 an empty ImageObjects returns itself); everything else must simply return the right
SHAPE, launch nothing that faults, and leave the library usable for the next call."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device('cuda', 0)


def _rot_boxes(n, seed=0, span=200.0):
    g = torch.Generator().manual_seed(seed)
    return torch.cat([torch.rand(n, 2, generator=g) * span + 50, torch.rand(n, 2, generator=g) * 60 + 8,
                      torch.rand(n, 1, generator=g) * 360 - 180], dim=1)


def test_nms_rot_empty_and_single():
    from mydetection_b200 import ops
    from oracle import iou as oi
    d = dev()
    # no images
    keep, cnt = ops.nms_rot(torch.zeros(0, 16, 5, device=d), torch.zeros(0, 16, device=d), 0.45)
    assert keep.shape[0] == 0 and cnt.shape == (0,)
    # images without boxes
    keep, cnt, votes = ops.nms_rot(torch.zeros(3, 0, 5, device=d), torch.zeros(3, 0, device=d), 0.45, want_votes=True)
    assert cnt.tolist() == [0, 0, 0]
    # one box: kept, one vote
    b = _rot_boxes(1)
    keep, cnt, votes = ops.nms_rot(b[None].to(d), torch.tensor([[0.7]], device=d), 0.45, want_votes=True)
    assert cnt.tolist() == [1] and keep[0, 0].item() == 0 and votes[0, 0].item() == 1
    # counts of zero next to a full image, on the large-N path (n > 1024)
    n = 1500
    bx = _rot_boxes(n, 3, 400.0)
    sc = torch.rand(n, generator=torch.Generator().manual_seed(4))
    counts = torch.tensor([0, n, 1], dtype=torch.int32, device=d)
    keep, cnt = ops.nms_rot(bx[None].repeat(3, 1, 1).to(d), sc[None].repeat(3, 1).to(d), 0.45, counts=counts)
    want = oi.nms_rot(bx, sc, 0.45)
    assert cnt.tolist() == [0, len(want), 1]
    assert keep[1, :len(want)].cpu().tolist() == want.tolist() and keep[2, 0].item() == 0
    # and the library still answers an ordinary call afterwards
    keep, cnt = ops.nms_rot(bx[None].to(d), sc[None].to(d), 0.45)
    assert keep[0, :int(cnt[0])].cpu().tolist() == want.tolist()


def test_nms_rot_all_identical_boxes():
    """Every box the same (the worst case of the lazy narrow phase and of the entry list: one root suppresses all)."""
    from mydetection_b200 import ops
    d = dev()
    n = 3000
    b = torch.tensor([[300.0, 300.0, 80.0, 40.0, 30.0]]).repeat(n, 1)
    sc = torch.linspace(0.9, 0.1, n)
    keep, cnt, votes = ops.nms_rot(b[None].to(d), sc[None].to(d), 0.45, want_votes=True)
    assert cnt.tolist() == [1] and keep[0, 0].item() == 0 and votes[0, 0].item() == n
    # equal scores as well: the tie goes to the lower index (stable descending sort of the reference)
    keep, cnt = ops.nms_rot(b[None].to(d), torch.full((1, n), 0.5, device=d), 0.45)
    assert cnt.tolist() == [1] and keep[0, 0].item() == 0


def test_pairwise_iou_empty_sides():
    from mydetection_b200 import ops
    d = dev()
    a4, a5 = torch.rand(6, 4, device=d) * 50 + 10, _rot_boxes(6).to(d)
    assert ops.iou_aabb(a4[:0], a4).shape == (0, 6) and ops.iou_aabb(a4, a4[:0]).shape == (6, 0)
    assert ops.iou_rot(a5[:0], a5).shape == (0, 6) and ops.iou_rot(a5, a5[:0]).shape == (6, 0)
    assert ops.iou_raster(a5[:0], a5).shape == (0, 6) and ops.iou_raster(a5, a5[:0]).shape == (6, 0)
    flat, out0 = ops.iou_rot_segments(a5, a5, torch.zeros(0, 4, dtype=torch.int64))
    assert flat.numel() == 0 and out0 == []
    # segments with an empty side between two ordinary ones keep their (empty) slot
    seg = torch.tensor([[0, 2, 0, 3], [2, 0, 3, 2], [2, 4, 3, 0], [2, 4, 3, 3]])
    flat, out0 = ops.iou_rot_segments(a5, a5, seg)
    assert out0 == [0, 6, 6, 6] and flat.numel() == 18
    full = ops.iou_rot(a5, a5)
    assert torch.allclose(flat[:6].view(2, 3), full[0:2, 0:3], rtol=0, atol=1e-12)
    assert torch.allclose(flat[6:].view(4, 3), full[2:6, 3:6], rtol=0, atol=1e-12)


def test_raster_iou_degenerate_boxes_vs_oracle():
    """Zero-size boxes, boxes entirely off the canvas, a box covering the whole canvas: bit-exact against oracle/raster.c
    (pycocotools gives 0 / 0 -> the oracle's convention for an empty union is what the kernel must reproduce)."""
    from mydetection_b200 import ops
    from oracle import iou as oi
    h = w = 64
    bx = torch.tensor([[20.0, 20.0, 0.0, 0.0, 0.0],        # zero size
                       [20.0, 20.0, 10.0, 0.0, 45.0],      # zero height
                       [-50.0, -50.0, 10.0, 10.0, 10.0],   # off canvas
                       [500.0, 20.0, 30.0, 30.0, 0.0],     # off canvas to the right
                       [32.0, 32.0, 400.0, 400.0, 0.0],    # covers everything
                       [32.0, 32.0, 400.0, 400.0, 45.0],
                       [0.0, 0.0, 20.0, 20.0, 0.0],        # corner-clipped
                       [63.9, 63.9, 5.0, 5.0, 30.0],
                       [20.0, 20.0, 0.3, 0.3, 0.0],        # sub-pixel
                       [20.5, 20.5, 1.0, 1.0, 0.0]])
    rad = bx.clone()
    rad[:, 4] = oi.deg2rad_f32(rad[:, 4])
    cs = np.ascontiguousarray(oi.xywha2vertex(rad).reshape(-1, 8).double().numpy())
    n = bx.shape[0]
    want = np.empty((n, n))
    f64p = ctypes.POINTER(ctypes.c_double)
    oi.lib().oracle_raster_iou_pairwise(cs.ctypes.data_as(f64p), n, cs.ctypes.data_as(f64p), n, h, w, want.ctypes.data_as(f64p))
    got = ops.iou_raster(bx.to(dev()), bx.to(dev()), (h, w)).cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want)), (got, want)
    assert np.array_equal(np.nan_to_num(got, nan=-1.0), np.nan_to_num(want, nan=-1.0)), (got, want)
    assert got[4, 5] == 1.0 and got[4, 4] == 1.0                   # both cover the whole canvas


def test_tracklet_bank_empty_and_no_measurements():
    from mydetection_b200.tracking import TrackletBank
    bank = TrackletBank(torch.zeros(0, 5), torch.zeros(0), img_hw=(480, 640))
    assert len(bank) == 0
    assert bank.predict().shape == (0, 5)
    assert bank.update(torch.zeros(0, 5), torch.zeros(0)).shape == (0, 5)
    assert bank.likelihood(torch.zeros(4, 5, dtype=torch.float64)).shape == (0, 4)
    assert bank.is_feasible().shape == (0,)
    assert bank.association_iou(torch.zeros(3, 5)).shape == (0, 3)
    # tracklets but no candidates / no measurement for anyone: state advances by the prediction only
    b = _rot_boxes(5, 7)
    bank = TrackletBank(b, torch.full((5,), 0.9), img_hw=(480, 640))
    with pytest.raises(AssertionError):
        bank.update(b, torch.full((5,), 0.5))                       # structures.py:489: predict() first
    pred = bank.predict()
    assert bank.likelihood(torch.zeros(0, 5, dtype=torch.float64)).shape == (5, 0)
    x_before, score_before = bank.x.clone(), bank.score.clone()
    out = bank.update(torch.zeros(5, 5), torch.zeros(5), has=torch.zeros(5, dtype=torch.bool))
    assert torch.equal(out, torch.zeros_like(out))
    assert torch.equal(bank.x, x_before) and torch.equal(bank.score, score_before)
    assert torch.equal(bank.bbox, pred)


def test_exchange_publish_of_nothing_and_bad_arguments():
    """batch == 0 publishes nothing and launches nothing; a bad peer description is refused with an error code and a
    message, never a launch."""
    from mydetection_b200 import _lib
    L = _lib.lib()
    d = dev()
    nbytes = L.mydet_exchange_buffer_bytes(8, 16, 4)
    assert nbytes > 0 and L.mydet_exchange_buffer_bytes(8, 16, 7) == 0 and L.mydet_exchange_buffer_bytes(-1, 16, 4) == 0
    buf = torch.zeros(nbytes // 4, dtype=torch.int32, device=d)
    peers = (ctypes.c_void_p * 1)(buf.data_ptr())
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.mydet_exchange_publish(peers, 1, None, 0, 0, 0, 8, 16, 4, stream) == 0
    torch.cuda.synchronize()
    assert int(buf.abs().sum()) == 0
    assert L.mydet_exchange_publish(peers, 1, None, 3, 0, 4, 8, 16, 4, stream) != 0          # self index outside the peers
    assert L.mydet_exchange_publish(peers, 1, None, 0, 6, 4, 8, 16, 4, stream) != 0          # rows past the end
    assert L.mydet_exchange_publish(peers, 9, None, 0, 0, 4, 8, 16, 4, stream) != 0          # more than 8 peers
    assert L.mydet_exchange_wait(None, 8, 16, 4, None, None, stream) != 0
    torch.cuda.synchronize()
    assert int(buf.abs().sum()) == 0


@pytest.mark.parametrize('k', [203, 204])
@pytest.mark.parametrize('xyxy', [False, True])
def test_iou_aabb_bit_pattern_on_disjoint_and_non_finite_pairs(k, xyxy):
    """The pairwise kernels skip the divide of a disjoint pair (its quotient is the +-0 / NaN numerator itself when
    area_a + area_b > 0): the BIT PATTERN of every element -- signed zeros, NaN positions -- equals the reference
    formula's (bbox_ops.py:25-50, restated in oracle/iou.py on torch CPU), for the vector-store kernel (k % 4 == 0) and
    the scalar one, with zero-size, negative-size, infinite and NaN boxes among ordinary ones; the row-max kernel
    agrees with torch.max over that matrix."""
    from mydetection_b200 import ops
    from oracle import iou as oi
    g = torch.Generator().manual_seed(k)
    def boxes(n):
        t = torch.cat([torch.rand(n, 2, generator=g) * 600, torch.rand(n, 2, generator=g) * 50 + 2], 1)
        if xyxy:
            t = torch.cat([t[:, :2], t[:, :2] + t[:, 2:]], 1)
        return t
    a, b = boxes(150), boxes(k)
    a[3, 2:] = a[3, :2] if xyxy else 0.0                 # zero size
    b[5, 2:] = b[5, :2] if xyxy else 0.0
    a[4] = a[3]                                          # two zero-size boxes at the same place: 0 / 0
    b[6] = a[3]
    a[7, 2] = float('inf'); b[9, 3] = float('inf')
    a[11, 0] = float('nan'); b[13, 2] = float('nan')
    if xyxy:
        a[15, 2:] = a[15, :2] - 5.0                      # negative extent: negative "area"
    else:
        a[15, 2] = -a[15, 2]
    want = oi.bboxes_iou(a, b, xyxy=xyxy)
    got = ops.iou_aabb(a.to(dev()), b.to(dev()), xyxy=xyxy).cpu()
    nan_w, nan_g = torch.isnan(want), torch.isnan(got)
    assert torch.equal(nan_w, nan_g)
    assert torch.equal(got.view(torch.int32)[~nan_g], want.view(torch.int32)[~nan_w])
    assert (want == 0).float().mean() > 0.9 and int(nan_w.sum()) > 0
    # row-wise max / arg-max without the matrix: rows without a NaN (torch.max with NaN: covered in test_gpu_parity)
    rows = ~nan_w.any(dim=1)
    mx, arg = ops.iou_rowmax(a[None].to(dev()), b[None].to(dev()), xyxy=xyxy)
    wmx, warg = want.max(dim=1)
    assert torch.equal(mx[0].cpu()[rows], wmx[rows]) and torch.equal(arg[0].cpu()[rows], warg[rows])


def test_atss_all_levels_in_one_call_equals_level_by_level():
    """mydet_atss_assign_levels (one GT ordering, one threshold pass, ONE grid over the cells of every level) against
    five calls of mydet_atss_assign (itself pinned to the reference's targets in test_gpu_configs / test_zz_fullmodel):
    every output bit-identical, with images of 0, 1 and many GT boxes and non-contiguous (permuted NCHW) inputs."""
    from mydetection_b200 import ops
    d = dev()
    g = torch.Generator().manual_seed(11)
    strides, sides, img = [8, 16, 32, 64, 128], [24, 48, 96, 192, 384], (384, 512)
    B, G, C = 4, 23, 7
    gt_box = torch.cat([torch.rand(B, G, 1, generator=g) * img[1], torch.rand(B, G, 1, generator=g) * img[0],
                        torch.rand(B, G, 2, generator=g) * 150 + 6], dim=2)
    gt_cls = torch.randint(0, C, (B, G), generator=g)
    gt_count = torch.tensor([G, 0, 1, 9], dtype=torch.int32)
    gt_box[3, 0] = torch.tensor([-40.0, 100.0, 60.0, 60.0])          # a centre outside the image: the exhaustive candidate scan
    gt_box[3, 1] = torch.tensor([2.0, 3.0, 30.0, 20.0])              # in the corner cell: the clipped 11 x 11 window
    gt_box, gt_cls, gt_count = gt_box.to(d), gt_cls.to(d), gt_count.to(d)
    ts = [(torch.randn(B, 4, img[0] // s, img[1] // s, generator=g) * 0.6).to(d).permute(0, 2, 3, 1) for s in strides]
    want, thr = [], None
    for li in range(len(strides)):
        o = ops.atss_assign(ts[li], li, strides, sides, img, gt_box, gt_cls, gt_count, 9, 0.6, C, thr=thr)
        thr = o['thr']
        want.append(o)
    got = ops.atss_assign_levels(ts, strides, sides, img, gt_box, gt_cls, gt_count, 9, 0.6, C)
    assert len(got) == len(want)
    n_pos = 0
    for li, (a, b) in enumerate(zip(got, want)):
        for k in ('PositiveMask', 'IgnoredMask', 'TargetLTRB', 'TargetConf', 'TargetCls'):
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, (li, k)
            assert torch.equal(a[k], b[k]), (li, k)
        n_pos += int(a['PositiveMask'].sum())
    valid = torch.arange(G)[None, :] < gt_count.cpu()[:, None]
    assert torch.equal(got[0]['thr'].cpu()[valid], want[0]['thr'].cpu()[valid])
    assert n_pos > 50


def test_atss_threshold_on_grids_too_thin_for_the_candidate_window():
    """A pyramid level of 1 x 20 or 2 x 12 cells (found by tests/test_kernel_claims_cpu.py): the k = 9 nearest anchors of a
    corner GT reach 8.5 cells, outside the 11 x 11 window -- such grids take the exhaustive scan.  Against the oracle's
    torch.topk over all anchors."""
    from mydetection_b200 import ops
    from oracle import atss as oa
    for img, strides, sides in (((8, 160), [8], [24]), ((16, 96), [8], [24]), ((32, 256), [8, 16, 32], [24, 48, 96])):
        pts = [(0, 0), (img[1], img[0]), (3.3, 2.2), (img[1] / 2 + 0.7, img[0] / 2 + 0.3), (img[1] - 1.9, 1.1), (17.1, img[0] - 0.6)]
        gt = torch.tensor([[x, y, 30.0 + 3 * i, 12.0 + 2 * i] for i, (x, y) in enumerate(pts)], dtype=torch.float32)[None]
        cls = torch.zeros(1, len(pts), dtype=torch.int64)
        counts = torch.tensor([len(pts)], dtype=torch.int32)
        k = 9 if len(strides) == 1 else 4                      # every level must hold at least k anchors
        t = torch.zeros(1, img[0] // strides[0], img[1] // strides[0], 4)
        out = ops.atss_assign(t.to(dev()), 0, strides, sides, img, gt.to(dev()), cls.to(dev()), counts.to(dev()), k, 0.7, 3)
        anchors = oa.all_level_anchors(img, strides, sides)
        for i in range(len(pts)):
            got = float(out['thr'][0, i])
            want = float(oa.atss_threshold_index_ties(gt[0, i], anchors, k))
            assert abs(got - want) <= 2e-6 * max(1.0, abs(want)), (img, pts[i], got, want)
