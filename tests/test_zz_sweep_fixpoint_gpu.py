"""GPU tests of the fixed-point sweep (sweep_kernel<2>) and of the broad / narrow phase rotated mask against the round-1 kernels.

STATUS: green on a B200 in the round-1 driver run (GPUTEST_r01.json); hard tests since round 2.  The phases the kernel
runs between its barriers are also verified on the CPU (tests/test_sweep_fixpoint_host.py).  Each test runs the block
sweep and the fixed-point sweep on the same inputs and demands identical results.
"""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu]


class legacy_path:
    """The round-1 kernels: serial block sweep (MYDET_SWEEP_FIXPOINT=0) and the single tile kernel of the rotated mask
    (MYDET_ROT_MASK_TILES=1).  The C library reads both variables at every large-N call; the defaults since round 2 are
    the fixed-point sweep and the broad / narrow phase split."""
    def __enter__(self):
        os.environ['MYDET_SWEEP_FIXPOINT'] = '0'
        os.environ['MYDET_ROT_MASK_TILES'] = '1'

    def __exit__(self, *exc):
        os.environ.pop('MYDET_SWEEP_FIXPOINT', None)
        os.environ.pop('MYDET_ROT_MASK_TILES', None)


def rotated_boxes(gen, batch, n, span):
    b = torch.cat([torch.rand(batch, n, 2, generator=gen) * span, torch.rand(batch, n, 2, generator=gen) * 100 + 10,
                   torch.rand(batch, n, 1, generator=gen) * 180 - 90], dim=2)
    return b, torch.rand(batch, n, generator=gen)


def test_rotated_nms_same_result_and_votes():
    from mydetection_b200 import ops
    from oracle import iou as oi
    gen = torch.Generator().manual_seed(41)
    dev = torch.device('cuda', 0)
    # (1100, 120): heavy overlap, long suppression lists; (3000, 40): nearly every pair overlaps, so the pair list of the
    # broad phase overflows (> 256 listed partners per box) and the tile kernel takes the image over
    for n, span in ((3000, 700.0), (10000, 1024.0), (1100, 120.0), (3000, 40.0)):
        b, s = rotated_boxes(gen, 3, n, span)
        counts = torch.tensor([n, n - 37, n // 2], dtype=torch.int32)
        with legacy_path():
            ref = ops.nms_rot(b.to(dev), s.to(dev), 0.45, counts=counts.to(dev), want_votes=True)
        got = ops.nms_rot(b.to(dev), s.to(dev), 0.45, counts=counts.to(dev), want_votes=True)
        torch.cuda.synchronize()
        assert torch.equal(ref[1], got[1])
        for i in range(3):
            k = int(ref[1][i])
            assert torch.equal(ref[0][i, :k], got[0][i, :k]) and torch.equal(ref[2][i, :k], got[2][i, :k])
        want = oi.nms_rot(b[2, :n // 2], s[2, :n // 2], 0.45)
        assert torch.equal(got[0][2, :int(got[1][2])].cpu(), want)


@pytest.mark.parametrize('img,batch', [(704, 4), (1024, 2)])
def test_dense_scene_same_result(img, batch):
    from mydetection_b200 import ops
    from mydetection_b200.heads import yolo_head_views
    gen = torch.Generator().manual_seed(1005 + img)
    dev = torch.device('cuda', 0)
    raws = []
    for s in (8, 16, 32):
        n = img // s
        t = torch.randn(batch, 6, n, n, generator=gen) * 0.5
        t[:, 4] = torch.randn(batch, n, n, generator=gen) * 1.5 + 2.0
        raws.append({k: v[:, 0].to(dev) for k, v in yolo_head_views(t, 1, 4, 1).items()})
    ls = ops.LevelSet(raws, (8, 16, 32))
    with legacy_path():
        ref = ops.detect(ops.KIND_FCOS, ls, (img, img), 0.005, 0.45, topk=None)
    got = ops.detect(ops.KIND_FCOS, ls, (img, img), 0.005, 0.45, topk=None)
    torch.cuda.synchronize()
    assert torch.equal(ref['count'], got['count']) and int(got['status'].abs().sum()) == 0
    for b in range(batch):
        k = int(ref['count'][b])
        assert k > 1000 and torch.equal(ref['idx'][b, :k], got['idx'][b, :k]) and torch.equal(ref['box'][b, :k], got['box'][b, :k])
