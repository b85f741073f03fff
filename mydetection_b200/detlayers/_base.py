"""Shared plumbing of the det-layer mirrors: stage the raw views on the GPU (no copy when they
already are) and call the one-launch dense decode."""
import torch

from .. import _lib, ops


def stage_raw(raw, keys, detach=True):
    """The raw dict with every tensor on a CUDA device.  `.to()` keeps the strides of the permuted
    NCHW views, so the kernel still reads channel planes.  detach=False keeps the autograd graph (the
    training branches compute their loss on these tensors)."""
    if not torch.cuda.is_available():
        raise _lib.MydetError('mydetection_b200 needs a CUDA device (B200); there is no CPU fallback')
    out = {}
    for k in keys:
        t = raw[k].detach() if detach else raw[k]
        out[k] = t if t.is_cuda else t.to(torch.device('cuda', torch.cuda.current_device()))
    return out


def decode_level(kind, raw, stride, img_size, anchors=None, conf_key='conf', keys=('bbox', 'conf', 'class')):
    """One level -> the preds dict of the reference layers: bbox (B,N,P), class_idx (B,N) int64,
    score (B,N).  Unlike yolov3.py:51 / rapid.py:64 nothing is copied to the host here."""
    staged = stage_raw(raw, [k for k in keys if k in raw])
    levels = ops.LevelSet([staged], [stride], None if anchors is None else [anchors], conf_key)
    box, cls, score = ops.decode_dense(kind, levels, img_size)
    return {'bbox': box, 'class_idx': cls, 'score': score}


def pack_labels(labels, n_param, device):
    """List of per-image GT (objects with .bboxes (n,P) / .cats (n,)) -> padded device tensors
    (gt_box (B,G,P) f32, gt_cls (B,G) i64, gt_count (B) i32) with ONE host-to-device copy each."""
    n_b = len(labels)
    max_gt = max([len(l) for l in labels] + [1])
    gt_box = torch.zeros(n_b, max_gt, n_param, dtype=torch.float32)
    gt_cls = torch.zeros(n_b, max_gt, dtype=torch.int64)
    counts = torch.zeros(n_b, dtype=torch.int32)
    for b, l in enumerate(labels):
        n = len(l)
        if n:
            gt_box[b, :n] = l.bboxes.detach().cpu()[:, :n_param]
            gt_cls[b, :n] = l.cats.detach().cpu()
        counts[b] = n
    return gt_box.to(device), gt_cls.to(device), counts.to(device)


def last_writer(lin, n_cells):
    """Mask over the entries of `lin` (flat target-cell indices, in GT order) that are the LAST one aimed at their
    cell: the reference assigns targets GT by GT, so when two GTs share a cell the later one owns it.  A plain
    vectorised index assignment with duplicates has no defined order on the GPU."""
    order = torch.arange(lin.numel(), device=lin.device)
    winner = torch.full((n_cells,), -1, dtype=torch.int64, device=lin.device)
    winner.scatter_reduce_(0, lin, order, reduce='amax', include_self=True)
    return winner[lin] == order


def periodic_angle_loss(name):
    """models/losses.py:7-98 get_angle_loss(name, reduction='sum'): the angle difference is folded into
    [-pi/2, pi/2) (a box rotated by 180 degrees is the same box) before the L1 / L2 / smooth-L1 (beta 0.4) penalty."""
    import numpy as np

    def fold(pred, gt):
        return torch.remainder(pred - gt - np.pi / 2, np.pi) - np.pi / 2

    if name == 'Periodic_L1':
        return lambda p, g: torch.abs(fold(p, g)).sum()
    if name == 'Periodic_L2':
        return lambda p, g: (fold(p, g) ** 2).sum()
    if name == 'Periodic_smoothL1':
        def sl1(p, g, beta=0.4):
            n = torch.abs(fold(p, g))
            return torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta).sum()
        return sl1
    raise NotImplementedError()


def no_training(name):
    raise NotImplementedError(
        f'{name}: training-time target assignment is outside the post-processing hot path '
        '(SURVEY.md section 8f, rank 2); the YOLO, Ultralytics, FCOS2, FCOS2-ATSS, RetinaNet and RAPiD layers implement '
        'forward(..., labels)')
