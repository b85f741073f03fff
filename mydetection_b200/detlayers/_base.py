"""Shared plumbing of the det-layer mirrors: stage the raw views on the GPU (no copy when they
already are) and call the one-launch dense decode."""
import torch

from .. import _lib, ops


def stage_raw(raw, keys):
    """The raw dict with every tensor on a CUDA device.  `.to()` keeps the strides of the permuted
    NCHW views, so the kernel still reads channel planes."""
    if not torch.cuda.is_available():
        raise _lib.MydetError('mydetection_b200 needs a CUDA device (B200); there is no CPU fallback')
    out = {}
    for k in keys:
        t = raw[k].detach()
        out[k] = t if t.is_cuda else t.to(torch.device('cuda', torch.cuda.current_device()))
    return out


def decode_level(kind, raw, stride, img_size, anchors=None, conf_key='conf', keys=('bbox', 'conf', 'class')):
    """One level -> the preds dict of the reference layers: bbox (B,N,P), class_idx (B,N) int64,
    score (B,N).  Unlike yolov3.py:51 / rapid.py:64 nothing is copied to the host here."""
    staged = stage_raw(raw, [k for k in keys if k in raw])
    levels = ops.LevelSet([staged], [stride], None if anchors is None else [anchors], conf_key)
    box, cls, score = ops.decode_dense(kind, levels, img_size)
    return {'bbox': box, 'class_idx': cls, 'score': score}


def no_training(name):
    raise NotImplementedError(
        f'{name}: training-time target assignment is outside the post-processing hot path '
        '(SURVEY.md section 8f, rank 2); only the ATSS layer implements forward(..., labels)')
