"""Batched post-processing pipeline: decode -> score threshold -> per-image top-k -> per-class NMS
for a whole batch in one C call, optionally replayed as a CUDA graph, and image-sharded over the
GPUs of one node with a single gather of the final detections.

This is the B200-first shape of what the reference does one image and one level at a time
(api/detection.py:168-172 -> models/general.py:69-84 -> utils/structures.py:92-173).
"""
import torch

from . import _lib, ops

KINDS = {'YOLO': ops.KIND_YOLO, 'FCOS': ops.KIND_FCOS, 'FCOS2': ops.KIND_FCOS, 'FCOS2_ATSS': ops.KIND_FCOS,
         'RAPiD': ops.KIND_RAPID, 'RetinaNet': ops.KIND_RETINA, 'Ultralytics': ops.KIND_UV5}


class BoundCall:
    """The pipeline bound to fixed input tensors, outputs and workspace: launch() is one ctypes call."""

    def __init__(self, pipe, raws):
        self.pipe = pipe
        self.levels = ops.LevelSet(raws, pipe.strides, pipe.anchors, pipe.conf_key)
        ls = self.levels
        k = pipe.topk or 0
        cap = min(k, ls.n_total) if k > 0 else ls.n_total
        self.out = ops._alloc_dets(ls.batch, max(cap, 1), ls.n_param, ls.device)
        self.workspace = ops.detect_workspace(ls, pipe.topk)
        self.cand = None
        self.graph = None

    def launch(self):
        p = self.pipe
        return ops.detect(p.kind, self.levels, p.img_hw, p.conf_thres, p.nms_thres, p.topk, out=self.out,
                          workspace=self.workspace)

    # stage-wise entry points (same kernels as launch(); used to time the decode kernel alone)
    def launch_decode(self):
        p = self.pipe
        self.cand = ops.decode_compact(p.kind, self.levels, p.img_hw, p.conf_thres) if self.cand is None else \
            self._decode_into(self.cand)
        return self.cand

    def _decode_into(self, c):
        p, ls = self.pipe, self.levels
        with torch.cuda.device(ls.device):
            rc = _lib.lib().mydet_decode_compact(p.kind, ls.array, ls.n_levels, ls.batch, ls.n_cls, ls.n_param,
                                                 float(p.img_hw[0]), float(p.img_hw[1]), float(p.conf_thres),
                                                 ops._ptr(c['box']), ops._ptr(c['score']), ops._ptr(c['cls']),
                                                 ops._ptr(c['idx']), ops._ptr(c['count']), c['box'].shape[1],
                                                 ops._stream())
        _lib.check(rc, 'mydet_decode_compact')
        return c

    def launch_postprocess(self):
        c, p = self.cand, self.pipe
        return ops.postprocess(c['box'], c['score'], c['cls'], float('-inf'), p.nms_thres, topk=p.topk,
                               counts=c['count'], src_idx=c['idx'], out=self.out)

    def capture(self):
        """Record launch() into a CUDA graph (fixed shapes, fixed buffers)."""
        self.launch()                      # warm-up outside capture (lazy module load, attributes)
        torch.cuda.synchronize(self.levels.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.launch()
        self.graph = g
        return self

    def replay(self):
        self.graph.replay()
        return self.out


class DetectionPipeline:
    def __init__(self, pred_layer, strides, n_cls, img_hw, conf_thres, nms_thres, topk=512, anchors=None,
                 conf_key='conf'):
        if pred_layer not in KINDS:
            raise NotImplementedError(pred_layer)
        self.kind = KINDS[pred_layer]
        self.strides = list(strides)
        self.anchors = anchors
        self.n_cls = n_cls
        self.img_hw = tuple(img_hw)
        self.conf_thres, self.nms_thres, self.topk = float(conf_thres), float(nms_thres), topk
        self.conf_key = conf_key

    def bind(self, raws):
        return BoundCall(self, raws)

    def __call__(self, raws):
        """raws: list (one per level) of raw dicts of CUDA head views -> detections dict."""
        return self.bind(raws).launch()


def unpack(out, to_cpu=True):
    """Detections dict -> list of (boxes, scores, classes) per image (one D2H of the counts)."""
    counts = out['count'].tolist()
    res = []
    for b, n in enumerate(counts):
        item = (out['box'][b, :n], out['score'][b, :n], out['cls'][b, :n])
        res.append(tuple(t.cpu() for t in item) if to_cpu else item)
    return res


# --------------------------------------------------------------------------------------- multi-GPU
def shard_range(n_images, rank, world):
    """Contiguous block of ceil(n/world) images per rank."""
    per = (n_images + world - 1) // world
    lo = min(rank * per, n_images)
    return lo, min(lo + per, n_images)


def pack_detections(out):
    """(B,K,P) boxes + scores + classes -> one (B,K,P+2) float32 tensor for the exchange.
    Class ids < 2^24 are exact in float32."""
    return torch.cat([out['box'], out['score'].unsqueeze(-1), out['cls'].to(torch.float32).unsqueeze(-1)], dim=-1)


def gather_detections(out, group=None):
    """The path's only exchange: every rank receives every rank's final detections.
    Two collectives on fixed-capacity buffers (counts, padded detections), no host sync.
    Returns (packed (world*B, K, P+2), counts (world*B,))."""
    import torch.distributed as dist
    packed = pack_detections(out).contiguous()
    counts = out['count'].contiguous()
    world = dist.get_world_size(group)
    all_packed = packed.new_empty((world * packed.shape[0],) + tuple(packed.shape[1:]))
    all_counts = counts.new_empty(world * counts.shape[0])
    if dist.get_backend(group) == 'nccl':
        dist.all_gather_into_tensor(all_counts, counts, group=group)
        dist.all_gather_into_tensor(all_packed, packed, group=group)
    else:  # gloo (CPU tests of the host logic)
        dist.all_gather(list(all_counts.chunk(world)), counts, group=group)
        dist.all_gather(list(all_packed.chunk(world)), packed, group=group)
    return all_packed, all_counts
