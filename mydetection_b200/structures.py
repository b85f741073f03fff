"""Host-side mirror of the reference's utils/structures.py::ImageObjects for the hot path.

Same constructor, attributes and method names; `post_process` / `nms` / `non_max_suppression`
(utils/structures.py:92-173) run in libmydet's CUDA kernels instead of boolean indexing +
torch.topk + a Python loop over classes around torchvision.ops.nms on the CPU.  The small helpers
the API layer calls (`category_filter_`, `mask_to_bbox_`, `draw_on_np`, `to_json`, ...) are here too;
the tracklet classes of the reference file (:295-529, host-side Kalman filters with no call site)
are re-exported from the reference's own file by `dropin.install()`.
"""
import torch

from . import _lib, ops

TOPK_CAP = 512  # utils/structures.py:99-101

# The reference post-processes a batch one image at a time (models/general.py:78-84 builds one ImageObjects per row of
# the level-concatenated (B,N,.) tensors, api/detection.py:172 / examples call post_process on each).  Each of those
# per-image tensors is a VIEW of the batch tensor, and the view metadata says which row it is -- so the first image's
# post_process runs ONE mydet_postprocess launch (and one device-to-host copy) for the whole batch, and the other
# images of the same batch take their rows from that result: no launch, no synchronisation.  One entry: the batch in
# flight.  It holds strong references to the three batch tensors (their addresses cannot be recycled while it lives)
# and is only honoured for the same tensors at the same version counters and the same thresholds.
_BATCH = {}


def _row_of_batch(t):
    """(batch tensor, row) when `t` is row `row` of a contiguous CUDA batch tensor (what iterating / unbinding /
    indexing dim 0 yields), else None."""
    base = t._base
    if base is None or not t.is_cuda or base.dim() != t.dim() + 1 or tuple(base.shape[1:]) != tuple(t.shape):
        return None
    per = t.numel()
    if per == 0 or base.shape[0] < 2 or not (t.is_contiguous() and base.is_contiguous()):
        return None
    off = t.storage_offset() - base.storage_offset()
    if off % per or not 0 <= off // per < base.shape[0]:
        return None
    return base, off // per


def _cuda_device():
    if not torch.cuda.is_available():
        raise _lib.MydetError('mydetection_b200 needs a CUDA device (B200); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


class ImageObjects():
    '''
    A group of image bounding boxes

    Args:
        bboxes: 2-d tensor, torch.float32
        cats: 1-d tensor, torch.int64, categories
        scores (optional): 1-d tensor, torch.float32, scores
        bb_format (optional): 'cxcywh' | 'x1y1x2y2' | 'cxcywhd' | 'cxcywhr'
        img_hw: tuple-like, image (height, width)
    '''
    def __init__(self, bboxes, cats, masks=None, scores=None, bb_format='cxcywh', img_hw=None):
        self.bboxes = bboxes
        self.cats = cats
        self.masks = masks
        self.scores = scores
        self._bb_format = bb_format
        self.img_hw = img_hw
        self.sanity_check()

    # ------------------------------------------------------------------ container protocol
    def __getitem__(self, idx):
        if isinstance(idx, int):
            idx = slice(idx, idx + 1)
        pick = lambda t: None if t is None else t[idx]
        return ImageObjects(self.bboxes[idx, :], self.cats[idx], pick(self.masks), pick(self.scores),
                            self._bb_format, self.img_hw)

    def __len__(self):
        return self.bboxes.shape[0]

    def cpu_(self):
        '''Move all attributes to CPU in-place'''
        for name in ('bboxes', 'cats', 'scores', 'masks'):
            t = getattr(self, name)
            if t is not None:
                setattr(self, name, t.cpu())

    def sanity_check(self):
        '''Integrity check (utils/structures.py:191-213).'''
        assert self.bboxes.dtype == torch.float and self.bboxes.dim() == 2
        if self._bb_format in ('cxcywh', 'x1y1x2y2'):
            assert self.bboxes.shape[1] == 4
        elif self._bb_format in ('cxcywhd', 'cxcywhr'):
            assert self.bboxes.shape[1] == 5
        else:
            raise NotImplementedError()
        assert self.cats.dtype == torch.int64, 'Incorrect data type of categories'
        assert self.cats.dim() == 1 and self.cats.shape[0] == self.bboxes.shape[0]
        if self.masks is not None:
            assert self.masks.dtype == torch.bool and self.masks.dim() == 3
        if self.scores is not None:
            assert self.scores.shape[0] == self.bboxes.shape[0]
        assert self.img_hw is None or len(self.img_hw) == 2

    def mask_to_bbox_(self):
        '''utils/structures.py:59-66: the reference checks its preconditions, walks the masks and updates
        nothing (the body was never finished); same observable behaviour here.'''
        assert self.masks is not None
        assert self._bb_format == 'cxcywh'

    def category_filter_(self, categories) -> None:
        '''Keep the objects of the given categories, in place (utils/structures.py:78-90).'''
        assert self.masks is None, 'filtering with masks is not currently supported'
        wanted = torch.as_tensor(list(categories), dtype=torch.int64, device=self.cats.device)
        assert self.cats.dim() == 1 and wanted.dim() == 1
        keep = torch.isin(self.cats, wanted)
        self.bboxes, self.cats = self.bboxes[keep], self.cats[keep]
        if self.scores is not None:
            self.scores = self.scores[keep]

    def draw_on_np(self, im, class_map='COCO', **kwargs):
        '''Draw the boxes on a numpy image in place (utils/structures.py:215-219).  Drawing is outside the hot
        path: this delegates to the reference's own utils.visualization (present wherever the drop-in is used).'''
        assert self.bboxes.dim() == 2
        from importlib import import_module
        import_module('utils.visualization').draw_bboxes_on_np(im, self, class_map=class_map, **kwargs)

    # ------------------------------------------------------------------ the hot path
    def _run(self, conf_thres, nms_thres, topk):
        assert self.masks is None, 'nms with masks is not currently supported'
        assert self.scores is not None
        if self._bb_format not in ops.BOX_FORMATS:
            raise NotImplementedError()
        if len(self) == 0:
            return self                      # structures.py:120-121
        hit = self._run_batched(conf_thres, nms_thres, topk) if topk else None    # the un-capped nms() stays per image
        if hit is not None:
            return hit
        home = self.bboxes.device
        dev = home if home.type == 'cuda' else _cuda_device()
        out = ops.postprocess(self.bboxes.detach().to(dev)[None], self.scores.detach().to(dev, torch.float32)[None],
                              self.cats.to(dev)[None], conf_thres, nms_thres, topk=topk, box_format=self._bb_format)
        n, status = (int(v) for v in torch.stack([out['count'][0], out['status'][0]]).tolist())  # one D2H sync
        self._check_status(status)
        return ImageObjects(out['box'][0, :n], out['cls'][0, :n], None, out['score'][0, :n], self._bb_format,
                            img_hw=self.img_hw), out['idx'][0, :n]

    @staticmethod
    def _check_status(status):
        if status & 1:
            raise _lib.MydetError(f'category ids must lie in [0, {_lib.MAX_CLASS_ID}]')
        if status & (2 | 4):   # cannot happen through this wrapper (out_cap and counts are derived from the input); never silent
            raise _lib.MydetError(f'mydet_postprocess reported status {status} (2 = output truncated, 4 = count > capacity)')
        if status & 8:
            import warnings
            warnings.warn('more than 2^20 candidates in one image: equal scores are ordered by the low 20 bits of the index')

    def _run_batched(self, conf_thres, nms_thres, topk, to_cpu=False):
        """This image as a row of its batch (see _BATCH): one launch per batch instead of one per image."""
        if self.bboxes.requires_grad or self.scores.requires_grad:
            return None
        rows = [_row_of_batch(t) for t in (self.bboxes, self.scores, self.cats)]
        if any(r is None for r in rows) or len({r[1] for r in rows}) != 1 or self.scores.dtype != torch.float32:
            return None
        (bb, b), (sc, _), (ct, _) = rows
        if not (bb.shape[0] == sc.shape[0] == ct.shape[0]):
            return None
        key = (bb.data_ptr(), sc.data_ptr(), ct.data_ptr(), bb._version, sc._version, ct._version, tuple(bb.shape),
               float(conf_thres), float(nms_thres), topk, self._bb_format)
        ent = _BATCH.get('entry')
        if ent is None or ent['key'] != key:
            out = ops.postprocess(bb.detach(), sc.detach(), ct, conf_thres, nms_thres, topk=topk, box_format=self._bb_format)
            meta = torch.stack([out['count'], out['status']]).cpu()               # the batch's one synchronisation
            ent = {'key': key, 'hold': (bb, sc, ct), 'out': out, 'count': meta[0].tolist(), 'status': meta[1].tolist(), 'host': None}
            _BATCH['entry'] = ent
        n = ent['count'][b]
        self._check_status(ent['status'][b])
        src = ent['out']
        if to_cpu:
            if ent['host'] is None:            # one device-to-host copy of the batch's rows, cut to the longest image
                m = max(ent['count'] + [1])
                ent['host'] = {k: src[k][:, :m].cpu() for k in ('box', 'cls', 'score', 'idx')}
            src = ent['host']
        # clones: the caller owns its result (bboxes_to_original_ edits it in place), the cached batch stays intact
        return ImageObjects(src['box'][b, :n].clone(), src['cls'][b, :n].clone(), None, src['score'][b, :n].clone(),
                            self._bb_format, img_hw=self.img_hw), src['idx'][b, :n].clone()

    def post_process(self, conf_thres, nms_thres):
        '''Confidence threshold + top-512 + per-class NMS (utils/structures.py:92-106).
        Like the reference, the result lives on the CPU.'''
        res = None
        if len(self) and self.masks is None and self.scores is not None and self._bb_format in ops.BOX_FORMATS:
            res = self._run_batched(conf_thres, nms_thres, TOPK_CAP, to_cpu=True)
        if res is None:
            res = self._run(conf_thres, nms_thres, TOPK_CAP)
        if res is self:
            return self
        dts, _ = res
        dts.cpu_()
        return dts

    def nms(self, nms_thres=0.45):
        return ImageObjects.non_max_suppression(self, nms_thres)

    @staticmethod
    def non_max_suppression(dts, nms_thres: float):
        '''Per-class NMS without threshold or cap (utils/structures.py:111-173); the result stays on
        the device of the input.'''
        assert isinstance(dts, ImageObjects)
        res = dts._run(float('-inf'), nms_thres, None)
        if res is dts:
            return dts
        out, _ = res
        home = dts.bboxes.device
        if home.type != 'cuda':
            out.cpu_()
        return out

    # ------------------------------------------------------------------ small helpers the API layer calls
    def bboxes_to_original_(self, pad_info):
        '''Undo the resize/pad of the input image (utils/structures.py:175-189).'''
        assert self.masks is None and len(pad_info) == 6
        ori_w, ori_h, tl_x, tl_y, imw, imh = pad_info
        self.bboxes[:, 0] = (self.bboxes[:, 0] - tl_x) / imw * ori_w
        self.bboxes[:, 1] = (self.bboxes[:, 1] - tl_y) / imh * ori_h
        self.bboxes[:, 2] = self.bboxes[:, 2] / imw * ori_w
        self.bboxes[:, 3] = self.bboxes[:, 3] / imh * ori_h
        self.img_hw = (ori_h, ori_w)

    def sort_by_score_(self, descending=True):
        assert self.scores is not None and self.masks is None
        order = torch.argsort(self.scores, descending=descending)
        self.bboxes, self.cats, self.scores = self.bboxes[order, :], self.cats[order], self.scores[order]

    def to_json(self, img_id, eval_type='x1y1wh', catIdx2id=None) -> list:
        '''COCO-like detection dicts {image_id, category_id, bbox, score} (utils/structures.py:221-259).'''
        assert self.bboxes.dim() == 2
        assert self.bboxes.shape[0] == self.cats.shape[0] == self.scores.shape[0]
        if eval_type == 'x1y1wh':
            assert self._bb_format == 'cxcywh'
        elif eval_type == 'cxcywhd':
            assert self._bb_format == 'cxcywhd'
        else:
            raise NotImplementedError()
        rows = []
        for bb, c, s in zip(self.bboxes.tolist(), self.cats.tolist(), self.scores.tolist()):
            if eval_type == 'x1y1wh':
                cx, cy, w, h = bb
                bb = [cx - w / 2, cy - h / 2, w, h]
            cat_id = catIdx2id[c] if catIdx2id is not None else COCO_CATEGORY_IDS[c]
            rows.append({'image_id': img_id, 'category_id': cat_id, 'bbox': bb, 'score': s})
        return rows


# category index -> COCO 2017 category id (the 'id' fields of the reference's utils/constants.py list)
# 80 "thing" ids followed by the 53 panoptic "stuff" ids
COCO_CATEGORY_IDS = ([*range(1, 12), *range(13, 26), 27, 28, *range(31, 45), *range(46, 66), 67, 70,
                      *range(72, 83), *range(84, 91)] +
                     [92, 93, 95, 100, 107, 109, 112, 118, 119, 122, 125, 128, 130, 133, 138, 141, 144, 145, 147,
                      148, 149, 151, 154, 155, 156, 159, 161, 166, 168, 171, *range(175, 179), 180, 181,
                      *range(184, 201)])
assert len(COCO_CATEGORY_IDS) == 133

from .tracking import TrackletBank  # noqa: E402,F401  batched counterpart of the reference's KFTracklet (utils/structures.py:447-529)
