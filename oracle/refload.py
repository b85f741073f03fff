"""Loader for the UNMODIFIED reference copied by oracle/fetch_ref.py.  TEST INFRASTRUCTURE ONLY.

    from oracle import refload
    if refload.available():
        refload.activate()                   # oracle/_ref on sys.path + stubs for the three absent third-party imports
        from models.general import OneStageBBox          # the reference's own module, byte-identical to /root/reference

Used by tests/ (the drop-in runs under the unmodified callers; the reference runs beside the CUDA path on the same
inputs) and by bench.py's CPU arm (`--impl reference`, `cpu_baseline`), which times the reference's own det-layer
`forward` + `ImageObjects.post_process`.  Never read from the product path, and never reads /root/reference.

Third-party modules the reference imports and this image lacks are answered with stand-ins:
  * `pycocotools(.mask/.coco/.cocoeval)`: imported at module scope by utils/bbox_ops.py:3; only `iou_rle`/`nms_rotbb`
    call it.  The stand-in's `mask.iou` is the oracle's exact polygon clipping (oracle/rotiou.c) -- rotated IoU VALUES
    therefore stay "parity unpinned" (DESIGN.md section 3); everything axis-aligned never touches it.
  * `matplotlib(.pyplot)`: imported by api/detection.py:5, utils/visualization.py, utils/image_ops.py; plotting only.
  * `fvcore.nn`: `smooth_l1_loss` of RetinaLayer's training loss, restated from its published definition.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

from . import fetch_ref

ROOT = fetch_ref.DST
_PKGS = ('api', 'models', 'utils', 'external', 'settings')


def available():
    return os.path.exists(os.path.join(ROOT, fetch_ref.MANIFEST))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__mydet_stub__ = True
    sys.modules[name] = m
    return m


def _have(name):
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


class _MiniParams:
    """The fields of pycocotools.cocoeval.Params that CEPDOFeval touches (defaults of the published class)."""
    def __init__(self, iouType='bbox'):
        self.iouType, self.useCats, self.maxDets = iouType, 1, [1, 10, 100]
        self.imgIds, self.catIds = [], []
        self.iouThrs = np.linspace(.5, 0.95, 10)
        self.areaRngLbl = ['all', 'small', 'medium', 'large']


class _MiniCOCOeval:
    """Stand-in base class for CEPDOFeval when pycocotools is absent: `evaluate()` up to and including the IoU table
    (`self.ious[(imgId, catId)] = self.computeIoU(imgId, catId)`, as the published COCOeval.evaluate does); matching
    and accumulation are third-party code outside the path and are not restated."""
    def evaluate(self):
        p = self.params
        p.imgIds, p.catIds = list(np.unique(p.imgIds)), list(np.unique(p.catIds))
        self._prepare()
        cat_ids = p.catIds if p.useCats else [-1]
        self.ious = {(i, c): self.computeIoU(i, c) for i in p.imgIds for c in cat_ids}


def install_stubs():
    """Stand-ins for pycocotools / matplotlib / fvcore -- only for those that are really absent."""
    if not _have('pycocotools'):
        def fr(polys, h, w):
            return polys

        def iou(d, g, crowd):
            from . import iou as oi
            import ctypes
            f64p = ctypes.POINTER(ctypes.c_double)
            a = np.ascontiguousarray(np.asarray(d, dtype=np.float64).reshape(-1, 8))
            b = np.ascontiguousarray(np.asarray(g, dtype=np.float64).reshape(-1, 8))
            out = np.empty((a.shape[0], b.shape[0]), dtype=np.float64)
            oi.lib().oracle_quad_iou_pairwise(a.ctypes.data_as(f64p), a.shape[0], b.ctypes.data_as(f64p), b.shape[0],
                                              out.ctypes.data_as(f64p))
            return out
        mask = _mod('pycocotools.mask', frPyObjects=fr, iou=iou)
        _mod('pycocotools', mask=mask)
        _mod('pycocotools.coco', COCO=object)
        _mod('pycocotools.cocoeval', COCOeval=_MiniCOCOeval, Params=_MiniParams)
    if not _have('matplotlib'):
        plt = _mod('matplotlib.pyplot')
        _mod('matplotlib', pyplot=plt)
    if not _have('fvcore'):
        def smooth_l1_loss(input, target, beta, reduction='none'):
            n = torch.abs(input - target)
            loss = torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta) if beta >= 1e-5 else n
            return loss.sum() if reduction == 'sum' else (loss.mean() if reduction == 'mean' else loss)
        nn = _mod('fvcore.nn', smooth_l1_loss=smooth_l1_loss, sigmoid_focal_loss=None)
        _mod('fvcore', nn=nn)


def purge():
    """Forget every imported reference module (and any drop-in alias), so the next import starts from disk."""
    for name in list(sys.modules):
        if name.split('.')[0] in _PKGS:
            del sys.modules[name]


def activate(fresh=False):
    """Make `import models`, `import utils`, `import api` resolve to oracle/_ref.  fresh=True drops cached modules
    first (e.g. after a test installed the drop-in aliases)."""
    if not available():
        raise RuntimeError('oracle/_ref is absent: run `python -m oracle.fetch_ref` where /root/reference exists')
    install_stubs()
    if fresh:
        purge()
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import warnings
    warnings.filterwarnings('ignore', message='.*meshgrid.*')
    return ROOT


def deactivate():
    purge()
    while ROOT in sys.path:
        sys.path.remove(ROOT)


def build_model(cfg_name, seed=2024, device='cpu'):
    """OneStageBBox(configs/<cfg_name>.json) of the reference with RANDOM weights.  The two places that would fetch
    pre-trained weights (models/registry.py:15 torch.load of weights/dark53_imgnet.pth, absent; external/efficientnet's
    download) are answered with the freshly initialised parameters.  Returns (model.eval(), cfg)."""
    import json
    from models.general import OneStageBBox
    from models.backbones import Darknet53
    import external.efficientnet.model as efn_model
    cfg = json.load(open(os.path.join(ROOT, 'configs', cfg_name + '.json')))
    torch.manual_seed(seed)
    real_load, real_pre = torch.load, efn_model.load_pretrained_weights
    torch.load = lambda path, *a, **k: (Darknet53(cfg).state_dict() if str(path).endswith('dark53_imgnet.pth')
                                        else real_load(path, *a, **k))
    efn_model.load_pretrained_weights = lambda *a, **k: None
    try:
        model = OneStageBBox(cfg).eval()
    finally:
        torch.load, efn_model.load_pretrained_weights = real_load, real_pre
    return model.to(device), cfg
