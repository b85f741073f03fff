"""Register the mirror modules under the reference's module names, so that reference code that
does `from utils.bbox_ops import ...`, `from utils.structures import ImageObjects` or
`models.registry.get_det_layer(cfg)` (api/detection.py, models/general.py, examples/eval_det.py)
picks up the CUDA path without being edited.  See INTEGRATION.md.

    import mydetection_b200.dropin as dropin
    dropin.install()            # before importing the reference's `models` / `api`
"""
import sys
import types


def _parent(name):
    """The real parent package if the reference is importable, else an empty stand-in."""
    import importlib
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        pkg = types.ModuleType(name)
        pkg.__path__ = []
        sys.modules[name] = pkg
        return pkg


def install():
    """Alias the mirror modules as utils.bbox_ops, utils.structures and models.detlayers[.*].

    Only these sub-modules are replaced: with the reference root on sys.path the rest of its `utils`
    and `models` packages (image_ops, registry, general, ...) keeps importing from disk, and
    `models.registry.get_det_layer` (:119-146) resolves its `from .detlayers.X import Y` to the mirror.
    """
    from . import bbox_ops, structures, detlayers
    from .detlayers import yolov3, fcos, fcos2, rapid, retinanet, uv5

    sys.modules['utils.bbox_ops'] = bbox_ops
    sys.modules['utils.structures'] = structures
    sys.modules['models.detlayers'] = detlayers
    for name, mod in (('yolov3', yolov3), ('fcos', fcos), ('fcos2', fcos2), ('rapid', rapid),
                      ('retinanet', retinanet), ('uv5', uv5)):
        sys.modules['models.detlayers.' + name] = mod
    utils_pkg = _parent('utils')
    utils_pkg.bbox_ops, utils_pkg.structures = bbox_ops, structures
    # importing `models` runs models/registry.py, whose det-layer imports are lazy (inside the function)
    models_pkg = _parent('models')
    models_pkg.detlayers = detlayers
    return {'utils.bbox_ops': bbox_ops, 'utils.structures': structures, 'models.detlayers': detlayers}


def install_preprocess(detector_cls=None):
    """Route Detector._predict_pil (api/detection.py:142-175) through the device pre-processing of
    mydetection_b200.image_ops (resize / pad / normalise in libmydet instead of Pillow on the CPU).
    `detector_cls` defaults to the reference's api.detection.Detector (the reference root must be on sys.path).
    Returns the patched class; detect_one / evaluation_predict / detect_imgs call _predict_pil unchanged."""
    from . import image_ops
    if detector_cls is None:
        import importlib
        detector_cls = importlib.import_module('api.detection').Detector
    detector_cls._predict_pil = image_ops.predict_pil
    return detector_cls
