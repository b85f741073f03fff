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


def _reference_structures(utils_pkg):
    """The reference's real utils/structures.py, loaded from disk under a private name inside the `utils` package
    (its relative imports `.bbox_ops`, `.tracking`, `.visualization`, `.constants`, `.kalman_filter` then resolve --
    `.bbox_ops` to the mirror).  None when the reference is not on disk."""
    import importlib.util
    import os
    for root in getattr(utils_pkg, '__path__', []):
        path = os.path.join(root, 'structures.py')
        if os.path.exists(path):
            name = 'utils._structures_reference'
            if name in sys.modules:
                return sys.modules[name]
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            mod.__package__ = 'utils'
            sys.modules[name] = mod
            try:
                spec.loader.exec_module(mod)
            except Exception:
                del sys.modules[name]
                return None
            return mod
    return None


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
    # everything of the reference's structures module that is NOT on the hot path keeps its own implementation:
    # the tracklet classes (utils/structures.py:295-529) and any ImageObjects method the mirror does not define
    ref = _reference_structures(utils_pkg)
    if ref is not None:
        for name, val in vars(ref).items():
            if isinstance(val, type) and val.__module__ == ref.__name__ and not hasattr(structures, name):
                setattr(structures, name, val)
        for name, val in vars(ref.ImageObjects).items():
            if not name.startswith('__') and name not in vars(structures.ImageObjects):
                setattr(structures.ImageObjects, name, val)
        ref.ImageObjects = structures.ImageObjects     # the tracklets' isinstance checks (:306, :344) see the class in use
    # importing `models` runs models/registry.py, whose det-layer imports are lazy (inside the function)
    models_pkg = _parent('models')
    models_pkg.detlayers = detlayers
    # the evaluator's matching IoU (utils/evaluation/cepdof.py): patched when that module is importable here
    from . import evaluation
    evaluation.install_cepdof()
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
