"""ctypes binding of libmydet.so (C ABI: include/mydet.h).

The library is loaded lazily and there is NO fallback: if it is missing, or a call returns a
non-zero status, a Python exception is raised.  Nothing in this package computes detections on
the CPU or with torch operators.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MYDET_LIB') or os.path.join(HERE, 'libmydet.so')   # MYDET_LIB: developer A/B runs of a variant build

MAX_LEVELS = 8
MAX_ANCHORS = 16
SMALL_K = 1024
MAX_CLASS_ID = 4095
MAX_CANDIDATES = 1048575

KIND_YOLO, KIND_FCOS, KIND_RAPID, KIND_RETINA, KIND_UV5 = range(5)
BOX_CXCYWH, BOX_X1Y1X2Y2 = 0, 1
INPUT_FORMATS = {'RGB_1': 0, 'RGB_1_norm': 1, 'BGR_255_norm': 2}   # utils/image_ops.py:174-186

c_int, c_i64, c_f32, c_f64, c_vp, c_sz = (ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double,
                                          ctypes.c_void_p, ctypes.c_size_t)


class Level(ctypes.Structure):
    """mydet_level_t"""
    _fields_ = [('bbox', c_vp), ('conf', c_vp), ('cls', c_vp),
                ('bbox_stride', c_i64 * 5), ('conf_stride', c_i64 * 4), ('cls_stride', c_i64 * 5),
                ('n_anchor', ctypes.c_int32), ('n_h', ctypes.c_int32), ('n_w', ctypes.c_int32),
                ('stride', c_f32), ('anchor_w', c_f32 * MAX_ANCHORS), ('anchor_h', c_f32 * MAX_ANCHORS)]


class AtssLevel(ctypes.Structure):
    """mydet_atss_level_t"""
    _fields_ = [('t_ltrb', c_vp), ('t_stride', c_i64 * 4), ('positive', c_vp), ('ignored', c_vp),
                ('target_ltrb', c_vp), ('target_conf', c_vp), ('target_cls', c_vp)]


# name -> (restype, argtypes); mirrors include/mydet.h one to one
SIGNATURES = {
    'mydet_version': (c_int, []),
    'mydet_last_error': (ctypes.c_char_p, []),
    'mydet_decode_dense': (c_int, [c_int, ctypes.POINTER(Level), c_int, c_int, c_int, c_int, c_f32, c_f32,
                                   c_vp, c_vp, c_vp, c_i64, c_vp]),
    'mydet_decode_compact': (c_int, [c_int, ctypes.POINTER(Level), c_int, c_int, c_int, c_int, c_f32, c_f32, c_f32,
                                     c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int32, c_int, c_vp]),
    'mydet_postprocess_workspace_bytes': (c_sz, [c_int, c_int, c_int]),
    'mydet_postprocess': (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_int, c_i64, c_int, c_int, c_int, c_f32,
                                  c_int, c_f64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_sz,
                                  c_int, c_vp]),
    'mydet_postprocess_scatter': (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_int, c_i64, c_int, c_int, c_int, c_f32,
                                          c_int, c_f64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_sz,
                                          ctypes.POINTER(c_vp), c_int, c_i64, c_i64, c_int, c_vp]),
    'mydet_postprocess_exchange': (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_int, c_i64, c_int, c_int, c_int, c_f32,
                                           c_int, c_f64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_sz,
                                           ctypes.POINTER(c_vp), c_int, c_vp, c_int, c_i64, c_i64, c_int, c_int, c_vp]),
    'mydet_exchange_buffer_bytes': (c_sz, [c_i64, c_int, c_int]),
    'mydet_exchange_wait': (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp]),
    'mydet_exchange_release': (c_int, [ctypes.POINTER(c_vp), c_int, c_vp, c_int, c_i64, c_int, c_int, c_vp]),
    'mydet_exchange_publish': (c_int, [ctypes.POINTER(c_vp), c_int, c_vp, c_int, c_i64, c_int, c_i64, c_int, c_int, c_vp]),
    'mydet_exchange_consume_counts': (c_int, [ctypes.POINTER(c_vp), c_int, c_vp, c_int, c_i64, c_int, c_i64, c_int, c_int, c_vp, c_vp, c_vp]),
    'mydet_detect_workspace_bytes': (c_sz, [c_int, c_i64, c_int, c_int]),
    'mydet_detect': (c_int, [c_int, ctypes.POINTER(Level), c_int, c_int, c_int, c_int, c_f32, c_f32, c_f32, c_int,
                             c_f64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_sz, c_vp]),
    'mydet_detect_ws': (c_int, [c_int, ctypes.POINTER(Level), c_int, c_int, c_int, c_int, c_f32, c_f32, c_f32, c_int,
                                c_f64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_sz, c_int, c_vp]),
    'mydet_pack_detections': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp]),
    'mydet_nms_rot_workspace_bytes': (c_sz, [c_int, c_int]),
    'mydet_nms_rot': (c_int, [c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_f64, c_int, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'mydet_nms_rot_ws': (c_int, [c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_f64, c_int, c_vp, c_vp, c_vp, c_vp, c_sz, c_int, c_vp]),
    'mydet_iou_aabb_pairwise': (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_vp]),
    'mydet_iou_aabb_rowmax': (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    'mydet_iou_rot_pairwise': (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp]),
    'mydet_iou_raster_workspace_bytes': (c_sz, [c_i64, c_i64, c_int]),
    'mydet_iou_raster_pairwise': (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_sz, c_vp]),
    'mydet_iou_rot_segments': (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_i64, c_vp, c_vp]),
    'mydet_cxcywh_to_x1y1x2y2': (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    'mydet_xywha2vertex': (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    'mydet_atss_workspace_bytes': (c_sz, [c_int, c_int]),
    'mydet_atss_assign': (c_int, [c_vp, ctypes.POINTER(c_i64), c_int, c_int, c_int, ctypes.POINTER(ctypes.c_int32),
                                  ctypes.POINTER(c_f32), c_int, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_f32, c_int,
                                  c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_sz, c_vp]),
    'mydet_atss_assign_levels': (c_int, [ctypes.POINTER(AtssLevel), c_int, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(c_f32),
                                         c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_f32, c_int, c_vp, c_vp, c_sz, c_vp]),
    'mydet_fcos_assign': (c_int, [c_vp, ctypes.POINTER(c_i64), c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int,
                                  c_f32, c_f32, c_f32, c_f32, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'mydet_kf_initiate': (c_int, [c_vp, c_int, ctypes.POINTER(c_f64), c_vp, c_vp, c_vp, c_vp]),
    'mydet_kf_predict': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, ctypes.POINTER(c_f64), c_vp, c_vp]),
    'mydet_kf_update': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, ctypes.POINTER(c_f64), c_vp, c_vp]),
    'mydet_kf_likelihood': (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_vp]),
    'mydet_preprocess_workspace_bytes': (c_sz, [c_int, c_int, c_int, c_int, c_int]),
    'mydet_preprocess': (c_int, [c_vp, c_int, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_vp, c_vp, c_sz, c_vp]),
}

_LIB = None


class MydetError(RuntimeError):
    pass


def lib():
    """Load libmydet.so (once).  Raises if the native library has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise MydetError(f'{LIB_PATH} is missing: build it with `python -m mydetection_b200.build` '
                             '(there is no CPU / torch fallback)')
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the library does not export it
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


def check(rc, what):
    if rc != 0:
        msg = lib().mydet_last_error().decode('utf-8', 'replace')
        kind = 'CUDA error' if rc > 0 else 'libmydet error'
        raise MydetError(f'{what}: {kind} {rc}: {msg}')
