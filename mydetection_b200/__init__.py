"""mydetection_b200 -- the detection post-processing hot path of duanzhiihao/myDetection, rebuilt
as hand-written CUDA for B200 (sm_100a) behind a C ABI (include/mydet.h).

Layout
    csrc/        CUDA kernels + the extern "C" entry points  -> libmydet.so (python -m mydetection_b200.build)
    _lib.py      ctypes binding (no fallback: missing library or failing call raises)
    ops.py       tensor-level wrappers
    bbox_ops.py, structures.py, detlayers/   host-side mirror of the reference's interface
                 (utils/bbox_ops.py, utils/structures.py, models/detlayers/*) for the hot path
    pipeline.py  batched decode -> threshold -> top-k -> NMS, image-sharded over GPUs
    dropin.py    registers the mirror under the reference's module names
"""
__version__ = '0.1.0'
