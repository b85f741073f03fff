"""Developer tool: per-phase SM-clock breakdown of postprocess_small_kernel (CTA 0) on the bench
workload.  `build` here (cross-compile with -DMYDET_PP_PROFILE), `run` on the GPU box."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIBP = os.path.join(ROOT, 'mydetection_b200', '_tune', 'libmydet_ppprof.so')
NAMES = ['A0 stage keys', 'A radix select', 'B1 slot assign', 'B2 gather loads', 'B3 class sort', 'corners', 'C mask',
         'D sweep', 'E output']

if sys.argv[1] == 'build':
    from mydetection_b200 import build as b
    os.makedirs(os.path.dirname(LIBP), exist_ok=True)
    b.build(force=True, extra_flags=['-DMYDET_PP_PROFILE'], out=LIBP)
else:
    import torch
    import bench
    from mydetection_b200 import _lib, pipeline as pl
    _lib.LIB_PATH = LIBP
    dev = torch.device('cuda', 0)
    gen = torch.Generator(device=dev).manual_seed(2000)
    wl = bench.Workload(640, 2.0)
    raws, _ = wl.make_batch(gen, dev)
    pipe = pl.DetectionPipeline('FCOS2', bench.STRIDES, bench.N_CLS, (wl.img, wl.img), bench.CONF_THRES,
                                bench.NMS_THRES, bench.TOPK)
    bc = pipe.bind(raws)
    for _ in range(3):
        bc.launch_decode(); bc.launch_postprocess()
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 32)()
    _lib.lib().mydet_debug_pp_clocks.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
    _lib.lib().mydet_debug_pp_clocks(buf)
    t = list(buf)
    # sampled front end (marks 0,16,17,18,3) instead of A0 / A / B1 (marks 0,1,2,3)
    print(f'{"S1 sample + edge":18s} {t[16] - t[0]:8d} cycles')
    print(f'{"S2 one pass":18s} {t[17] - t[16]:8d} cycles')
    print(f'{"S3 list select":18s} {t[18] - t[17]:8d} cycles')
    for i, nm in enumerate(NAMES):
        if i >= 3:
            print(f'{nm:18s} {t[i + 1] - t[i]:8d} cycles')
    print(f'{"total":18s} {t[9] - t[0]:8d} cycles')
