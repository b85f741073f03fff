"""Developer tool: build decode-kernel variants (-D tunables) into scratch libraries and time the
compacting FCOS decode alone on the bench workload.  Usage (on a GPU box):
    python scripts/tune_decode.py build      # here, cross-compile the variants into build/tune/
    python scripts/tune_decode.py run        # on the GPU: time every variant found
"""
import ctypes
import glob
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, 'mydetection_b200', '_tune')

VARIANTS = [dict(w=w, u=u, m=m) for w, u, m in itertools.product((4, 8), (4, 8, 16), (4, 8))
            if w * 32 * m <= 2048]


def build():
    from mydetection_b200 import build as b
    os.makedirs(OUT, exist_ok=True)
    for v in VARIANTS:
        name = os.path.join(OUT, 'libmydet_w{w}_u{u}_m{m}.so'.format(**v))
        flags = [f'-DMYDET_DECODE_WARPS={v["w"]}', f'-DMYDET_CLS_UNROLL={v["u"]}', f'-DMYDET_DECODE_MINBLOCKS={v["m"]}']
        b.build(force=True, extra_flags=flags, out=name)
        print('built', name)


def run():
    import torch
    import bench
    from mydetection_b200 import _lib, pipeline as pl
    dev = torch.device('cuda', 0)
    gen = torch.Generator(device=dev).manual_seed(2000)
    batches = [bench.make_batch(gen, dev) for _ in range(3)]
    libs = sorted(glob.glob(os.path.join(OUT, '*.so'))) + [_lib.LIB_PATH]
    for path in libs:
        _lib._LIB = None
        _lib.LIB_PATH = path
        pipe = pl.DetectionPipeline('FCOS2', bench.STRIDES, bench.N_CLS, (bench.IMG, bench.IMG), bench.CONF_THRES,
                                    bench.NMS_THRES, bench.TOPK)
        bound = [pipe.bind(r, self_cleaning=False) for _, r, _ in batches]
        for i in range(6):
            bound[i % 3].launch_decode()
        torch.cuda.synchronize()
        n = 60
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            bound[i % 3].launch_decode()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / n * 1e3
        print(f'{os.path.basename(path):40s} {us:8.2f} us/launch  {200.78e6 / us / 1e3:8.1f} GB/s')


if __name__ == '__main__':
    {'build': build, 'run': run}[sys.argv[1]]()
