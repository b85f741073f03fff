"""Build libmydet.so (hand-written CUDA for sm_100a behind the C ABI of include/mydet.h).

    python -m mydetection_b200.build [--force] [--verbose]

The library is built IN-TREE (mydetection_b200/libmydet.so) with nvcc; it links the CUDA runtime
statically, so it loads on a machine without a GPU (symbol checks) and travels with the repo
snapshot to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libmydet.so')
SOURCES = ['api.cu', 'decode.cu', 'postprocess_small.cu', 'nms_large.cu', 'iou.cu', 'atss.cu', 'preprocess.cu', 'exchange.cu', 'kalman.cu', 'raster.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',   # B200 only: no other arch, no PTX fallback
    '-O3', '-std=c++17', '-lineinfo',
    '-fmad=false',            # no silent FMA contraction: IoU / decode arithmetic must round like the reference
    '-Xcompiler', '-fPIC,-fvisibility=hidden,-O2',
    '--shared', '-cudart', 'static',
]


def _stale(srcs):
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'mydet.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=(), out=None):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    out = out or LIB
    if out == LIB and not force and not _stale(srcs):
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (['-Xptxas', '-v'] if verbose else []) + ['-o', out] + srcs
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed building libmydet.so')
    return out


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
