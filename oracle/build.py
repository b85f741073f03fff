"""Build recipe for the C part of the oracle.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.build        # -> oracle/_build/liboracle.so

The reference (duanzhiihao/myDetection) is pure Python, so there is nothing of
its own to compile into oracle/_ref/ ("Oracle"): oracle/_ref/ stays
empty and the CPU baseline kind is "port".
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'liboracle.so')
SOURCES = ['nms.c', 'rotiou.c', 'raster.c']


def build(force=False):
    srcs = [os.path.join(HERE, s) for s in SOURCES]
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs)):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ['gcc', '-O2', '-ffp-contract=off', '-fPIC', '-shared', '-Wall', '-o', LIB] + srcs + ['-lm']
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
