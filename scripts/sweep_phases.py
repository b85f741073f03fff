"""Developer tool: phase breakdown of the spatial sweep kernel (needs the -DMYDET_SWEEP_PROFILE build)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mydetection_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, 'mydetection_b200', '_tune', 'libmydet_sweepprof.so')
import torch, bench
print(bench.rotated_nms_metric(torch.device('cuda', 0), iters=1)['us_per_image'])
