"""Developer tool: phase breakdown of the sweep kernels (needs the -DMYDET_SWEEP_PROFILE build:
python -c "from mydetection_b200 import build; build.build(extra_flags=['-DMYDET_SWEEP_PROFILE'], out='mydetection_b200/_tune/libmydet_sweepprof.so')")."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mydetection_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, 'mydetection_b200', '_tune', 'libmydet_sweepprof.so')
import torch
from mydetection_b200 import ops
gen = torch.Generator().manual_seed(11)
n = 10000
b = torch.cat([torch.rand(1, n, 2, generator=gen) * 1024, torch.rand(1, n, 2, generator=gen) * 60 + 12, torch.rand(1, n, 1, generator=gen) * 180 - 90], dim=2).cuda()
s = torch.rand(1, n, generator=gen).cuda()
for _ in range(3):
    ops.nms_rot(b, s, 0.45)
    torch.cuda.synchronize()
