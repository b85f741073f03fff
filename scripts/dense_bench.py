"""Developer tool: the dense-scene configuration (BASELINE configs[4]: single class, no top-k cap) on its own.
    python scripts/dense_bench.py [img_size] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mydetection_b200 import _lib, pipeline as pl
from mydetection_b200.heads import yolo_head_views
if os.environ.get('MYDET_LIB'):
    _lib.LIB_PATH = os.environ['MYDET_LIB']

img_s = int(sys.argv[1]) if len(sys.argv) > 1 else 704
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device('cuda', 0)
gen = torch.Generator().manual_seed(1)
raws = []
for s in (8, 16, 32):
    n = img_s // s
    t = torch.randn(B, 6, n, n, generator=gen) * 0.5
    t[:, 4] = torch.randn(B, n, n, generator=gen) * 1.5 + 2.0
    raws.append({k: v[:, 0].to(dev) for k, v in yolo_head_views(t, 1, 4, 1).items()})
pipe = pl.DetectionPipeline('FCOS2', (8, 16, 32), 1, (img_s, img_s), 0.005, 0.45, None)
bc = pipe.bind(raws)
for _ in range(2):
    bc.launch()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    out = bc.launch()
b.record()
torch.cuda.synchronize()
print(f'dense @{img_s} x {B}: {a.elapsed_time(b) / 5 * 1e3 / B:.1f} us / image, {bc.levels.n_total} candidates, '
      f'kept mean {float(out["count"].float().mean()):.0f}')
