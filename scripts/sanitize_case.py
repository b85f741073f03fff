"""Small end-to-end case for compute-sanitizer (racecheck / memcheck): every kernel family once."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mydetection_b200 import ops, pipeline as pl
from mydetection_b200.heads import efdet_head_views

d = torch.device('cuda', 0)
g = torch.Generator().manual_seed(1)
strides, img = (8, 16, 32, 64, 128), (128, 256)
raws = []
for s in strides:
    bb = torch.randn(2, 4, img[0] // s, img[1] // s, generator=g) * 0.5
    cc = torch.randn(2, 1 + 5, img[0] // s, img[1] // s, generator=g) * 1.5 + 1.0
    raws.append({k: v.to(d) for k, v in efdet_head_views(bb, cc).items()})
out = pl.DetectionPipeline('FCOS2', strides, 5, img, 0.05, 0.5, 128)(raws)            # radix select + small path
n = 1500
boxes = torch.cat([torch.rand(2, n, 2, generator=g) * 300, torch.rand(2, n, 2, generator=g) * 60 + 4], 2).to(d)
scores = torch.rand(2, n, generator=g).to(d)
cls = torch.randint(0, 3, (2, n), generator=g).to(d)
big = ops.postprocess(boxes, scores, cls, 0.1, 0.5, topk=None)                        # large AABB path
rb = torch.cat([boxes, (torch.rand(2, n, 1, generator=g) * 360 - 180).to(d)], 2).contiguous()
keep, cnt, votes = ops.nms_rot(rb, scores, 0.3, want_votes=True)                      # rotated path + votes
iou = ops.iou_rot(rb[0, :200], rb[1, :100])
torch.cuda.synchronize()
print('ok', out['count'].tolist(), big['count'].tolist(), cnt.tolist(), float(iou.max()))
