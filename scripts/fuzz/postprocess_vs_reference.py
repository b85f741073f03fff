"""One-off differential fuzz (CPU, needs the reference checkout): oracle/postprocess.py against the UNMODIFIED
ImageObjects.post_process / .nms (utils/structures.py:92-173) -- 150 cases, 1-3000 boxes, 1-80 classes, the three box
formats, confidence thresholds around the score quantiles, the top-512 cap on both sides.  Scores are tie-free (the
reference's torch.topk is unstable)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for sub in ('tests/golden', ''):
    sys.path.insert(0, os.path.join(ROOT, sub))
import torch                            # noqa: E402
import make_golden as m                 # noqa: E402

m.import_reference()
torch.set_grad_enabled(False)
from utils.structures import ImageObjects                # noqa: E402
from oracle import postprocess as opp                    # noqa: E402

bad = total = 0
for seed in range(150):
    gen = torch.Generator().manual_seed(9000 + seed)
    n = int(torch.randint(1, [40, 600, 3000][seed % 3], (1,), generator=gen))
    n_cls = [1, 3, 80][(seed // 3) % 3]
    fmt = ['cxcywh', 'cxcywhd'][(seed // 9) % 2]      # the reference's sanity_check rejects 'x1y1x2y2' objects
    span = [60.0, 400.0][seed % 2]
    b = torch.rand(n, 2, generator=gen) * span
    wh = torch.rand(n, 2, generator=gen) * 80 + 2
    if fmt == 'x1y1x2y2':
        boxes = torch.cat([b, b + wh], 1)
    elif fmt == 'cxcywhd':
        boxes = torch.cat([b, wh, torch.rand(n, 1, generator=gen) * 360 - 180], 1)
    else:
        boxes = torch.cat([b, wh], 1)
    scores = torch.rand(n, generator=gen)
    if scores.unique().numel() != n:
        continue
    cats = torch.randint(0, n_cls, (n,), generator=gen)
    conf = float([0.0, 0.05, 0.5, 0.95][seed % 4])
    nms = float([0.3, 0.45, 0.5, 0.7][(seed // 4) % 4])
    keys = torch.cat([boxes, scores[:, None], cats[:, None].float()], 1)
    for direct in (False, True):
        obj = ImageObjects(boxes.clone(), cats.clone(), scores=scores.clone(), bb_format=fmt, img_hw=(512, 512))
        res = obj.nms(nms) if direct else obj.post_process(conf, nms)
        got = torch.cat([res.bboxes, res.scores[:, None], res.cats[:, None].float()], 1)
        want_idx = torch.tensor([int(torch.nonzero((keys == g).all(dim=1))[0, 0]) for g in got], dtype=torch.int64)
        if direct:
            mine = opp.class_nms(boxes, scores, cats, nms, fmt)
        else:
            mine = opp.post_process(boxes, cats, scores, conf, nms, fmt, 512)
        total += 1
        if not torch.equal(mine, want_idx):
            bad += 1
            print('MISMATCH seed', seed, 'direct' if direct else 'post_process', n, n_cls, fmt, conf, nms, mine.numel(), want_idx.numel())
print('compared', total, 'bad', bad)
