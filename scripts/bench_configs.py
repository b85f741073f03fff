"""Timing of the BASELINE.json configurations other than the headline one (configs 0, 2, 3, 4) on one GPU,
each beside the oracle port on the host cores.  Prints a markdown table (kept in profiles/).
    python scripts/bench_configs.py            # on a GPU box
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mydetection_b200 import ops, pipeline as pl
from mydetection_b200.heads import yolo_head_views, efdet_head_views
from oracle import decode as od, postprocess as opp, atss as oa

DEV = torch.device('cuda', 0)
YOLO_A = [[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198], [373, 326]]
RAPID_A = [[18.7807, 33.4659], [28.8912, 61.7536], [48.6849, 68.3897], [45.0668, 101.4673], [63.0952, 113.5382],
           [81.3909, 134.4554], [91.7364, 144.9949], [137.5189, 178.4791], [194.4429, 250.7985]]


def gpu_time(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3      # us


def cpu_time(fn, reps=2):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6


def to_dev(raws):
    return [{k: v.to(DEV) for k, v in r.items()} for r in raws]


rows = []
torch.set_num_threads(os.cpu_count())
gen = torch.Generator().manual_seed(1)

# ---- config 0: YOLOv3-80 head, one 608x608 image
raws = []
for s in (8, 16, 32):
    n = 608 // s
    t = torch.randn(1, 255, n, n, generator=gen)
    t.view(1, 3, 85, n, n)[:, :, 5:] -= 2.0
    raws.append(yolo_head_views(t, 3, 4, 80))
groups = [YOLO_A[0:3], YOLO_A[3:6], YOLO_A[6:9]]
pipe = pl.DetectionPipeline('YOLO', (8, 16, 32), 80, (608, 608), 0.005, 0.45, 512, anchors=groups)
bc = pipe.bind(to_dev(raws))
g_us = gpu_time(bc.launch)


def cpu0():
    ref = od.merge_levels([od.decode_yolo(r, torch.tensor(a, dtype=torch.float32), s, 80)
                           for r, a, s in zip(raws, groups, (8, 16, 32))])
    opp.post_process(ref[0][0], ref[1][0], ref[2][0], 0.005, 0.45, 'cxcywh', 512)


rows.append(('0 YOLOv3-80 @608, 1 image (22 743 candidates)', 'decode + post_process', 1, g_us, cpu_time(cpu0, 5)))

# ---- config 2: RAPiD @1024, batch 32: decode + post_process (AABB, top-512)
B = 32
raws = []
for s in (8, 16, 32):
    n = 1024 // s
    t = torch.randn(B, 18, n, n, generator=gen) * 0.5
    v = t.view(B, 3, 6, n, n)
    v[:, :, 4] = torch.rand(B, 3, n, n, generator=gen) * 6 - 3
    v[:, :, 5] = torch.randn(B, 3, n, n, generator=gen) * 1.5 - 1.5
    raws.append(yolo_head_views(t, 3, 5, 0))
groups = [RAPID_A[0:3], RAPID_A[3:6], RAPID_A[6:9]]
pipe = pl.DetectionPipeline('RAPiD', (8, 16, 32), 0, (1024, 1024), 0.3, 0.45, 512, anchors=groups)
bc = pipe.bind(to_dev(raws))
g_us = gpu_time(bc.launch)


def cpu2():
    sub = [{k: v[:4] for k, v in r.items()} for r in raws]
    ref = od.merge_levels([od.decode_rapid(r, torch.tensor(a, dtype=torch.float32), s, 0)
                           for r, a, s in zip(sub, groups, (8, 16, 32))])
    for b in range(4):
        opp.post_process(ref[0][b], ref[1][b], ref[2][b], 0.3, 0.45, 'cxcywhd', 512)


rows.append(('2 RAPiD @1024, batch 32 (64 512 candidates / image)', 'xywha decode + post_process', B, g_us, cpu_time(cpu2) * 8))

# ---- config 3: ATSS assignment, 100 GT / image, 5 levels @640, batch 64
strides, sides, img = [8, 16, 32, 64, 128], [24, 48, 96, 192, 384], (640, 640)
B, G = 64, 100
gt_box = torch.empty(B, G, 4)
gt_box[..., 0:2] = torch.rand(B, G, 2, generator=gen) * 600 + 20
gt_box[..., 2:4] = torch.rand(B, G, 2, generator=gen) * 200 + 16
gt_cls = torch.randint(0, 80, (B, G), generator=gen)
cnt = torch.full((B,), G, dtype=torch.int32)
ts = [(torch.randn(B, 4, 640 // s, 640 // s, generator=gen) * 0.5).permute(0, 2, 3, 1) for s in strides]
ts_d = [t.to(DEV) for t in ts]
gb, gc, gn = gt_box.to(DEV), gt_cls.to(DEV), cnt.to(DEV)


def gpu3():
    thr = None
    for li in range(5):
        thr = ops.atss_assign(ts_d[li], li, strides, sides, img, gb, gc, gn, 9, 0.7, 80, thr=thr)['thr']


g_us = gpu_time(gpu3, iters=5, warm=1)


def cpu3():
    gts = [(gt_box[0], gt_cls[0])]
    for li in range(5):
        oa.assign_level(li, ts[li][:1], gts, img, strides, sides, 9, 0.7, 80)


rows.append(('3 D1 + FCOS2 + ATSS @640, 100 GT / image, batch 64', 'ATSS targets of all 5 levels', B, g_us, cpu_time(cpu3, 1) * B))

# ---- config 4: dense scenes, single class, no top-k cap
for img_s, B in ((704, 64), (1024, 32), (1536, 8)):
    raws = []
    for s in (8, 16, 32):
        n = img_s // s
        t = torch.randn(B, 6, n, n, generator=gen) * 0.5
        t[:, 4] = torch.randn(B, n, n, generator=gen) * 1.5 + 2.0
        raws.append({k: v[:, 0] for k, v in yolo_head_views(t, 1, 4, 1).items()})
    pipe = pl.DetectionPipeline('FCOS2', (8, 16, 32), 1, (img_s, img_s), 0.005, 0.45, None)
    bc = pipe.bind(to_dev(raws))
    n_total = bc.levels.n_total
    g_us = gpu_time(bc.launch, iters=5, warm=1)

    def cpu4():
        sub = [{k: v[:1] for k, v in r.items()} for r in raws]
        ref = od.merge_levels([od.decode_fcos(r, s, (img_s, img_s)) for r, s in zip(sub, (8, 16, 32))])
        opp.post_process(ref[0][0], ref[1][0], ref[2][0], 0.005, 0.45, 'cxcywh', None)

    rows.append((f'4 dense scene @{img_s}, batch {B} ({n_total} candidates / image, 1 class, no cap)',
                 'decode + un-capped NMS', B, g_us, cpu_time(cpu4, 1) * B))

print('| config | what | batch | GPU us / batch | GPU us / image | oracle port on %d host threads, us / image | ratio |' % torch.get_num_threads())
print('|---|---|---|---|---|---|---|')
for name, what, b, g, c in rows:
    print(f'| {name} | {what} | {b} | {g:.0f} | {g / b:.1f} | {c / b:.0f} | {c / g:.0f}x |')
