"""Every kernel family of libmydet.so outside the two of the bench step, on the BASELINE workload that exercises it:
one JSON line per section with CUDA-event timings, the stated bound and the achieved figure against it.

    python scripts/kernel_zoo.py                 # all sections, timed (events, warm-up, L2-sized rotation where it matters)
    python scripts/kernel_zoo.py --once rot      # ONE un-timed invocation of a section: the command ncu wraps
Sections: rot, rotbench (bench.py's rotated workload), rot1 (single image), rotclu, atss, fcos, rowmax, iou, iourot, dense, pre, decode_yolo, decode_rapid
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from mydetection_b200 import ops, pipeline as pl, image_ops
from mydetection_b200.heads import yolo_head_views

DEV = torch.device('cuda', 0)
PEAKS = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
HBM = float(PEAKS.get('hbm_gbs', 6534.8))
ONCE = '--once' in sys.argv
ARGS = [a for a in sys.argv[1:] if not a.startswith('--')]
gen = torch.Generator().manual_seed(11)


def timed(fn, iters=10, warm=2):
    if ONCE:
        fn(0)
        torch.cuda.synchronize()
        return float('nan')
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3          # us


def emit(section, **kw):
    print(json.dumps(dict(section=section, **kw)), flush=True)


def rot_boxes(B, n, clustered=False):
    if not clustered:
        xy = torch.rand(B, n, 2, generator=gen) * 1024
        wh = torch.rand(B, n, 2, generator=gen) * 60 + 12
    else:                                           # 40 objects x 250 mutually overlapping boxes
        obj = torch.rand(B, 40, 1, 2, generator=gen) * 900 + 60
        xy = (obj + torch.randn(B, 40, n // 40, 2, generator=gen) * 6).reshape(B, n, 2)
        wh = (torch.rand(B, 40, 1, 2, generator=gen) * 80 + 40 + torch.randn(B, 40, n // 40, 2, generator=gen) * 4).reshape(B, n, 2).abs() + 4
    ang = torch.rand(B, n, 1, generator=gen) * 180 - 90
    return torch.cat([xy, wh, ang], dim=2).to(DEV), torch.rand(B, n, generator=gen).to(DEV)


def sec_rot(clustered=False, B=32):
    n = 10000
    sets = [rot_boxes(B, n, clustered) for _ in range(2)]
    us = timed(lambda i: ops.nms_rot(sets[i % 2][0], sets[i % 2][1], 0.45), iters=6)
    emit(('rotclu' if clustered else 'rot') + ('' if B == 32 else str(B)), workload=f'{B} images x {n} rotated boxes' + (' (40 objects x 250 boxes)' if clustered else ''),
         us_per_batch=us, us_per_image=us / B, bound='ALU / latency (pairwise)', algorithmic_pairs_per_s=B * n * (n - 1) / 2 / us * 1e6)


def sec_rotbench():
    """bench.py's own rotated workload: the best 10 000 of 64 512 RAPiD candidates @1024 per image."""
    import bench
    rb, rs = bench.rapid_boxes(DEV, 32, 10000)
    us = timed(lambda i: ops.nms_rot(rb, rs, 0.45), iters=6)
    emit('rotbench', workload='32 images x the best 10000 of 64512 RAPiD candidates @1024', us_per_batch=us, us_per_image=us / 32,
         bound='ALU / latency (pairwise)', algorithmic_pairs_per_s=32 * 10000 * 9999 / 2 / us * 1e6)


def sec_atss(fcos=False):
    strides, sides, img = [8, 16, 32, 64, 128], [24, 48, 96, 192, 384], (640, 640)
    B, G = 64, 100
    gt_box = torch.empty(B, G, 4)
    gt_box[..., 0:2] = torch.rand(B, G, 2, generator=gen) * 600 + 20
    gt_box[..., 2:4] = torch.rand(B, G, 2, generator=gen) * 200 + 16
    gb, gc = gt_box.to(DEV), torch.randint(0, 80, (B, G), generator=gen).to(DEV)
    gn = torch.full((B,), G, dtype=torch.int32, device=DEV)
    ts = [(torch.randn(B, 4, 640 // s, 640 // s, generator=gen) * 0.5).permute(0, 2, 3, 1).to(DEV) for s in strides]
    cells = sum((640 // s) ** 2 for s in strides)
    # bytes the call MUST move: the five target maps written once (C + 4 + 1 floats + 2 mask bytes per cell) + ltrb read
    algo = B * cells * ((80 + 4 + 1) * 4 + 2 + 16)

    def atss(i):
        thr = None
        for li in range(5):
            thr = ops.atss_assign(ts[li], li, strides, sides, img, gb, gc, gn, 9, 0.7, 80, thr=thr)['thr']

    def fc(i):
        lims = [0, 64, 128, 256, 512, 1e8]
        for li in range(5):
            ops.fcos_assign(ts[li], strides[li], img, gb, gc, gn, 0.5, lims[li], lims[li + 1], 0.7, 80)

    if not fcos:
        us_all = timed(lambda i: ops.atss_assign_levels(ts, strides, sides, img, gb, gc, gn, 9, 0.7, 80), iters=5, warm=1)
        emit('atss_all_levels', workload=f'batch {B} @640, {G} GT / image, 5 levels ({cells} cells), mydet_atss_assign_levels', us_per_batch=us_all,
             us_per_image=us_all / B, anchor_gt_pairs_per_s=B * cells * G / us_all * 1e6, algorithmic_bytes=algo,
             achieved_gbs=algo / us_all / 1e3, frac_of_hbm=algo / us_all / 1e3 / HBM)
    us = timed(fc if fcos else atss, iters=5, warm=1)
    emit('fcos' if fcos else 'atss', workload=f'batch {B} @640, {G} GT / image, 5 levels ({cells} cells)', us_per_batch=us,
         us_per_image=us / B, anchor_gt_pairs_per_s=B * cells * G / us * 1e6, bound='HBM write of the target maps (memset-like) + ALU for the pair tests',
         algorithmic_bytes=algo, achieved_gbs=algo / us / 1e3, frac_of_hbm=algo / us / 1e3 / HBM)


def sec_rowmax():
    B, n, G = 64, 8525 * 3, 100
    a = (torch.rand(B, n, 4, generator=gen) * 300 + 20).to(DEV)
    gt = (torch.rand(B, G, 4, generator=gen) * 300 + 20).to(DEV)
    us = timed(lambda i: ops.iou_rowmax(a, gt))
    algo = B * n * (16 + 12)
    emit('rowmax', workload=f'{B} x {n} boxes vs {G} GT', us_per_batch=us, pairs_per_s=B * n * G / us * 1e6, bound='ALU (100 IoUs per 28 B)',
         algorithmic_bytes=algo, achieved_gbs=algo / us / 1e3, frac_of_hbm=algo / us / 1e3 / HBM)


def sec_iou(rot=False):
    n, k = (8192, 2048) if rot else (16384, 8192)
    P = 5 if rot else 4
    a = torch.rand(n, P, generator=gen) * 200 + 20
    b = torch.rand(k, P, generator=gen) * 200 + 20
    a, b = a.to(DEV), b.to(DEV)
    us = timed(lambda i: (ops.iou_rot if rot else ops.iou_aabb)(a, b), iters=5)
    algo = n * k * (8 if rot else 4)
    emit('iourot' if rot else 'iou', workload=f'({n},{P}) x ({k},{P}) -> {"f64" if rot else "f32"} matrix', us_per_call=us,
         pairs_per_s=n * k / us * 1e6, bound='HBM (output write)' if not rot else 'ALU (polygon clip per pair) / HBM write',
         algorithmic_bytes=algo, achieved_gbs=algo / us / 1e3, frac_of_hbm=algo / us / 1e3 / HBM,
         note='includes the torch allocation of the output matrix')
    if not rot:                                     # anchor-vs-GT like: small boxes on a large canvas, almost every pair disjoint
        a3 = torch.cat([torch.rand(n, 2, generator=gen) * 1000, torch.rand(n, 2, generator=gen) * 60 + 4], 1).to(DEV)
        b3 = torch.cat([torch.rand(k, 2, generator=gen) * 1000, torch.rand(k, 2, generator=gen) * 60 + 4], 1).to(DEV)
        us3 = timed(lambda i: ops.iou_aabb(a3, b3), iters=5)
        emit('iou_sparse', workload=f'({n},4) x ({k},4), 4-64 px boxes on a 1000 px canvas', us_per_call=us3, pairs_per_s=n * k / us3 * 1e6,
             bound='HBM (output write)', algorithmic_bytes=algo, achieved_gbs=algo / us3 / 1e3, frac_of_hbm=algo / us3 / 1e3 / HBM)
        a4 = a3[:, :].contiguous(); b4 = b3[:k - 3].contiguous()          # k % 4 != 0: the scalar-store kernel
        us4 = timed(lambda i: ops.iou_aabb(a4, b4), iters=5)
        emit('iou_sparse_scalar', workload=f'({n},4) x ({k - 3},4)', us_per_call=us4, pairs_per_s=n * (k - 3) / us4 * 1e6)
    if not rot:                                     # the configs[3] shape: 8 525 anchors x 100 GT
        a2, b2 = a[:8525].contiguous(), b[:100].contiguous()
        us2 = timed(lambda i: ops.iou_aabb(a2, b2), iters=20)
        emit('iou_cfg4', workload='(8525,4) x (100,4)', us_per_call=us2, pairs_per_s=8525 * 100 / us2 * 1e6, bound='launch latency (3.4 MB out)')


def sec_dense():
    for img_s, B in ((704, 64), (1536, 8)):
        raws = []
        for s in (8, 16, 32):
            m = img_s // s
            t = torch.randn(B, 6, m, m, generator=gen) * 0.5
            t[:, 4] = torch.randn(B, m, m, generator=gen) * 1.5 + 2.0
            raws.append({k: v[:, 0].to(DEV) for k, v in yolo_head_views(t, 1, 4, 1).items()})
        pipe = pl.DetectionPipeline('FCOS2', (8, 16, 32), 1, (img_s, img_s), 0.005, 0.45, None)
        bc = pipe.bind(raws)
        n = bc.levels.n_total
        us = timed(lambda i: bc.launch(), iters=5, warm=1)
        emit('dense', workload=f'dense scene @{img_s}, batch {B}, {n} candidates / image, 1 class, no cap', us_per_batch=us,
             us_per_image=us / B, bound='ALU / latency (pairwise)', algorithmic_pairs_per_s=B * n * (n - 1) / 2 / us * 1e6)


def sec_pre():
    rng = np.random.default_rng(0)
    for batch, in_h, in_w, size in ((64, 1080, 1920, 608), (32, 2048, 2048, 1024)):
        dev = [torch.from_numpy(rng.integers(0, 256, (batch, in_h, in_w, 3), dtype=np.uint8)).to(DEV) for _ in range(2)]
        rs_h, rs_w, left, top, out_h, out_w, _ = image_ops.plan(in_h, in_w, 'resize_pad_square', size, 32)
        out = torch.empty(batch, 3, out_h, out_w, device=DEV)
        us = timed(lambda i: image_ops.preprocess(dev[i % 2], 'resize_pad_square', size, 32, 'RGB_1_norm', out=out))
        algo = batch * (in_h * in_w * 3 + out_h * out_w * 12)
        emit('pre', workload=f'{batch} x {in_h}x{in_w} uint8 -> {out_h}x{out_w} f32', us_per_batch=us, frames_per_s=batch / us * 1e6,
             bound='HBM', algorithmic_bytes=algo, achieved_gbs=algo / us / 1e3, frac_of_hbm=algo / us / 1e3 / HBM)
        del dev, out


def sec_decode(kind):
    YOLO_A = [[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198], [373, 326]]
    if kind == 'yolo':
        B, img, C, P, thr = 16, 608, 80, 4, 0.005
    else:
        B, img, C, P, thr = 32, 1024, 0, 5, 0.3
    sets = []
    for _ in range(3):
        raws = []
        for s in (8, 16, 32):
            m = img // s
            t = torch.randn(B, 3 * (P + 1 + C), m, m, generator=gen)
            raws.append({k: v.to(DEV) for k, v in yolo_head_views(t, 3, P, C).items()})
        sets.append(raws)
    groups = [YOLO_A[0:3], YOLO_A[3:6], YOLO_A[6:9]]
    pipe = pl.DetectionPipeline('YOLO' if kind == 'yolo' else 'RAPiD', (8, 16, 32), C, (img, img), thr, 0.45, 512, anchors=groups)
    bcs = [pipe.bind(r) for r in sets]
    cells = 3 * sum((img // s) ** 2 for s in (8, 16, 32))

    def dec(i):
        bcs[i % 3].reset_state()
        bcs[i % 3].launch_decode()

    us = timed(dec, iters=12, warm=3)
    torch.cuda.synchronize()
    cand = int(bcs[0].cand['count'].clamp(max=cells).sum()) if not ONCE else 0
    algo = B * cells * (P + 1 + C) * 4 + cand * (28 if P == 4 else 32)
    emit('decode_' + kind, workload=f'{kind} head @{img}, batch {B}: {cells} candidates / image, {C} classes; 3 rotating batches of {algo / 1e6:.0f} MB',
         us_per_launch=us, bound='HBM', algorithmic_bytes=algo, achieved_gbs=algo / us / 1e3, frac_of_hbm=algo / us / 1e3 / HBM,
         note='eager launch + a 4-byte-per-image memset; the lone-launch ramp is included')


def sec_eager():
    """The one-call path (mydet_detect: memset + decode + post-process on one stream), single image and batch 64:
    GPU time per call with the calls queued back to back, and host wall clock of a call that is waited for."""
    import time
    YOLO_A = [[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198], [373, 326]]
    for name, B in (('yolo', 1), ('yolo', 16), ('fcos2', 1), ('fcos2', 64)):
        if name == 'yolo':
            img, C, P = 608, 80, 4
            raws = []
            for s_ in (8, 16, 32):
                m = img // s_
                t = torch.randn(B, 3 * (P + 1 + C), m, m, generator=gen)
                raws.append({k: v.to(DEV) for k, v in yolo_head_views(t, 3, P, C).items()})
            pipe = pl.DetectionPipeline('YOLO', (8, 16, 32), C, (img, img), 0.005, 0.45, 512, anchors=[YOLO_A[0:3], YOLO_A[3:6], YOLO_A[6:9]])
        else:
            import bench
            wl = bench.Workload(640, 2.0)
            g2 = torch.Generator(device=DEV).manual_seed(5)
            raws, _ = wl.make_batch(g2, DEV)
            raws = [{k: v[:B] for k, v in r.items()} for r in raws]
            pipe = pl.DetectionPipeline('FCOS2', bench.STRIDES, bench.N_CLS, (wl.img, wl.img), bench.CONF_THRES, bench.NMS_THRES, bench.TOPK)
        bc = pipe.bind(raws)
        us = timed(lambda i: bc.launch(), iters=50, warm=5)
        torch.cuda.synchronize()
        walls = []
        for _ in range(50):
            t0 = time.perf_counter()
            bc.launch()
            torch.cuda.synchronize()
            walls.append((time.perf_counter() - t0) * 1e6)
        walls.sort()
        emit('eager', workload=f'{name} batch {B}: mydet_detect (memset + decode + post-process)', us_per_call_queued=us,
             us_wall_sync_median=walls[len(walls) // 2], us_wall_sync_min=walls[0], pdl=os.environ.get('MYDET_PDL', '1'))


SECTIONS = {'rot': sec_rot, 'rotbench': sec_rotbench, 'rot1': lambda: sec_rot(False, 1), 'rotclu': lambda: sec_rot(True), 'atss': sec_atss, 'fcos': lambda: sec_atss(True), 'rowmax': sec_rowmax,
            'iou': sec_iou, 'iourot': lambda: sec_iou(True), 'dense': sec_dense, 'pre': sec_pre,
            'eager': sec_eager, 'decode_yolo': lambda: sec_decode('yolo'), 'decode_rapid': lambda: sec_decode('rapid')}

if __name__ == '__main__':
    for name in (ARGS or list(SECTIONS)):
        SECTIONS[name]()
        torch.cuda.empty_cache()
