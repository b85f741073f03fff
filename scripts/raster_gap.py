"""Report how far greedy rotated NMS on the exact polygon IoU (what libmydet computes) lies from greedy NMS on a
rasterised IoU in the manner of the reference's pycocotools path (oracle/raster.c, PARITY UNPINNED).  CPU only.
DESIGN.md section 3 quotes the output."""
import sys, time, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import iou as oi
def greedy(iou, order, thr):
    keep=[]
    for i in order:
        if not keep or not (iou[i, keep] >= thr).any(): keep.append(i)
    return keep
gen = torch.Generator().manual_seed(3)
for name, lo, hi, span, n in (('RAPiD-like 20-120 px', 20, 120, 1024, 1500), ('small 8-30 px', 8, 30, 400, 1500), ('large 60-250 px', 60, 250, 1500, 1500)):
    b = torch.cat([torch.rand(n,2,generator=gen)*span+100, torch.rand(n,2,generator=gen)*(hi-lo)+lo, torch.rand(n,1,generator=gen)*180-90],1)
    s = torch.rand(n, generator=gen)
    t0=time.time(); e = oi.iou_rot(b,b).numpy(); t1=time.time(); r = oi.iou_rle_raster(b,b).numpy(); t2=time.time()
    order = np.argsort(-s.numpy(), kind='stable')
    ke, kr = greedy(e, order, 0.45), greedy(r, order, 0.45)
    se, sr = set(ke), set(kr)
    ov = e>0.05
    print(f'{name}: n={n} overlapping pairs={int(ov.sum()-n)//2} kept exact={len(ke)} raster={len(kr)} symmetric difference={len(se^sr)} | |raster-exact| median={np.median(np.abs(r-e)[ov]):.2e} max={np.abs(r-e)[ov].max():.2e} | times exact {t1-t0:.1f}s raster {t2-t1:.1f}s')
