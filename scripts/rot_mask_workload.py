"""Design probe (CPU): how much work each stage of the rotated mask kernel (nms_large.cu: mask_rot_spatial_kernel) does on
the bench workload -- 10 000 rotated boxes of one image in Morton order, 64-box tiles.  Counts tile pairs whose hulls
overlap, the branch-free circle + area-ratio tests they cost (64 x 64 each), the pairs that survive to the hull bound, and
the pairs that reach the polygon clip; the same for other tile sizes, to see what a finer spatial structure would buy."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
import numpy as np                                  # noqa: E402
import torch                                        # noqa: E402
from fixpoint_depth import rapid_10k, morton        # noqa: E402
from oracle import iou as oi                        # noqa: E402

box, score = rapid_10k()
n = box.shape[0]
b = box.numpy().astype(np.float64)
rad = np.deg2rad(b[:, 4])
c, s = np.abs(np.cos(rad)), np.abs(np.sin(rad))
hw, hh = b[:, 2] / 2, b[:, 3] / 2
ex, ey = hw * c + hh * s, hw * s + hh * c           # half extents of the axis-aligned hull of the rotated box
x0, x1, y0, y1 = b[:, 0] - ex, b[:, 0] + ex, b[:, 1] - ey, b[:, 1] + ey
r = 0.5 * np.hypot(b[:, 2], b[:, 3])
area = b[:, 2] * b[:, 3]
pos = np.argsort(morton(b[:, 0], b[:, 1], 1024), kind='stable')
thr = 0.45
exact = None
for tile in (64, 32, 16):
    T = (n + tile - 1) // tile
    idx = [pos[t * tile:(t + 1) * tile] for t in range(T)]
    hx0 = np.array([x0[i].min() for i in idx]); hx1 = np.array([x1[i].max() for i in idx])
    hy0 = np.array([y0[i].min() for i in idx]); hy1 = np.array([y1[i].max() for i in idx])
    ov = (hx0[:, None] <= hx1[None, :]) & (hx0[None, :] <= hx1[:, None]) & (hy0[:, None] <= hy1[None, :]) & (hy0[None, :] <= hy1[:, None])
    ti, tj = np.nonzero(np.triu(ov))
    circle = hull = 0
    for a_, b_ in zip(ti, tj):
        ia, ib = idx[a_], idx[b_]
        d2 = (b[ia, 0][:, None] - b[ib, 0][None, :]) ** 2 + (b[ia, 1][:, None] - b[ib, 1][None, :]) ** 2
        rr = r[ia][:, None] + r[ib][None, :]
        lo = np.minimum(area[ia][:, None], area[ib][None, :]); hi = np.maximum(area[ia][:, None], area[ib][None, :])
        p1 = (d2 <= rr * rr) & (lo >= thr * hi)
        if a_ == b_:
            p1 = np.triu(p1, k=1)
        circle += int(p1.sum())
        ix = np.minimum(x1[ia][:, None], x1[ib][None, :]) - np.maximum(x0[ia][:, None], x0[ib][None, :])
        iy = np.minimum(y1[ia][:, None], y1[ib][None, :]) - np.maximum(y0[ia][:, None], y0[ib][None, :])
        ub = np.clip(ix, 0, None) * np.clip(iy, 0, None)
        p2 = p1 & (ix > 0) & (iy > 0) & (ub >= thr * (area[ia][:, None] + area[ib][None, :] - ub))
        hull += int(p2.sum())
    print(f'tile {tile:3d}: {T} tiles, {len(ti)} of {T * (T + 1) // 2} tile pairs overlap ({100 * len(ti) / (T * (T + 1) // 2):.1f} %), '
          f'{len(ti) * tile * tile / 1e6:.2f} M circle tests, {circle} pass circle + area ratio, {hull} reach the clip')
