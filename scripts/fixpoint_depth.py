"""Design probe for the large-N sweep; CPU only.

The greedy NMS result is the unique fixed point of   keep[i] = valid[i] and not any(keep[j] and M[j, i] for j ranked above i)
and a JACOBI iteration from keep = valid reaches it in (longest decisive suppression chain + 1) rounds -- every round is
embarrassingly parallel over the non-empty words of the suppression bit matrix, where the present sweep walks the 64-box
blocks of an image one after the other (157 blocks at 10 k boxes, 757 at 48 k).  This script measures, on the BASELINE
workloads, how many rounds the iteration needs and how many non-empty 64-bit mask words (in Morton order of the box
centres, as nms_large.cu lays the matrix out) it would touch per round."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np                       # noqa: E402
import scipy.sparse as sp                # noqa: E402
import torch                             # noqa: E402
from oracle import decode as od, iou as oi                        # noqa: E402
from helpers import yolo_views, level_anchors, RAPID_ANCHORS      # noqa: E402


def morton(cx, cy, span):
    def spread(v):
        v = v.astype(np.uint64) & 0xFFFF
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        return (v | (v << 1)) & 0x55555555
    q = 65535.0 / span
    return spread(np.clip(cx * q, 0, 65535)) | (spread(np.clip(cy * q, 0, 65535)) << 1)


def report(name, boxes, scores, pairs_fn, thr, span):
    n = boxes.shape[0]
    order = torch.argsort(scores, descending=True, stable=True)
    b = boxes[order]
    rows, cols = [], []
    for s in range(0, n, 500):
        iou = pairs_fn(b[s:s + 500], b)
        r, c = np.nonzero(iou >= thr)
        r = r + s
        keep = r < c                                   # the higher-ranked box suppresses
        rows.append(r[keep]); cols.append(c[keep])
    r, c = np.concatenate(rows), np.concatenate(cols)
    M = sp.csr_matrix((np.ones(len(r), dtype=np.int32), (r, c)), shape=(n, n))
    MT = M.T.tocsr()
    greedy = np.ones(n, dtype=bool)
    for i in range(n):
        if greedy[i]:
            greedy[M.indices[M.indptr[i]:M.indptr[i + 1]]] = False
    k, rounds, changes = np.ones(n, dtype=np.int32), 0, []
    while True:
        k2 = (~(MT.dot(k) > 0)).astype(np.int32)
        rounds += 1
        changes.append(int((k2 != k).sum()))
        if changes[-1] == 0:
            break
        k = k2
    assert np.array_equal(k.astype(bool), greedy)
    spos = np.empty(n, dtype=np.int64)                 # rank -> position in Morton order of the centres
    spos[np.argsort(morton(b[:, 0].numpy(), b[:, 1].numpy(), span), kind='stable')] = np.arange(n)
    words = len(set(zip(spos[r].tolist(), (spos[c] >> 6).tolist())))
    print(f'{name}: n={n} suppressing pairs={len(r)} kept={int(greedy.sum())} | Jacobi rounds={rounds} (changes {changes}) | '
          f'non-empty mask words={words} | serial 64-box blocks today={(n + 63) // 64}')


def aabb_pairs(a, b):
    return oi.bboxes_iou(a[:, :4], b[:, :4]).numpy()


def rot_pairs(a, b):
    return oi.iou_rot(a, b).numpy()


def dense_scene(img, seed=0):
    gen = torch.Generator().manual_seed(seed)
    lv = []
    for s in (8, 16, 32):
        n = img // s
        bb = torch.randn(1, 4, n, n, generator=gen) * 0.5
        cc = torch.randn(1, 2, n, n, generator=gen) * 1.5
        cc[:, 0] += 2.0
        c = cc.permute(0, 2, 3, 1)
        lv.append(od.decode_fcos({'bbox': bb.permute(0, 2, 3, 1), 'conf': c[..., 0:1], 'class': c[..., 1:]}, s, (img, img)))
    box, _, score = (t[0] for t in od.merge_levels(lv))
    return box, score


def rapid_10k():
    gen = torch.Generator().manual_seed(3003)
    lv = []
    for li, s in enumerate((8, 16, 32)):
        m = 1024 // s
        t = torch.randn(1, 18, m, m, generator=gen) * 0.5
        v = t.view(1, 3, 6, m, m)
        v[:, :, 4] = torch.rand(1, 3, m, m, generator=gen) * 6 - 3
        v[:, :, 5] = torch.randn(1, 3, m, m, generator=gen) * 1.5 - 1.5
        lv.append(od.decode_rapid(yolo_views(t, 3, 5, 0), level_anchors(RAPID_ANCHORS, li), s, 0))
    box, _, score = (t[0] for t in od.merge_levels(lv))
    top = score.topk(10000).indices
    return box[top].contiguous(), score[top].contiguous()


def clustered(n_obj, per_obj, seed=1):
    """Trained-detector-like: every object draws a cloud of jittered boxes (heavy mutual overlap inside a cloud)."""
    gen = torch.Generator().manual_seed(seed)
    c = torch.rand(n_obj, 2, generator=gen) * 900 + 60
    wh = torch.rand(n_obj, 2, generator=gen) * 100 + 30
    ang = torch.rand(n_obj, 1, generator=gen) * 180 - 90
    base = torch.cat([c, wh, ang], 1).repeat_interleave(per_obj, 0)
    jit = torch.randn(base.shape, generator=gen) * torch.tensor([6.0, 6.0, 8.0, 8.0, 6.0])
    box = base + jit
    box[:, 2:4] = box[:, 2:4].clamp(min=4)
    return box, torch.rand(box.shape[0], generator=gen)


if __name__ == '__main__':
    b, s = rapid_10k()
    report('configs[2] rotated NMS, 10 000 boxes (the bench workload)', b, s, rot_pairs, 0.45, 1024)
    for img in (704, 1024):
        b, s = dense_scene(img)
        report(f'configs[4] dense scene @{img}', b, s, aabb_pairs, 0.45, img)
    b, s = clustered(100, 100)
    report('clustered: 100 objects x 100 jittered rotated boxes', b, s, rot_pairs, 0.45, 1024)
    b, s = clustered(20, 500)
    report('clustered: 20 objects x 500 jittered rotated boxes', b, s, rot_pairs, 0.45, 1024)
