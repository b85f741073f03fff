"""CPU: the arithmetic / geometric claims the kernels' shortcuts rest on, checked by brute force in float32 numpy.
Nothing here runs device code; each test restates the shortcut exactly as the kernel source does and compares it with the
unabridged computation (the reference's formula) on adversarial and random inputs."""
import numpy as np

f32 = np.float32


# ------------------------------------------------------------------------------------------------------------------
# csrc/atss.cu, atss_threshold_kernel: candidate window of the k nearest anchor centres
def _topk_by_distance_index(gx, gy, rows, cols, n_w, stride, k):
    """(distance, flat index) order of the given cells, float32 arithmetic of the kernel / of fcos2.py:396."""
    s, half = f32(stride), f32(0.5) * f32(stride)
    ax = cols.astype(f32) * s + half
    ay = rows.astype(f32) * s + half
    dx, dy = f32(gx) - ax, f32(gy) - ay
    d = dx * dx + dy * dy
    idx = rows * n_w + cols
    order = np.lexsort((idx, d))
    return idx[order][:k]


def _window(gx, gy, n_h, n_w, stride, k):
    """The kernel's window: 5x5 when k <= 9 and the 3x3 block around the centre's cell fits, else 11x11, clipped; the whole
    grid when it is thinner than ceil(sqrt(k)) (found by this test: on a 1 x 20 grid the 9 nearest of a corner reach 8.5 cells)."""
    col0 = min(max(int(np.floor(f32(gx) / f32(stride))), 0), n_w - 1)
    row0 = min(max(int(np.floor(f32(gy) / f32(stride))), 0), n_h - 1)
    m_blk = 1 if k <= 1 else (2 if k <= 4 else (3 if k <= 9 else 4))
    if n_h < m_blk or n_w < m_blk:                      # grid too thin for the window argument: exhaustive scan
        rr, cc = np.meshgrid(np.arange(n_h), np.arange(n_w), indexing='ij')
        return rr.ravel(), cc.ravel(), False
    small = k <= 9 and col0 >= 1 and col0 + 1 <= n_w - 1 and row0 >= 1 and row0 + 1 <= n_h - 1
    R = 2 if small else 5
    c_lo, r_lo = max(col0 - R, 0), max(row0 - R, 0)
    c_hi, r_hi = min(col0 + R, n_w - 1), min(row0 + R, n_h - 1)
    rr, cc = np.meshgrid(np.arange(r_lo, r_hi + 1), np.arange(c_lo, c_hi + 1), indexing='ij')
    return rr.ravel(), cc.ravel(), small


def test_atss_candidate_window_holds_the_k_nearest_anchors():
    rng = np.random.default_rng(0)
    n_small = n_big = 0
    for stride, n_h, n_w in ((8, 48, 80), (16, 24, 40), (32, 12, 20), (64, 6, 10), (128, 3, 5), (8, 3, 3), (8, 1, 20), (8, 20, 2), (8, 3, 30), (8, 4, 4), (8, 2, 9), (8, 30, 4)):
        all_r, all_c = (a.ravel() for a in np.meshgrid(np.arange(n_h), np.arange(n_w), indexing='ij'))
        W, H = n_w * stride, n_h * stride
        pts = [(0, 0), (W, H), (0, H), (W, 0), (W / 2, 0), (stride, stride), (stride * 1.5, stride * 1.5), (W - 1e-3, H - 1e-3),
               (stride * 2, stride * 2.5), (np.nextafter(f32(stride * 3), f32(0)), stride * 3)]
        pts += [(rng.uniform(0, W), rng.uniform(0, H)) for _ in range(150)]
        # cell boundaries and half cells, where distances tie
        pts += [(stride * rng.integers(0, n_w + 1), stride * rng.integers(0, n_h + 1) + stride / 2 * rng.integers(0, 2)) for _ in range(60)]
        for k in (1, 4, 9, 12, 16):
            if n_h * n_w < k:
                continue
            for gx, gy in pts:
                if not (0 <= gx <= W and 0 <= gy <= H):
                    continue
                wr, wc, small = _window(gx, gy, n_h, n_w, stride, k)
                want = _topk_by_distance_index(gx, gy, all_r, all_c, n_w, stride, k)
                got = _topk_by_distance_index(gx, gy, wr, wc, n_w, stride, k)
                assert np.array_equal(got, want), (stride, n_h, n_w, k, gx, gy, small)
                n_small += small
                n_big += not small
    assert n_small > 1000 and n_big > 1000


# ------------------------------------------------------------------------------------------------------------------
# csrc/atss.cu, assign_body: per-CTA GT cull against the hull of the CTA's predicted boxes and cell centres
def test_atss_gt_cull_never_drops_a_gt_that_could_matter():
    rng = np.random.default_rng(1)
    dropped = kept = 0
    for trial in range(300):
        stride = f32(rng.choice([8, 16, 32]))
        n_w = int(rng.integers(3, 40))
        cells = np.arange(int(rng.integers(1, 129))) + int(rng.integers(0, 50))
        row, col = cells // n_w, cells % n_w
        half = stride * f32(0.5)
        gx = col.astype(f32) * stride + half
        gy = row.astype(f32) * stride + half
        t = rng.normal(0, 0.8, (cells.size, 4)).astype(f32)
        if trial % 7 == 0:
            t[rng.integers(0, cells.size)] = [np.nan, 0, 0, 0]
        if trial % 11 == 0:
            t[rng.integers(0, cells.size)] = [200.0, 0, 0, 0]            # exp overflows to inf
        with np.errstate(over='ignore', invalid='ignore'):
            l, tp, r, bt = (np.exp(t[:, i]).astype(f32) * stride for i in range(4))
            pcx, pcy = gx + (r - l) * f32(0.5), gy + (bt - tp) * f32(0.5)
            pw, ph = l + r, tp + bt
            phw, phh = pw * f32(0.5), ph * f32(0.5)
            px1, py1, px2, py2 = pcx - phw, pcy - phh, pcx + phw, pcy + phh
            # hull: fminf / fmaxf drop NaN operands
            hx1, hy1 = np.fmin(gx, px1).min(), np.fmin(gy, py1).min()
            hx2, hy2 = np.fmax(gx, px2).max(), np.fmax(gy, py2).max()
            G = 60
            gt = np.concatenate([rng.uniform(-50, 400, (G, 2)), rng.uniform(0, 120, (G, 2))], axis=1).astype(f32)
            gt[0] = [np.nan, 10, 20, 20]
            gt[1, 2:] = 0
            ghw, ghh = gt[:, 2] * f32(0.5), gt[:, 3] * f32(0.5)
            g1x, g1y, g2x, g2y = gt[:, 0] - ghw, gt[:, 1] - ghh, gt[:, 0] + ghw, gt[:, 1] + ghh
            keep = (np.fmax(hx1, g1x) <= np.fmin(hx2, g2x)) & (np.fmax(hy1, g1y) <= np.fmin(hy2, g2y))
            for g in np.nonzero(~keep)[0]:
                # iou_cxcywh_or_zero's overlap test (NaN-propagating max / min: np.maximum / np.minimum) fails for every cell
                tlx, tly = np.maximum(px1, g1x[g]), np.maximum(py1, g1y[g])
                brx, bry = np.minimum(px2, g2x[g]), np.minimum(py2, g2y[g])
                assert not ((tlx < brx) & (tly < bry)).any(), (trial, g)
                # and no cell centre lies strictly inside the GT (fcos2.py:321)
                inside = (gx - g1x[g] > 0) & (gy - g1y[g] > 0) & (g2x[g] - gx > 0) & (g2y[g] - gy > 0)
                assert not inside.any(), (trial, g)
            # FCOS mode (fcos2.py:113-133): the positive test is "centre strictly inside the GT scaled by center_region", and
            # the cull box is the union of the GT and that region (center_region may exceed 1)
            for cr in (f32(0.5), f32(1.0), f32(1.7)):
                chw, chh = (gt[:, 2] * cr) * f32(0.5), (gt[:, 3] * cr) * f32(0.5)
                uhw, uhh = np.fmax(ghw, chw), np.fmax(ghh, chh)
                keep_f = (np.fmax(hx1, gt[:, 0] - uhw) <= np.fmin(hx2, gt[:, 0] + uhw)) & \
                         (np.fmax(hy1, gt[:, 1] - uhh) <= np.fmin(hy2, gt[:, 1] + uhh))
                for g in np.nonzero(~keep_f)[0]:
                    centre = (gx > gt[g, 0] - chw[g]) & (gx < gt[g, 0] + chw[g]) & (gy > gt[g, 1] - chh[g]) & (gy < gt[g, 1] + chh[g])
                    assert not centre.any(), (trial, g, cr)
                    tlx, tly = np.maximum(px1, gt[g, 0] - ghw[g]), np.maximum(py1, gt[g, 1] - ghh[g])
                    brx, bry = np.minimum(px2, gt[g, 0] + ghw[g]), np.minimum(py2, gt[g, 1] + ghh[g])
                    assert not ((tlx < brx) & (tly < bry)).any(), (trial, g, cr)
        dropped += int((~keep).sum())
        kept += int(keep.sum())
    assert dropped > 3000 and kept > 1000


# ------------------------------------------------------------------------------------------------------------------
# csrc/common.cuh, iou_from_parts / iou_corners_gt: the divide is skipped where the numerator is a zero
def test_zero_numerator_quotient_is_the_numerator():
    sums = np.array([1e-38, 1e-30, 1.0, 3.5, 1e30, np.inf, np.finfo(f32).tiny, np.finfo(f32).max], dtype=f32)
    for z in (f32(0.0), f32(-0.0)):
        with np.errstate(all='ignore'):
            q = z / (sums - z)
        assert np.array_equal(q.view(np.int32), np.full_like(sums, z).view(np.int32)), (z, q)
    # ... and NOT where area_a + area_b is not positive: the kernels divide there
    with np.errstate(all='ignore'):
        assert np.isnan(f32(0.0) / (f32(0.0) - f32(0.0)))
        assert (f32(0.0) / (f32(-2.0) - f32(0.0))).view(np.int32) == f32(-0.0).view(np.int32)
    # iou_corners_gt: inter == 0 gives 0, -0 or NaN -- never above a non-negative threshold
    unis = np.array([-3.0, -0.0, 0.0, 2.0, np.inf, np.nan], dtype=f32)
    with np.errstate(all='ignore'):
        for thr in (f32(0.0), f32(0.45), f32(1.0)):
            assert not (f32(0.0) / unis > thr).any()


# ------------------------------------------------------------------------------------------------------------------
# csrc/nms_large.cu, rot_broad_kernel / rot_narrow_kernel: a pair is dropped before the polygon clip by the circle test, the
# area-ratio bound, the hull-overlap bound or the oriented-extent bound -- none may drop a pair whose exact IoU reaches thr
def _rot_quantities(b):
    """Per-box quantities as the gather kernel stores them (float32): corners, centre, radius, area, hull, half axes."""
    import torch
    from oracle import iou as oi
    rad = b.clone()
    rad[:, 4] = oi.deg2rad_f32(rad[:, 4])
    v = oi.xywha2vertex(rad).numpy().astype(f32)                          # (N,4,2) tl,tr,br,bl
    x, y = v[:, :, 0], v[:, :, 1]
    bn = b.numpy().astype(f32)
    r = f32(0.5) * np.sqrt(bn[:, 2] * bn[:, 2] + bn[:, 3] * bn[:, 3]).astype(f32)
    xd, yd = x.astype(np.float64), y.astype(np.float64)
    a2 = sum(xd[:, k] * yd[:, (k + 1) % 4] - xd[:, (k + 1) % 4] * yd[:, k] for k in range(4))
    area = f32(0.5) * np.abs(a2.astype(f32))
    hull = np.stack([x.min(1), y.min(1), x.max(1), y.max(1)], 1)
    axes = np.stack([f32(0.5) * (x[:, 1] - x[:, 0]), f32(0.5) * (y[:, 1] - y[:, 0]),
                     f32(0.5) * (x[:, 0] - x[:, 3]), f32(0.5) * (y[:, 0] - y[:, 3])], 1)
    return bn[:, 0], bn[:, 1], r, area, hull, axes


def _oriented_bound(ca, xa, cb, xb):
    """oriented_overlap_bound of nms_large.cu, vectorised over pairs; c = (cx, cy, r, area), x = (Hx, Hy, Vx, Vy)."""
    dx, dy = cb[0] - ca[0], cb[1] - ca[1]
    hh, vh = np.abs(xb[0] * xa[0] + xb[1] * xa[1]), np.abs(xb[2] * xa[0] + xb[3] * xa[1])
    hv, vv = np.abs(xb[0] * xa[2] + xb[1] * xa[3]), np.abs(xb[2] * xa[2] + xb[3] * xa[3])

    def frame(h2, v2, ph, pv, eh, ev, area):
        ox = np.minimum(h2, ph + eh) - np.maximum(-h2, ph - eh)
        oy = np.minimum(v2, pv + ev) - np.maximum(-v2, pv - ev)
        return np.maximum(ox, f32(0)) * np.maximum(oy, f32(0)) * f32(4) / area
    ua = frame(xa[0] * xa[0] + xa[1] * xa[1], xa[2] * xa[2] + xa[3] * xa[3], dx * xa[0] + dy * xa[1], dx * xa[2] + dy * xa[3],
               hh + vh, hv + vv, ca[3])
    ub = frame(xb[0] * xb[0] + xb[1] * xb[1], xb[2] * xb[2] + xb[3] * xb[3], dx * xb[0] + dy * xb[1], dx * xb[2] + dy * xb[3],
               hh + hv, vh + vv, cb[3])
    return np.minimum(ua, ub)


def test_rotated_culls_never_drop_a_pair_that_reaches_the_threshold():
    import torch
    from oracle import iou as oi
    g = torch.Generator().manual_seed(7)
    n = 700
    boxes = []
    # clusters of similar boxes (IoUs on both sides of every threshold), thin boxes, tiny and huge ones, all angles
    for _ in range(14):
        c = torch.rand(2, generator=g) * 300 + 100
        wh = torch.rand(2, generator=g) * torch.tensor([180.0, 60.0]) + 4
        ang = torch.rand(1, generator=g) * 360 - 180
        k = n // 14
        boxes.append(torch.cat([c + torch.randn(k, 2, generator=g) * wh.min() * 0.4, wh * torch.exp(torch.randn(k, 2, generator=g) * 0.25),
                                ang + torch.randn(k, 1, generator=g) * 25], 1))
    b = torch.cat(boxes).float()
    b[::97, 4] = 0.0
    b[5::97, 4] = 90.0
    b[11::97, 2:4] = b[11::97, 2:4].flip(1)
    cx, cy, r, area, hull, axes = _rot_quantities(b)
    exact = oi.iou_rot(b, b).numpy()
    i, j = np.triu_indices(b.shape[0], 1)
    ca = (cx[i], cy[i], r[i], area[i])
    cb = (cx[j], cy[j], r[j], area[j])
    with np.errstate(all='ignore'):
        # (1) circumscribed circles: rot_broad_kernel, mr = r * 1.00001 + 1e-3, rr = r_other * 1.00001 + mr
        dx, dy = ca[0] - cb[0], ca[1] - cb[1]
        rr = cb[2] * f32(1.00001) + (ca[2] * f32(1.00001) + f32(1e-3))
        circle_drop = ~(dx * dx + dy * dy <= rr * rr)
        assert exact[i, j][circle_drop].max(initial=0.0) == 0.0
        ub_o = _oriented_bound(ca, axes[i].T, cb, axes[j].T)
        ix = np.minimum(hull[i, 2], hull[j, 2]) - np.maximum(hull[i, 0], hull[j, 0])
        iy = np.minimum(hull[i, 3], hull[j, 3]) - np.maximum(hull[i, 1], hull[j, 1])
        ub_h = (ix + f32(2e-3)) * (iy + f32(2e-3))
        n_dropped = 0
        for thr in (0.05, 0.3, 0.45, 0.5, 0.7, 0.95):
            t = f32(thr)
            ratio_drop = ~(np.minimum(ca[3], cb[3]) * f32(1.0001) >= t * np.maximum(ca[3], cb[3]))           # (2)
            hull_drop = ~((ix > f32(-1e-3)) & (iy > f32(-1e-3))) | (ub_h * f32(1.0001) < t * (ca[3] + cb[3] - ub_h))   # (3)
            orient_drop = ub_o * f32(1.001) + f32(1e-2) < t * (ca[3] + cb[3] - ub_o)                           # (4)
            for name, drop in (('ratio', ratio_drop), ('hull', hull_drop), ('oriented', orient_drop)):
                worst = exact[i, j][drop].max(initial=0.0)
                assert worst < thr, (name, thr, worst)
                n_dropped += int(drop.sum())
        # the oriented bound is an upper bound of the intersection area itself
        inter = exact[i, j] * (area[i].astype(np.float64) + area[j]) / (1.0 + exact[i, j])
        assert (inter <= ub_o.astype(np.float64) * 1.001 + 1e-2).all()
    assert n_dropped > 100000 and (exact[i, j] > 0.5).sum() > 500 and (exact[i, j] > 0.05).sum() > 5000


# ------------------------------------------------------------------------------------------------------------------
# csrc/common.cuh: float_at_or_below / float_at_or_above (torchvision compares a float32 IoU with a float64 threshold; the
# kernels compare with ONE float) and float_key (order-preserving float -> uint32 of the sort / select keys)
def test_float_threshold_and_sort_key_identities():
    rng = np.random.default_rng(3)
    ts = np.concatenate([rng.uniform(-1, 2, 2000), [0.0, 0.5, 0.45, 0.3, 1.0, 1e-45, -1e-45, 0.1, 0.7, 1 / 3]])
    for t in ts:
        lo = f32(t)
        if np.float64(lo) > t:
            lo = np.nextafter(lo, f32(-np.inf))                      # float_at_or_below
        hi = f32(t)
        if np.float64(hi) < t:
            hi = np.nextafter(hi, f32(np.inf))                       # float_at_or_above
        xs = np.array([lo, hi, np.nextafter(lo, f32(-np.inf)), np.nextafter(lo, f32(np.inf)), np.nextafter(hi, f32(np.inf)),
                       np.nextafter(hi, f32(-np.inf)), f32(0), f32(1), f32(-0.0)], dtype=f32)
        xs = np.concatenate([xs, rng.uniform(-1, 2, 50).astype(f32)])
        assert np.array_equal(xs.astype(np.float64) > t, xs > lo), t
        assert np.array_equal(xs.astype(np.float64) >= t, xs >= hi), t

    def float_key(s):
        u = (s + f32(0.0)).view(np.uint32)
        return np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000))
    vals = np.concatenate([rng.normal(0, 1, 5000).astype(f32), rng.uniform(0, 1, 5000).astype(f32),
                           np.array([0.0, -0.0, np.inf, -np.inf, 1e-45, -1e-45, np.finfo(f32).max, np.finfo(f32).min, 1.0, -1.0], dtype=f32)])
    keys = float_key(vals)
    order = np.argsort(vals, kind='stable')
    sv, sk = vals[order], keys[order]
    assert (np.diff(sk.astype(np.int64)) >= 0).all()                              # monotone
    assert ((np.diff(sk.astype(np.int64)) == 0) == (np.diff(sv) == 0)).all()      # equal keys <=> equal floats (-0 == +0)


# ------------------------------------------------------------------------------------------------------------------
# csrc/nms_large.cu, lazy narrow phase (rot_filter / rot_clip<0> / rot_clip<1>): greedy NMS on the REDUCED suppression matrix
# -- pairs under a root first, pairs with a dead member never evaluated -- keeps the same boxes as on the full matrix
def test_lazy_root_first_matrix_gives_the_greedy_set():
    rng = np.random.default_rng(5)

    def greedy(M):
        n = M.shape[0]
        kept = np.zeros(n, dtype=bool)
        for i in range(n):                       # rank order: 0 = best score; M[j, i], j < i: j suppresses i if kept
            kept[i] = not (M[:i, i] & kept[:i]).any()
        return kept

    total_pairs = evaluated = 0
    for trial in range(120):
        n = int(rng.integers(2, 160))
        n_obj = int(rng.integers(1, 8))
        obj = rng.integers(0, n_obj, n)
        # true overlaps: mostly inside an object's cluster, some across; listed pairs = a superset (the bound is conservative)
        p_in, p_out = rng.uniform(0.2, 0.95), rng.uniform(0.0, 0.05)
        same = obj[:, None] == obj[None, :]
        O = np.triu((rng.random((n, n)) < np.where(same, p_in, p_out)), 1)
        L = O | np.triu(rng.random((n, n)) < 0.1, 1)
        want = greedy(O)
        has_higher = L.any(axis=0)               # filter: the lower box of every listed pair is marked
        root = ~has_higher
        M = np.zeros_like(O)
        dead = np.zeros(n, dtype=bool)
        for j, i in zip(*np.nonzero(L)):         # clip<0>: pairs whose higher-scored box is a root
            if root[j]:
                evaluated += 1
                if O[j, i]:
                    M[j, i] = True
                    dead[i] = True
        for j, i in zip(*np.nonzero(L)):         # clip<1>: the rest, except pairs with a dead member
            if not root[j] and not dead[j] and not dead[i]:
                evaluated += 1
                if O[j, i]:
                    M[j, i] = True
        total_pairs += int(L.sum())
        assert np.array_equal(greedy(M), want), trial
        assert want[root].all()                  # roots are certainly kept
    assert evaluated < 0.6 * total_pairs         # and the reduction is real on clustered input


# ------------------------------------------------------------------------------------------------------------------
# csrc/postprocess_small.cu, register-resident front end: the exact top-K through a 2048-bin histogram that is LINEAR over
# the score range -- bins above the K-th score's bin are taken whole, that bin's candidates are ranked by (score, index)
def test_histogram_select_is_the_exact_top_k():
    rng = np.random.default_rng(9)
    BINS = 2048

    def select(s, thr, K):
        valid = s >= thr                                                   # NaN fails
        if valid.sum() <= K:
            return np.nonzero(valid)[0]
        smin, smax = s[valid].min(), s[valid].max()
        with np.errstate(all='ignore'):
            scale = f32(BINS - 1) / (smax - smin) if smax > smin else f32(0)
            b = np.nan_to_num((s - smin) * scale, nan=0.0, posinf=3e9, neginf=-3e9)      # cvt.rzi: NaN -> 0, saturating
        b = np.clip(b.astype(np.int64), 0, BINS - 1)
        hist = np.bincount(b[valid], minlength=BINS)
        above = hist[::-1].cumsum()[::-1] - hist                           # candidates in bins above each bin
        T = int(np.nonzero((above < K) & (K <= above + hist))[0][0])
        need = K - above[T]
        sure = np.nonzero(valid & (b > T))[0]
        und = np.nonzero(valid & (b == T))[0]
        und = und[np.lexsort((und, -s[und].astype(np.float64)))][:need]    # score descending, index ascending
        return np.concatenate([sure, und])

    for trial in range(400):
        n = int(rng.integers(1, 9217))
        kind = trial % 6
        if kind == 0:
            s = rng.random(n).astype(f32)
        elif kind == 1:
            s = (1 / (1 + np.exp(-rng.normal(-3, 2, n)))).astype(f32)      # detector scores: piled up near 0
        elif kind == 2:
            s = rng.choice(np.array([0.1, 0.2, 0.2000001, 0.5, 0.9], dtype=f32), n)      # heavy ties
        elif kind == 3:
            s = (f32(0.5) + rng.integers(0, 3, n).astype(f32) * np.finfo(f32).eps).astype(f32)   # range of two ulps
        elif kind == 4:
            s = rng.normal(0, 1e-38, n).astype(f32)                        # denormal range: the scale overflows
        else:
            s = rng.random(n).astype(f32)
            s[rng.integers(0, n, max(1, n // 50))] = np.nan
        thr = f32(rng.choice([-1.0, 0.0, 0.005, 0.3]))
        K = int(rng.choice([1, 7, 100, 512, 1024]))
        got = np.sort(select(s, thr, K))
        valid = np.nonzero(s >= thr)[0]
        want = np.sort(valid[np.lexsort((valid, -s[valid].astype(np.float64)))][:K])
        assert np.array_equal(got, want), (trial, kind, n, K)
