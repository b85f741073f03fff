"""CPU-only test of the fixed-point sweep (sweep_kernel<2> of nms_large.cu, the default sweep of the large-N NMS;
its GPU counterparts are tests/test_zz_sweep_fixpoint_gpu.py and every large-N parity test).

tests/host_harness/sweep_fixpoint_host.cpp runs the phases of csrc/sweep_fixpoint.cuh -- the code the kernel executes
between its barriers -- on the CPU under AddressSanitizer.  Suppression matrices are built here in the spatial layout
of nms_large.cu (rows / bits by Morton position, bit only in the row of the higher-ranked box, per-tile adjacency map,
rank -> position map) and the survivors are compared with the serial greedy sweep, for several emulated CTA sizes."""
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def harness(tmp_path_factory):
    d = tmp_path_factory.mktemp('fx')
    exe = str(d / 'sweep_fixpoint_host')
    res = subprocess.run(['g++', '-O1', '-g', '-std=c++17', '-fsanitize=address', '-o', exe,
                          os.path.join(ROOT, 'tests', 'host_harness', 'sweep_fixpoint_host.cpp')],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout

    def run(mask, adj, spos_of_rank, mb, words_total, aw, nt, list_cap=-1):
        src, dst = str(d / 'in.bin'), str(d / 'out.bin')
        with open(src, 'wb') as f:
            np.array([mb, words_total, aw], dtype=np.int32).tofile(f)
            mask.astype(np.uint64).tofile(f)
            adj.astype(np.uint64).tofile(f)
            spos_of_rank.astype(np.int32).tofile(f)
        res = subprocess.run([exe, src, dst, str(nt), str(list_cap)], env=dict(os.environ, ASAN_OPTIONS='detect_leaks=0'),
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert res.returncode == 0, res.stderr[-2000:]
        raw = open(dst, 'rb').read()
        rounds, entries, used_list = (int(v) for v in np.frombuffer(raw[:12], dtype=np.int32))
        kept = np.frombuffer(raw[12:], dtype=np.uint64)
        return rounds, kept, entries, used_list
    return run


def morton_order(cx, cy):
    def spread(v):
        v = v.astype(np.uint64) & np.uint64(0xFFFF)
        for sh, msk in ((8, 0x00FF00FF), (4, 0x0F0F0F0F), (2, 0x33333333), (1, 0x55555555)):
            v = (v | (v << np.uint64(sh))) & np.uint64(msk)
        return v
    q = 65535.0 / max(float(cx.max()), float(cy.max()), 1.0)
    key = spread(np.clip(cx * q, 0, 65535)) | (spread(np.clip(cy * q, 0, 65535)) << np.uint64(1))
    return np.argsort(key, kind='stable')


def layout(sup_pairs, mb, cap, spos_of_rank, extra_adj=False):
    """(rank_hi, rank_lo) suppression pairs -> mask / tile_adj in the layout of the spatial path."""
    words_total = (cap + 63) // 64
    aw = (words_total + 63) // 64
    tiles = (mb + 63) // 64
    mask = np.zeros((mb, words_total), dtype=np.uint64)
    adj = np.zeros((max(tiles, 1), aw), dtype=np.uint64)
    for hi, lo in sup_pairs:
        p, c = int(spos_of_rank[hi]), int(spos_of_rank[lo])
        mask[p, c >> 6] |= np.uint64(1) << np.uint64(c & 63)
        adj[p >> 6, (c >> 6) >> 6] |= np.uint64(1) << np.uint64((c >> 6) & 63)
    if extra_adj:                            # a superset adjacency map must be harmless (words it points at are zero)
        adj[:, 0] |= np.uint64(0b101)
    return mask, adj[:tiles], words_total, aw


def greedy(sup_pairs, mb):
    below = [[] for _ in range(mb)]
    for hi, lo in sup_pairs:
        below[hi].append(lo)
    keep = np.ones(mb, dtype=bool)
    for i in range(mb):
        if keep[i]:
            keep[below[i]] = False
    return keep


def unpack(kept_words, mb):
    bits = np.unpackbits(kept_words.view(np.uint8), bitorder='little')
    return bits[:mb].astype(bool), bits[mb:]


def boxes_case(rng, n, span, lo, hi, thr):
    from oracle import iou as oi
    b = torch.from_numpy(np.concatenate([rng.uniform(0, span, (n, 2)), rng.uniform(lo, hi, (n, 2))], 1).astype(np.float32))
    s = torch.from_numpy(rng.permutation(n).astype(np.float32))
    order = torch.argsort(s, descending=True, stable=True)
    b = b[order]                                            # rank order
    iou = oi.bboxes_iou(b, b).numpy()
    hi_i, lo_i = np.nonzero(np.triu(iou > thr, k=1))
    return b.numpy(), list(zip(hi_i.tolist(), lo_i.tolist()))


@pytest.mark.parametrize('nt', [512, 7])
def test_fixpoint_sweep_equals_greedy(harness, nt):
    rng = np.random.default_rng(17 + nt)
    cases = []
    b, pairs = boxes_case(rng, 1500, 600, 8, 80, 0.45); cases.append(('dense random', b, pairs, 1500, 1500, False))
    b, pairs = boxes_case(rng, 900, 150, 20, 60, 0.3); cases.append(('heavy overlap', b, pairs, 900, 1024, True))
    b, pairs = boxes_case(rng, 130, 80, 8, 40, 0.45); cases.append(('capacity above count', b, pairs, 130, 4200, False))
    b, pairs = boxes_case(rng, 65, 30, 8, 40, 0.45); cases.append(('65 boxes', b, pairs, 65, 65, False))
    b, pairs = boxes_case(rng, 1, 30, 8, 40, 0.45); cases.append(('1 box', b, pairs, 1, 64, False))
    # adversarial chain: box i overlaps only box i+1 -> the decisive chain is as long as the image (rounds ~ n)
    n = 200
    chain = np.stack([np.arange(n) * 10.0 + 50, np.full(n, 50.0), np.full(n, 16.0), np.full(n, 16.0)], 1).astype(np.float32)
    cases.append(('chain', chain, [(i, i + 1) for i in range(n - 1)], n, 256, False))
    for name, b, pairs, mb, cap, extra in cases:
        for perm in ('morton', 'random'):
            pos_to_rank = morton_order(b[:, 0], b[:, 1]) if perm == 'morton' else rng.permutation(mb)
            spos_of_rank = np.empty(mb, dtype=np.int64)
            spos_of_rank[pos_to_rank] = np.arange(mb)
            mask, adj, words_total, aw = layout(pairs, mb, cap, spos_of_rank, extra)
            want = greedy(pairs, mb)
            nonzero_words = int(np.count_nonzero(mask))
            for list_cap in (-1, max(nonzero_words // 2, 0)):      # the library's capacity, and an overflowing list
                rounds, kept_words, entries, used_list = harness(mask, adj, spos_of_rank, mb, words_total, aw, nt, list_cap)
                got, tail = unpack(kept_words, mb)
                assert np.array_equal(got, want), (name, perm, list_cap, int((got != want).sum()))
                assert not tail.any(), (name, 'bits beyond the valid boxes')
                fits = nonzero_words <= (4 * mb if list_cap < 0 else list_cap)
                assert entries == nonzero_words and used_list == int(fits), (name, entries, nonzero_words, list_cap)
            if name == 'chain':
                assert want.sum() == mb // 2 and rounds >= mb // 2       # alternate survivors; one chain link per round or two
            elif mb > 100:
                assert rounds <= 12, (name, rounds)


def test_fixpoint_sweep_empty_image(harness):
    rounds, kept, entries, used_list = harness(np.zeros((0, 2), np.uint64), np.zeros((0, 1), np.uint64), np.zeros(0, np.int64), 0, 2, 1, 64)
    assert rounds == 1 and entries == 0 and used_list == 1 and not kept.any()
