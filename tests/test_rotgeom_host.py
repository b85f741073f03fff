"""CPU-only fuzz of the rotated-NMS pair decision.

tests/host_harness/rotgeom_host.cpp compiles mydetection_b200/csrc/rotgeom.cuh -- the device functions behind
mydet_nms_rot and mydet_iou_rot_pairwise -- for the host and puts the cull chain of the rotated mask kernel in front
of them.  Adversarial pair families (near-duplicates, pairs constructed to sit AT the threshold, extreme aspect
ratios, sub-pixel boxes far from the origin, image-sized boxes, special angles, zero-sized boxes) are pushed through
it and compared with the oracle's float64 polygon clipping (oracle/rotiou.c):
  * no cull stage ever drops a pair the oracle suppresses,
  * the decision equals `IoU >= thr` (or `>`) of the exact float64 IoU of the same float32 corners,
  * the float32 clip stays far inside the 1e-3 band that triggers the float64 re-check.
The GPU tests cover ~1e8 pairs of realistic boxes; this covers the corners of the input space they never visit.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = np.dtype([('decision', '<i4'), ('stage', '<i4'), ('iou32', '<f4'), ('iou64', '<f8')])
N = 20000


@pytest.fixture(scope='module')
def harness(tmp_path_factory):
    d = tmp_path_factory.mktemp('rotgeom')
    exe = str(d / 'rotgeom_host')
    cuda_inc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')
    res = subprocess.run(['g++', '-std=c++17', '-O2', '-ffp-contract=off', '-I', cuda_inc, '-o', exe,
                          os.path.join(ROOT, 'tests', 'host_harness', 'rotgeom_host.cpp')],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout

    def run(a, b, thr, ge):
        src, dst = str(d / 'p.bin'), str(d / 'o.bin')
        np.concatenate([a, b], 1).astype(np.float32).tofile(src)
        subprocess.run([exe, src, dst, str(a.shape[0]), repr(float(thr)), str(int(ge))], check=True)
        return np.fromfile(dst, dtype=OUT)
    return run


def oracle_pairs(a, b, blk=32):
    """Element-wise oracle IoU (diagonals of small blocks of the pairwise routine)."""
    from oracle import iou as oi
    L = oi.lib()
    f32p, f64p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
    out = np.empty(a.shape[0])
    for s in range(0, a.shape[0], blk):
        aa, bb = np.ascontiguousarray(a[s:s + blk]), np.ascontiguousarray(b[s:s + blk])
        t = np.empty((aa.shape[0], bb.shape[0]))
        L.oracle_rot_iou_pairwise(aa.ctypes.data_as(f32p), aa.shape[0], bb.ctypes.data_as(f32p), bb.shape[0], t.ctypes.data_as(f64p))
        out[s:s + blk] = np.diag(t)
    return out


def families(rng, thr):
    def boxes(lo, hi, span, ang=90.0):
        return np.stack([rng.uniform(0, span, N), rng.uniform(0, span, N), np.exp(rng.uniform(np.log(lo), np.log(hi), N)),
                         np.exp(rng.uniform(np.log(lo), np.log(hi), N)), rng.uniform(-ang, ang, N)], 1).astype(np.float32)

    def near(a, pos, size, ang):
        b = a.copy()
        b[:, 0:2] += rng.normal(0, 1, (N, 2)).astype(np.float32) * pos * a[:, 2:4].min(1, keepdims=True)
        b[:, 2:4] *= np.exp(rng.normal(0, size, (N, 2))).astype(np.float32)
        b[:, 4] += rng.normal(0, ang, N).astype(np.float32) if ang else 0
        return b

    def at_threshold(a):              # the same box shifted along its own width axis so that IoU = thr (+- 1e-4 relative)
        d = a[:, 2] * (1 - thr) / (1 + thr) + rng.normal(0, 1e-4, N) * a[:, 2]
        rad = np.deg2rad(a[:, 4].astype(np.float64))
        b = a.copy()
        b[:, 0] += (d * np.cos(rad)).astype(np.float32)
        b[:, 1] += (d * np.sin(rad)).astype(np.float32)
        return b
    a = boxes(8, 300, 1024); yield 'near', a, near(a, 0.5, 0.3, 20), 1e-4
    a = boxes(8, 300, 1024); yield 'at_threshold', a, at_threshold(a), 1e-4
    a = boxes(0.5, 2000, 2048); yield 'extreme_aspect', a, near(a, 0.3, 0.2, 5), 2e-4
    a = boxes(0.05, 1.0, 2048); yield 'subpixel_far', a, near(a, 0.3, 0.2, 10), 2e-4
    a = boxes(50, 2000, 4096); yield 'image_sized', a, near(a, 0.3, 0.2, 10), 1e-4
    a = boxes(8, 300, 1024); yield 'duplicates', a, near(a, 1e-4, 1e-5, 1e-3), 1e-4
    yield 'random_dense', boxes(8, 300, 200), boxes(8, 300, 200), 1e-4
    a = boxes(8, 300, 1024, ang=0); a[:, 4] = rng.choice([0, 90, -90, 45, 180, -180, 360], N); yield 'special_angles', a, near(a, 0.4, 0.2, 0), 1e-4
    a = boxes(1, 300, 1024); a[::3, 2] = 0; a[1::3, 3] = 0; yield 'zero_size', a, near(a, 0.2, 0.1, 5), 1e-4


@pytest.mark.parametrize('thr,ge', [(0.45, 1), (0.45, 0), (0.7, 1), (0.1, 1)])
def test_pair_decision_against_oracle(harness, thr, ge):
    rng = np.random.default_rng(int(thr * 100) + ge)
    for name, a, b, f32_bound in families(rng, thr):
        o = harness(a, b, thr, ge)
        ref = oracle_pairs(a, b)
        exact = (o['iou64'] >= thr) if ge else (o['iou64'] > thr)
        decided = o['decision'].astype(bool)
        assert np.array_equal(decided, exact), (name, int((decided != exact).sum()))      # incl. every culled pair
        # against the oracle: same decision unless the IoU is within the sin/cos last-bit band of the threshold
        want = (ref >= thr) if ge else (ref > thr)
        band = np.maximum(1e-6, 2 * np.abs(o['iou64'] - ref))
        assert not ((decided != want) & (np.abs(ref - thr) > band)).any(), name
        # the exact path reproduces the oracle (identical corners -> identical bits; a last-bit sin/cos difference
        # between libm's sinf and the double-rounded sine moves thin boxes by up to ~1e-4 in IoU)
        d = np.abs(o['iou64'] - ref)
        assert np.quantile(d, 0.98) < 1e-9 and d.max() < 5e-4, (name, float(d.max()))
        # float32 clip error of every pair that reaches it: well inside the 1e-3 re-check band
        alive = o['stage'] >= 2
        if alive.any():
            assert float(np.abs(o['iou32'][alive] - o['iou64'][alive]).max()) < f32_bound, name
        if name == 'at_threshold':
            assert (o['stage'] == 3).mean() > 0.99 and 0.3 < want.mean() < 0.7        # the family does sit at the threshold


def test_float_vs_double_threshold_helpers(tmp_path):
    """csrc/common.cuh: float_at_or_below / float_at_or_above turn torchvision's float-IoU-vs-double-threshold comparison
    into a float-only one.  tests/host_harness/threshold_host.cpp checks the defining equivalence on the floats around
    the cut for 10 M thresholds (random, exactly representable, one double ulp off a float)."""
    exe = str(tmp_path / 'threshold_host')
    cuda_inc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')
    res = subprocess.run(['g++', '-O2', '-std=c++17', '-I', cuda_inc, '-o', exe,
                          os.path.join(ROOT, 'tests', 'host_harness', 'threshold_host.cpp')],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    res = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert res.returncode == 0 and res.stdout.strip().endswith(' 0 violations'), res.stdout
