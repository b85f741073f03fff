"""CPU-only tests of the image pre-processing path.

The CUDA kernels of mydetection_b200/csrc/preprocess.cu are thin wrappers: all arithmetic, the thread-index ->
element mapping and the workspace plan live in csrc/preprocess_core.cuh as host/device functions.  Here that header
is compiled with g++ (tests/host_harness/preprocess_host.cpp, AddressSanitizer on, -ffp-contract=off) and every
work item of a call is executed on the CPU, then compared BIT FOR BIT with
  * the unmodified reference's Detector._preprocess_pil + to_tensor + format_tensor_img (tests/golden/preprocess.npz),
  * the installed Pillow / torchvision on random sizes,
so the arithmetic the GPU runs is pinned without a GPU.  The launch itself is covered by tests/test_zz_preprocess_gpu.py.
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FORMATS = {'RGB_1': 0, 'RGB_1_norm': 1, 'BGR_255_norm': 2}


@pytest.fixture(scope='module')
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp('harness') / 'preprocess_host')
    cmd = ['g++', '-O1', '-g', '-std=c++17', '-fsanitize=address', '-ffp-contract=off', '-o', exe,
           os.path.join(ROOT, 'tests', 'host_harness', 'preprocess_host.cpp')]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    work = tmp_path_factory.mktemp('io')

    def run(frames, rs_h, rs_w, left, top, out_h, out_w, code, row_pad=0, expect_ok=True):
        """frames (B,H,W,3) uint8 -> (B,3,out_h,out_w) float32 through the host build of the kernel code."""
        n_b, in_h, in_w = frames.shape[:3]
        padded = np.zeros((n_b, in_h, in_w * 3 + row_pad), dtype=np.uint8)
        padded[:, :, :in_w * 3] = frames.reshape(n_b, in_h, in_w * 3)
        if row_pad:
            padded[:, :, in_w * 3:] = 0xAB                      # pitch bytes must never be read into the result
        src, dst = str(work / 'src.bin'), str(work / 'dst.bin')
        padded.tofile(src)
        args = [exe, src, dst] + [str(int(v)) for v in (n_b, in_h, in_w, rs_h, rs_w, left, top, out_h, out_w, FORMATS[code], row_pad)]
        res = subprocess.run(args, env=dict(os.environ, ASAN_OPTIONS='detect_leaks=0'), stdout=subprocess.PIPE,
                             stderr=subprocess.PIPE, text=True)
        if not expect_ok:
            return res
        assert res.returncode == 0, res.stderr[-2000:]
        return np.fromfile(dst, dtype=np.float32).reshape(n_b, 3, out_h, out_w)
    return run


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.int32), b.view(np.int32))


def test_kernel_code_matches_reference_fixture(harness, golden):
    """Every fixture case of the unmodified reference, geometry from the mirror's plan()."""
    from mydetection_b200 import image_ops
    g = golden('preprocess')
    for i, case in enumerate(g['cases']):
        name, size, div, code = str(case).split('|')
        img = g[f'pre{i}_img']
        rs_h, rs_w, left, top, out_h, out_w, pad = image_ops.plan(img.shape[0], img.shape[1], name,
                                                                 None if size == 'None' else int(size), int(div))
        got = harness(img[None], rs_h, rs_w, left, top, out_h, out_w, code)
        assert same_bits(got[0], g[f'pre{i}_out']), case
        assert (list(pad) if pad is not None else [-1] * 6) == g[f'pre{i}_pad'].tolist(), case


def test_kernel_code_matches_pillow_random_sizes(harness):
    """Random geometries, a batch of 2 with a padded row pitch, against Pillow (through torchvision, as the reference
    calls it) followed by the oracle's to_tensor / format step (itself pinned to the reference)."""
    import PIL.Image
    import torchvision.transforms.functional as tvf
    from oracle import preprocess as op
    rng = np.random.default_rng(11)
    codes = list(FORMATS)
    for it in range(14):
        in_h, in_w, rs_h, rs_w = (int(v) for v in rng.integers(2, 160, 4))
        if it == 0:
            rs_h, rs_w = in_h, in_w                              # no resize: direct path
        if it == 1:
            rs_w = in_w                                          # one axis unchanged
        if it == 2:
            in_h, in_w, rs_h, rs_w = 7, 300, 90, 3               # extreme down-scale on one axis, up-scale on the other
        left, top = int(rng.integers(0, 9)), int(rng.integers(0, 9))
        out_h, out_w = top + rs_h + int(rng.integers(0, 9)), left + rs_w + int(rng.integers(0, 9))
        frames = rng.integers(0, 256, (2, in_h, in_w, 3), dtype=np.uint8)
        code = codes[it % 3]
        got = harness(frames, rs_h, rs_w, left, top, out_h, out_w, code, row_pad=int(it % 2) * 5)
        for b in range(2):
            canvas = np.zeros((out_h, out_w, 3), dtype=np.uint8)
            canvas[top:top + rs_h, left:left + rs_w] = np.array(tvf.resize(PIL.Image.fromarray(frames[b]), (rs_h, rs_w)))
            assert same_bits(got[b], op.format_u8(canvas, code)), (it, b, in_h, in_w, rs_h, rs_w)


def test_kernel_code_realistic_frame(harness):
    """A 1080p-shaped frame to the 608 x 608 YOLOv3 input ('resize_pad_square'): the configuration the reference's
    Detector runs, out_w a multiple of 4 (the vector-store layout)."""
    import PIL.Image
    import torchvision.transforms.functional as tvf
    from mydetection_b200 import image_ops
    from oracle import preprocess as op
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:540, 0:960]
    img = np.stack([(yy + xx) % 256, (xx * 3) % 256, (yy * 5 + xx) % 256], -1).astype(np.uint8)
    img ^= rng.integers(0, 32, img.shape, dtype=np.uint8)
    rs_h, rs_w, left, top, out_h, out_w, pad = image_ops.plan(540, 960, 'resize_pad_square', 608, 32)
    assert (rs_h, rs_w, left, top, out_h, out_w) == (342, 608, 0, 133, 608, 608) and pad == (960, 540, 0, 133, 608, 342)
    got = harness(img[None], rs_h, rs_w, left, top, out_h, out_w, 'RGB_1_norm')
    canvas = np.zeros((608, 608, 3), dtype=np.uint8)
    canvas[top:top + rs_h] = np.array(tvf.resize(PIL.Image.fromarray(img), (rs_h, rs_w)))
    assert same_bits(got[0], op.format_u8(canvas, 'RGB_1_norm'))


def test_kernel_code_tall_images_follow_pillows_pass_order(harness):
    """Pillow's Image.resize (PIL/Image.py, 12.2.0) resizes an image more than 100 times taller than wide vertically
    FIRST when the height shrinks; the uint8 intermediate is then rounded on the other axis and the bits differ.  The
    plan applies the same rule (Geometry.v_first); geometries on both sides of it, against Pillow."""
    import PIL.Image
    from oracle import preprocess as op
    rng = np.random.default_rng(5)
    for in_h, in_w, rs_h, rs_w in [(332, 2, 244, 72), (201, 2, 200, 30), (200, 2, 199, 30), (1601, 16, 800, 8), (801, 8, 802, 30),
                                   (301, 3, 4, 200), (1001, 10, 1000, 10), (501, 5, 100, 2)]:
        frames = rng.integers(0, 256, (2, in_h, in_w, 3), dtype=np.uint8)
        got = harness(frames, rs_h, rs_w, 1, 2, rs_h + 3, rs_w + 3, 'RGB_1', row_pad=3)
        for b in range(2):
            canvas = np.zeros((rs_h + 3, rs_w + 3, 3), dtype=np.uint8)
            canvas[2:2 + rs_h, 1:1 + rs_w] = np.array(PIL.Image.fromarray(frames[b]).resize((rs_w, rs_h), PIL.Image.BILINEAR))
            assert same_bits(got[b], op.format_u8(canvas, 'RGB_1')), (in_h, in_w, rs_h, rs_w)
            assert np.array_equal(op.resize_bilinear_u8(frames[b], rs_h, rs_w), canvas[2:2 + rs_h, 1:1 + rs_w])   # the oracle too


def test_plan_rejects_bad_geometry(harness):
    frames = np.zeros((1, 8, 8, 3), dtype=np.uint8)
    res = harness(frames, 8, 8, 4, 0, 8, 8, 'RGB_1', expect_ok=False)       # left + rs_w > out_w
    assert res.returncode == 3 and 'does not fit' in res.stderr


def test_mirror_plan_matches_oracle_plan():
    """image_ops.plan (host logic of the product) against the oracle's restatement of _preprocess_pil's geometry."""
    from mydetection_b200 import image_ops
    from oracle import preprocess as op
    rng = np.random.default_rng(2)
    for _ in range(300):
        h, w = (int(v) for v in rng.integers(8, 3000, 2))
        size, div = int(rng.integers(32, 1400)), int(rng.choice([1, 32, 64, 128]))
        for name in ('pad_divisible', 'resize_pad_divisible', 'resize_pad_square'):
            assert image_ops.plan(h, w, name, size, div) == op.plan(h, w, name, size, div)
    with pytest.raises(Exception, match='Unknown preprocessing name'):
        image_ops.plan(10, 10, 'nope', 32, 32)


def test_preprocess_needs_cuda_and_validates():
    """No CPU fallback, reference-style errors; the C entry point rejects bad arguments before any launch."""
    import torch
    from mydetection_b200 import _lib, image_ops
    with pytest.raises(NotImplementedError):
        image_ops.preprocess(np.zeros((4, 4, 3), np.uint8), 'pad_divisible', None, 32, 'RGB_255')
    with pytest.raises(TypeError):
        image_ops.preprocess(np.zeros((4, 4, 3), np.float32), 'pad_divisible', None, 32, 'RGB_1')
    if not torch.cuda.is_available():
        with pytest.raises(_lib.MydetError, match='no CPU fallback'):
            image_ops.preprocess(np.zeros((4, 4, 3), np.uint8), 'pad_divisible', None, 32, 'RGB_1')
    L = _lib.lib()
    rc = L.mydet_preprocess(None, 1, 0, 12, 4, 4, 4, 4, 2, 0, 4, 4, 0, None, None, 0, None)
    assert rc == -1 and b'does not fit' in L.mydet_last_error()
    rc = L.mydet_preprocess(None, 1, 0, 11, 4, 4, 4, 4, 0, 0, 4, 4, 0, None, None, 0, None)
    assert rc == -1 and b'pitch' in L.mydet_last_error()
    rc = L.mydet_preprocess(None, 1, 0, 12, 4, 4, 4, 4, 0, 0, 4, 4, 7, None, None, 0, None)
    assert rc == -1 and b'format' in L.mydet_last_error()
    assert L.mydet_preprocess_workspace_bytes(1, 1080, 1920, 342, 608) >= 1080 * 608 * 3
    assert L.mydet_preprocess_workspace_bytes(1, 608, 608, 608, 608) == 256        # padding only: no workspace


def test_predict_pil_control_flow(monkeypatch):
    """predict_pil mirrors Detector._predict_pil (api/detection.py:142-175): kwargs override the detector's defaults,
    the model sees the pre-processed batch, post_process and bboxes_to_original_ are applied in that order.
    (The device call itself is replaced here: no GPU.)"""
    import torch
    from mydetection_b200 import dropin, image_ops
    calls = []

    class Dts:
        def post_process(self, conf, nms):
            calls.append(('post_process', conf, nms))
            return self

        def bboxes_to_original_(self, pad_info):
            calls.append(('to_original', pad_info))

    class Model:
        input_format = 'RGB_1_norm'

        def __call__(self, x):
            calls.append(('model', tuple(x.shape)))
            return [Dts()]

    class Detector:                                   # the attributes Detector.__init__ sets (api/detection.py:44-53)
        divisibe, input_size, preprocess, conf_thres, nms_thres = 32, 96, 'resize_pad_square', 0.3, 0.45
        model = Model()

    def fake_preprocess(images, name, size, div, code):
        calls.append(('preprocess', name, size, div, code))
        return torch.zeros(1, 3, size, size), image_ops.plan(60, 80, name, size, div)[-1]
    monkeypatch.setattr(image_ops, 'preprocess', fake_preprocess)
    assert dropin.install_preprocess(Detector) is Detector
    out = Detector()._predict_pil(object(), input_size=64, conf_thres=0.1)
    assert isinstance(out, Dts)
    assert calls == [('preprocess', 'resize_pad_square', 64, 32, 'RGB_1_norm'), ('model', (1, 3, 64, 64)),
                     ('post_process', 0.1, 0.45), ('to_original', (80, 60, 0, 8, 64, 48))]
