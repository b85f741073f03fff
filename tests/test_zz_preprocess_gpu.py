"""GPU parity tests of the image pre-processing path (mydet_preprocess).

STATUS: green on a B200 in the round-1 driver run (GPUTEST_r01.json); hard tests since round 2.  The arithmetic, index
mapping and workspace plan the kernels execute are also pinned on the CPU (tests/test_preprocess_host.py compiles the
same header for the host and compares bit for bit with the reference and with Pillow).
"""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu]


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.int32), b.view(np.int32))


def test_preprocess_matches_reference_fixture(golden):
    from mydetection_b200 import image_ops
    g = golden('preprocess')
    for i, case in enumerate(g['cases']):
        name, size, div, code = str(case).split('|')
        out, pad = image_ops.preprocess(g[f'pre{i}_img'], name, None if size == 'None' else int(size), int(div), code)
        torch.cuda.synchronize()
        assert same_bits(out[0].cpu().numpy(), g[f'pre{i}_out']), case
        assert (list(pad) if pad is not None else [-1] * 6) == g[f'pre{i}_pad'].tolist(), case


def test_preprocess_batch_and_layouts():
    """A batch of frames, CPU and CUDA-resident inputs, a strided (cropped) view, a caller-provided output -- against
    the oracle (pinned to the reference and to Pillow)."""
    from mydetection_b200 import image_ops
    from oracle import preprocess as op
    rng = np.random.default_rng(21)
    frames = rng.integers(0, 256, (3, 270, 480, 3), dtype=np.uint8)
    for name, size, div, code in (('resize_pad_square', 160, 32, 'RGB_1_norm'), ('resize_pad_divisible', 200, 32, 'BGR_255_norm'),
                                  ('pad_divisible', None, 32, 'RGB_1'), ('resize_pad_square', 608, 32, 'RGB_1')):
        want = np.stack([op.preprocess(f, name, size, div, code)[0] for f in frames])
        got, _ = image_ops.preprocess(frames, name, size, div, code)
        assert same_bits(got.cpu().numpy(), want), (name, 'cpu input')
        dev_frames = torch.from_numpy(frames).cuda()
        out = torch.full_like(got, float('nan'))
        got2, _ = image_ops.preprocess(dev_frames, name, size, div, code, out=out)
        assert got2 is out and same_bits(out.cpu().numpy(), want), (name, 'cuda input')
    crop = torch.from_numpy(frames).cuda()[:, 10:200, 7:300]          # row pitch and image stride larger than the crop
    want = np.stack([op.preprocess(np.ascontiguousarray(f[10:200, 7:300]), 'resize_pad_square', 96, 32, 'RGB_1_norm')[0] for f in frames])
    got, _ = image_ops.preprocess(crop, 'resize_pad_square', 96, 32, 'RGB_1_norm')
    assert same_bits(got.cpu().numpy(), want)


def test_preprocess_full_hd_frame():
    from mydetection_b200 import image_ops
    from oracle import preprocess as op
    rng = np.random.default_rng(22)
    img = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    want, pad = op.preprocess(img, 'resize_pad_square', 608, 32, 'RGB_1')
    got, pad2 = image_ops.preprocess(img, 'resize_pad_square', 608, 32, 'RGB_1')
    assert pad == pad2 and same_bits(got[0].cpu().numpy(), want)


def test_preprocess_tall_image_vertical_first():
    """More than 100 times taller than wide and shrinking in height: Pillow's vertical-first rule (Geometry.v_first)."""
    from mydetection_b200 import _lib, ops
    from oracle import preprocess as op
    rng = np.random.default_rng(23)
    frames = rng.integers(0, 256, (2, 1601, 16, 3), dtype=np.uint8)
    rs_h, rs_w = 800, 24
    want = np.stack([op.format_u8(op.resize_bilinear_u8(f, rs_h, rs_w), 'RGB_1_norm') for f in frames])
    dev = torch.device('cuda', 0)
    src = torch.from_numpy(frames).to(dev)
    out = torch.empty(2, 3, rs_h, rs_w, device=dev)
    L = _lib.lib()
    ws = ops._workspace(L.mydet_preprocess_workspace_bytes(2, 1601, 16, rs_h, rs_w), dev)
    _lib.check(L.mydet_preprocess(ops._ptr(src), 2, src.stride(0), src.stride(1), 1601, 16, rs_h, rs_w, 0, 0, rs_h, rs_w,
                                  _lib.INPUT_FORMATS['RGB_1_norm'], ops._ptr(out), ops._ptr(ws), ws.numel(), ops._stream()),
               'mydet_preprocess')
    torch.cuda.synchronize()
    assert same_bits(out.cpu().numpy(), want)
