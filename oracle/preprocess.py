"""Oracle: the image pre-processing in front of the model.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, with numpy integer arithmetic, what the reference's Detector does to a PIL image before the forward
pass (api/detection.py:158-162, :177-205):

    _preprocess_pil  -> utils/image_ops.py: resize_pil (:22-35) / pad_to_divisible (:38-52) / rect_to_square (:55-106)
                        = torchvision.transforms.functional.resize(PIL image) [Pillow's Image.resize, BILINEAR,
                        always anti-aliased] and tvf.pad(fill=0)
    tvf.to_tensor    -> uint8 HWC -> float32 CHW, divided by 255
    format_tensor_img (utils/image_ops.py:165-188): 'RGB_1' | 'RGB_1_norm' | 'BGR_255_norm'

The only non-trivial arithmetic is Pillow's resampler, a third-party dependency the reference does not pin
(installed here: Pillow 12.2.0).  Its published algorithm (libImaging/Resample.c: precompute_coeffs,
normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc) is restated below:
    * per output coordinate: centre = (i + 0.5) * scale, support = max(scale, 1) (triangle filter of radius 1),
      taps [int(centre - support + 0.5), int(centre + support + 0.5)) clipped to the image, triangle weights
      normalised to sum 1 in double precision;
    * weights converted to 22-bit fixed point with round-half-away;
    * horizontal pass over the source rows the vertical pass needs, result rounded and clipped to uint8;
      then the vertical pass on that uint8 intermediate.
PINNED: tests/test_oracle_golden.py checks this restatement bit-for-bit against Pillow itself (random sizes, up- and
down-scaling) and against the unmodified reference's Detector._preprocess_pil + format_tensor_img
(tests/golden/preprocess.npz).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resample_coeffs(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR filter and the box (0, in_size).
    Returns (bounds (out,2) int32 [first tap, tap count], kk (out, ksize) int32)."""
    scale = float(in_size) / out_size                      # (double)(in1 - in0) / outSize
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        ws, ww = [], 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w = 1.0 - a if a < 1.0 else 0.0
            ws.append(w)
            ww += w
        for x in range(xmax):
            k = ws[x] / ww if ww != 0.0 else ws[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bilinear_u8(img, out_h, out_w):
    """Image.resize((out_w, out_h), BILINEAR) of an (H, W, C) uint8 array -- Pillow's two-pass 8 bits-per-channel path."""
    in_h, in_w = img.shape[:2]
    if (in_h, in_w) == (out_h, out_w):
        return img.copy()                                   # Image.resize returns a copy when nothing changes
    if in_h > in_w * 100 and out_h < in_h and in_w != out_w:
        # PIL/Image.py (Pillow 12.2.0), Image.resize: `if self.size[1] > self.size[0] * 100 and size[1] < self.size[1]`
        # a very tall image is resized vertically first, then horizontally -- two separate resize calls
        return resize_bilinear_u8(resize_bilinear_u8(img, out_h, in_w), out_h, out_w)
    src = img.astype(np.int64)
    need_h, need_v = out_w != in_w, out_h != in_h
    bv, kv = resample_coeffs(in_h, out_h)
    first, last = int(bv[0, 0]), int(bv[-1, 0] + bv[-1, 1])
    if need_h:
        bh, kh = resample_coeffs(in_w, out_w)
        tmp = np.empty((last - first, out_w) + img.shape[2:], dtype=np.uint8)
        for xx in range(out_w):
            x0, n = int(bh[xx, 0]), int(bh[xx, 1])
            acc = np.tensordot(src[first:last, x0:x0 + n], kh[xx, :n].astype(np.int64), axes=([1], [0]))
            tmp[:, xx] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
        src, bv = tmp.astype(np.int64), bv - np.array([first, 0], dtype=np.int32)
    if not need_v:
        return src.astype(np.uint8)
    out = np.empty((out_h,) + src.shape[1:], dtype=np.uint8)
    for yy in range(out_h):
        y0, n = int(bv[yy, 0]), int(bv[yy, 1])
        acc = np.tensordot(kv[yy, :n].astype(np.int64), src[y0:y0 + n], axes=([0], [0]))
        out[yy] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
    return out


def plan(ori_h, ori_w, pre_proc_name, input_size, divisible):
    """Geometry of Detector._preprocess_pil (api/detection.py:177-205):
    returns (resized_h, resized_w, left, top, out_h, out_w, pad_info)."""
    def up(v):
        return int(np.ceil(v / divisible) * divisible)      # utils/image_ops.py:48-49
    if pre_proc_name == 'pad_divisible':
        return ori_h, ori_w, 0, 0, up(ori_h), up(ori_w), None
    if pre_proc_name == 'resize_pad_divisible':
        factor = input_size / max(ori_h, ori_w)             # resize_pil(shorter=False), utils/image_ops.py:30-33
        th, tw = round(ori_h * factor), round(ori_w * factor)
        return th, tw, 0, 0, up(th), up(tw), (ori_w, ori_h, 0, 0, tw, th)
    if pre_proc_name == 'resize_pad_square':
        scale = input_size / max(ori_w, ori_h)              # rect_to_square(aug=False), utils/image_ops.py:85-104
        rw, rh = int(ori_w * scale), int(ori_h * scale)
        left, top = (input_size - rw) // 2, (input_size - rh) // 2
        return rh, rw, left, top, input_size, input_size, (ori_w, ori_h, left, top, rw, rh)
    raise Exception('Unknown preprocessing name')


MEANS = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STDS = np.array([0.229, 0.224, 0.225], dtype=np.float32)
BGR_MEANS = np.array([102.9801, 115.9465, 122.7717], dtype=np.float32)


def format_u8(img_u8, code):
    """tvf.to_tensor + format_tensor_img (utils/image_ops.py:165-188) of an (H, W, 3) uint8 array -> (3, H, W) float32.
    All operations are single float32 IEEE operations in the reference's order."""
    t = img_u8.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    if code == 'RGB_1':
        return t
    if code == 'RGB_1_norm':
        return (t - MEANS[:, None, None]) / STDS[:, None, None]
    if code == 'BGR_255_norm':
        t = t[[2, 1, 0]] * np.float32(255)
        return (t - BGR_MEANS[:, None, None]) / np.float32(1)
    raise NotImplementedError()


def preprocess(img_u8, pre_proc_name, input_size, divisible, code):
    """(H, W, 3) uint8 -> ((3, out_h, out_w) float32, pad_info): the tensor the reference feeds to the model."""
    rh, rw, left, top, out_h, out_w, pad_info = plan(img_u8.shape[0], img_u8.shape[1], pre_proc_name, input_size, divisible)
    canvas = np.zeros((out_h, out_w, 3), dtype=np.uint8)     # tvf.pad(fill=0) on the uint8 image, BEFORE normalisation
    canvas[top:top + rh, left:left + rw] = resize_bilinear_u8(img_u8, rh, rw)
    return format_u8(canvas, code), pad_info
