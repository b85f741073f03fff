"""Image pre-processing in front of the model, on the device.

Mirrors what the reference's Detector does to a PIL image before the forward pass --
`_preprocess_pil` (api/detection.py:177-205: resize_pil / pad_to_divisible / rect_to_square of
utils/image_ops.py:22-106), `tvf.to_tensor` and `format_tensor_img` (utils/image_ops.py:165-188) -- with the
same names for the pre-processing modes and input formats.  The geometry (`plan`) is the reference's own Python
arithmetic; the pixels are produced by libmydet (`mydet_preprocess`: Pillow's anti-aliased BILINEAR resize,
bit-exact, zero padding, /255, normalisation) from the uint8 frame, so only 3 bytes per source pixel cross PCIe
instead of 12 bytes per padded output pixel.  There is no CPU implementation here.
"""

import numpy as np
import torch

from . import _lib, ops


def plan(ori_h, ori_w, pre_proc_name, input_size=None, divisible=1):
    """Geometry of Detector._preprocess_pil (api/detection.py:177-205).

    Returns (resized_h, resized_w, left, top, out_h, out_w, pad_info); pad_info is what the reference hands to
    ImageObjects.bboxes_to_original_ (None for 'pad_divisible')."""
    assert isinstance(divisible, int)

    def up(v):                                                   # pad_to_divisible, utils/image_ops.py:48-50
        return int(np.ceil(v / divisible) * divisible)
    if pre_proc_name == 'pad_divisible':
        return ori_h, ori_w, 0, 0, up(ori_h), up(ori_w), None
    if pre_proc_name == 'resize_pad_divisible':
        assert input_size is not None
        factor = input_size / max(ori_h, ori_w)                  # resize_pil(shorter=False), utils/image_ops.py:30-33
        th, tw = round(ori_h * factor), round(ori_w * factor)
        return th, tw, 0, 0, up(th), up(tw), (ori_w, ori_h, 0, 0, tw, th)
    if pre_proc_name == 'resize_pad_square':
        assert input_size is not None
        scale = input_size / max(ori_w, ori_h)                   # rect_to_square(aug=False), utils/image_ops.py:85-104
        assert scale > 0
        rw, rh = int(ori_w * scale), int(ori_h * scale)
        left, top = (input_size - rw) // 2, (input_size - rh) // 2
        return rh, rw, left, top, input_size, input_size, (ori_w, ori_h, left, top, rw, rh)
    raise Exception('Unknown preprocessing name')


def _as_u8_frames(images):
    """PIL image | (H,W,3) / (B,H,W,3) uint8 ndarray or tensor  ->  (B,H,W,3) uint8 tensor (CPU or CUDA)."""
    if hasattr(images, 'mode') and hasattr(images, 'size') and not torch.is_tensor(images):   # PIL.Image.Image
        assert images.mode == 'RGB', 'input must be an RGB image'
        images = np.array(images)                                # a writable copy: torch.from_numpy warns on read-only arrays
    if isinstance(images, np.ndarray):
        images = torch.from_numpy(np.ascontiguousarray(images))
    if not torch.is_tensor(images) or images.dtype != torch.uint8:
        raise TypeError('expected a PIL image or a uint8 array / tensor of shape (H,W,3) or (B,H,W,3)')
    if images.dim() == 3:
        images = images.unsqueeze(0)
    if images.dim() != 4 or images.shape[-1] != 3:
        raise ValueError(f'expected (H,W,3) or (B,H,W,3), got {tuple(images.shape)}')
    return images


def preprocess(images, pre_proc_name, input_size=None, divisible=1, input_format='RGB_1', out=None):
    """uint8 RGB frame(s) -> ((B,3,H,W) float32 CUDA tensor ready for the model, pad_info).

    `images`: PIL image, or uint8 (H,W,3) / (B,H,W,3) array or tensor, on the CPU (copied to the device as uint8)
    or already on the device.  All frames of a batch share one geometry.  Raises NotImplementedError for an unknown
    `input_format` and Exception('Unknown preprocessing name') like the reference."""
    if input_format not in _lib.INPUT_FORMATS:
        raise NotImplementedError()
    frames = _as_u8_frames(images)
    n_b, in_h, in_w = frames.shape[:3]
    rs_h, rs_w, left, top, out_h, out_w, pad_info = plan(in_h, in_w, pre_proc_name, input_size, divisible)
    if not torch.cuda.is_available():
        raise _lib.MydetError('mydetection_b200 needs a CUDA device (B200); there is no CPU fallback')
    if not frames.is_cuda:
        frames = frames.to(torch.device('cuda', torch.cuda.current_device()), non_blocking=True)
    if frames.stride(3) != 1 or frames.stride(2) != 3:
        frames = frames.contiguous()
    dev = frames.device
    if out is None:
        out = torch.empty(n_b, 3, out_h, out_w, dtype=torch.float32, device=dev)
    else:
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (n_b, 3, out_h, out_w)
    L = _lib.lib()
    ws = ops._workspace(L.mydet_preprocess_workspace_bytes(n_b, in_h, in_w, rs_h, rs_w), dev)
    with torch.cuda.device(dev):
        _lib.check(L.mydet_preprocess(ops._ptr(frames), n_b, frames.stride(0) if n_b > 1 else in_h * frames.stride(1),
                                      frames.stride(1), in_h, in_w, rs_h, rs_w, left, top, out_h, out_w,
                                      _lib.INPUT_FORMATS[input_format], ops._ptr(out), ops._ptr(ws), ws.numel(),
                                      ops._stream()), 'mydet_preprocess')
    return out, pad_info


def predict_pil(self, pil_img, **kwargs):
    """Drop-in for Detector._predict_pil (api/detection.py:142-175) with the pre-processing on the device:
    `Detector._predict_pil = mydetection_b200.image_ops.predict_pil`.  `self` is the reference's
    Detector: its model, defaults and `divisibe` attribute are used as they are."""
    pre_proc = kwargs.get('preprocessing', self.preprocess)
    input_size = kwargs.get('input_size', self.input_size)
    conf_thres = kwargs.get('conf_thres', self.conf_thres)
    nms_thres = kwargs.get('nms_thres', self.nms_thres)
    input_, pad_info = preprocess(pil_img, pre_proc, input_size, self.divisibe, self.model.input_format)
    assert input_.dim() == 4
    with torch.no_grad():
        dts = self.model(input_)
    assert isinstance(dts, list)
    dts = dts[0].post_process(conf_thres, nms_thres)
    if pad_info is not None:
        dts.bboxes_to_original_(pad_info)
    return dts
