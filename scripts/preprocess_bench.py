"""Time mydet_preprocess (image pre-processing on the device) and check it against the oracle.

    python scripts/preprocess_bench.py [batch=64] [in_h=1080] [in_w=1920] [input_size=608]

Prints one JSON line: frames/s with the uint8 frames resident in HBM, the algorithmic bytes per frame
(3 B per source pixel read + 12 B per output pixel written) against MEASURED_PEAKS.json, the end-to-end rate with the
frames in pinned host memory (uint8 H2D inside the timed region), and the reference's own host path (Pillow resize + pad,
to_tensor, normalize) for one frame.  Inputs larger than L2: 3 rotating batches.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mydetection_b200 import image_ops          # noqa: E402
from oracle import preprocess as op             # noqa: E402

batch, in_h, in_w, size = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 64), (2, 1080), (3, 1920), (4, 608)))
name, div, code = 'resize_pad_square', 32, 'RGB_1_norm'
rng = np.random.default_rng(0)
host = [torch.from_numpy(rng.integers(0, 256, (batch, in_h, in_w, 3), dtype=np.uint8)).pin_memory() for _ in range(3)]
dev = [h.cuda() for h in host]
rs_h, rs_w, left, top, out_h, out_w, _ = image_ops.plan(in_h, in_w, name, size, div)
out = torch.empty(batch, 3, out_h, out_w, device='cuda')

want = op.preprocess(host[0][0].numpy(), name, size, div, code)[0]
got, _ = image_ops.preprocess(dev[0][:1], name, size, div, code)
ok = bool(np.array_equal(got[0].cpu().numpy().view(np.int32), want.view(np.int32)))

def timed(fn, steps=30, warmup=5):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

ms = timed(lambda i: image_ops.preprocess(dev[i % 3], name, size, div, code, out=out))
stage = [torch.empty_like(d) for d in dev]
def e2e(i):
    stage[i % 3].copy_(host[i % 3], non_blocking=True)
    image_ops.preprocess(stage[i % 3], name, size, div, code, out=out)
ms_e2e = timed(e2e)

def reference_cpu(frame):
    """What the reference's Detector runs on the host (api/detection.py:158-162): Pillow resize + pad, to_tensor, normalize."""
    import PIL.Image
    import torchvision.transforms.functional as tvf
    pil = tvf.resize(PIL.Image.fromarray(frame), (rs_h, rs_w))
    pil = tvf.pad(pil, padding=(left, top, out_w - rs_w - left, out_h - rs_h - top), fill=0)
    return tvf.normalize(tvf.to_tensor(pil), [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])

frame0 = host[0][0].numpy()
ok = ok and bool(np.array_equal(reference_cpu(frame0).numpy().view(np.int32), want.view(np.int32)))
t0 = time.perf_counter()
for _ in range(10):
    reference_cpu(frame0)
cpu_s = (time.perf_counter() - t0) / 10
peaks_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
algo = batch * (in_h * in_w * 3 + out_h * out_w * 12)
print(json.dumps({'workload': f'{batch} x {in_h}x{in_w} uint8 -> {name} {size} ({out_h}x{out_w}), {code}', 'matches_oracle': ok,
                  'frames_per_s': batch / ms * 1e3, 'ms_per_batch': ms, 'algorithmic_bytes': algo,
                  'achieved_gbs': algo / ms / 1e6, 'peak_gbs': peaks.get('hbm_gbs'),
                  'e2e_frames_per_s': batch / ms_e2e * 1e3, 'h2d_bytes_per_batch': batch * in_h * in_w * 3,
                  'reference_cpu_frames_per_s': 1.0 / cpu_s, 'reference_cpu': 'Pillow + torchvision, 1 thread'}))
