"""BASELINE.json configs[4]: the dense-scene sweep (single class, no top-k cap) with a batch of 256 images sharded
over the GPUs of one node.  Every rank decodes and suppresses its own contiguous block of images (pipeline.shard_range),
then ONE all-gather of the packed variable-length detections gives every rank the whole result.

    python scripts/dense_sharded.py [img_size]                                     # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        scripts/dense_sharded.py [img_size]

Prints one JSON line (rank 0): images/s (max over ranks, CUDA events), and a digest of the gathered result that must
be identical for every N (the images are generated from per-image seeds)."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from mydetection_b200 import pipeline as pl
from mydetection_b200.heads import yolo_head_views

TOTAL = 256
img_s = int(sys.argv[1]) if len(sys.argv) > 1 else 704
world, rank, local = (int(os.environ.get(k, d)) for k, d in (('WORLD_SIZE', '1'), ('RANK', '0'), ('LOCAL_RANK', '0')))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
lo, hi = pl.shard_range(TOTAL, rank, world)
strides = (8, 16, 32)
raws = []
for s in strides:
    n = img_s // s
    t = torch.empty(hi - lo, 6, n, n)
    for i in range(lo, hi):                                   # per-image seeds: the batch does not depend on N
        g = torch.Generator().manual_seed(100000 + 10 * i + strides.index(s))
        t[i - lo] = torch.randn(6, n, n, generator=g) * 0.5
        t[i - lo, 4] = torch.randn(n, n, generator=g) * 1.5 + 2.0
    raws.append({k: v[:, 0].to(dev) for k, v in yolo_head_views(t, 1, 4, 1).items()})
pipe = pl.DetectionPipeline('FCOS2', strides, 1, (img_s, img_s), 0.005, 0.45, None)
bc = pipe.bind(raws)
B, K, P = bc.out['box'].shape
packed = torch.empty(pl.packed_numel(B, K, P), dtype=torch.float32, device=dev)
gathered = torch.empty(world * packed.numel(), dtype=torch.float32, device=dev)


def step():
    out = bc.launch()
    if world > 1:
        pl.gather_detections(out, packed=packed, all_packed=gathered)
    return out


for _ in range(2):
    step()
torch.cuda.synchronize(dev)
if world > 1:
    dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 5
a.record()
for _ in range(iters):
    out = step()
b.record()
torch.cuda.synchronize(dev)
ms = torch.tensor([a.elapsed_time(b) / iters], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    rows, counts = pl.unpack_gathered(gathered, world, B, K, P)
else:
    rows, counts = pl.unpack_gathered(pl.pack_detections(out), 1, B, K, P)
if rank == 0:
    live = (torch.arange(K, device=dev)[None, :] < counts[:, None])
    digest = hashlib.sha1(rows[live].cpu().numpy().tobytes() + counts.cpu().numpy().tobytes()).hexdigest()[:16]
    print(json.dumps({'config': f'dense scene @{img_s}, {TOTAL} images, 1 class, no cap', 'n_gpus': world,
                      'candidates_per_image': bc.levels.n_total, 'ms_per_batch': float(ms), 'images_per_s': TOTAL / float(ms) * 1e3,
                      'kept_total': int(counts.sum()), 'digest': digest}))
if world > 1:
    dist.destroy_process_group()
