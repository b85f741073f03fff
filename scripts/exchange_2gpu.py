"""The fused detections exchange with its device-side protocol on N >= 2 GPUs (torchrun): every rank produces S steps
into ONE gathered buffer per rank (the tightest back-pressure: a producer may not start step s+1 before every rank has
consumed step s) and consumes every rank's rows with NO host barrier and no synchronisation inside the loop; ranks are
pushed out of step with device-side sleeps.  Each rank then checks its snapshot of every step against detections it
computes locally from the other ranks' (seeded) inputs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/exchange_2gpu.py [--no-multicast]
Prints one JSON line on rank 0.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from mydetection_b200 import pipeline as pl
from mydetection_b200.heads import efdet_head_views

STRIDES, IMG, N_CLS, B, S, K = (8, 16, 32, 64, 128), (256, 384), 12, 8, 40, 512


def make_inputs(rank, j, dev):
    gen = torch.Generator().manual_seed(10_000 + 97 * rank + j)
    raws = []
    for s in STRIDES:
        bb = torch.randn(B, 4, IMG[0] // s, IMG[1] // s, generator=gen) * 0.5
        cc = torch.randn(B, 1 + N_CLS, IMG[0] // s, IMG[1] // s, generator=gen) * 1.5
        raws.append({k: v.to(dev) for k, v in efdet_head_views(bb, cc).items()})
    return raws


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    pipe = pl.DetectionPipeline('FCOS2', STRIDES, N_CLS, IMG, 0.05, 0.5, K)
    ex = pl.PeerExchange(B, K, 4, dev)
    use_mc = '--no-multicast' not in sys.argv
    calls = [pipe.bind(make_inputs(rank, j, dev)).bind_exchange(ex, protocol=True, multicast=use_mc) for j in range(3)]
    # what every rank must deliver, computed locally from its seeded inputs through the plain (no exchange) path
    want = {}
    for r in range(world):
        for j in range(3):
            out = pipe.bind(make_inputs(r, j, dev)).launch()
            torch.cuda.synchronize()
            want[(r, j)] = pl.unpack_gathered(pl.pack_detections(out), 1, B, K, 4)
    rows_v, _ = ex.views()
    snap_rows = torch.empty((S,) + tuple(rows_v.shape), device=dev)
    snap_counts = torch.empty(S, world * B, dtype=torch.int32, device=dev)
    snap_status = torch.zeros(S, 2, dtype=torch.int32, device=dev)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(S):                                # no barrier, no synchronize in here
        bc = calls[s % 3]
        if (s + rank) % 5 == 0:
            torch.cuda._sleep(3_000_000)              # ~1.5 ms: this rank falls behind as producer AND consumer
        bc.launch_decode()
        bc.launch_postprocess_scatter()
        ex.publish(multicast=use_mc)
        snap_status[s, 0].copy_((bc.out['status'] & 16).max())
        counts = ex.wait()
        if (s + 2 * rank) % 7 == 0:
            torch.cuda._sleep(2_000_000)              # a slow consumer: holds the buffer while producers want to move on
        snap_rows[s].copy_(rows_v)
        snap_counts[s].copy_(counts)
        snap_status[s, 1].copy_(ex.wait_status[0])
        ex.release(multicast=use_mc)
    e1.record()
    torch.cuda.synchronize()
    ok, bad = True, []
    for s in range(S):
        for r in range(world):
            w_rows, w_counts = want[(r, s % 3)]
            c = snap_counts[s, r * B:(r + 1) * B]
            good = torch.equal(c, w_counts)
            for b in range(B):
                n = int(w_counts[b])
                good = good and torch.equal(snap_rows[s, r * B + b, :n], w_rows[b, :n])
            if not good:
                ok = False
                bad.append((s, r))
    late = int(snap_status.sum())
    flag = torch.tensor([1 if (ok and late == 0) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({'world': world, 'steps': S, 'images_per_rank': B, 'multicast': bool(calls[0].exchange_multicast),
                          'multicast_ptr_available': bool(ex.multicast_ptr), 'all_ranks_ok': bool(flag.item()),
                          'rank0_bad_steps': bad[:8], 'timeouts': late, 'ms_per_step': e0.elapsed_time(e1) / S,
                          'how': 'one gathered buffer per rank, protocol on, device-side sleeps, no host barrier in the loop'}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if bool(flag.item()) else 1)


if __name__ == '__main__':
    main()
