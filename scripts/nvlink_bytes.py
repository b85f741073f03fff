"""NVLink bytes per step of the fused detections exchange (torchrun, N >= 2): `nvidia-smi nvlink -gt d` data counters of
every rank's GPU before and after K steps of decode -> post-process with the exchange (bench workload: 64 images per rank,
top-512, all-pass point), with the NVLS multicast mapping and with unicast peer stores.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 scripts/nvlink_bytes.py
Prints one JSON line per variant on rank 0: tx / rx KiB per step and rank, against the algorithmic row bytes.
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
from mydetection_b200 import pipeline as pl


def counters(gpu):
    """(tx KiB, rx KiB) summed over the links of one GPU, or None when nvidia-smi does not report them."""
    try:
        txt = subprocess.run(['nvidia-smi', 'nvlink', '-gt', 'd', '-i', str(gpu)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                             text=True, timeout=30).stdout
    except Exception:
        return None
    tx = [int(v) for v in re.findall(r'Data Tx:\s*(\d+)\s*KiB', txt)]
    rx = [int(v) for v in re.findall(r'Data Rx:\s*(\d+)\s*KiB', txt)]
    return (sum(tx), sum(rx)) if tx and rx else None


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    wl = bench.Workload(640, 2.0)
    gen = torch.Generator(device=dev).manual_seed(3000 + rank)
    raws, _ = wl.make_batch(gen, dev)
    pipe = pl.DetectionPipeline('FCOS2', bench.STRIDES, bench.N_CLS, (wl.img, wl.img), bench.CONF_THRES, bench.NMS_THRES, bench.TOPK)
    steps = 2000
    for multicast in (True, False):
        ex = pl.PeerExchange(bench.BATCH, bench.TOPK, 4, dev)
        bc = pipe.bind(raws).bind_exchange(ex, protocol=True, multicast=multicast)
        for _ in range(20):
            bc.launch_decode(); bc.launch_postprocess_scatter(); ex.consume_counts(multicast=multicast)
        torch.cuda.synchronize(); dist.barrier()
        c0 = counters(local)
        dist.barrier()
        for _ in range(steps):
            bc.launch_decode(); bc.launch_postprocess_scatter(); ex.consume_counts(multicast=multicast)
        torch.cuda.synchronize(); dist.barrier()
        c1 = counters(local)
        mine = [-1.0, -1.0] if (c0 is None or c1 is None) else [(c1[0] - c0[0]) / steps, (c1[1] - c0[1]) / steps]
        allv = [None] * world
        dist.all_gather_object(allv, mine)
        kept = float(bc.out['count'].sum())
        if rank == 0:
            row_kib = kept * 24 / 1024            # rows this rank publishes per step: (box, score, class) = 24 B each
            print(json.dumps({'world': world, 'multicast': bool(bc.exchange_multicast), 'steps': steps,
                              'tx_kib_per_step_by_rank': [round(v[0], 1) for v in allv],
                              'rx_kib_per_step_by_rank': [round(v[1], 1) for v in allv],
                              'rows_kib_per_step_per_rank': round(row_kib, 1),
                              'expected': ('tx = rows once (the switch replicates), rx = (N-1) x rows' if bc.exchange_multicast
                                           else 'tx = rx = (N-1) x rows'),
                              'counters': 'nvidia-smi nvlink -gt d' if allv[0][0] >= 0 else 'unavailable on this box'}))
        del bc, ex
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
