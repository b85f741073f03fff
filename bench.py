#!/usr/bin/env python
"""bench.py -- decode+NMS images/sec on BASELINE.json configs[1] (EfficientDet-D1 + FCOS2,
batch 64 at 640x640, synthetic head outputs), 1..8 B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (decode -> score threshold -> top-512 -> per-class NMS) over
one batch of 64 images' head outputs that are already resident in HBM.  Prints ONE JSON line
(rank 0).  See DESIGN.md "Measurement" for what every key means and how it is measured.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# ---- workload: configs/d1_fcs2.json geometry
BATCH = 64
IMG = 640
STRIDES = (8, 16, 32, 64, 128)
N_CLS = 80
CONF_THRES = 0.005   # test.ap_conf_thres
NMS_THRES = 0.5      # test.nms_thres
TOPK = 512           # utils/structures.py:99
CONF_MU = 2.0        # "all-pass" operating point: every location is a candidate (8 525 / image)
N_ROTATE = 3         # distinct input batches cycled through, so no step re-reads L2-resident data
METRIC = 'decode+NMS images/sec (EfficientDet-D1/FCOS2 head outputs, 640x640, top-512, 80-class NMS)'
UNIT = 'images/s'


def workload_config(n_gpus):
    return {'workload': 'd1_fcs2 decode+NMS: batch 64 per GPU @640x640, 5 levels (8525 loc/img), 80 classes, '
                        'all-pass operating point (conf logit ~N(2,1.5^2)), conf 0.005, top-512, nms 0.5',
            'images_per_step_per_gpu': BATCH, 'candidates_per_image': 8525,
            'l2_policy': f'inputs larger than L2: {N_ROTATE} rotating 185.5 MB batches',
            'sharding': f'images, {n_gpus} x {BATCH}'}


def make_batch(gen, device, batch=BATCH):
    """Synthetic head outputs of SURVEY.md section 8d cfg2 + the permuted views the EfficientDet head emits."""
    from mydetection_b200.heads import efdet_head_views as efdet_views
    store, raws = [], []
    for s in STRIDES:
        n = IMG // s
        bb = torch.randn(batch, 4, n, n, generator=gen, device=device) * 0.5
        cc = torch.randn(batch, 1 + N_CLS, n, n, generator=gen, device=device) * 1.5
        cc[:, 0] += CONF_MU
        cc[:, 1:] -= 2.0
        store.append((bb, cc))
        raws.append(efdet_views(bb, cc))
    return store, raws


def algorithmic_bytes(batch, candidates_written):
    """every logit read once (4*(4+1+C) B per location) + 28 B per candidate written."""
    loc = sum((IMG // s) ** 2 for s in STRIDES)
    return batch * loc * 4 * (4 + 1 + N_CLS) + 28 * candidates_written


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={index}', f'--query-gpu={self.QUERY}',
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for t, r in self.rows if (t0 is None or t >= t0 - 0.05) and (t1 is None or t <= t1 + 0.05)] or \
               [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------- CPU arms
def cpu_pass(raws_cpu):
    """One pass of the reference algorithm (oracle port, torch CPU ops + C NMS) over a batch:
    det-layer decode of all levels, level concat, then post_process image by image."""
    from oracle import decode as od, postprocess as opp
    levels = [od.decode_fcos(r, s, (IMG, IMG)) for r, s in zip(raws_cpu, STRIDES)]
    box, cls, score = od.merge_levels(levels)
    kept = 0
    for b in range(box.shape[0]):
        kept += int(opp.post_process(box[b], cls[b], score[b], CONF_THRES, NMS_THRES, 'cxcywh', TOPK).numel())
    return kept


def cpu_inputs(batch):
    from mydetection_b200.heads import efdet_head_views as efdet_views
    gen = torch.Generator().manual_seed(1002)
    raws = []
    for s in STRIDES:
        n = IMG // s
        bb = torch.randn(batch, 4, n, n, generator=gen) * 0.5
        cc = torch.randn(batch, 1 + N_CLS, n, n, generator=gen) * 1.5
        cc[:, 0] += CONF_MU
        cc[:, 1:] -= 2.0
        raws.append(efdet_views(bb, cc))
    return raws


def cpu_baseline(budget_s=12.0, batch=BATCH):
    """Bounded sample of the same workload on the host cores (kind 'port': the reference is pure
    Python/torch and cannot travel to the GPU box; the oracle restates it with the same torch CPU ops)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    raws = cpu_inputs(batch)
    cpu_pass(raws)  # warm-up
    t0 = time.perf_counter()
    reps = 0
    while True:
        cpu_pass(raws)
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 10:
            break
    dt = time.perf_counter() - t0
    return {'value': batch * reps / dt, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'{reps} passes over one {batch}-image batch of the bench workload ({dt:.1f} s)'}


def run_reference(args, budget_s=150.0):
    """The reference's own algorithm on the host cores (oracle port; the Python reference cannot
    travel to the GPU box).  Each step is a bounded sample -- the first `n` images of the bench
    batch -- sized from a calibration pass so that warmup + steps end within a few minutes."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    full = cpu_inputs(BATCH)
    cpu_pass(full)                                   # untimed: page in, build the C oracle
    t0 = time.perf_counter()
    cpu_pass(full)
    per_image = (time.perf_counter() - t0) / BATCH   # calibration
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    n = int(max(1, min(BATCH, budget_s / per_image / (steps + warmup))))
    raws = [{k: v[:n] for k, v in r.items()} for r in full]
    for _ in range(warmup):
        cpu_pass(raws)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_pass(raws)
    dt = time.perf_counter() - t0
    val = n * steps / dt
    sample = (f'{steps} steps x the first {n} images of the {BATCH}-image bench batch through the oracle port '
              f'(torch CPU decode, all {cores} host threads + C greedy NMS), {dt:.1f} s')
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
            'warmup': warmup, 'ms_per_step': dt / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args.gpus),
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                             'sample': sample},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch.distributed as dist
    from mydetection_b200 import pipeline as pl

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device visible (the GPU arm has no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    steps, warmup = max(args.steps, 1), max(args.warmup, 3)

    pipe = pl.DetectionPipeline('FCOS2', STRIDES, N_CLS, (IMG, IMG), CONF_THRES, NMS_THRES, TOPK)
    gen = torch.Generator(device=dev).manual_seed(2000 + rank)
    batches = [make_batch(gen, dev) for _ in range(N_ROTATE)]
    bound = [pipe.bind(raws) for _, raws in batches]
    comm = torch.cuda.Stream(dev) if world > 1 else None
    gathered = None

    def step(i):
        nonlocal gathered
        bc = bound[i % N_ROTATE]
        bc.launch_decode()
        bc.launch_postprocess()
        if world > 1:
            # the path's only exchange: final detections of every rank, on a side stream so that the
            # next batch's decode overlaps it
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                gathered = pl.gather_detections(bc.out)
        return bc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(warmup):
        step(i)
    barrier()

    # ---- timed region: whole step + the decode kernel alone (events on the launching stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dec_a = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    dec_b = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    t_wall0 = time.perf_counter()
    ev0.record()
    for i in range(steps):
        bc = bound[i % N_ROTATE]
        dec_a[i].record()
        bc.launch_decode()
        dec_b[i].record()
        bc.launch_postprocess()
        if world > 1:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                gathered = pl.gather_detections(bc.out)
    if world > 1:
        torch.cuda.current_stream().wait_stream(comm)
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    dec_ms = sum(a.elapsed_time(b) for a, b in zip(dec_a, dec_b)) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * BATCH * steps / (ms * 1e-3)

    # ---- e2e through the public API with HOST buffers: H2D of the step's head outputs from pinned
    # memory + the same launches + D2H of the detections, all inside the timed region
    store0 = batches[0][0]
    host_in = [(bb.cpu().pin_memory(), cc.cpu().pin_memory()) for bb, cc in store0]
    h2d = sum(bb.numel() * 4 + cc.numel() * 4 for bb, cc in host_in)
    bc = bound[0]
    host_out = {k: torch.empty_like(v, device='cpu').pin_memory() for k, v in bc.out.items() if k != 'status'}
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())
    e2e_steps = max(3, min(steps, 50))

    def e2e_step():
        for (hb, hc), (db, dc) in zip(host_in, store0):
            db.copy_(hb, non_blocking=True)
            dc.copy_(hc, non_blocking=True)
        bc.launch()
        for k, v in host_out.items():
            v.copy_(bc.out[k], non_blocking=True)

    for _ in range(3):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * BATCH * e2e_steps / (e2e_ms * 1e-3)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    if rank == 0:
        bound[0].launch_decode()
        torch.cuda.synchronize(dev)
        cand = int(bound[0].cand['count'].clamp(max=bound[0].cand['box'].shape[1]).sum().item())
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        alg = algorithmic_bytes(BATCH, cand)
        achieved = alg / (dec_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'decode_traffic.json'))).get('dram_bytes_per_launch')
        except (OSError, ValueError):
            pass
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': warmup,
                'ms_per_step': ms / steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(world),
                'roofline': {'bound': 'hbm', 'kernel': 'decode_kernel<FCOS,compact>', 'achieved': achieved,
                             'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                             'algorithmic_bytes_per_launch': alg, 'kernel_ms': dec_ms,
                             'peak_source': 'MEASURED_PEAKS.json hbm_gbs (measured copy)' if 'hbm_gbs' in peaks
                             else 'fallback 6650 GB/s (B200_PROFILING.md)'},
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'steps': e2e_steps, 'ms_per_step': e2e_ms / e2e_steps},
                'gpu_launches': 2 * steps, 'clocks': clocks}
        if world == 1 and not args.no_cpu:
            line['cpu_baseline'] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg (profiling runs)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
