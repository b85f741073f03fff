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
    point = 'all-pass operating point' if CONF_MU >= 2.0 else 'operating point'
    loc = sum((IMG // s) ** 2 for s in STRIDES)                   # 8 525 at 640 (the default), 12 276 at 768 (--img-size)
    mb = BATCH * loc * 4 * (4 + 1 + N_CLS) / 1e6
    return {'workload': f'd1_fcs2 decode+NMS: batch 64 per GPU @{IMG}x{IMG}, 5 levels ({loc} loc/img), 80 classes, '
                        f'{point} (conf logit ~N({CONF_MU:g},1.5^2)), conf 0.005, top-512, nms 0.5',
            'images_per_step_per_gpu': BATCH, 'candidates_per_image': loc if CONF_MU >= 2.0 else 'measured: see roofline',
            'l2_policy': f'inputs larger than L2: {N_ROTATE} rotating {mb:.1f} MB batches',
            'sharding': f'images, {n_gpus} x {BATCH}'}


def make_batch(gen, device, batch=BATCH):
    """Synthetic head outputs of SURVEY.md section 8d cfg2 + the permuted views the EfficientDet head emits.
    The level tensors are carved out of ONE slab (returned last), so that the e2e leg can move a whole batch
    with a single host-to-device copy."""
    from mydetection_b200.heads import efdet_head_views as efdet_views
    sizes = []
    for s in STRIDES:
        n = IMG // s
        sizes += [batch * 4 * n * n, batch * (1 + N_CLS) * n * n]
    slab = torch.empty(sum(sizes), dtype=torch.float32, device=device)
    store, raws, off = [], [], 0
    for li, s in enumerate(STRIDES):
        n = IMG // s
        bb = slab[off:off + sizes[2 * li]].view(batch, 4, n, n); off += sizes[2 * li]
        cc = slab[off:off + sizes[2 * li + 1]].view(batch, 1 + N_CLS, n, n); off += sizes[2 * li + 1]
        bb.normal_(0.0, 0.5, generator=gen)
        cc.normal_(0.0, 1.5, generator=gen)
        cc[:, 0] += CONF_MU
        cc[:, 1:] -= 2.0
        store.append((bb, cc))
        raws.append(efdet_views(bb, cc))
    return store, raws, slab


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
        if time.perf_counter() - t0 > budget_s or reps >= 500:     # ~12 s of host work
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


# ------------------------------------------------------------------------------------- rotated NMS
def rotated_nms_metric(dev, batch=32, n=10000, iters=10, chunks=None):
    """Second half of BASELINE.json's metric: rotated-NMS us/image at 10 000 boxes (configs[2]: RAPiD
    @1024, batch 32).  Boxes are the 10 000 best of 64 512 decoded candidates per image; timed with CUDA
    events; the oracle (exact polygon clipping, nms_rotbb control flow) is timed on one image beside it."""
    from mydetection_b200 import ops
    from mydetection_b200.heads import yolo_head_views
    anchors = [[18.7807, 33.4659], [28.8912, 61.7536], [48.6849, 68.3897], [45.0668, 101.4673], [63.0952, 113.5382],
               [81.3909, 134.4554], [91.7364, 144.9949], [137.5189, 178.4791], [194.4429, 250.7985]]
    gen = torch.Generator(device=dev).manual_seed(3003)
    raws = []
    for s in (8, 16, 32):
        m = 1024 // s
        t = torch.randn(batch, 18, m, m, generator=gen, device=dev) * 0.5
        v = t.view(batch, 3, 6, m, m)
        v[:, :, 4] = torch.rand(batch, 3, m, m, generator=gen, device=dev) * 6 - 3
        v[:, :, 5] = torch.randn(batch, 3, m, m, generator=gen, device=dev) * 1.5 - 1.5
        raws.append(yolo_head_views(t, 3, 5, 0))
    ls = ops.LevelSet(raws, (8, 16, 32), [anchors[0:3], anchors[3:6], anchors[6:9]])
    box, _, score = ops.decode_dense(ops.KIND_RAPID, ls, (1024, 1024))
    top = score.topk(n, dim=1).indices
    rb = torch.gather(box, 1, top[..., None].expand(-1, -1, 5)).contiguous()
    rs = torch.gather(score, 1, top).contiguous()
    for _ in range(3):
        keep, cnt = ops.nms_rot(rb, rs, 0.45, chunks=chunks)
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        keep, cnt = ops.nms_rot(rb, rs, 0.45, chunks=chunks)
    b.record()
    torch.cuda.synchronize(dev)
    us = a.elapsed_time(b) * 1e3 / iters / batch
    from oracle import iou as oi
    t0 = time.perf_counter()
    want = oi.nms_rot(rb[0].cpu(), rs[0].cpu(), 0.45)
    cpu_us = (time.perf_counter() - t0) * 1e6
    ok = bool(int(cnt[0]) == want.numel() and torch.equal(keep[0, :int(cnt[0])].cpu(), want))
    pairs = n * (n - 1) / 2
    return {'us_per_image': us, 'boxes_per_image': n, 'batch': batch, 'thr': 0.45, 'kept_mean': float(cnt.float().mean()),
            'algorithmic_pairs_per_s': pairs / (us * 1e-6), 'oracle_cpu_us_per_image': cpu_us, 'matches_oracle': ok,
            'target_us_per_image': 100.0}


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
    bound = [pipe.bind(raws) for _, raws, _ in batches]
    comm = torch.cuda.Stream(dev) if world > 1 else None
    P = 4
    exchange_mode = 'none'
    if world > 1 and args.no_exchange:
        exchange_mode = 'none (--no-exchange: diagnostic run, detections stay on their rank)'
    if world > 1 and not args.no_exchange:
        exchange_mode = 'nccl_all_gather'
        if not args.nccl_exchange:
            try:   # fused exchange: the post-process kernel stores its rows into every rank's buffer (NVLink)
                peers = [pl.PeerExchange(BATCH, TOPK, P, dev) for _ in range(N_ROTATE)]
                for bc, ex in zip(bound, peers):
                    bc.bind_exchange(ex)
                exchange_mode = 'p2p_store_fused_in_postprocess'
            except Exception as e:  # symmetric memory not available on this box: keep the NCCL exchange
                if rank == 0:
                    sys.stderr.write(f'[bench] peer-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL\n')
        ok = torch.tensor([1 if exchange_mode.startswith('p2p') else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            exchange_mode = 'nccl_all_gather'
        pk = [torch.empty(pl.packed_numel(BATCH, TOPK, P), dtype=torch.float32, device=dev) for _ in range(N_ROTATE)]
        gathered = [torch.empty(world * pk[0].numel(), dtype=torch.float32, device=dev) for _ in range(N_ROTATE)]
    fused = exchange_mode.startswith('p2p')
    done = [torch.cuda.Event() for _ in range(N_ROTATE)]
    sent = [None] * N_ROTATE     # exchange of buffer j finished (its out/packed buffers may be overwritten)

    def exchange(i):
        """The path's only exchange: one all-gather of the packed detections, on a side
        stream behind an event, so that the next batch's decode overlaps it."""
        j = i % N_ROTATE
        done[j].record()
        with torch.cuda.stream(comm):
            comm.wait_event(done[j])
            pl.gather_detections(bound[j].out, packed=pk[j], all_packed=gathered[j])
            sent[j] = torch.cuda.Event()
            sent[j].record()

    def reuse_guard(i):
        j = i % N_ROTATE
        if world > 1 and sent[j] is not None:
            torch.cuda.current_stream().wait_event(sent[j])

    def step(i):
        bc = bound[i % N_ROTATE]
        reuse_guard(i)
        bc.launch_decode()
        if fused:
            bc.launch_postprocess_scatter()
        else:
            bc.launch_postprocess()
            if world > 1 and not args.no_exchange:
                exchange(i)
        return bc

    graphs = None
    if args.launch == 'graph' and world == 1:
        # whole step (decode, post-process, pack, all-gather) as one CUDA graph per input batch
        graphs = []
        for j in range(N_ROTATE):
            step_body = (lambda j=j: (bound[j].launch_decode(), bound[j].launch_postprocess(),
                                      pl.gather_detections(bound[j].out, packed=pk[j], all_packed=gathered[j])
                                      if world > 1 else None))
            step_body()
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step_body()
            graphs.append(g)

    # steps per pipelined graph: a multiple of the buffer rotation, no longer than the run itself, so that a short
    # `--steps K` is still measured in the pipelined mode; the remainder (< PIPE_STEPS) gets its own graph
    pipe_graph, tail_graph = None, None
    PIPE_STEPS = max(N_ROTATE, min(args.pipe_steps, steps) // N_ROTATE * N_ROTATE)
    if args.launch == 'pipelined' and (world == 1 or fused or args.no_exchange):
        # one CUDA graph spanning PIPE_STEPS steps with a fork: decode(k+1) runs on the capture stream while
        # post-process(k) runs on a second, higher-priority stream (only the candidate-buffer reuse and the
        # final join order them)
        pp_launch = (lambda bc: bc.launch_postprocess_scatter()) if fused else (lambda bc: bc.launch_postprocess())
        for bc in bound:
            bc.launch_decode(); pp_launch(bc)
        barrier_early = dist.barrier if world > 1 else (lambda: None)
        torch.cuda.synchronize(dev); barrier_early()
        s_cap, s_pp = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1)
        # decode(k+1) does not depend on decode(k) (different candidate buffers): with --decode-streams 2 the
        # decodes alternate between two streams, so the launch ramp / drain tail of one overlaps the next
        n_dec = max(1, min(args.decode_streams, N_ROTATE))
        s_dec = [s_cap] + [torch.cuda.Stream(dev) for _ in range(n_dec - 1)]

        def capture_pipeline(n_steps):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s_cap):
                pp_ev = []
                for s in s_dec[1:]:
                    s.wait_stream(s_cap)
                for k in range(n_steps):
                    j = k % N_ROTATE
                    sd = s_dec[k % n_dec]
                    with torch.cuda.stream(sd):
                        if k >= N_ROTATE:
                            sd.wait_event(pp_ev[k - N_ROTATE])
                        bound[j].launch_decode()
                        ev = torch.cuda.Event()
                        ev.record(sd)
                    with torch.cuda.stream(s_pp):
                        s_pp.wait_event(ev)
                        pp_launch(bound[j])
                        e2 = torch.cuda.Event()
                        e2.record(s_pp)
                        pp_ev.append(e2)
                for s in s_dec[1:]:
                    s_cap.wait_stream(s)
                s_cap.wait_stream(s_pp)
            return g

        pipe_graph = capture_pipeline(PIPE_STEPS)
        if steps % PIPE_STEPS:
            tail_graph = capture_pipeline(steps % PIPE_STEPS)   # PIPE_STEPS is a multiple of N_ROTATE: it starts at buffer 0 too

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(warmup):
        step(i)
    for g in (pipe_graph, tail_graph):     # first replay of a graph uploads it: keep that out of the timed region
        if g is not None:
            g.replay()
    barrier()

    # ---- timed region: whole step + the decode kernel alone (events on the launching stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dec_a = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    dec_b = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    t_wall0 = time.perf_counter()
    ev0.record()
    if pipe_graph is not None:
        for i in range(steps // PIPE_STEPS):
            pipe_graph.replay()
        if tail_graph is not None:
            tail_graph.replay()
    elif graphs is not None:
        for i in range(steps):
            graphs[i % N_ROTATE].replay()
    elif args.launch == 'two-streams' and world == 1:
        # decode(i+1) on one stream, post-process(i) on a second, higher-priority stream: the 64 one-CTA-per-
        # image post-process blocks take their SMs first, the bandwidth-bound decode fills the rest
        s_dec, s_pp = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1)
        main = torch.cuda.current_stream()
        dec_done = [torch.cuda.Event() for _ in range(steps)]
        pp_done = [None] * N_ROTATE
        s_dec.wait_stream(main); s_pp.wait_stream(main)
        for i in range(steps):
            j = i % N_ROTATE
            bc = bound[j]
            with torch.cuda.stream(s_dec):
                if pp_done[j] is not None:
                    s_dec.wait_event(pp_done[j])          # candidate buffers of batch j are free again
                dec_a[i].record()
                bc.launch_decode()
                dec_b[i].record()
                dec_done[i].record()
            with torch.cuda.stream(s_pp):
                s_pp.wait_event(dec_done[i])
                bc.launch_postprocess()
                pp_done[j] = torch.cuda.Event()
                pp_done[j].record()
        main.wait_stream(s_dec); main.wait_stream(s_pp)
    else:
        for i in range(steps):
            bc = bound[i % N_ROTATE]
            reuse_guard(i)
            dec_a[i].record()
            bc.launch_decode()
            dec_b[i].record()
            if fused:
                bc.launch_postprocess_scatter()
            else:
                bc.launch_postprocess()
                if world > 1 and not args.no_exchange:
                    exchange(i)
        if world > 1:
            torch.cuda.current_stream().wait_stream(comm)
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    if graphs is not None or pipe_graph is not None:
        # kernel time for the roofline from a short eager pass over the same inputs
        for i in range(min(steps, 60)):
            dec_a[i].record(); bound[i % N_ROTATE].launch_decode(); dec_b[i].record(); bound[i % N_ROTATE].launch_postprocess()
        torch.cuda.synchronize(dev)
        dec_a, dec_b = dec_a[:min(steps, 60)], dec_b[:min(steps, 60)]
    ms = ev0.elapsed_time(ev1)
    dec_ms = sum(a.elapsed_time(b) for a, b in zip(dec_a, dec_b)) / len(dec_a)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * BATCH * steps / (ms * 1e-3)

    # ---- e2e through the public API with HOST buffers: H2D of the step's head outputs from pinned
    # memory + the same launches + D2H of the detections, all inside the timed region
    slab0 = batches[0][2]
    host_in = slab0.cpu().pin_memory()                # the batch's 10 head tensors, one pinned slab
    h2d = host_in.numel() * 4
    bc = bound[0]
    host_out = {k: torch.empty_like(v, device='cpu').pin_memory() for k, v in bc.out.items() if k != 'status'}
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())
    e2e_steps = max(3, min(steps, 50))

    def e2e_step():
        slab0.copy_(host_in, non_blocking=True)
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
    exchange_ok = None
    if world > 1 and not args.no_exchange:
        # outside the timed region: what the exchange delivered must equal an NCCL all-gather of the
        # separately packed detections of every rank
        bc = bound[0]
        bc.launch_decode()
        (bc.launch_postprocess_scatter() if fused else bc.launch_postprocess())
        ref = pl.gather_detections(bc.out)
        barrier()
        want_rows, want_counts = pl.unpack_gathered(ref, world, BATCH, TOPK, P)
        if fused:
            got_rows, got_counts = bc.exchange.views()
        else:
            got_rows, got_counts = want_rows, want_counts
        live = torch.arange(TOPK, device=dev)[None, :] < want_counts[:, None]
        good = torch.equal(got_counts, want_counts) and torch.equal(got_rows[live], want_rows[live])
        flag = torch.tensor([1 if good else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        exchange_ok = bool(flag.item())
    rot = rotated_nms_metric(dev) if (rank == 0 and not args.no_rot) else None

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
        # Decode-kernel time per launch.  Un-pipelined modes: CUDA events around the launch, inside the timed
        # region.  Pipelined mode: decodes of consecutive steps overlap on two streams, so a per-launch
        # start-to-end time would count the shared interval twice; the decode stream(s) are busy for the whole
        # timed region, hence time per launch = timed region / launches.  That can only UNDERSTATE the
        # kernel (the region also holds the post-process kernels).  The event-bracketed single eager launch
        # (which includes the ~4 us launch ramp / drain a lone launch cannot hide) is reported next to it.
        kernel_ms = (ms / steps) if pipe_graph is not None else dec_ms
        achieved = alg / (kernel_ms * 1e-3) / 1e9
        traffic = None
        try:                                                      # the ncu capture was taken on the default geometry
            traffic = None if IMG != 640 else json.load(open(os.path.join(ROOT, 'profiles', 'decode_traffic.json'))).get('dram_bytes_per_launch')
        except (OSError, ValueError):
            pass
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': warmup,
                'ms_per_step': ms / steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(world),
                'roofline': {'bound': 'hbm', 'kernel': 'decode_kernel<FCOS,compact>', 'achieved': achieved,
                             'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                             'algorithmic_bytes_per_launch': alg, 'candidates_written_per_launch': cand, 'kernel_ms': kernel_ms,
                             'kernel_ms_how': ('timed region / launches (overlapped decode launches, see DESIGN.md section 6)'
                                               if pipe_graph is not None else 'CUDA events around each launch in the timed region'),
                             'single_eager_launch_ms': dec_ms, 'single_eager_launch_gbs': alg / (dec_ms * 1e-3) / 1e9,
                             'peak_source': 'MEASURED_PEAKS.json hbm_gbs (measured copy)' if 'hbm_gbs' in peaks
                             else 'fallback 6650 GB/s (B200_PROFILING.md)'},
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'steps': e2e_steps, 'ms_per_step': e2e_ms / e2e_steps},
                'gpu_launches': (3 if (world > 1 and not fused) else 2) * steps, 'exchange': exchange_mode, 'exchange_verified': exchange_ok, 'launch_mode': 'cuda_graph_pipelined' if pipe_graph else 'cuda_graph' if graphs else ('eager_two_streams' if (args.launch == 'two-streams' and world == 1) else 'eager'),
                'clocks': clocks}
        if rot is not None:
            line['rotated_nms'] = rot
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
    ap.add_argument('--launch', default='pipelined', choices=['pipelined', 'graph', 'eager', 'two-streams'],
                    help='pipelined: multi-step CUDA graph with decode(k+1) || post-process(k) [default]')
    ap.add_argument('--pipe-steps', type=int, default=96,
                    help='steps per pipelined CUDA graph (the pipeline drains once per graph: 24 -> 96 steps is +2 %%)')
    ap.add_argument('--decode-streams', type=int, default=2, help='pipelined mode: streams the decode launches alternate on')
    ap.add_argument('--nccl-exchange', action='store_true', help='N>1: use the NCCL all-gather instead of peer stores')
    ap.add_argument('--no-exchange', action='store_true', help='N>1 diagnostic: skip the detections exchange')
    ap.add_argument('--no-rot', action='store_true', help='skip the rotated-NMS side metric')
    ap.add_argument('--conf-mu', type=float, default=None,
                    help='mean of the objectness logits (default 2.0: every cell is a candidate; -4: "trained-like", SURVEY 8d)')
    ap.add_argument('--img-size', type=int, default=640,
                    help='square input size (default 640 = BASELINE configs[1], 8 525 cells; 768 gives 12 276 cells: '
                         'the ">= 10k pre-NMS candidates" statement of the north star, SURVEY 8d)')
    args = ap.parse_args()
    global CONF_MU, IMG, METRIC
    if args.conf_mu is not None:
        CONF_MU = args.conf_mu
    if args.img_size != IMG:
        assert args.img_size % 128 == 0, '--img-size must be a multiple of the coarsest stride (128)'
        IMG = args.img_size
        METRIC = METRIC.replace('640x640', f'{IMG}x{IMG}')
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
