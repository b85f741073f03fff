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
STRIDES = (8, 16, 32, 64, 128)
N_CLS = 80
CONF_THRES = 0.005   # test.ap_conf_thres
NMS_THRES = 0.5      # test.nms_thres
TOPK = 512           # utils/structures.py:99
N_ROTATE = 3         # distinct input batches cycled through, so no step re-reads L2-resident data
UNIT = 'images/s'
FCOS_ANCHORS = [0, 64, 128, 256, 512, 100000000]     # configs/d1_fcs2.json "model.fcos.anchors"


class Workload:
    """Geometry + synthetic inputs of SURVEY.md section 8d cfg2 at one image size (640: BASELINE configs[1], 8 525
    cells per image; 768: 12 276 cells, the literal ">= 10k pre-NMS candidates" point of the north star)."""

    def __init__(self, img=640, conf_mu=2.0):
        assert img % STRIDES[-1] == 0, 'image size must be a multiple of the coarsest stride (128)'
        self.img, self.conf_mu = img, conf_mu       # conf_mu 2.0 = "all-pass": every location is a candidate
        self.loc = sum((img // s) ** 2 for s in STRIDES)
        self.metric = f'decode+NMS images/sec (EfficientDet-D1/FCOS2 head outputs, {img}x{img}, top-512, 80-class NMS)'

    def config(self, n_gpus):
        point = 'all-pass operating point' if self.conf_mu >= 2.0 else 'operating point'
        mb = BATCH * self.loc * 4 * (4 + 1 + N_CLS) / 1e6
        return {'workload': f'd1_fcs2 decode+NMS: batch 64 per GPU @{self.img}x{self.img}, 5 levels ({self.loc} loc/img), '
                            f'80 classes, {point} (conf logit ~N({self.conf_mu:g},1.5^2)), conf 0.005, top-512, nms 0.5',
                'images_per_step_per_gpu': BATCH,
                'candidates_per_image': self.loc if self.conf_mu >= 2.0 else 'measured: see roofline',
                'l2_policy': f'inputs larger than L2: {N_ROTATE} rotating {mb:.1f} MB batches',
                'sharding': f'images, {n_gpus} x {BATCH}'}

    def make_batch(self, gen, device, batch=BATCH):
        """Head outputs + the permuted views the EfficientDet head emits.  The level tensors are carved out of ONE
        slab (returned last), so that the e2e leg can move a whole batch with a single host-to-device copy."""
        from mydetection_b200.heads import efdet_head_views
        sizes = []
        for s in STRIDES:
            n = self.img // s
            sizes += [batch * 4 * n * n, batch * (1 + N_CLS) * n * n]
        slab = torch.empty(sum(sizes), dtype=torch.float32, device=device)
        raws, off = [], 0
        for li, s in enumerate(STRIDES):
            n = self.img // s
            bb = slab[off:off + sizes[2 * li]].view(batch, 4, n, n); off += sizes[2 * li]
            cc = slab[off:off + sizes[2 * li + 1]].view(batch, 1 + N_CLS, n, n); off += sizes[2 * li + 1]
            bb.normal_(0.0, 0.5, generator=gen)
            cc.normal_(0.0, 1.5, generator=gen)
            cc[:, 0] += self.conf_mu
            cc[:, 1:] -= 2.0
            raws.append(efdet_head_views(bb, cc))
        return raws, slab

    def algorithmic_bytes(self, candidates_written, batch=BATCH):
        """every logit read once (4*(4+1+C) B per location) + 28 B per candidate written."""
        return batch * self.loc * 4 * (4 + 1 + N_CLS) + 28 * candidates_written


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
def cpu_inputs(wl, batch):
    from mydetection_b200.heads import efdet_head_views
    gen = torch.Generator().manual_seed(1002)
    raws = []
    for s in STRIDES:
        n = wl.img // s
        bb = torch.randn(batch, 4, n, n, generator=gen) * 0.5
        cc = torch.randn(batch, 1 + N_CLS, n, n, generator=gen) * 1.5
        cc[:, 0] += wl.conf_mu
        cc[:, 1:] -= 2.0
        raws.append(efdet_head_views(bb, cc))
    return raws


def port_pass(wl):
    """One pass of the reference ALGORITHM restated (oracle port: torch CPU ops + C greedy NMS)."""
    from oracle import decode as od, postprocess as opp

    def run(raws_cpu):
        levels = [od.decode_fcos(r, s, (wl.img, wl.img)) for r, s in zip(raws_cpu, STRIDES)]
        box, cls, score = od.merge_levels(levels)
        kept = 0
        for b in range(box.shape[0]):
            kept += int(opp.post_process(box[b], cls[b], score[b], CONF_THRES, NMS_THRES, 'cxcywh', TOPK).numel())
        return kept
    return run


def reference_pass(wl):
    """One pass of the UNMODIFIED reference over a batch of head outputs: its own FCOSLayer.forward per level
    (models/detlayers/fcos2.py:24-69), the level concatenation and per-image ImageObjects of OneStageBBox.forward
    (models/general.py:67-84) and ImageObjects.post_process of every image (api/detection.py:172 ->
    utils/structures.py:92-173: threshold, torch.topk(512), per-class torchvision.ops.nms) -- imported from the
    byte-identical copy oracle/fetch_ref.py made (oracle/_ref).  None when that copy is absent."""
    from oracle import refload
    if not refload.available():
        return None
    refload.activate()
    from models.detlayers.fcos2 import FCOSLayer
    from utils.structures import ImageObjects
    cfg = json.load(open(os.path.join(refload.ROOT, 'configs', 'd1_fcs2.json')))
    cfg['model.fpn.out_strides'] = STRIDES          # what models/registry.py fills in when it builds the BiFPN
    layers = [FCOSLayer(level_i=i, cfg=cfg) for i in range(len(STRIDES))]
    bb_format = cfg['general.pred_bbox_format']
    img_size = (wl.img, wl.img)

    def run(raws_cpu):
        with torch.no_grad():
            dts_all = [layers[i](raw, img_size, None)[0] for i, raw in enumerate(raws_cpu)]
            batch_bbs = torch.cat([d['bbox'] for d in dts_all], dim=1)
            batch_cls_idx = torch.cat([d['class_idx'] for d in dts_all], dim=1)
            batch_scores = torch.cat([d['score'] for d in dts_all], dim=1)
            kept = 0
            for bbs, cls_idx, scores in zip(batch_bbs, batch_cls_idx, batch_scores):
                objs = ImageObjects(bboxes=bbs, cats=cls_idx, scores=scores, bb_format=bb_format, img_hw=img_size)
                kept += len(objs.post_process(CONF_THRES, NMS_THRES))
        return kept
    return run


def _timed_passes(run, raws, budget_s, max_reps=500):
    run(raws)                                        # warm-up: page in, build the C oracle
    t0 = time.perf_counter()
    reps = 0
    while True:
        run(raws)
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= max_reps:
            break
    return reps, time.perf_counter() - t0


def cpu_baseline(wl, budget_s=12.0, batch=BATCH):
    """Bounded sample of the same workload on the host cores: the unmodified reference (kind 'reference') when
    oracle/_ref is present, else the oracle port; the port's figure is kept beside it either way."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    raws = cpu_inputs(wl, batch)
    ref = reference_pass(wl)
    p_reps, p_dt = _timed_passes(port_pass(wl), raws, budget_s / 3 if ref else budget_s)
    port = {'value': batch * p_reps / p_dt, 'sample': f'{p_reps} passes over one {batch}-image batch ({p_dt:.1f} s)'}
    if ref is None:
        return {'value': port['value'], 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                'sample': port['sample'] + ' of the bench workload through the oracle port (oracle/_ref absent)'}
    reps, dt = _timed_passes(ref, raws, budget_s)
    return {'value': batch * reps / dt, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'reference',
            'sample': f'{reps} passes over one {batch}-image batch of the bench workload ({dt:.1f} s) through the unmodified '
                      'reference (FCOSLayer.forward x 5 levels, torch.cat, per-image ImageObjects.post_process with '
                      'torchvision NMS; oracle/_ref)',
            'port': port}


def run_reference(args, wl, budget_s=150.0):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores (the unmodified
    reference from oracle/_ref; the oracle port only if that copy is missing).  Each step is a bounded sample -- the
    first `n` images of the bench batch -- sized from a calibration pass so that warmup + steps end within minutes."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    full = cpu_inputs(wl, BATCH)
    run = reference_pass(wl)
    kind = 'reference' if run is not None else 'port'
    run = run or port_pass(wl)
    run(full)                                        # untimed: page in
    t0 = time.perf_counter()
    run(full)
    per_image = (time.perf_counter() - t0) / BATCH   # calibration
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    n = int(max(1, min(BATCH, budget_s / per_image / (steps + warmup))))
    raws = [{k: v[:n] for k, v in r.items()} for r in full]
    for _ in range(warmup):
        run(raws)
    t0 = time.perf_counter()
    for _ in range(steps):
        run(raws)
    dt = time.perf_counter() - t0
    val = n * steps / dt
    what = ('the unmodified reference (FCOSLayer.forward x 5 levels, torch.cat, per-image ImageObjects.post_process with '
            'torchvision NMS; oracle/_ref)' if kind == 'reference' else
            'the oracle port (torch CPU decode + C greedy NMS; oracle/_ref absent)')
    sample = f'{steps} steps x the first {n} images of the {BATCH}-image bench batch through {what}, all {cores} host threads, {dt:.1f} s'
    line = {'impl': 'reference', 'metric': wl.metric, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
            'warmup': warmup, 'ms_per_step': dt / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': wl.config(args.gpus),
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': kind, 'sample': sample},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------- rotated NMS
RAPID_ANCHORS = [[18.7807, 33.4659], [28.8912, 61.7536], [48.6849, 68.3897], [45.0668, 101.4673], [63.0952, 113.5382],
                 [81.3909, 134.4554], [91.7364, 144.9949], [137.5189, 178.4791], [194.4429, 250.7985]]


def rapid_boxes(dev, batch, n, seed=3003):
    """The n best of the 64 512 decoded RAPiD candidates of a 1024 x 1024 image (configs[2]) per image."""
    from mydetection_b200 import ops
    from mydetection_b200.heads import yolo_head_views
    gen = torch.Generator(device=dev).manual_seed(seed)
    raws = []
    for s in (8, 16, 32):
        m = 1024 // s
        t = torch.randn(batch, 18, m, m, generator=gen, device=dev) * 0.5
        v = t.view(batch, 3, 6, m, m)
        v[:, :, 4] = torch.rand(batch, 3, m, m, generator=gen, device=dev) * 6 - 3
        v[:, :, 5] = torch.randn(batch, 3, m, m, generator=gen, device=dev) * 1.5 - 1.5
        raws.append(yolo_head_views(t, 3, 5, 0))
    ls = ops.LevelSet(raws, (8, 16, 32), [RAPID_ANCHORS[0:3], RAPID_ANCHORS[3:6], RAPID_ANCHORS[6:9]])
    box, _, score = ops.decode_dense(ops.KIND_RAPID, ls, (1024, 1024))
    top = score.topk(n, dim=1).indices
    return (torch.gather(box, 1, top[..., None].expand(-1, -1, 5)).contiguous(), torch.gather(score, 1, top).contiguous())


def clustered_boxes(dev, batch, n, per_object=250, seed=4004):
    """Detector-like scene: n / per_object objects per 1024 x 1024 image, each proposed per_object times with jittered
    centre, size and angle -- 100-500 mutually overlapping boxes per object, the regime a trained detector's
    pre-NMS candidates live in (the uniform scene above has ~1.4 overlapping partners per box)."""
    gen = torch.Generator(device=dev).manual_seed(seed)
    n_obj = max(1, n // per_object)
    ctr = torch.rand(batch, n_obj, 2, generator=gen, device=dev) * 900 + 62
    wh = torch.rand(batch, n_obj, 2, generator=gen, device=dev) * 120 + 30
    ang = torch.rand(batch, n_obj, 1, generator=gen, device=dev) * 180 - 90
    obj = torch.cat([ctr, wh, ang], dim=-1).repeat_interleave(per_object, dim=1)[:, :n]
    jit = torch.randn(batch, obj.shape[1], 5, generator=gen, device=dev) * torch.tensor([6., 6., 8., 8., 6.], device=dev)
    box = obj + jit
    box[..., 2:4].clamp_(min=8.0)
    score = torch.rand(batch, obj.shape[1], generator=gen, device=dev)
    return box.contiguous(), score.contiguous()


def _time_nms_rot(rb, rs, iters, chunks=None):
    from mydetection_b200 import ops
    for _ in range(3):
        keep, cnt = ops.nms_rot(rb, rs, 0.45, chunks=chunks)
    torch.cuda.synchronize(rb.device)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        keep, cnt = ops.nms_rot(rb, rs, 0.45, chunks=chunks)
    b.record()
    torch.cuda.synchronize(rb.device)
    return a.elapsed_time(b) * 1e3 / iters, keep, cnt


def rotated_nms_metric(dev, batch=32, n=10000, iters=10, chunks=None):
    """Second half of BASELINE.json's metric: rotated-NMS us/image at 10 000 boxes (configs[2]: RAPiD @1024,
    batch 32), timed with CUDA events and checked against the oracle (exact polygon clipping, nms_rotbb control
    flow) in the same run.  Beside the batch throughput: the latency of ONE image (the serial chain sort -> mask ->
    sweep cannot hide behind other images) and a clustered, detector-like scene."""
    from oracle import iou as oi
    rb, rs = rapid_boxes(dev, batch, n)
    us_batch, keep, cnt = _time_nms_rot(rb, rs, iters, chunks)
    us = us_batch / batch
    t0 = time.perf_counter()
    want = oi.nms_rot(rb[0].cpu(), rs[0].cpu(), 0.45)
    cpu_us = (time.perf_counter() - t0) * 1e6
    ok = bool(int(cnt[0]) == want.numel() and torch.equal(keep[0, :int(cnt[0])].cpu(), want))
    us_single, _, _ = _time_nms_rot(rb[:1].contiguous(), rs[:1].contiguous(), iters, chunks=1)
    cb, cs = clustered_boxes(dev, batch, n)
    us_cl, ckeep, ccnt = _time_nms_rot(cb, cs, iters, chunks)
    cwant = oi.nms_rot(cb[0].cpu(), cs[0].cpu(), 0.45)
    cok = bool(int(ccnt[0]) == cwant.numel() and torch.equal(ckeep[0, :int(ccnt[0])].cpu(), cwant))
    pairs = n * (n - 1) / 2
    return {'us_per_image': us, 'boxes_per_image': n, 'batch': batch, 'thr': 0.45, 'kept_mean': float(cnt.float().mean()),
            'algorithmic_pairs_per_s': pairs / (us * 1e-6), 'oracle_cpu_us_per_image': cpu_us, 'matches_oracle': ok,
            'target_us_per_image': 100.0, 'single_image_latency_us': us_single,
            'clustered': {'us_per_image': us_cl / batch, 'objects_per_image': max(1, n // 250), 'boxes_per_object': 250,
                          'kept_mean': float(ccnt.float().mean()), 'matches_oracle': cok}}


# ------------------------------------------------------------------------------------- GPU arm
def check_against_oracle(wl, raws, out, images=(0, 1)):
    """The detections the timed launches left in `out` against the oracle on the same head outputs: kept indices and
    their order exact, boxes / scores within 1e-5 relative (+ 2 ulp of the image extent for FCOS's x2 - x1)."""
    from oracle import decode as od, postprocess as opp
    ok, worst = True, 0.0
    atol = 2 * float(torch.finfo(torch.float32).eps * wl.img)
    for b in images:
        cpu = [{k: v[b:b + 1].cpu() for k, v in r.items()} for r in raws]
        box, cls, score = od.merge_levels([od.decode_fcos(r, s, (wl.img, wl.img)) for r, s in zip(cpu, STRIDES)])
        want = opp.post_process(box[0], cls[0], score[0], CONF_THRES, NMS_THRES, 'cxcywh', TOPK)
        n = int(out['count'][b])
        got = out['idx'][b, :n].cpu().long()
        if n != want.numel() or not torch.equal(got, want):
            return {'ok': False, 'images': list(images), 'why': f'kept indices of image {b} differ'}
        gb, gs = out['box'][b, :n].cpu(), out['score'][b, :n].cpu()
        ok &= bool(torch.allclose(gb, box[0][want], rtol=1e-5, atol=atol) and torch.allclose(gs, score[0][want], rtol=1e-5, atol=0)
                   and torch.equal(out['cls'][b, :n].cpu(), cls[0][want]))
        worst = max(worst, float(((gs - score[0][want]).abs() / score[0][want]).max()))
    return {'ok': ok, 'images': list(images), 'max_rel_err_score': worst}


def graph_chunk(steps, pipe_steps):
    """Steps per pipelined CUDA graph: the whole run when it fits, else the largest divisor of `steps` in
    (pipe_steps/2, pipe_steps] so that the run is a whole number of replays with no tail graph; (chunk, tail)."""
    if steps <= pipe_steps:
        return steps, 0
    for c in range(pipe_steps, pipe_steps // 2, -1):
        if steps % c == 0:
            return c, 0
    return pipe_steps, steps % pipe_steps


def median(xs):
    s = sorted(xs)
    return s[len(s) // 2] if len(s) % 2 else 0.5 * (s[len(s) // 2 - 1] + s[len(s) // 2])


class Runner:
    """decode+NMS steps of one workload on this rank's GPU: buffers, launch modes, timing."""

    def __init__(self, wl, dev, rank, world, args, exchange=True):
        from mydetection_b200 import pipeline as pl
        import torch.distributed as dist
        self.wl, self.dev, self.rank, self.world, self.args, self.pl, self.dist = wl, dev, rank, world, args, pl, dist
        self.pipe = pl.DetectionPipeline('FCOS2', STRIDES, N_CLS, (wl.img, wl.img), CONF_THRES, NMS_THRES, TOPK)
        gen = torch.Generator(device=dev).manual_seed(2000 + rank)
        self.batches = [wl.make_batch(gen, dev) for _ in range(N_ROTATE)]
        self.bound = [self.pipe.bind(raws) for raws, _ in self.batches]
        self.comm = torch.cuda.Stream(dev) if world > 1 else None
        self.exchange_mode, P = 'none', 4
        no_ex = args.no_exchange or not exchange
        if world == 1 and getattr(args, 'local_exchange', False):
            # diagnostic: the fused exchange and its protocol against a LOCAL gathered buffer (one rank = producer and
            # consumer): isolates what the flags, the fence and the consumer kernels cost, without NVLink
            peers = [pl.PeerExchange(BATCH, TOPK, P, dev, local_only=True) for _ in range(N_ROTATE)]
            self.protocol = not args.no_protocol
            for bc, ex in zip(self.bound, peers):
                bc.bind_exchange(ex, protocol=self.protocol)
            self.exchange_mode = 'p2p_local_buffer_diagnostic' + ('+seq_ack_protocol' if self.protocol else '')
        if world > 1 and no_ex:
            self.exchange_mode = 'none (diagnostic: detections stay on their rank)'
        if world > 1 and not no_ex:
            self.exchange_mode = 'nccl_all_gather'
            if not args.nccl_exchange:
                try:   # fused exchange: the post-process kernel stores its rows into every rank's buffer (NVLink)
                    peers = [pl.PeerExchange(BATCH, TOPK, P, dev) for _ in range(N_ROTATE)]
                    self.protocol = not args.no_protocol
                    for bc, ex in zip(self.bound, peers):
                        bc.bind_exchange(ex, protocol=self.protocol, multicast=not args.no_multicast)
                    self.multicast = bool(self.bound[0].exchange_multicast)
                    self.exchange_mode = ('p2p_store_fused_in_postprocess' + ('+nvls_multicast' if self.multicast else '+unicast')
                                          + ('+seq_ack_protocol' if self.protocol else ''))
                except Exception as e:  # symmetric memory not available on this box: keep the NCCL exchange
                    if rank == 0:
                        sys.stderr.write(f'[bench] peer-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL\n')
            ok = torch.tensor([1 if self.exchange_mode.startswith('p2p') else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                self.exchange_mode = 'nccl_all_gather'
            self.pk = [torch.empty(pl.packed_numel(BATCH, TOPK, P), dtype=torch.float32, device=dev) for _ in range(N_ROTATE)]
            self.gathered = [torch.empty(world * self.pk[0].numel(), dtype=torch.float32, device=dev) for _ in range(N_ROTATE)]
        self.fused = self.exchange_mode.startswith('p2p')
        self.nccl = self.exchange_mode == 'nccl_all_gather'
        self.protocol = self.fused and getattr(self, 'protocol', False)
        # the consumer of the gathered detections: tiny kernels, high priority so that they are not queued behind a
        # grid-filling decode
        self.s_con = torch.cuda.Stream(dev, priority=-1) if self.protocol else None
        self.done = [torch.cuda.Event() for _ in range(N_ROTATE)]
        self.sent = [None] * N_ROTATE     # exchange of buffer j finished (its out/packed buffers may be overwritten)
        self.P = P

    # ---- one eager step
    def pp_launch(self, bc):
        return bc.launch_postprocess_scatter() if self.fused else bc.launch_postprocess()

    def consume(self, bc, after=None):
        """Exchange protocol on: this rank's consumer of the gathered detections -- a device-side wait until every
        rank has published the step (no host barrier, no cross-process stream dependency), a snapshot of the counts,
        then the acknowledgement that lets the producers overwrite the buffer.  Runs on its own stream beside the
        next steps' kernels; the producers' back-pressure is the only thing that couples it to them."""
        if not self.protocol:
            return
        cur = torch.cuda.current_stream()
        ev = after
        if ev is None:
            ev = torch.cuda.Event()
            ev.record(cur)
        with torch.cuda.stream(self.s_con):
            self.s_con.wait_event(ev)
            if self.args.consumer == 'fused':
                bc.exchange.consume_counts()
            else:
                bc.exchange.publish()
                bc.exchange.wait()
                bc.exchange.release()

    def nccl_exchange(self, i):
        """The path's only exchange as one all-gather of the packed detections, on a side
        stream behind an event, so that the next batch's decode overlaps it."""
        j = i % N_ROTATE
        self.done[j].record()
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.done[j])
            self.pl.gather_detections(self.bound[j].out, packed=self.pk[j], all_packed=self.gathered[j])
            self.sent[j] = torch.cuda.Event()
            self.sent[j].record()

    def step(self, i, ev=None):
        j = i % N_ROTATE
        bc = self.bound[j]
        if self.nccl and self.sent[j] is not None:
            torch.cuda.current_stream().wait_event(self.sent[j])
        if ev:
            ev[0].record()
        bc.launch_decode()
        if ev:
            ev[1].record()
        self.pp_launch(bc)
        self.consume(bc)
        if self.nccl:
            self.nccl_exchange(i)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    # ---- pipelined CUDA graph: decode(k+1) beside post-process(k), decodes alternating between streams
    def capture_pipeline(self, n_steps, n_dec, with_pp=True, bound=None):
        bound = bound or self.bound
        dev = self.dev
        if not hasattr(self, 's_cap'):
            self.s_cap, self.s_pp = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=self.args.pp_priority)
            self.s_dec = [self.s_cap] + [torch.cuda.Stream(dev) for _ in range(N_ROTATE - 1)]
        s_cap, s_pp, s_dec = self.s_cap, self.s_pp, self.s_dec[:n_dec]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s_cap):
            pp_ev = []
            for s in s_dec[1:]:
                s.wait_stream(s_cap)
            for k in range(n_steps):
                j = k % N_ROTATE
                sd = s_dec[k % n_dec]
                with torch.cuda.stream(sd):
                    if with_pp and k >= N_ROTATE:
                        sd.wait_event(pp_ev[k - N_ROTATE])       # candidate buffers of batch j are free again
                    bound[j].launch_decode()
                    ev = torch.cuda.Event()
                    ev.record(sd)
                if with_pp:
                    with torch.cuda.stream(s_pp):
                        s_pp.wait_event(ev)
                        self.pp_launch(bound[j])
                        e2 = torch.cuda.Event()
                        e2.record(s_pp)
                        pp_ev.append(e2)
                    self.consume(bound[j], after=e2)
            for s in s_dec[1:]:
                s_cap.wait_stream(s)
            if with_pp:
                s_cap.wait_stream(s_pp)
                if self.protocol:
                    s_cap.wait_stream(self.s_con)
        return g

    def measure(self, steps, warmup, repeats):
        """`repeats` timed regions of exactly `steps` steps each (barrier + synchronize on both sides, CUDA events,
        max over ranks); returns the per-region times and what the roofline needs."""
        args, dev, world = self.args, self.dev, self.world
        pipelined = args.launch == 'pipelined' and not self.nccl
        n_dec = max(1, min(args.decode_streams, N_ROTATE))
        chunk, tail = graph_chunk(steps, args.pipe_steps)
        pipe_graph = tail_graph = None
        for i in range(warmup):
            self.step(i)
        if self.comm is not None:
            torch.cuda.current_stream().wait_stream(self.comm)
        if self.protocol:
            torch.cuda.current_stream().wait_stream(self.s_con)
        self.barrier()
        if pipelined:
            pipe_graph = self.capture_pipeline(chunk, n_dec)
            tail_graph = self.capture_pipeline(tail, n_dec) if tail else None
            for g in (pipe_graph, tail_graph):     # first replay of a graph uploads it: keep that out of the timed region
                if g is not None:
                    g.replay()
        self.barrier()

        regions = []
        n_ev = steps if not pipelined else 0
        dec_a = [torch.cuda.Event(enable_timing=True) for _ in range(n_ev)]
        dec_b = [torch.cuda.Event(enable_timing=True) for _ in range(n_ev)]
        t_wall0 = time.perf_counter()
        for _ in range(repeats):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.barrier()
            ev0.record()
            if pipelined:
                for _ in range(steps // chunk):
                    pipe_graph.replay()
                if tail_graph is not None:
                    tail_graph.replay()
            else:
                for i in range(steps):
                    self.step(i, (dec_a[i], dec_b[i]))
                if self.comm is not None:
                    torch.cuda.current_stream().wait_stream(self.comm)
                if self.protocol:
                    torch.cuda.current_stream().wait_stream(self.s_con)
            ev1.record()
            self.barrier()
            ms = ev0.elapsed_time(ev1)
            if world > 1:
                t = torch.tensor([ms], device=dev)
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
                ms = float(t.item())
            regions.append(ms)
        t_wall1 = time.perf_counter()
        res = {'regions_ms': regions, 'wall': (t_wall0, t_wall1), 'pipelined': pipelined, 'graph_steps': chunk, 'graph_tail': tail}
        if not pipelined:
            res['decode_event_ms'] = sum(a.elapsed_time(b) for a, b in zip(dec_a, dec_b)) / len(dec_a)
        return res

    def decode_only(self, n_launch=60, n_dec=2, repeats=5):
        """The decode kernel on its own, back to back: a CUDA graph of n_launch decode launches (no post-process)
        alternating between n_dec streams as in the pipelined step, and a lone event-bracketed eager launch.  Uses
        bound calls whose decode zeroes the candidate count itself (a 256-byte memset node per launch)."""
        own = [self.pipe.bind(raws, self_cleaning=False) for raws, _ in self.batches]
        for bc in own:
            bc.launch_decode()
        torch.cuda.synchronize(self.dev)
        g = self.capture_pipeline(n_launch, n_dec, with_pp=False, bound=own)
        g.replay()
        torch.cuda.synchronize(self.dev)
        times = []
        for _ in range(repeats):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record()
            torch.cuda.synchronize(self.dev)
            times.append(a.elapsed_time(b) / n_launch)
        lone = []
        for i in range(12):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); own[i % N_ROTATE].launch_decode(); b.record()
            torch.cuda.synchronize(self.dev)
            lone.append(a.elapsed_time(b))
        cand = int(own[0].cand['count'].clamp(max=own[0].cand['box'].shape[1]).sum().item())
        del own
        return median(times), median(lone), cand

    def verify_exchange(self):
        """Outside the timed region: what the exchange delivered must equal an NCCL all-gather of the separately
        packed detections of every rank."""
        pl, dist, dev = self.pl, self.dist, self.dev
        bc = self.bound[0]
        bc.launch_decode()
        self.pp_launch(bc)
        if self.protocol:            # the consumer's snapshot is taken behind the device-side wait, before any barrier
            self.consume(bc)
            torch.cuda.current_stream().wait_stream(self.s_con)
        ref = pl.gather_detections(bc.out)
        self.barrier()
        want_rows, want_counts = pl.unpack_gathered(ref, self.world, BATCH, TOPK, self.P)
        got_rows, got_counts = bc.exchange.views() if self.fused else (want_rows, want_counts)
        live = torch.arange(TOPK, device=dev)[None, :] < want_counts[:, None]
        good = torch.equal(got_counts, want_counts) and torch.equal(got_rows[live], want_rows[live])
        if self.protocol:
            ex = bc.exchange
            good = good and torch.equal(ex.counts_snapshot, want_counts) and int(ex.wait_status) == 0
            late = sum(int((b.out['status'] & 16).sum()) for b in self.bound)
            good = good and late == 0
        flag = torch.tensor([1 if good else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item())

    def e2e(self, steps):
        """The public API (DetectionPipeline.bind(...).launch() = one mydet_detect call) with HOST buffers: H2D of the
        step's head outputs from one pinned slab + the same kernels + D2H of the detections into pinned memory, all
        inside the timed region."""
        slab0 = self.batches[0][1]
        host_in = slab0.cpu().pin_memory()
        bc = self.bound[0]
        host_out = {k: torch.empty_like(v, device='cpu').pin_memory() for k, v in bc.out.items() if k != 'status'}
        h2d = host_in.numel() * 4
        d2h = sum(v.numel() * v.element_size() for v in host_out.values())

        def one():
            slab0.copy_(host_in, non_blocking=True)
            bc.launch()
            for k, v in host_out.items():
                v.copy_(bc.out[k], non_blocking=True)

        for _ in range(3):
            one()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return {'value': self.world * BATCH * steps / (ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'steps': steps, 'ms_per_step': ms / steps}


def dropin_flow(wl, runner, iters=5):
    """The reference-shaped call sequence on the mirror classes, device tensors in, per-image host results out:
    `det_layers[i](raw, img_size, None)` per level -> three torch.cat -> one ImageObjects per image ->
    `post_process(conf, nms)` per image (models/general.py:67-84 + api/detection.py:172), i.e. what the UNEDITED
    reference does once dropin.install() is active -- beside DetectionPipeline (one mydet_detect call) + unpack()."""
    from mydetection_b200 import detlayers, structures, pipeline as pl
    cfg = {'model.fcos.anchors': FCOS_ANCHORS, 'model.fpn.out_strides': STRIDES, 'general.num_class': N_CLS,
           'model.fcos2.ignored_threshold': 0.7, 'general.pred_bbox_format': 'cxcywh', 'model.pred_layer': 'FCOS2',
           'model.fpn.out_channels': [88] * len(STRIDES)}
    layer_cls = detlayers.get_det_layer(cfg)
    layers = [layer_cls(level_i=i, cfg=cfg) for i in range(len(STRIDES))]
    raws = runner.batches[0][0]
    img_size = (wl.img, wl.img)

    def flow():
        dts_all = [layers[i](raw, img_size, None)[0] for i, raw in enumerate(raws)]
        batch_bbs = torch.cat([d['bbox'] for d in dts_all], dim=1)
        batch_cls_idx = torch.cat([d['class_idx'] for d in dts_all], dim=1)
        batch_scores = torch.cat([d['score'] for d in dts_all], dim=1)
        objs = [structures.ImageObjects(bboxes=b, cats=c, scores=s, bb_format='cxcywh', img_hw=img_size)
                for b, c, s in zip(batch_bbs, batch_cls_idx, batch_scores)]
        return [o.post_process(CONF_THRES, NMS_THRES) for o in objs]

    def direct():
        return pl.unpack(runner.bound[0].launch())

    def wall(fn):
        fn()
        torch.cuda.synchronize(runner.dev)
        t0 = time.perf_counter()
        for _ in range(iters):
            res = fn()
        torch.cuda.synchronize(runner.dev)
        return (time.perf_counter() - t0) / iters, res

    t_flow, res_flow = wall(flow)
    t_dir, res_dir = wall(direct)
    same = all(len(a) == b[0].shape[0] and torch.equal(a.bboxes, b[0]) and torch.equal(a.scores, b[1]) and torch.equal(a.cats, b[2])
               for a, b in zip(res_flow, res_dir))
    return {'images_per_s': BATCH / t_flow, 'ms_per_batch': t_flow * 1e3, 'pipeline_unpack_images_per_s': BATCH / t_dir,
            'pipeline_unpack_ms_per_batch': t_dir * 1e3, 'ratio_to_pipeline': t_flow / t_dir, 'identical_results': bool(same),
            'how': 'wall clock, device-resident head outputs in, per-image CPU ImageObjects out; batch 64'}


def decode_roofline(wl, region_ms_per_step, pipelined, decode_event_ms, dec_only_ms, lone_ms, cand):
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    alg = wl.algorithmic_bytes(cand)
    # Decode-kernel time per launch.  Un-pipelined modes: CUDA events around the launch, inside the timed region.
    # Pipelined mode: decodes of consecutive steps overlap on two streams, so a per-launch start-to-end time would
    # count the shared interval twice; the decode streams are busy for the whole timed region, hence time per launch
    # = timed region / launches, which can only UNDERSTATE the kernel (the region also holds the post-process
    # kernels).  `kernel_ms_decode_only` is the same launch pattern WITHOUT the post-process, measured directly.
    kernel_ms = region_ms_per_step if pipelined else decode_event_ms
    achieved = alg / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'decode_traffic.json')))
        if wl.img == 640:
            traffic = tj.get('dram_bytes_per_launch')
            traffic_src = ('profiles/decode_traffic.json: ncu --set full capture of this kernel on this workload '
                           f'({tj.get("source", "committed")}); not measured by this run')
    except (OSError, ValueError):
        pass
    return {'bound': 'hbm', 'kernel': 'decode_kernel<FCOS,compact>', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
            'frac': achieved / peak, 'traffic': traffic, 'traffic_source': traffic_src,
            'algorithmic_bytes_per_launch': alg, 'candidates_written_per_launch': cand, 'kernel_ms': kernel_ms,
            'kernel_ms_how': ('timed region / launches (overlapped decode launches, see DESIGN.md section 6)'
                              if pipelined else 'CUDA events around each launch in the timed region'),
            'kernel_ms_decode_only': dec_only_ms, 'decode_only_gbs': alg / (dec_only_ms * 1e-3) / 1e9,
            'decode_only_frac': alg / (dec_only_ms * 1e-3) / 1e9 / peak,
            'decode_only_how': 'CUDA graph of 60 decode launches alternating between 2 streams, no post-process, median of 5',
            'single_eager_launch_ms': lone_ms, 'single_eager_launch_gbs': alg / (lone_ms * 1e-3) / 1e9,
            'peak_source': 'MEASURED_PEAKS.json hbm_gbs (measured copy)' if 'hbm_gbs' in peaks
            else 'fallback 6650 GB/s (B200_PROFILING.md)'}


def run_gpu(args, wl):
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device visible (the GPU arm has no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    steps, warmup, repeats = max(args.steps, 1), max(args.warmup, 3), max(args.repeats, 1)

    run = Runner(wl, dev, rank, world, args)
    sampler = ClockSampler(local) if rank == 0 else None
    m = run.measure(steps, warmup, repeats)
    ms = median(m['regions_ms'])
    value = world * BATCH * steps / (ms * 1e-3)
    clocks = sampler.stop(*m['wall']) if sampler else None
    # the timed launches left their detections in the output buffers: check them against the oracle (rank 0)
    parity = check_against_oracle(wl, run.batches[0][0], run.bound[0].out) if rank == 0 else None
    exchange_mode, nccl, run_protocol = run.exchange_mode, run.nccl, run.protocol
    exchange_ok = run.verify_exchange() if (world > 1 and not exchange_mode.startswith('none')) else None
    e2e = run.e2e(max(3, min(steps, 50)))
    dec_only_ms, lone_ms, cand = run.decode_only() if rank == 0 else (None, None, None)
    flow = dropin_flow(wl, run) if (rank == 0 and world == 1 and not args.no_flow) else None
    rot = rotated_nms_metric(dev) if (rank == 0 and not args.no_rot) else None

    ge10k = None
    if rank == 0 and world == 1 and wl.img == 640 and not args.no_ge10k:
        del run
        torch.cuda.empty_cache()
        wl2 = Workload(768, wl.conf_mu)
        run2 = Runner(wl2, dev, rank, world, args)
        s2 = min(steps, 480)
        m2 = run2.measure(s2, warmup, min(repeats, 3))
        ms2 = median(m2['regions_ms'])
        d2, l2, c2 = run2.decode_only()
        roof2 = decode_roofline(wl2, ms2 / s2, m2['pipelined'], m2.get('decode_event_ms'), d2, l2, c2)
        par2 = check_against_oracle(wl2, run2.batches[0][0], run2.bound[0].out, images=(0,))
        ge10k = {'candidates_per_image': wl2.loc, 'img_size': 768, 'images_per_s': BATCH * s2 / (ms2 * 1e-3), 'steps': s2,
                 'ms_per_step': ms2 / s2, 'frac': roof2['frac'], 'achieved_gbs': roof2['achieved'],
                 'decode_only_frac': roof2['decode_only_frac'], 'matches_oracle': par2['ok']}
        del run2

    if rank == 0:
        mode = 'cuda_graph_pipelined' if m['pipelined'] else 'eager'
        per_step = 3 if nccl else ((3 if args.consumer == 'fused' else 5) if run_protocol else 2)
        line = {'metric': wl.metric, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': warmup,
                'ms_per_step': ms / steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic', 'config': wl.config(world),
                'timing': {'repeats': repeats, 'region_ms': m['regions_ms'], 'value_from': 'median region of exactly `steps` steps',
                           'graph_steps': m['graph_steps'], 'graph_tail_steps': m['graph_tail']},
                'roofline': decode_roofline(wl, ms / steps, m['pipelined'], m.get('decode_event_ms'), dec_only_ms, lone_ms, cand),
                'e2e': e2e, 'matches_oracle': parity['ok'], 'parity': parity,
                # kernels of libmydet launched inside the timed regions: decode + post-process per step (+ the pack kernel
                # in front of an NCCL exchange; + the consumer's wait and release kernels under the exchange protocol),
                # over all `repeats` regions
                'gpu_launches': per_step * steps * repeats, 'gpu_launches_per_step': per_step,
                'exchange': exchange_mode, 'exchange_verified': exchange_ok, 'launch_mode': mode, 'clocks': clocks}
        if ge10k is not None:
            line['ge10k'] = ge10k
        if flow is not None:
            line['dropin_flow'] = flow
        if rot is not None:
            line['rotated_nms'] = rot
        if world == 1 and not args.no_cpu:
            line['cpu_baseline'] = cpu_baseline(wl)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--repeats', type=int, default=5, help='timed regions of exactly --steps steps; the median is reported')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg (profiling runs)')
    ap.add_argument('--launch', default='pipelined', choices=['pipelined', 'eager'],
                    help='pipelined: multi-step CUDA graph with decode(k+1) || post-process(k) [default]')
    ap.add_argument('--pipe-steps', type=int, default=96,
                    help='max steps per pipelined CUDA graph (the pipeline drains once per graph)')
    ap.add_argument('--decode-streams', type=int, default=2, help='pipelined mode: streams the decode launches alternate on')
    ap.add_argument('--pp-priority', type=int, default=-1, help='pipelined mode: priority of the post-process stream (-1 = high)')
    ap.add_argument('--nccl-exchange', action='store_true', help='N>1: use the NCCL all-gather instead of peer stores')
    ap.add_argument('--no-exchange', action='store_true', help='N>1 diagnostic: skip the detections exchange')
    ap.add_argument('--local-exchange', action='store_true',
                    help='N=1 diagnostic: run the fused exchange (and its protocol) against a local gathered buffer')
    ap.add_argument('--consumer', default='fused', choices=['fused', 'two-kernel'],
                    help='exchange protocol: publish+wait+snapshot+ack in one launch [default] or as three launches')
    ap.add_argument('--no-multicast', action='store_true', help='N>1: unicast peer stores even where an NVLS multicast mapping exists')
    ap.add_argument('--no-protocol', action='store_true',
                    help='N>1: no sequence flags / acknowledgements / consumer kernels (the round-1 behaviour: rows only)')
    ap.add_argument('--no-rot', action='store_true', help='skip the rotated-NMS side metric')
    ap.add_argument('--no-flow', action='store_true', help='skip the drop-in call-sequence leg')
    ap.add_argument('--no-ge10k', action='store_true', help='skip the 768 x 768 (12 276 candidates/image) sub-record')
    ap.add_argument('--conf-mu', type=float, default=2.0,
                    help='mean of the objectness logits (default 2.0: every cell is a candidate; -4: "trained-like", SURVEY 8d)')
    ap.add_argument('--img-size', type=int, default=640,
                    help='square input size (default 640 = BASELINE configs[1], 8 525 cells; 768 gives 12 276 cells)')
    args = ap.parse_args()
    wl = Workload(args.img_size, args.conf_mu)
    if args.impl == 'reference':
        run_reference(args, wl)
    else:
        run_gpu(args, wl)


if __name__ == '__main__':
    main()
