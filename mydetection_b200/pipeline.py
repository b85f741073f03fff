"""Batched post-processing pipeline: decode -> score threshold -> per-image top-k -> per-class NMS
for a whole batch in one C call, optionally replayed as a CUDA graph, and image-sharded over the
GPUs of one node with a single gather of the final detections.

This is the B200-first shape of what the reference does one image and one level at a time
(api/detection.py:168-172 -> models/general.py:69-84 -> utils/structures.py:92-173).
"""
import torch

from . import _lib, ops

KINDS = {'YOLO': ops.KIND_YOLO, 'FCOS': ops.KIND_FCOS, 'FCOS2': ops.KIND_FCOS, 'FCOS2_ATSS': ops.KIND_FCOS,
         'RAPiD': ops.KIND_RAPID, 'RetinaNet': ops.KIND_RETINA, 'Ultralytics': ops.KIND_UV5}


class BoundCall:
    """The pipeline bound to fixed input tensors, outputs and workspace.

    Every buffer and every ctypes argument is prepared once; launch() / launch_decode() /
    launch_postprocess() are then single C calls with no Python-side allocation, so the host stays
    far ahead of the GPU (a step is two kernel launches)."""

    def __init__(self, pipe, raws, self_cleaning=True):
        self.pipe = pipe
        self.levels = ops.LevelSet(raws, pipe.strides, pipe.anchors, pipe.conf_key)
        ls, L, ptr = self.levels, _lib.lib(), ops._ptr
        dev = ls.device
        k = int(pipe.topk) if pipe.topk else 0
        cap = min(k, ls.n_total) if k > 0 else ls.n_total
        self.out = ops._alloc_dets(ls.batch, max(cap, 1), ls.n_param, dev)
        # the bound call owns its workspace for its whole life: persistent and clean (include/mydet.h), so the large-N
        # path never clears its suppression matrix wholesale after the first launch
        self.workspace = ops.detect_workspace(ls, pipe.topk, zeroed=True)
        B, N, P = ls.batch, ls.n_total, ls.n_param
        self.cand = {'box': torch.empty(B, N, P, dtype=torch.float32, device=dev),
                     'score': torch.empty(B, N, dtype=torch.float32, device=dev),
                     'cls': torch.empty(B, N, dtype=torch.int32, device=dev),
                     'idx': torch.empty(B, N, dtype=torch.int32, device=dev),
                     # zeroed here once; every launch_postprocess*() consumes the count and leaves it zero for
                     # the next decode (state_clean / consume of include/mydet.h): no memset in the step
                     'count': torch.zeros(B, dtype=torch.int32, device=dev)}
        single = (k if 0 < k < N else N) <= _lib.SMALL_K      # only the single-kernel post-process cleans it
        clean = 1 if (single and self_cleaning) else 0        # else: launch_decode() memsets the count itself
        self.self_cleaning = bool(clean)
        self.pp_workspace = ops._workspace(L.mydet_postprocess_workspace_bytes(B, N, k), dev)
        o, c = self.out, self.cand
        img_h, img_w = float(pipe.img_hw[0]), float(pipe.img_hw[1])
        self._L = L
        self._detect_args = (pipe.kind, ls.array, ls.n_levels, B, ls.n_cls, P, img_h, img_w, pipe.conf_thres, k,
                             pipe.nms_thres, ptr(o['box']), ptr(o['score']), ptr(o['cls']), ptr(o['idx']),
                             ptr(o['count']), ptr(o['status']), o['box'].shape[1], ptr(self.workspace),
                             self.workspace.numel(), 1)
        self._decode_args = (pipe.kind, ls.array, ls.n_levels, B, ls.n_cls, P, img_h, img_w, pipe.conf_thres,
                             ptr(c['box']), ptr(c['score']), ptr(c['cls']), ptr(c['idx']), ptr(c['count']), N, clean)
        self._pp_args = (ptr(c['box']), ptr(c['score']), ptr(c['cls']), 0, ptr(c['idx']), ptr(c['count']), B, N, N, P,
                         ops.BOX_CXCYWH, float('-inf'), k, pipe.nms_thres, ptr(o['box']), ptr(o['score']),
                         ptr(o['cls']), ptr(o['idx']), ptr(o['count']), ptr(o['status']), o['box'].shape[1],
                         ptr(self.pp_workspace), self.pp_workspace.numel())
        self._consume_args = (clean,)
        self.graph = None

    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    def launch(self):
        """decode + threshold + top-k + NMS: one C call (mydet_detect), asynchronous."""
        rc = self._L.mydet_detect_ws(*self._detect_args, self._stream())
        if rc:
            _lib.check(rc, 'mydet_detect_ws')
        return self.out

    # stage-wise entry points: the same two kernels as launch(), exposed so that bench.py can put
    # CUDA events around the decode kernel.  They must alternate: the decode relies on the candidate
    # count being zero, which the post-process that consumed it guarantees (state_clean / consume of
    # include/mydet.h) -- no memset in the step.  reset_state() re-arms it after a decode whose
    # candidates were not post-processed.
    def reset_state(self):
        self.cand['count'].zero_()

    def launch_decode(self):
        rc = self._L.mydet_decode_compact(*self._decode_args, self._stream())
        if rc:
            _lib.check(rc, 'mydet_decode_compact')
        return self.cand

    def launch_postprocess(self):
        rc = self._L.mydet_postprocess(*self._pp_args, *self._consume_args, self._stream())
        if rc:
            _lib.check(rc, 'mydet_postprocess')
        return self.out

    def bind_exchange(self, exchange, protocol=False, multicast=True):
        """Fuse the multi-GPU exchange into the post-process kernel: detections are stored straight
        into every rank's gathered buffer (peer memory) by the kernel's output stage.
        multicast: use the buffer's NVLS multicast mapping when the exchange has one (one store per vector instead of
        one per peer).  protocol: publish per-image sequence numbers and wait for the consumers' acknowledgements
        before overwriting (include/mydet.h); every rank must then consume each step with exchange.wait() ... release()."""
        import ctypes
        o = self.out
        B, K, P = o['box'].shape
        if exchange.batch != B or exchange.cap != K or exchange.n_param != P:
            raise ValueError('exchange buffer geometry does not match the bound call')
        mc = ctypes.c_void_p(exchange.multicast_ptr if (multicast and exchange.multicast_ptr) else None)
        self._scatter_args = self._pp_args + (exchange.peer_array, exchange.world, mc, exchange.rank, exchange.rank * B,
                                              exchange.world * B, 1 if protocol else 0)
        self.exchange = exchange
        self.exchange_multicast = bool(mc.value)
        self.exchange_protocol = bool(protocol)
        return self

    def launch_postprocess_scatter(self):
        rc = self._L.mydet_postprocess_exchange(*self._scatter_args, *self._consume_args, self._stream())
        if rc:
            _lib.check(rc, 'mydet_postprocess_exchange')
        return self.out

    def capture(self):
        """Record launch() into a CUDA graph (fixed shapes, fixed buffers)."""
        self.launch()                      # warm-up outside capture (lazy module load, attributes)
        torch.cuda.synchronize(self.levels.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.launch()
        self.graph = g
        return self

    def replay(self):
        self.graph.replay()
        return self.out


class DetectionPipeline:
    def __init__(self, pred_layer, strides, n_cls, img_hw, conf_thres, nms_thres, topk=512, anchors=None,
                 conf_key='conf'):
        if pred_layer not in KINDS:
            raise NotImplementedError(pred_layer)
        self.kind = KINDS[pred_layer]
        self.strides = list(strides)
        self.anchors = anchors
        self.n_cls = n_cls
        self.img_hw = tuple(img_hw)
        self.conf_thres, self.nms_thres, self.topk = float(conf_thres), float(nms_thres), topk
        self.conf_key = conf_key

    def bind(self, raws, self_cleaning=True):
        """self_cleaning=False: launch_decode() zeroes the candidate count itself (a memset), so it may be
        called without a matching launch_postprocess()."""
        return BoundCall(self, raws, self_cleaning)

    def __call__(self, raws):
        """raws: list (one per level) of raw dicts of CUDA head views -> detections dict."""
        return self.bind(raws).launch()


def unpack(out, to_cpu=True):
    """Detections dict -> list of (boxes, scores, classes) per image (one D2H of the counts)."""
    counts = out['count'].tolist()
    res = []
    for b, n in enumerate(counts):
        item = (out['box'][b, :n], out['score'][b, :n], out['cls'][b, :n])
        res.append(tuple(t.cpu() for t in item) if to_cpu else item)
    return res


# --------------------------------------------------------------------------------------- multi-GPU
class PeerExchange:
    """Gathered-detections buffer of every rank, mapped into this process (torch symmetric memory over
    CUDA IPC / NVLink, plus its NVLS multicast mapping where the fabric offers one), for the fused exchange of
    mydet_postprocess_exchange.

    Layout per rank (include/mydet.h): float32 rows[world*batch][cap][P+2], int32 counts[world*batch], then the
    protocol words (sequence numbers per image, acknowledgements per rank).
    `local_only` builds a one-rank exchange on a plain local tensor -- "pack into one buffer" -- which is how the
    layout and the protocol are unit-tested on one GPU."""

    def __init__(self, batch, cap, n_param, device, group=None, local_only=False):
        import ctypes
        self.batch, self.cap, self.n_param = batch, cap, n_param
        self.multicast_ptr = 0
        if local_only:
            self.world, self.rank = 1, 0
            self.buffer = torch.zeros(self.numel(1), dtype=torch.float32, device=device)
            ptrs = [self.buffer.data_ptr()]
            self.handle = None
        else:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
            self.buffer = symm.empty(self.numel(self.world), dtype=torch.float32, device=device)
            self.buffer.zero_()
            self.handle = symm.rendezvous(self.buffer, group if group is not None else dist.group.WORLD)
            off = int(getattr(self.handle, 'offset', 0) or 0)       # of this tensor inside its symmetric allocation
            ptrs = [int(p) + off for p in self.handle.buffer_ptrs]
            try:
                mc = int(self.handle.multicast_ptr or 0)
                self.multicast_ptr = mc + off if mc else 0
            except Exception:        # symmetric memory without multicast support
                self.multicast_ptr = 0
            torch.cuda.synchronize(device)
            dist.barrier(group)      # every copy is zeroed before any rank publishes into it
        self.images_total = self.world * batch
        self.peer_array = (ctypes.c_void_p * len(ptrs))(*ptrs)
        self._mc = ctypes.c_void_p(self.multicast_ptr or None)
        self.wait_status = torch.zeros(1, dtype=torch.int32, device=device)
        self.counts_snapshot = torch.zeros(self.images_total, dtype=torch.int32, device=device)

    def numel(self, world):
        nbytes = _lib.lib().mydet_exchange_buffer_bytes(world * self.batch, self.cap, self.n_param)
        return nbytes // 4

    def views(self):
        """(rows (world*batch, cap, P+2) f32, counts (world*batch,) i32) views of the local gathered buffer."""
        n_img = self.world * self.batch
        n_rows = n_img * self.cap * (self.n_param + 2)
        rows = self.buffer[:n_rows].view(n_img, self.cap, self.n_param + 2)
        counts = self.buffer[n_rows:n_rows + n_img].view(torch.int32)
        return rows, counts

    # producer side: after launch_postprocess_scatter(), on the same stream
    def publish(self, multicast=True):
        """Publish the step the post-process kernel just wrote (this rank's images): sequence numbers with release
        semantics, after the kernel boundary.  consume_counts() includes it."""
        rc = _lib.lib().mydet_exchange_publish(self.peer_array, self.world, self._mc if multicast else None, self.rank,
                                               self.rank * self.batch, self.batch, self.images_total, self.cap, self.n_param,
                                               torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, 'mydet_exchange_publish')

    # consumer side of the protocol: wait() ... kernels reading views()[0] on the same stream ... release()
    def wait(self):
        """Device-side wait (no host involvement) until every image of every rank carries the next publication;
        snapshots the counts into self.counts_snapshot.  self.wait_status becomes 1 if the bounded wait expired."""
        rc = _lib.lib().mydet_exchange_wait(ops._ptr(self.buffer), self.images_total, self.cap, self.n_param,
                                            ops._ptr(self.counts_snapshot), ops._ptr(self.wait_status),
                                            torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, 'mydet_exchange_wait')
        return self.counts_snapshot

    def consume_counts(self, multicast=True):
        """publish() + wait() + release() in one launch, for a consumer that needs only the counts of the step."""
        rc = _lib.lib().mydet_exchange_consume_counts(self.peer_array, self.world, self._mc if multicast else None, self.rank,
                                                      self.rank * self.batch, self.batch, self.images_total, self.cap, self.n_param,
                                                      ops._ptr(self.counts_snapshot), ops._ptr(self.wait_status),
                                                      torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, 'mydet_exchange_consume_counts')
        return self.counts_snapshot

    def release(self, multicast=True):
        """Acknowledge the publication the last wait() returned: its rows may now be overwritten by the producers."""
        rc = _lib.lib().mydet_exchange_release(self.peer_array, self.world, self._mc if multicast else None, self.rank,
                                               self.images_total, self.cap, self.n_param,
                                               torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, 'mydet_exchange_release')


def shard_range(n_images, rank, world):
    """Contiguous block of ceil(n/world) images per rank."""
    per = (n_images + world - 1) // world
    lo = min(rank * per, n_images)
    return lo, min(lo + per, n_images)


def packed_numel(batch, cap, n_param):
    return batch * (cap * (n_param + 2) + 1)


def pack_detections(out, packed=None):
    """Detections dict -> ONE float32 buffer (mydet_pack_detections): rows of (box, score, class) followed
    by the bit patterns of the per-image counts, so the exchange is a single collective."""
    B, K, P = out['box'].shape
    if packed is None:
        packed = torch.empty(packed_numel(B, K, P), dtype=torch.float32, device=out['box'].device)
    if out['box'].is_cuda:
        rc = _lib.lib().mydet_pack_detections(ops._ptr(out['box']), ops._ptr(out['score']), ops._ptr(out['cls']),
                                              ops._ptr(out['count']), B, K, P, ops._ptr(packed),
                                              torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, 'mydet_pack_detections')
    else:   # host-side restatement, used by the gloo tests of the exchange logic only
        live = (torch.arange(K)[None, :] < out['count'][:, None]).unsqueeze(-1)
        rows = torch.cat([out['box'], out['score'].unsqueeze(-1), out['cls'].to(torch.float32).unsqueeze(-1)], dim=-1)
        packed[:B * K * (P + 2)] = (rows * live).reshape(-1)
        packed[B * K * (P + 2):] = out['count'].to(torch.int32).view(torch.float32)
    return packed


def unpack_gathered(all_packed, world, batch, cap, n_param):
    """(world * packed_numel,) -> (rows (world*batch, cap, P+2) f32, counts (world*batch,) i32)."""
    per = packed_numel(batch, cap, n_param)
    chunks = all_packed.view(world, per)
    rows = chunks[:, :batch * cap * (n_param + 2)].reshape(world * batch, cap, n_param + 2)
    counts = chunks[:, batch * cap * (n_param + 2):].contiguous().view(torch.int32).reshape(world * batch)
    return rows, counts


def gather_detections(out, group=None, packed=None, all_packed=None):
    """The path's only exchange: every rank receives every rank's final detections.
    One collective on one fixed-capacity buffer, no host sync.  Returns the gathered flat buffer;
    unpack_gathered() gives (rows, counts) views."""
    import torch.distributed as dist
    packed = pack_detections(out, packed)
    world = dist.get_world_size(group)
    if all_packed is None:
        all_packed = packed.new_empty(world * packed.numel())
    if dist.get_backend(group) == 'nccl':
        dist.all_gather_into_tensor(all_packed, packed, group=group)
    else:  # gloo (CPU tests of the host logic)
        dist.all_gather(list(all_packed.chunk(world)), packed, group=group)
    return all_packed
