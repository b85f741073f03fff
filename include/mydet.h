/* mydet.h -- C ABI of libmydet.so: the myDetection post-processing hot path on B200 (sm_100a).
 *
 * The reference (duanzhiihao/myDetection) is pure Python and has NO native/FFI boundary; its
 * plug-in boundary is Python duck typing.  This header is the boundary a
 * native replacement needs: every entry point names the reference routine it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes binding and the
 * three reference modules a maintainer would swap.
 *
 * Conventions
 *   - All data pointers are DEVICE pointers unless a parameter is documented as host memory.
 *     The library never allocates device memory: outputs and workspaces are caller-owned.
 *   - Every call is asynchronous and stream-ordered on `stream` (a cudaStream_t passed as
 *     void*; NULL = the legacy default stream).  No call synchronises the device.
 *   - Re-entrant, no global mutable state besides a thread-local last-error string.
 *   - Return value: 0 = ok; > 0 = a cudaError_t; < 0 = one of the MYDET_ERR_* codes.
 *   - Strides are in ELEMENTS, so the permuted NCHW views the reference heads hand to the
 *     det layers (models/rpns.py:29-41, :175-189) are consumed in place, without a copy.
 *   - Alignment: arrays of 4-float boxes (boxes / out_box of mydet_postprocess with n_param == 4, cand_box of
 *     mydet_decode_compact, a / b / gt of the IoU entry points) move as 128-bit vectors and must be 16-byte
 *     aligned -- a violation returns MYDET_ERR_INVALID, it never faults on the device.  Every other pointer
 *     needs only the alignment of its element type (dense-decode outputs fall back to scalar stores).
 *     Workspaces: 256 bytes.  Anything torch allocates satisfies all of this unless it is a view with an odd
 *     storage offset.
 *   - Persistent workspaces.  The large-N NMS (more than MYDET_SMALL_K survivors, and mydet_nms_rot) keeps an
 *     n x ceil(n/64)-word suppression matrix per image in its workspace: 12.6 MB at 10 000 boxes, 293 MB at 48 384 --
 *     sparse, but it must start all-zero, and clearing it costs more than any kernel of the path.  A caller that passes
 *     the SAME workspace buffer, for the SAME (batch, n_per_image, entry point), to every call may hand it over zeroed
 *     once (cudaMemset of the whole buffer) and set MYDET_PP_WS_CLEAN / workspace_clean: the library then skips the
 *     wholesale clear and, as its last step, zeroes exactly the words the call set, so the buffer is clean again for the
 *     next call.  Without the flag every call clears the matrix itself and makes no assumption about the buffer.
 *   - Tie policy (the reference leaves it open): wherever scores are ranked,
 *     equal scores are ordered by ascending candidate index.
 */
#ifndef MYDET_H_
#define MYDET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MYDET_VERSION 100
#define MYDET_MAX_LEVELS 8
#define MYDET_MAX_ANCHORS 16
#define MYDET_MAX_CLASS_ID 4095      /* class ids must lie in [0, 4095]                    */
#define MYDET_MAX_CANDIDATES 1048575 /* candidates per image must be < 2^20                */
#define MYDET_SMALL_K 1024           /* <= this many survivors: single fused kernel per image */

#define MYDET_PP_CONSUME 1           /* mydet_postprocess flags                            */
#define MYDET_PP_FORCE_SCAN 2
#define MYDET_PP_WS_CLEAN 4          /* persistent workspace in its clean state (see "persistent workspaces")  */

enum {
    MYDET_OK = 0,
    MYDET_ERR_INVALID = -1,     /* bad argument (see mydet_last_error)            */
    MYDET_ERR_WORKSPACE = -2,   /* workspace too small                             */
    MYDET_ERR_UNSUPPORTED = -3  /* shape outside the documented limits             */
};

/* Which det layer's decode arithmetic to apply (models/registry.py:119-146). */
typedef enum {
    MYDET_KIND_YOLO = 0,   /* models/detlayers/yolov3.py:41-69   YOLOLayer                   */
    MYDET_KIND_FCOS = 1,   /* models/detlayers/fcos2.py:40-69, :222-251; fcos.py:41-68       */
    MYDET_KIND_RAPID = 2,  /* models/detlayers/rapid.py:48-82    RAPiDLayer (cx,cy,w,h,deg)  */
    MYDET_KIND_RETINA = 3, /* models/detlayers/retinanet.py:63-82 RetinaLayer                */
    MYDET_KIND_UV5 = 4     /* models/detlayers/uv5.py:60-91      DetectLayer                 */
} mydet_kind_t;

typedef enum {
    MYDET_BOX_CXCYWH = 0,  /* also 'cxcywhd': the angle is carried but ignored by AABB NMS    */
    MYDET_BOX_X1Y1X2Y2 = 1
} mydet_box_format_t;

/* One pyramid level of raw head output: the tensors of the reference's raw dict
 * ('bbox', 'conf' or 'center', 'class').  Logical shapes (B,nA,nH,nW,P), (B,nA,nH,nW),
 * (B,nA,nH,nW,C); for single-anchor heads pass n_anchor = 1 and stride_a = 0. */
typedef struct {
    const float* bbox;        /* element (b=0,a=0,h=0,w=0,p=0)                                  */
    const float* conf;        /* objectness / centerness logits; NULL for MYDET_KIND_RETINA     */
    const float* cls;         /* class logits, class 0; NULL when n_cls == 0                    */
    int64_t bbox_stride[5];   /* b, a, h, w, p                                                  */
    int64_t conf_stride[4];   /* b, a, h, w                                                     */
    int64_t cls_stride[5];    /* b, a, h, w, c                                                  */
    int32_t n_anchor, n_h, n_w;
    float stride;             /* cfg['model.fpn.out_strides'][level]                            */
    float anchor_w[MYDET_MAX_ANCHORS]; /* pixels; unused by MYDET_KIND_FCOS                      */
    float anchor_h[MYDET_MAX_ANCHORS];
} mydet_level_t;

int mydet_version(void);
const char* mydet_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Decode, dense.  Replaces the test-mode branch of every det layer's forward(raw, img_size, None)
 * (files above) AND the level concatenation of OneStageBBox.forward (models/general.py:74-76):
 * all levels are decoded by ONE launch straight into the level-concatenated buffers.
 *   levels        HOST array of n_levels descriptors, stride-ascending (reference order)
 *   out_box       (B, n_total, P) float32     n_total = sum_l nA*nH*nW, candidate order a->h->w
 *   out_cls       (B, n_total)    int64       first index of the maximal class probability
 *   out_score     (B, n_total)    float32
 * n_param P is 4, or 5 for rotated kinds.  img_h/img_w: network input size (FCOS/Retina clamps). */
int mydet_decode_dense(int kind, const mydet_level_t* levels, int n_levels, int batch, int n_cls,
                       int n_param, float img_h, float img_w, float* out_box, int64_t* out_cls,
                       float* out_score, int64_t n_total, void* stream);

/* Decode fused with the score threshold of ImageObjects.post_process (utils/structures.py:98):
 * only candidates with score >= conf_thres (float32 compare) are written, compacted per image with
 * a warp-ballot + one atomic per warp.  Slot order is unspecified; cand_idx carries the flat
 * candidate index (position in the dense layout above), which later ranking uses as tie-break.
 *   cand_box (B,capacity,P) f32, cand_score (B,capacity) f32, cand_cls (B,capacity) i32,
 *   cand_idx (B,capacity) i32, cand_count (B) i32; the count may exceed capacity (the overflow is
 *   dropped, consumers clamp).
 *   state_clean 0: cand_count is zeroed by this call (a memset in front of the kernel); non-zero: the
 *               caller guarantees it is zero already -- true after a one-time memset and then after every
 *               mydet_postprocess(..., flags = MYDET_PP_CONSUME) that consumed the candidates. */
int mydet_decode_compact(int kind, const mydet_level_t* levels, int n_levels, int batch, int n_cls,
                         int n_param, float img_h, float img_w, float conf_thres, float* cand_box,
                         float* cand_score, int32_t* cand_cls, int32_t* cand_idx,
                         int32_t* cand_count, int32_t capacity, int state_clean, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Threshold -> top-k -> class-aware axis-aligned NMS, batched.  Replaces
 * ImageObjects.post_process / nms / non_max_suppression (utils/structures.py:92-173), including
 * torchvision.ops.nms's arithmetic (float32 IoU = inter/(a_i+a_j-inter), suppressed iff
 * (double)iou > nms_thres, visiting order = stable descending score).
 *   boxes  (B, pitch, P) f32; scores (B, pitch) f32; cls (B, pitch) int32 or int64 (cls_is_i64)
 *   src_idx  optional (B,pitch) i32: value reported in out_idx instead of the slot number
 *   counts   optional (B) i32: valid candidates per image (clamped to n_per_image); else all
 *   conf_thres: candidates with score < conf_thres (or NaN) are ignored; pass -INFINITY for none
 *   nms_thres  torchvision's rule: a box is dropped iff its IoU with a kept, higher-scored box of its class is > nms_thres.
 *            A negative threshold (which disjoint boxes pass: 0 > thr) is honoured; it takes the all-pairs kernels.
 *   topk     > 0: keep the topk best survivors before NMS (reference: 512); <= 0: no cap
 *   outputs  out_box (B,out_cap,P), out_score (B,out_cap), out_cls (B,out_cap) i64,
 *            out_idx (B,out_cap) i32, out_count (B) i32; rows ordered class ascending, then score
 *            descending, exactly as the reference concatenates its per-class groups.
 *   status   optional (B) i32 written with a bit mask: 1 = class id out of range, 2 = output
 *            truncated to out_cap, 4 = input count exceeded n_per_image, 8 = a src_idx
 *            value does not fit 20 bits (tie-break between equal scores then uses its low bits),
 *            16 = exchange protocol: a consumer's acknowledgement did not arrive within the bounded wait.
 *   flags    MYDET_PP_CONSUME: counts[b] is zeroed once read, which leaves the candidate state ready for
 *            the next mydet_decode_compact(..., state_clean = 1): a decode + post-process step is then two
 *            kernel launches and nothing else.  Single-kernel path only (effective top-k <= MYDET_SMALL_K).
 *            MYDET_PP_FORCE_SCAN: skip the histogram front ends of the top-k select (the kernel then scans
 *            every score with radix passes, as it does on its own when a front end cannot decide: heavy
 *            ties, a misleading sample); results are identical either way -- the flag exists so that tests
 *            can prove that.
 *            MYDET_PP_WS_CLEAN: the workspace is persistent and clean (conventions at the top); large-N path only.
 * Workspace: mydet_postprocess_workspace_bytes(...) bytes, 256-byte aligned. */
size_t mydet_postprocess_workspace_bytes(int batch, int n_per_image, int topk);
int mydet_postprocess(const float* boxes, const float* scores, const void* cls, int cls_is_i64,
                      const int32_t* src_idx, int32_t* counts, int batch, int64_t pitch,
                      int n_per_image, int n_param, int box_format, float conf_thres, int topk,
                      double nms_thres, float* out_box, float* out_score, int64_t* out_cls,
                      int32_t* out_idx, int32_t* out_count, int32_t* status, int out_cap,
                      void* workspace, size_t workspace_bytes, int flags, void* stream);

/* mydet_postprocess fused with the path's only multi-GPU exchange: besides the local
 * outputs, every surviving detection is stored -- packed as P box floats, score, (float)class -- into the
 * gathered buffer of EACH of the n_peers GPUs of the node (peer memory mapped into this process, e.g. CUDA
 * IPC / symmetric memory; the own buffer is one of them), and the per-image count into the int32 tail:
 *     peer_bufs[q]: float rows[images_total][out_cap][P+2];  int32 counts[images_total];
 * This rank's images occupy rows [image_offset, image_offset + batch).  Rows beyond an image's count
 * are unspecified (at most the three floats that complete the last 16-byte store are zeroed).  Peer stores are complete when the kernel has completed on `stream`; consumers
 * on other ranks order themselves with a stream/host barrier.  peer_bufs is a HOST array.
 * Only the single-kernel path (effective top-k <= MYDET_SMALL_K) supports it. */
int mydet_postprocess_scatter(const float* boxes, const float* scores, const void* cls, int cls_is_i64,
                              const int32_t* src_idx, int32_t* counts, int batch, int64_t pitch,
                              int n_per_image, int n_param, int box_format, float conf_thres, int topk,
                              double nms_thres, float* out_box, float* out_score, int64_t* out_cls,
                              int32_t* out_idx, int32_t* out_count, int32_t* status, int out_cap,
                              void* workspace, size_t workspace_bytes, void* const* peer_bufs, int n_peers,
                              int64_t image_offset, int64_t images_total, int flags, void* stream);

/* The fused exchange with (a) NVLS multicast and (b) a device-side protocol, so that it can feed a pipeline and
 * not only a benchmark.  Buffer of every rank (mydet_exchange_buffer_bytes bytes, 16-byte aligned, zeroed once by its
 * owner before the first use, followed by one barrier):
 *     float  rows[images_total][out_cap][P+2];  int32 counts[images_total];      as for mydet_postprocess_scatter
 *     uint32 seq[images_total]   how often image i has been PUBLISHED (release semantics at system scope)
 *     uint32 ack[8]              ack[q]: how many publications consumer rank q has CONSUMED; rank q writes it into every
 *                                rank's copy (mydet_exchange_release)
 *     uint32 want[8], prod[images_total]   local bookkeeping of this rank's consumer / producer
 *   multicast_buf  the multicast mapping of the same symmetric buffer (cuMulticast* / torch symmetric memory
 *                  `multicast_ptr`), or NULL.  With it every row vector, count and flag is ONE multimem.st that the
 *                  NVSwitch replicates into all copies (the own one included); without it the kernel stores to each
 *                  of the n_peers unicast mappings in turn.
 *   self_index     index of this rank's own buffer in peer_bufs (its flags are polled locally)
 *   protocol != 0  producer: before it overwrites image slot i for the k-th time it waits (polling LOCAL memory) until
 *                  ack[q] >= k-1 for every rank q -- back-pressure, so a fast producer never overwrites rows a slow
 *                  consumer is still reading.  The wait is bounded (~1 s); if it expires the image is stored anyway and
 *                  status bit 16 is raised.  The kernel itself does not publish: mydet_exchange_publish, next on the same
 *                  stream, sets seq[i] = k for this rank's images (one fence and a few flag stores after the kernel
 *                  boundary, instead of a system-scope drain at the end of each of the kernel's CTAs).  EVERY rank must
 *                  then consume every publication: mydet_exchange_wait (device-side acquire wait on the local copy until
 *                  all images_total images carry publication number consumed+1; optional snapshot of the counts; status
 *                  word: 1 = timed out), the consumer's own kernels reading rows on the same stream, then
 *                  mydet_exchange_release (ack into every copy).  mydet_exchange_consume_counts is publish + wait +
 *                  snapshot + release in one launch.  Nothing here involves the host or a stream dependency between
 *                  processes; all state lives in the buffer, so the calls can sit in a replayed CUDA graph.
 *   protocol == 0  rows / counts only, as mydet_postprocess_scatter. */
size_t mydet_exchange_buffer_bytes(int64_t images_total, int out_cap, int n_param);
int mydet_postprocess_exchange(const float* boxes, const float* scores, const void* cls, int cls_is_i64,
                               const int32_t* src_idx, int32_t* counts, int batch, int64_t pitch,
                               int n_per_image, int n_param, int box_format, float conf_thres, int topk,
                               double nms_thres, float* out_box, float* out_score, int64_t* out_cls,
                               int32_t* out_idx, int32_t* out_count, int32_t* status, int out_cap,
                               void* workspace, size_t workspace_bytes, void* const* peer_bufs, int n_peers,
                               void* multicast_buf, int self_index, int64_t image_offset, int64_t images_total,
                               int protocol, int flags, void* stream);
int mydet_exchange_publish(void* const* peer_bufs, int n_peers, void* multicast_buf, int self_index, int64_t image_offset,
                           int batch, int64_t images_total, int out_cap, int n_param, void* stream);
int mydet_exchange_wait(void* local_buf, int64_t images_total, int out_cap, int n_param, int32_t* counts_out,
                        int32_t* status, void* stream);
int mydet_exchange_release(void* const* peer_bufs, int n_peers, void* multicast_buf, int self_index,
                           int64_t images_total, int out_cap, int n_param, void* stream);
/* mydet_exchange_publish (this rank's images [image_offset, image_offset + batch); batch 0 = nothing to publish) +
 * mydet_exchange_wait + mydet_exchange_release in ONE launch, for a consumer that needs only the per-image counts of
 * the step (snapshotted into counts_out before the acknowledgement goes out). */
int mydet_exchange_consume_counts(void* const* peer_bufs, int n_peers, void* multicast_buf, int self_index,
                                  int64_t image_offset, int batch, int64_t images_total, int out_cap, int n_param,
                                  int32_t* counts_out, int32_t* status, void* stream);

/* Whole path in one call: decode_compact + postprocess (what api/detection.py:168-172 does per
 * image, here for the batch).  Workspace: mydet_detect_workspace_bytes(...). */
size_t mydet_detect_workspace_bytes(int batch, int64_t n_total, int n_param, int topk);
int mydet_detect(int kind, const mydet_level_t* levels, int n_levels, int batch, int n_cls,
                 int n_param, float img_h, float img_w, float conf_thres, int topk,
                 double nms_thres, float* out_box, float* out_score, int64_t* out_cls,
                 int32_t* out_idx, int32_t* out_count, int32_t* status, int out_cap,
                 void* workspace, size_t workspace_bytes, void* stream);

/* mydet_detect on a persistent workspace (see the conventions at the top): workspace_clean != 0 = the buffer was zeroed
 * once and has only been used by this entry point with this geometry since.  On the single-kernel path (<= 1024
 * survivors after top-k) the call is then two launches and no memset: the post-process zeroes the candidate counts it
 * consumed.  In both entry points the post-process kernel is launched as a programmatic dependent of the decode kernel
 * (it starts while the decode drains and waits before its first read); the environment variable MYDET_PDL=0 turns
 * that off. */
int mydet_detect_ws(int kind, const mydet_level_t* levels, int n_levels, int batch, int n_cls,
                    int n_param, float img_h, float img_w, float conf_thres, int topk,
                    double nms_thres, float* out_box, float* out_score, int64_t* out_cls,
                    int32_t* out_idx, int32_t* out_count, int32_t* status, int out_cap,
                    void* workspace, size_t workspace_bytes, int workspace_clean, void* stream);

/* Pack the detections of a batch into ONE float32 buffer for the multi-GPU exchange:
 * packed[(b*out_cap + k)*(P+2) + 0..P-1] = box, [+P] = score, [+P+1] = (float)class, followed by
 * batch floats holding the bit patterns of the int32 counts.  Size: batch*(out_cap*(P+2) + 1) floats. */
int mydet_pack_detections(const float* out_box, const float* out_score, const int64_t* out_cls,
                          const int32_t* out_count, int batch, int out_cap, int n_param, float* packed,
                          void* stream);

/* ---------------------------------------------------------------------------------------------
 * Rotated greedy NMS, batched, single class per image.  Replaces nms_rotbb
 * (utils/bbox_ops.py:250-306) with exact polygon clipping instead of the pycocotools raster
 * (parity unpinned for the IoU values).  Box i is dropped iff its IoU with an
 * already kept, higher-scored box is >= thr (ge_mode != 0, the reference) or > thr.
 *   boxes (B,pitch,5) (cx,cy,w,h,degrees clockwise), scores (B,pitch), counts optional (B)
 *   keep  (B,pitch) i64: kept indices into the image's boxes, descending score; keep_count (B)
 *   votes optional (B,pitch) i32: per kept box, 1 + number of dropped boxes whose best IoU was
 *         with it (the reference's majority-vote bookkeeping, :291-306) */
size_t mydet_nms_rot_workspace_bytes(int batch, int n_per_image);
int mydet_nms_rot(const float* boxes, const float* scores, const int32_t* counts, int batch,
                  int64_t pitch, int n_per_image, double thr, int ge_mode, int64_t* keep,
                  int32_t* keep_count, int32_t* votes, void* workspace, size_t workspace_bytes,
                  void* stream);

/* mydet_nms_rot on a persistent workspace (see the conventions at the top). */
int mydet_nms_rot_ws(const float* boxes, const float* scores, const int32_t* counts, int batch,
                     int64_t pitch, int n_per_image, double thr, int ge_mode, int64_t* keep,
                     int32_t* keep_count, int32_t* votes, void* workspace, size_t workspace_bytes,
                     int workspace_clean, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Pairwise IoU matrices. */
/* bboxes_iou (utils/bbox_ops.py:6-49): a (N,4), b (K,4) -> out (N,K) f32, bit-exact arithmetic. */
int mydet_iou_aabb_pairwise(const float* a, int64_t n, const float* b, int64_t k, int xyxy,
                            float* out, void* stream);
/* Row-wise max / arg-max of that matrix, batched, without materialising it: `bboxes_iou(a, gt).max(dim=1)` of the
 * training branches (models/detlayers/yolov3.py:94-95 and :106-107, fcos2.py:104-106, retinanet.py:106-107).
 *   a        row boxes: image b, row i at a + b*a_batch_stride + i*a_pitch (elements; a_pitch >= 4 lets (n,5) rows
 *            pass, a_batch_stride = 0 shares one set of boxes -- e.g. anchors -- between all images)
 *   gt       (B, max_gt, 4), gt_count (B) i32 or NULL (= max_gt everywhere)
 *   out_max  (B, n) f32, out_arg (B, n) i64 or NULL: first index of the maximum (torch.max); NaN propagates;
 *            images without GT get -1 / -1. */
int mydet_iou_aabb_rowmax(const float* a, int64_t a_batch_stride, int64_t a_pitch, int64_t n, const float* gt,
                          const int32_t* gt_count, int max_gt, int batch, int xyxy, float* out_max,
                          int64_t* out_arg, void* stream);
/* iou_rle (utils/bbox_ops.py:52-100): a (N,5), b (K,5) degrees -> out (N,K) f64, exact clipping. */
int mydet_iou_rot_pairwise(const float* a, int64_t n, const float* b, int64_t k, double* out,
                           void* stream);

/* iou_rle the way the reference computes it (utils/bbox_ops.py:84-96): both polygons rasterised on a canvas_h x
 * canvas_w canvas with pycocotools' polygon rule (maskApi.c rleFrPoly: 5x sub-sampled integer boundary walk, a pixel column
 * changes value where the walk crosses its sub-column 2|3 line) and IoU = common pixels / united pixels (rleIou).  The
 * library is not available here; the algorithm is restated in oracle/raster.c, pinned to run-length encodings derived by
 * hand, and this entry point reproduces that restatement bit for bit.  a (N,5), b (K,5) degrees -> out (N,K) f64.
 * Workspace: mydet_iou_raster_workspace_bytes(n, k, canvas_w) = one (lo, hi) run per box and pixel column. */
size_t mydet_iou_raster_workspace_bytes(int64_t n, int64_t k, int canvas_w);
int mydet_iou_raster_pairwise(const float* a, int64_t n, const float* b, int64_t k, int canvas_h, int canvas_w,
                              double* out, void* workspace, size_t workspace_bytes, void* stream);

/* The matching IoU of a whole CEPDOF evaluation in one launch: CEPDOFeval.computeIoU (utils/evaluation/cepdof.py:67-99)
 * is called once per (image, category) and each call builds a small dt x gt matrix with the evaluator's own iou_rle
 * (:210-243).  segments: DEVICE array of n_segments x 5 int64 {a0, na, b0, nb, out0}: rows a[a0..a0+na) against columns
 * b[b0..b0+nb), written row-major at out + out0; out0 must be the running sum of na*nb (non-decreasing), total its end.
 * boxes_are_f64 != 0: a / b are float64 (N,5) and the corners are built in float64, as the evaluator's numpy code does
 * with the JSON's Python floats (:181-207, :232-236); 0: float32 boxes and corners, as iou_rle of utils/bbox_ops.py. */
int mydet_iou_rot_segments(const void* a, const void* b, int boxes_are_f64, const int64_t* segments, int n_segments,
                           int64_t total, double* out, void* stream);

/* Format helpers of utils/bbox_ops.py, kept because callers outside the path use the names.
 * cxcywh_to_x1y1x2y2 (:309-316): rows of n_param >= 4 floats, columns 4.. are copied through.
 * xywha2vertex (:137-172): rows (cx,cy,w,h,RADIANS) with n_param >= 5 -> out (N,4,2) tl,tr,br,bl. */
int mydet_cxcywh_to_x1y1x2y2(const float* in, int64_t n, int n_param, float* out, void* stream);
int mydet_xywha2vertex(const float* in, int64_t n, int n_param, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ATSS assignment of one pyramid level.  Replaces the target construction of
 * FCOS_ATSS_Layer.forward (models/detlayers/fcos2.py:253-341) and _get_atss_threshold (:385-405).
 *   t_ltrb      this level's raw regression logits, logical (B,nH,nW,4), strides b,h,w,p
 *   gt_box      (B, max_gt, 4) cxcywh f32, gt_cls (B, max_gt) i64, gt_count (B) i32
 *   strides / anchor_sides  HOST arrays of n_levels entries (cfg 'model.fpn.out_strides',
 *                           'model.atss.anchors'); level = index of this level
 *   outputs     positive, ignored (B,nH,nW) u8; target_ltrb (B,nH,nW,4), target_conf (B,nH,nW),
 *               target_cls (B,nH,nW,C) f32 -- all fully written by the call
 *   thr_out     optional (B, max_gt) f32: the adaptive threshold of every GT (mean + std), in the caller's
 *               GT order.  The threshold of a GT does not depend on the level, but the reference
 *               recomputes it in every level's forward (fcos2.py:262-264 admits the waste): with
 *               thr_is_input != 0 the thresholds are READ from thr_out (as written by an earlier level's
 *               call for the same GT) and the nearest-anchor search is skipped. */
size_t mydet_atss_workspace_bytes(int batch, int max_gt);
int mydet_atss_assign(const float* t_ltrb, const int64_t t_stride[4], int batch, int level,
                      int n_levels, const int32_t* strides, const float* anchor_sides, int img_h,
                      int img_w, const float* gt_box, const int64_t* gt_cls,
                      const int32_t* gt_count, int max_gt, int topk, float ignore_thres, int n_cls,
                      uint8_t* positive, uint8_t* ignored, float* target_ltrb, float* target_conf,
                      float* target_cls, float* thr_out, int thr_is_input, void* workspace,
                      size_t workspace_bytes, void* stream);

/* mydet_atss_assign for ALL pyramid levels in one call: the GT ordering and the adaptive thresholds are computed once and
 * one grid covers the cells of every level (3 launches instead of 15 for five levels, and the few CTAs of the coarse
 * levels fill the tail of the finest one).  levels: HOST array of n_levels descriptors, stride-ascending like strides /
 * anchor_sides; every output is fully written, exactly as by n_levels calls of mydet_atss_assign.  thr_out optional. */
typedef struct {
    const float* t_ltrb;          /* this level's raw regression logits, logical (B,nH,nW,4)            */
    int64_t t_stride[4];          /* element strides of b, h, w, p                                      */
    uint8_t* positive;            /* (B,nH,nW)                                                          */
    uint8_t* ignored;             /* (B,nH,nW)                                                          */
    float* target_ltrb;           /* (B,nH,nW,4)                                                        */
    float* target_conf;           /* (B,nH,nW)                                                          */
    float* target_cls;            /* (B,nH,nW,C)                                                        */
} mydet_atss_level_t;
int mydet_atss_assign_levels(const mydet_atss_level_t* levels, int n_levels, const int32_t* strides,
                             const float* anchor_sides, int batch, int img_h, int img_w, const float* gt_box,
                             const int64_t* gt_cls, const int32_t* gt_count, int max_gt, int topk,
                             float ignore_thres, int n_cls, float* thr_out, void* workspace,
                             size_t workspace_bytes, void* stream);

/* Target assignment of FCOSLayer.forward (models/detlayers/fcos2.py:84-143), one pyramid level: the same
 * outputs and GT conventions as mydet_atss_assign, with FCOSLayer's rule for a positive cell -- the cell centre
 * lies strictly inside the GT shrunk by center_region (0.5 in the reference) and
 * anch_min < max(l,t,r,b) < anch_max (cfg 'model.fcos.anchors'[level], [level+1]).
 * Workspace: mydet_atss_workspace_bytes(batch, max_gt). */
int mydet_fcos_assign(const float* t_ltrb, const int64_t t_stride[4], int batch, int stride, int img_h,
                      int img_w, const float* gt_box, const int64_t* gt_cls, const int32_t* gt_count,
                      int max_gt, float center_region, float anch_min, float anch_max, float ignore_thres,
                      int n_cls, uint8_t* positive, uint8_t* ignored, float* target_ltrb,
                      float* target_conf, float* target_cls, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tracklets of rotated boxes, batched.  Replaces the per-object numpy Kalman filters of
 * utils/kalman_filter.py:77-142 (RotBBoxKalmanFilter) as driven by utils/structures.py:447-529 (KFTracklet.__init__,
 * predict, update, likelihood): all tracklets of a frame advance in one launch, float64 like the reference.
 *   x (N,10), P (N,10,10) f64: state (cx, cy, w, h, angle in [0,180), velocities) and covariance; score (N) f64 and
 *   pred_count (N) i32 (predictions since the last update) may be NULL.
 *   noise       HOST array of 26 doubles: 10 initial, 10 process, 5 measurement standard deviations, score momentum
 *               (structures.py:458-461, :471)
 *   initiate    boxes (N,5) -> x, P (the four size-like components and their velocities scaled by w*h), pred_count = 0
 *   predict     out_boxes (N,5) or NULL: the predicted boxes (angle NOT wrapped, as KFTracklet.predict returns it)
 *   update      meas (N,5), meas_score (N) or NULL, has (N) u8 or NULL (= all): tracklets without a measurement are left
 *               untouched and get a zero row in out_boxes
 *   likelihood  cand (M,5) -> out (N,M): the Gaussian density of every candidate under every tracklet (:519-528)
 * The association IoU between predicted tracklet boxes and detections is mydet_iou_rot_pairwise. */
int mydet_kf_initiate(const double* boxes, int n, const double* noise, double* x, double* P, int32_t* pred_count,
                      void* stream);
int mydet_kf_predict(double* x, double* P, double* score, int32_t* pred_count, int n, const double* noise,
                     double* out_boxes, void* stream);
int mydet_kf_update(double* x, double* P, double* score, int32_t* pred_count, const double* meas,
                    const double* meas_score, const uint8_t* has, int n, const double* noise, double* out_boxes,
                    void* stream);
int mydet_kf_likelihood(const double* x, const double* P, int n, const double* cand, int m, double* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Image pre-processing in front of the model.  Replaces, for a batch of equally
 * sized uint8 RGB frames, Detector._preprocess_pil + tvf.to_tensor + format_tensor_img
 * (api/detection.py:158-162, :177-205; utils/image_ops.py:22-52 resize_pil / pad_to_divisible, :55-106
 * rect_to_square(aug=False), :165-188 format_tensor_img): Pillow's anti-aliased BILINEAR Image.resize (two-pass
 * 8 bits-per-channel resampler, bit-exact, including Pillow's rule of resizing images more than 100 times taller
 * than wide vertically first), zero padding of the uint8 image, /255, normalisation / channel order.
 *   src         (batch, in_h, in_w, 3) uint8, row pitch and image stride in BYTES
 *   resized_*   size after the resize (the caller computes it with the reference's own Python expressions, see
 *               mydetection_b200/image_ops.py: plan); == in_* when the image is only padded
 *   left, top   position of the resized image inside the (out_h, out_w) output; everything else is the
 *               normalised value of a zero pixel
 *   format      MYDET_FORMAT_*: cfg 'model.input_format'
 *   dst         (batch, 3, out_h, out_w) float32, fully written
 * Workspace: mydet_preprocess_workspace_bytes(...), 256-byte aligned (filter banks + the uint8 intermediate of
 * the horizontal pass); not needed when resized_* == in_*. */
typedef enum {
    MYDET_FORMAT_RGB_1 = 0,        /* 'RGB_1'                                                  */
    MYDET_FORMAT_RGB_1_NORM = 1,   /* 'RGB_1_norm'   (x - mean) / std, ImageNet constants      */
    MYDET_FORMAT_BGR_255_NORM = 2  /* 'BGR_255_norm' channels reversed, x * 255 - mean         */
} mydet_input_format_t;
size_t mydet_preprocess_workspace_bytes(int batch, int in_h, int in_w, int resized_h, int resized_w);
int mydet_preprocess(const uint8_t* src, int batch, int64_t src_image_stride, int64_t src_row_pitch, int in_h,
                     int in_w, int resized_h, int resized_w, int left, int top, int out_h, int out_w, int format,
                     float* dst, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MYDET_H_ */
