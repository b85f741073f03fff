// extern "C" entry points that tie the kernels together (see include/mydet.h).
#include <stdarg.h>
#include <string.h>

#include "internal.cuh"

namespace mydet {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int effective_k(int n_per_image, int topk) { return (topk > 0 && topk < n_per_image) ? topk : n_per_image; }
static int next_pow2_min32(int v) { int p = 32; while (p < v) p <<= 1; return p; }

}  // namespace mydet

using namespace mydet;

MYDET_API int mydet_version(void) { return MYDET_VERSION; }
MYDET_API const char* mydet_last_error(void) { return g_err; }

MYDET_API size_t mydet_postprocess_workspace_bytes(int batch, int n_per_image, int topk) {
    if (batch <= 0 || n_per_image <= 0) return 256;
    if (effective_k(n_per_image, topk) <= MYDET_SMALL_K) return 256;
    return large_workspace_bytes(batch, n_per_image, false);
}

static int postprocess_impl(const float* boxes, const float* scores, const void* cls, int cls_is_i64,
                            const int32_t* src_idx, int32_t* counts, int batch, int64_t pitch,
                            int n_per_image, int n_param, int box_format, float conf_thres, int topk,
                            double nms_thres, float* out_box, float* out_score, int64_t* out_cls,
                            int32_t* out_idx, int32_t* out_count, int32_t* status, int out_cap,
                            void* workspace, size_t workspace_bytes, void* const* peers, int n_peers,
                            int64_t peer_row0, int64_t peer_rows_total, int pp_flags, void* stream,
                            void* peer_mc = nullptr, int peer_self = 0, int peer_protocol = 0) {
    cudaStream_t st = (cudaStream_t)stream;
    MYDET_REQUIRE(n_peers >= 0 && n_peers <= 8, "n_peers must be in [0,8]");
    MYDET_REQUIRE(n_peers == 0 || (peers && peer_row0 >= 0 && peer_row0 + batch <= peer_rows_total),
                  "bad peer buffer description");
    MYDET_REQUIRE(!peer_protocol || (n_peers > 0 && peer_self >= 0 && peer_self < n_peers),
                  "the exchange protocol needs the peer buffers and this rank's index among them");
    MYDET_REQUIRE(!peer_mc || n_peers > 0, "a multicast buffer without peer buffers");
    MYDET_REQUIRE((reinterpret_cast<uintptr_t>(peer_mc) & 15) == 0, "the multicast buffer must be 16-byte aligned");
    MYDET_REQUIRE(batch >= 0 && n_per_image >= 0 && pitch >= n_per_image, "bad batch / n_per_image / pitch");
    MYDET_REQUIRE(n_param == 4 || n_param == 5, "n_param must be 4 or 5");
    MYDET_REQUIRE(box_format == MYDET_BOX_CXCYWH || box_format == MYDET_BOX_X1Y1X2Y2, "unknown box format");
    MYDET_REQUIRE(n_per_image <= MYDET_MAX_CANDIDATES, "more than %d candidates per image", MYDET_MAX_CANDIDATES);
    MYDET_REQUIRE(batch == 0 || (out_count && out_cap > 0), "NULL out_count or out_cap <= 0");
    const int consume = pp_flags & MYDET_PP_CONSUME;
    MYDET_REQUIRE(!consume || counts, "consume needs counts");
    if (batch == 0) return 0;
    const bool single = effective_k(n_per_image, topk) <= MYDET_SMALL_K;
    MYDET_REQUIRE(!consume || (single && n_per_image > 0), "consume is implemented for the single-kernel path only");
    if (n_per_image == 0) {
        MYDET_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int32_t) * (size_t)batch, st));
        if (status) MYDET_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)batch, st));
        return 0;
    }
    MYDET_REQUIRE(boxes && scores && cls && out_box && out_score && out_cls && out_idx, "NULL tensor pointer");
    MYDET_REQUIRE(n_param != 4 || ((reinterpret_cast<uintptr_t>(boxes) | reinterpret_cast<uintptr_t>(out_box)) & 15) == 0,
                  "boxes and out_box must be 16-byte aligned (4-parameter boxes move as 128-bit vectors)");
    const int K = effective_k(n_per_image, topk);
    if (K <= MYDET_SMALL_K) {
        PPParams P;
        P.boxes = boxes; P.scores = scores; P.cls = cls; P.src_idx = src_idx; P.counts = counts;
        P.pitch = pitch; P.n_per_image = n_per_image; P.n_param = n_param; P.box_format = box_format;
        P.cls_is_i64 = cls_is_i64; P.conf_thres = conf_thres; P.topk = K;
        P.nms_thr_f = float_at_or_below(nms_thres); P.kpad = next_pow2_min32(K);
        P.out_box = out_box; P.out_score = out_score; P.out_cls = reinterpret_cast<long long*>(out_cls);
        P.out_idx = out_idx; P.out_count = out_count; P.status = status; P.out_cap = out_cap;
        P.n_peers = n_peers; P.peer_row0 = peer_row0; P.peer_rows_total = peer_rows_total;
        for (int q = 0; q < 8; ++q) P.peer[q] = q < n_peers ? static_cast<float*>(peers[q]) : nullptr;
        P.peer_vec = (n_peers > 0 && ((long long)out_cap * (n_param + 2)) % 4 == 0) ? 1 : 0;
        for (int q = 0; q < n_peers; ++q)
            if (reinterpret_cast<uintptr_t>(peers[q]) & 15) P.peer_vec = 0;
        P.peer_mc = static_cast<float*>(peer_mc); P.peer_self = peer_self; P.peer_protocol = peer_protocol ? 1 : 0;
        P.consume = consume; P.force_scan = (pp_flags & MYDET_PP_FORCE_SCAN) ? 1 : 0;
        return launch_postprocess_small(P, batch, st, (pp_flags & kPPFlagPdl) != 0);
    }
    MYDET_REQUIRE(n_peers == 0, "the fused exchange is implemented for the single-kernel path only (<= %d survivors)", MYDET_SMALL_K);
    LargeArgs A{boxes, scores, cls, cls_is_i64, src_idx, counts, batch, pitch, n_per_image, n_param, box_format,
                conf_thres, topk, nms_thres, 0, false, out_box, out_score, reinterpret_cast<long long*>(out_cls),
                out_idx, out_count, status, out_cap, nullptr, nullptr};
    A.ws_clean = (pp_flags & MYDET_PP_WS_CLEAN) != 0;
    return run_large(A, workspace, workspace_bytes, st);
}

MYDET_API int mydet_postprocess(const float* boxes, const float* scores, const void* cls, int cls_is_i64,
                                const int32_t* src_idx, int32_t* counts, int batch, int64_t pitch,
                                int n_per_image, int n_param, int box_format, float conf_thres, int topk,
                                double nms_thres, float* out_box, float* out_score, int64_t* out_cls,
                                int32_t* out_idx, int32_t* out_count, int32_t* status, int out_cap,
                                void* workspace, size_t workspace_bytes, int flags, void* stream) {
    return postprocess_impl(boxes, scores, cls, cls_is_i64, src_idx, counts, batch, pitch, n_per_image, n_param,
                            box_format, conf_thres, topk, nms_thres, out_box, out_score, out_cls, out_idx, out_count,
                            status, out_cap, workspace, workspace_bytes, nullptr, 0, 0, 0, flags, stream);
}

MYDET_API int mydet_postprocess_scatter(const float* boxes, const float* scores, const void* cls, int cls_is_i64,
                                        const int32_t* src_idx, int32_t* counts, int batch, int64_t pitch,
                                        int n_per_image, int n_param, int box_format, float conf_thres, int topk,
                                        double nms_thres, float* out_box, float* out_score, int64_t* out_cls,
                                        int32_t* out_idx, int32_t* out_count, int32_t* status, int out_cap,
                                        void* workspace, size_t workspace_bytes, void* const* peer_bufs, int n_peers,
                                        int64_t image_offset, int64_t images_total, int flags, void* stream) {
    return postprocess_impl(boxes, scores, cls, cls_is_i64, src_idx, counts, batch, pitch, n_per_image, n_param,
                            box_format, conf_thres, topk, nms_thres, out_box, out_score, out_cls, out_idx, out_count,
                            status, out_cap, workspace, workspace_bytes, peer_bufs, n_peers, image_offset, images_total,
                            flags, stream);
}

MYDET_API int mydet_postprocess_exchange(const float* boxes, const float* scores, const void* cls, int cls_is_i64,
                                         const int32_t* src_idx, int32_t* counts, int batch, int64_t pitch,
                                         int n_per_image, int n_param, int box_format, float conf_thres, int topk,
                                         double nms_thres, float* out_box, float* out_score, int64_t* out_cls,
                                         int32_t* out_idx, int32_t* out_count, int32_t* status, int out_cap,
                                         void* workspace, size_t workspace_bytes, void* const* peer_bufs, int n_peers,
                                         void* multicast_buf, int self_index, int64_t image_offset, int64_t images_total,
                                         int protocol, int flags, void* stream) {
    return postprocess_impl(boxes, scores, cls, cls_is_i64, src_idx, counts, batch, pitch, n_per_image, n_param,
                            box_format, conf_thres, topk, nms_thres, out_box, out_score, out_cls, out_idx, out_count,
                            status, out_cap, workspace, workspace_bytes, peer_bufs, n_peers, image_offset, images_total,
                            flags, stream, multicast_buf, self_index, protocol);
}

// ---- whole path: candidate buffers live in the workspace
namespace {
struct DetectWs { float* box; float* score; int32_t* cls; int32_t* idx; int32_t* count; void* rest; size_t rest_bytes; };
size_t carve_detect(DetectWs& w, void* base, size_t bytes, int batch, int64_t cap, int n_param, int topk) {
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
    const size_t bn = (size_t)batch * (size_t)cap;
    const size_t o_box = take(bn * n_param * 4), o_score = take(bn * 4), o_cls = take(bn * 4), o_idx = take(bn * 4);
    const size_t o_cnt = take((size_t)batch * 4);
    const size_t pp = mydet_postprocess_workspace_bytes(batch, (int)cap, topk);
    const size_t o_rest = take(pp);
    if (base) {
        char* p = static_cast<char*>(base);
        w.box = (float*)(p + o_box); w.score = (float*)(p + o_score); w.cls = (int32_t*)(p + o_cls);
        w.idx = (int32_t*)(p + o_idx); w.count = (int32_t*)(p + o_cnt); w.rest = p + o_rest; w.rest_bytes = pp;
    }
    return off;
}
}  // namespace

MYDET_API size_t mydet_detect_workspace_bytes(int batch, int64_t n_total, int n_param, int topk) {
    DetectWs w;
    return carve_detect(w, nullptr, 0, batch, n_total, n_param, topk);
}

static int detect_impl(int kind, const mydet_level_t* levels, int n_levels, int batch, int n_cls, int n_param,
                       float img_h, float img_w, float conf_thres, int topk, double nms_thres, float* out_box,
                       float* out_score, int64_t* out_cls, int32_t* out_idx, int32_t* out_count, int32_t* status,
                       int out_cap, void* workspace, size_t workspace_bytes, int pp_flags, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MYDET_REQUIRE(levels && n_levels >= 1 && n_levels <= MYDET_MAX_LEVELS, "n_levels must be in [1,%d]", MYDET_MAX_LEVELS);
    int64_t n_total = 0;
    for (int i = 0; i < n_levels; ++i) n_total += (int64_t)levels[i].n_anchor * levels[i].n_h * levels[i].n_w;
    DetectWs w;
    const size_t need = carve_detect(w, workspace, workspace_bytes, batch, n_total, n_param, topk);
    if (!workspace || need > workspace_bytes) {
        set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        return MYDET_ERR_WORKSPACE;
    }
    if (batch == 0) return 0;
    // persistent clean workspace + single-kernel post-process: the post-process zeroes the candidate counts it consumed,
    // so the counts are zero on entry of every call and the memset in front of the decode is not needed
    const bool self_clean = (pp_flags & MYDET_PP_WS_CLEAN) && n_total > 0 && effective_k((int)n_total, topk) <= MYDET_SMALL_K;
    int rc = decode_compact_impl(kind, levels, n_levels, batch, n_cls, n_param, img_h, img_w, conf_thres, w.box, w.score,
                                 w.cls, w.idx, w.count, (int32_t)n_total, /*state_clean=*/self_clean ? 1 : 0, nullptr, st);
    if (rc) return rc;
    if (self_clean) pp_flags |= MYDET_PP_CONSUME;
    // the decode already applied conf_thres; the post-process sees only survivors.  It is launched as a programmatic
    // dependent of the decode kernel (its launch latency and CTA ramp overlap the decode's tail); MYDET_PDL=0 turns that off.
    static const bool pdl = [] { const char* e = getenv("MYDET_PDL"); return !(e && e[0] == '0'); }();
    if (pdl) pp_flags |= kPPFlagPdl;
    return mydet_postprocess(w.box, w.score, w.cls, 0, w.idx, w.count, batch, n_total, (int)n_total, n_param,
                             MYDET_BOX_CXCYWH, -INFINITY, topk, nms_thres, out_box, out_score, out_cls, out_idx,
                             out_count, status, out_cap, w.rest, w.rest_bytes, pp_flags, st);
}

MYDET_API int mydet_detect(int kind, const mydet_level_t* levels, int n_levels, int batch, int n_cls, int n_param,
                           float img_h, float img_w, float conf_thres, int topk, double nms_thres, float* out_box,
                           float* out_score, int64_t* out_cls, int32_t* out_idx, int32_t* out_count, int32_t* status,
                           int out_cap, void* workspace, size_t workspace_bytes, void* stream) {
    return detect_impl(kind, levels, n_levels, batch, n_cls, n_param, img_h, img_w, conf_thres, topk, nms_thres, out_box, out_score,
                       out_cls, out_idx, out_count, status, out_cap, workspace, workspace_bytes, 0, stream);
}

MYDET_API int mydet_detect_ws(int kind, const mydet_level_t* levels, int n_levels, int batch, int n_cls, int n_param,
                              float img_h, float img_w, float conf_thres, int topk, double nms_thres, float* out_box,
                              float* out_score, int64_t* out_cls, int32_t* out_idx, int32_t* out_count, int32_t* status,
                              int out_cap, void* workspace, size_t workspace_bytes, int workspace_clean, void* stream) {
    return detect_impl(kind, levels, n_levels, batch, n_cls, n_param, img_h, img_w, conf_thres, topk, nms_thres, out_box, out_score,
                       out_cls, out_idx, out_count, status, out_cap, workspace, workspace_bytes,
                       workspace_clean ? MYDET_PP_WS_CLEAN : 0, stream);
}

namespace mydet {
__global__ void pack_kernel(const float* __restrict__ box, const float* __restrict__ score, const long long* __restrict__ cls,
                            const int* __restrict__ count, int batch, int cap, int n_param, float* __restrict__ packed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // detection row
    const long long rows = (long long)batch * cap;
    if (i < rows) {
        float* o = packed + i * (n_param + 2);
        const int b = (int)(i / cap), k = (int)(i - (long long)b * cap);
        const bool live = k < count[b];
        for (int p = 0; p < n_param; ++p) o[p] = live ? box[i * n_param + p] : 0.f;
        o[n_param] = live ? score[i] : 0.f;
        o[n_param + 1] = live ? (float)cls[i] : 0.f;
    }
    if (i < batch) packed[rows * (n_param + 2) + i] = __int_as_float(count[i]);
}
}  // namespace mydet

MYDET_API int mydet_pack_detections(const float* out_box, const float* out_score, const int64_t* out_cls,
                                    const int32_t* out_count, int batch, int out_cap, int n_param, float* packed,
                                    void* stream) {
    MYDET_REQUIRE(batch >= 0 && out_cap > 0 && (n_param == 4 || n_param == 5), "bad batch / out_cap / n_param");
    if (batch == 0) return 0;
    MYDET_REQUIRE(out_box && out_score && out_cls && out_count && packed, "NULL tensor pointer");
    const long long rows = (long long)batch * out_cap;
    pack_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        out_box, out_score, reinterpret_cast<const long long*>(out_cls), out_count, batch, out_cap, n_param, packed);
    return launch_status("pack_kernel");
}

MYDET_API size_t mydet_nms_rot_workspace_bytes(int batch, int n_per_image) {
    if (batch <= 0 || n_per_image <= 0) return 256;
    return large_workspace_bytes(batch, n_per_image, true);
}

static int nms_rot_impl(const float* boxes, const float* scores, const int32_t* counts, int batch, int64_t pitch,
                        int n_per_image, double thr, int ge_mode, int64_t* keep, int32_t* keep_count,
                        int32_t* votes, void* workspace, size_t workspace_bytes, int ws_clean, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MYDET_REQUIRE(batch >= 0 && n_per_image >= 0 && pitch >= n_per_image, "bad batch / n_per_image / pitch");
    MYDET_REQUIRE(n_per_image <= MYDET_MAX_CANDIDATES, "more than %d boxes per image", MYDET_MAX_CANDIDATES);
    if (batch == 0) return 0;
    MYDET_REQUIRE(keep_count, "NULL keep_count");
    if (n_per_image == 0) { MYDET_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(int32_t) * (size_t)batch, st)); return 0; }
    MYDET_REQUIRE(boxes && scores && keep, "NULL tensor pointer");
    LargeArgs A{boxes, scores, nullptr, 0, nullptr, counts, batch, pitch, n_per_image, 5, MYDET_BOX_CXCYWH,
                -INFINITY, 0, thr, ge_mode, true, nullptr, nullptr, nullptr, nullptr, keep_count, nullptr, 0,
                reinterpret_cast<long long*>(keep), votes};
    A.ws_clean = ws_clean != 0;
    return run_large(A, workspace, workspace_bytes, st);
}

MYDET_API int mydet_nms_rot(const float* boxes, const float* scores, const int32_t* counts, int batch, int64_t pitch,
                            int n_per_image, double thr, int ge_mode, int64_t* keep, int32_t* keep_count,
                            int32_t* votes, void* workspace, size_t workspace_bytes, void* stream) {
    return nms_rot_impl(boxes, scores, counts, batch, pitch, n_per_image, thr, ge_mode, keep, keep_count, votes, workspace,
                        workspace_bytes, 0, stream);
}

MYDET_API int mydet_nms_rot_ws(const float* boxes, const float* scores, const int32_t* counts, int batch, int64_t pitch,
                               int n_per_image, double thr, int ge_mode, int64_t* keep, int32_t* keep_count,
                               int32_t* votes, void* workspace, size_t workspace_bytes, int workspace_clean, void* stream) {
    return nms_rot_impl(boxes, scores, counts, batch, pitch, n_per_image, thr, ge_mode, keep, keep_count, votes, workspace,
                        workspace_bytes, workspace_clean, stream);
}
