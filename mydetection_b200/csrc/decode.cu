// Per-level head decode for every det-layer kind, all pyramid levels in ONE launch.
//
// Replaces (reference paths): models/detlayers/yolov3.py:41-69, fcos2.py:40-69 / :222-251,
// fcos.py:41-68, rapid.py:48-82, retinanet.py:63-82, uv5.py:60-91 and the level concatenation
// models/general.py:74-76.  The fused variant also applies the score threshold of
// utils/structures.py:98 and compacts the survivors.
//
// Design (HBM-bound streaming pass, DESIGN.md "decode"):
//   * the head tensors are channel-planar NCHW; a warp walks 32 consecutive 16-byte chunks of one
//     (image, anchor) plane, so every channel is one fully coalesced 512-byte request;
//   * each thread owns VEC=4 consecutive cells (128-bit loads) and keeps a running arg-max over
//     the class planes, 8 planes (8 independent LDG.128) in flight per thread;
//   * planes whose size or base is not 16-byte friendly (19x19, 5x5 ...) take the VEC=1 path;
//   * no shared memory, no re-reads: every logit is loaded exactly once.
#include "internal.cuh"

namespace mydet {

// tunables (overridable at build time by scripts/tune_decode.py)
#ifndef MYDET_DECODE_WARPS
#define MYDET_DECODE_WARPS 4
#endif
#ifndef MYDET_CLS_UNROLL
#define MYDET_CLS_UNROLL 8
#endif
#ifndef MYDET_DECODE_MINBLOCKS
#define MYDET_DECODE_MINBLOCKS 8
#endif
constexpr int kDecodeWarps = MYDET_DECODE_WARPS;   // warps per CTA
constexpr int kDecodeThreads = kDecodeWarps * 32;
constexpr int kClsUnroll = MYDET_CLS_UNROLL;       // class planes in flight per thread

struct LevelDev {
    const float* bbox;
    const float* conf;
    const float* cls;
    long long bs_b, bs_a, bs_h, bs_w, bs_p;
    long long cs_b, cs_a, cs_h, cs_w;
    long long ks_b, ks_a, ks_h, ks_w, ks_c;
    int n_a, n_h, n_w, n_hw;
    int vec;          // 4: hw-contiguous, 16-byte aligned planes; 1: generic strides
    int chunks;       // warps per (image, anchor) plane = ceil(ceil(n_hw / vec) / 32)
    int first_block;  // first CTA of this level
    int out_offset;   // candidate offset of this level inside an image
    float stride;
    float aw[MYDET_MAX_ANCHORS], ah[MYDET_MAX_ANCHORS];
};

struct DecodeParams {
    LevelDev lv[MYDET_MAX_LEVELS];
    int n_levels, batch, n_cls, n_param;
    float img_h, img_w, img_max;
    float thr;
    // dense outputs
    float* out_box;
    long long* out_cls;
    float* out_score;
    long long n_total;
    // compact outputs
    float* cand_box;
    float* cand_score;
    int* cand_cls;
    int* cand_idx;
    int* cand_count;
    int capacity;
};

// Running arg-max over class logits, plus the runner-up value.
// The reference takes torch.max over the class PROBABILITIES (first index on ties).  sigmoid is
// monotone, so the arg-max over logits is the same index EXCEPT where float32 sigmoid collapses two
// distinct logits onto one probability.  That needs a top-2 gap below ulp(sigma)/sigma'(x):
// < 2.4e-7 for x <= 0, < 1.3e-6 for x < 3, < 1.8e-4 for x < 8, anything above.  Cells whose top-2
// gap falls inside a (5-8x wider) guard band are re-scanned on the probabilities (rare, divergent).
__device__ __forceinline__ void class_update(float& best, float& second, int& best_c, float v, int c) {
    const bool gt = v > best;
    second = fmaxf(second, fminf(best, v));
    best = fmaxf(best, v);
    best_c = gt ? c : best_c;
}
__device__ __forceinline__ bool class_ambiguous(float best, float second) {
    const float gap = __fsub_rn(best, second);
    if (best > 8.0f) return second > 8.0f || gap < 1e-3f;   // both saturated, or close
    return gap < (best < 3.0f ? 1e-5f : 1e-3f);
}

// Exact restatement of torch.max(sigmoid(logits)) for the cells flagged in `flagged` (one bit per
// lane): the WHOLE warp re-reads that cell's class logits (lane k takes classes k, k+32, ...),
// evaluates the probabilities and reduces to the first index of the maximal probability.
// Must be called by all 32 lanes.  A serial per-thread re-scan here cost a 30 us tail (r1 profile).
__device__ __forceinline__ void class_rescan_warp(unsigned flagged, const float* pc, long long stride_c, int n_cls,
                                                  float& best, int& best_c) {
    const unsigned lane = lane_id();
    while (flagged) {
        const int src = __ffs(flagged) - 1;
        flagged &= flagged - 1;
        const unsigned long long base = __shfl_sync(0xffffffffu, (unsigned long long)pc, src);
        const float* q = reinterpret_cast<const float*>(base);
        float bp = -1.0f, bl = 0.f;
        int bc = 0x7fffffff;
        for (int c = (int)lane; c < n_cls; c += 32) {
            const float v = ld_stream(q + (long long)c * stride_c);
            const float p = sigmoid_f(v);
            if (p > bp) { bp = p; bl = v; bc = c; }      // ascending c: keeps the first maximum
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float op = __shfl_xor_sync(0xffffffffu, bp, o);
            const float ol = __shfl_xor_sync(0xffffffffu, bl, o);
            const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
            if (op > bp || (op == bp && oc < bc)) { bp = op; bl = ol; bc = oc; }
        }
        if ((int)lane == src) { best = bl; best_c = bc; }
    }
}

template <int KIND>
__device__ __forceinline__ float final_score(float conf_logit, float best_logit, int n_cls) {
    if (KIND == MYDET_KIND_RETINA) return sigmoid_f(best_logit);
    float pc = sigmoid_f(conf_logit);
    if (n_cls <= 0) return pc;  // yolov3.py:61-62, rapid.py:75
    float prod = __fmul_rn(pc, sigmoid_f(best_logit));
    if (KIND == MYDET_KIND_FCOS || KIND == MYDET_KIND_RAPID) return __fsqrt_rn(prod);  // fcos2.py:62, rapid.py:72
    return prod;  // yolov3.py:59, uv5.py:86
}

constexpr float kPiF = 3.14159265358979323846f;

// t[0..P) raw regression logits of one cell -> box[0..P)
template <int KIND>
__device__ __forceinline__ void decode_box(const float* t, float* box, int n_param, float col, float row,
                                           float stride, float aw, float ah, float img_h, float img_w,
                                           float img_max) {
    if (KIND == MYDET_KIND_YOLO || KIND == MYDET_KIND_RAPID) {
        box[0] = __fmul_rn(__fadd_rn(sigmoid_f(t[0]), col), stride);
        box[1] = __fmul_rn(__fadd_rn(sigmoid_f(t[1]), row), stride);
        box[2] = __fmul_rn(expf(t[2]), aw);
        box[3] = __fmul_rn(expf(t[3]), ah);
        if (KIND == MYDET_KIND_RAPID) {
            // rapid.py:49,63: ((sigmoid*2)*pi - pi) / pi * 180, every step rounded to float32
            float rad = __fsub_rn(__fmul_rn(__fmul_rn(sigmoid_f(t[4]), 2.0f), kPiF), kPiF);
            box[4] = __fmul_rn(__fdiv_rn(rad, kPiF), 180.0f);
        }
    } else if (KIND == MYDET_KIND_FCOS) {
        float half = __fmul_rn(stride, 0.5f);
        float cx = __fadd_rn(__fmul_rn(col, stride), half);  // fcos2.py:441-442
        float cy = __fadd_rn(__fmul_rn(row, stride), half);
        float l = __fmul_rn(expf(t[0]), stride), tp = __fmul_rn(expf(t[1]), stride);
        float r = __fmul_rn(expf(t[2]), stride), bt = __fmul_rn(expf(t[3]), stride);
        float x1 = fminf(fmaxf(__fsub_rn(cx, l), 0.0f), img_w);   // :51-54 clamp into the image
        float y1 = fminf(fmaxf(__fsub_rn(cy, tp), 0.0f), img_h);
        float x2 = fminf(fmaxf(__fadd_rn(cx, r), 0.0f), img_w);
        float y2 = fminf(fmaxf(__fadd_rn(cy, bt), 0.0f), img_h);
        box[0] = __fmul_rn(__fadd_rn(x1, x2), 0.5f);               // :420-423
        box[1] = __fmul_rn(__fadd_rn(y1, y2), 0.5f);
        box[2] = __fsub_rn(x2, x1);
        box[3] = __fsub_rn(y2, y1);
    } else if (KIND == MYDET_KIND_RETINA) {
        float half = __fmul_rn(stride, 0.5f);
        float acx = __fadd_rn(half, __fmul_rn(col, stride));      // retinanet.py:57-58
        float acy = __fadd_rn(half, __fmul_rn(row, stride));
        float v0 = __fadd_rn(acx, __fmul_rn(t[0], aw));
        float v1 = __fadd_rn(acy, __fmul_rn(t[1], ah));
        float v2 = __fmul_rn(expf(t[2]), aw);
        float v3 = __fmul_rn(expf(t[3]), ah);
        box[0] = fminf(fmaxf(v0, 1.0f), img_max);                  // :70
        box[1] = fminf(fmaxf(v1, 1.0f), img_max);
        box[2] = fminf(fmaxf(v2, 1.0f), img_max);
        box[3] = fminf(fmaxf(v3, 1.0f), img_max);
        if (n_param == 5) box[4] = __fsub_rn(__fmul_rn(sigmoid_f(t[4]), 360.0f), 180.0f);  // :72
    } else {  // MYDET_KIND_UV5, uv5.py:68-73
        float s0 = sigmoid_f(t[0]), s1 = sigmoid_f(t[1]), s2 = sigmoid_f(t[2]), s3 = sigmoid_f(t[3]);
        box[0] = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(s0, 2.0f), 0.5f), col), stride);
        box[1] = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(s1, 2.0f), 0.5f), row), stride);
        float w2 = __fmul_rn(s2, 2.0f), h2 = __fmul_rn(s3, 2.0f);
        box[2] = __fmul_rn(__fmul_rn(w2, w2), aw);
        box[3] = __fmul_rn(__fmul_rn(h2, h2), ah);
    }
}

template <int KIND, bool COMPACT, int VEC>
__device__ __forceinline__ void decode_unit(const DecodeParams& P, const LevelDev& L, int unit) {
    const unsigned lane = lane_id();
    // unit -> (image, anchor, chunk)
    const int chunk = unit % L.chunks;
    const int ba = unit / L.chunks;
    const int a = ba % L.n_a;
    const int b = ba / L.n_a;
    if (b >= P.batch) return;  // whole warp leaves together (unit is warp-uniform)

    const int q = chunk * 32 + (int)lane;     // this thread's group of VEC cells
    const int hw0 = q * VEC;
    const bool active = hw0 < L.n_hw;
    const int n_cls = P.n_cls, n_par = P.n_param;
    const int h0 = hw0 / L.n_w, w0 = hw0 - h0 * L.n_w;   // VEC == 1: the cell; VEC == 4: first cell

    // ---- class planes first (94 % of the bytes): running arg-max, kClsUnroll loads in flight
    float best[VEC], second[VEC];
    int best_c[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { best[j] = -INFINITY; second[j] = -INFINITY; best_c[j] = 0; }
    if (active && n_cls > 0) {
        if constexpr (VEC == 4) {
            const float* pc = L.cls + b * L.ks_b + a * L.ks_a + hw0;   // hw-contiguous planes
            int c = 0;
            for (; c + kClsUnroll <= n_cls; c += kClsUnroll) {
                float4 v[kClsUnroll];
#pragma unroll
                for (int u = 0; u < kClsUnroll; ++u) v[u] = ld_stream_v4(pc + (long long)(c + u) * L.ks_c);
#pragma unroll
                for (int u = 0; u < kClsUnroll; ++u) {
                    class_update(best[0], second[0], best_c[0], v[u].x, c + u);
                    class_update(best[1], second[1], best_c[1], v[u].y, c + u);
                    class_update(best[2], second[2], best_c[2], v[u].z, c + u);
                    class_update(best[3], second[3], best_c[3], v[u].w, c + u);
                }
            }
            for (; c < n_cls; ++c) {
                const float4 v = ld_stream_v4(pc + (long long)c * L.ks_c);
                class_update(best[0], second[0], best_c[0], v.x, c);
                class_update(best[1], second[1], best_c[1], v.y, c);
                class_update(best[2], second[2], best_c[2], v.z, c);
                class_update(best[3], second[3], best_c[3], v.w, c);
            }
        } else {
            const float* pc = L.cls + b * L.ks_b + a * L.ks_a + h0 * L.ks_h + w0 * L.ks_w;
            int c = 0;
            for (; c + kClsUnroll <= n_cls; c += kClsUnroll) {
                float v[kClsUnroll];
#pragma unroll
                for (int u = 0; u < kClsUnroll; ++u) v[u] = ld_stream(pc + (long long)(c + u) * L.ks_c);
#pragma unroll
                for (int u = 0; u < kClsUnroll; ++u) class_update(best[0], second[0], best_c[0], v[u], c + u);
            }
            for (; c < n_cls; ++c) class_update(best[0], second[0], best_c[0], ld_stream(pc + (long long)c * L.ks_c), c);
        }
    }

    if (n_cls > 1) {   // warp-uniform; rare cooperative re-scan of near-tied cells
        const float* pc0 = (VEC == 4) ? L.cls + b * L.ks_b + a * L.ks_a + hw0
                                      : L.cls + b * L.ks_b + a * L.ks_a + h0 * L.ks_h + w0 * L.ks_w;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const unsigned flagged = __ballot_sync(0xffffffffu, active && class_ambiguous(best[j], second[j]));
            if (flagged) class_rescan_warp(flagged, pc0 + j, L.ks_c, n_cls, best[j], best_c[j]);
        }
    }

    // ---- box and objectness planes (6 % of the bytes), loaded after the class loop so that they
    // do not occupy registers during it
    float t[5][VEC];
    float conf_logit[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) conf_logit[j] = 0.f;
    if (active) {
        if constexpr (VEC == 4) {
            const float* pb = L.bbox + b * L.bs_b + a * L.bs_a + hw0;
#pragma unroll
            for (int p = 0; p < 5; ++p) {
                if (p < n_par) {
                    const float4 v = ld_stream_v4(pb + p * L.bs_p);
                    t[p][0] = v.x; t[p][1] = v.y; t[p][2] = v.z; t[p][3] = v.w;
                }
            }
            if (KIND != MYDET_KIND_RETINA) {
                const float4 v = ld_stream_v4(L.conf + b * L.cs_b + a * L.cs_a + hw0);
                conf_logit[0] = v.x; conf_logit[1] = v.y; conf_logit[2] = v.z; conf_logit[3] = v.w;
            }
        } else {
            const float* pb = L.bbox + b * L.bs_b + a * L.bs_a + h0 * L.bs_h + w0 * L.bs_w;
#pragma unroll
            for (int p = 0; p < 5; ++p)
                if (p < n_par) t[p][0] = ld_stream(pb + p * L.bs_p);
            if (KIND != MYDET_KIND_RETINA)
                conf_logit[0] = ld_stream(L.conf + b * L.cs_b + a * L.cs_a + h0 * L.cs_h + w0 * L.cs_w);
        }
    }

    // ---- activations + box arithmetic
    float box[VEC][5];
    float score[VEC];
    const float aw = L.aw[a], ah = L.ah[a];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        score[j] = -INFINITY;
        if (active) {
            const int hw = hw0 + j;
            const int h = hw / L.n_w, w = hw - h * L.n_w;
            float tt[5];
#pragma unroll
            for (int p = 0; p < 5; ++p) tt[p] = (p < n_par) ? t[p][j] : 0.f;
            decode_box<KIND>(tt, box[j], n_par, (float)w, (float)h, L.stride, aw, ah, P.img_h, P.img_w, P.img_max);
            score[j] = final_score<KIND>(conf_logit[j], best[j], n_cls);
        }
    }

    const long long cand0 = (long long)L.out_offset + (long long)a * L.n_hw + hw0;  // flat index in the image
    if (!COMPACT) {
        if (!active) return;
        const long long row = (long long)b * P.n_total + cand0;
        float* ob = P.out_box + row * n_par;
        float* os = P.out_score + row;
        long long* oc = P.out_cls + row;
        if (VEC == 4 && n_par == 4 && ((reinterpret_cast<uintptr_t>(ob) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                reinterpret_cast<float4*>(ob)[j] = make_float4(box[j][0], box[j][1], box[j][2], box[j][3]);
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j)
#pragma unroll
                for (int p = 0; p < 5; ++p)
                    if (p < n_par) ob[j * n_par + p] = box[j][p];
        }
        if (VEC == 4 && ((reinterpret_cast<uintptr_t>(os) & 15) == 0) && ((reinterpret_cast<uintptr_t>(oc) & 15) == 0)) {
            *reinterpret_cast<float4*>(os) = make_float4(score[0], score[1 % VEC], score[2 % VEC], score[3 % VEC]);
            reinterpret_cast<longlong2*>(oc)[0] = make_longlong2(best_c[0], best_c[1 % VEC]);
            reinterpret_cast<longlong2*>(oc)[1] = make_longlong2(best_c[2 % VEC], best_c[3 % VEC]);
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) { os[j] = score[j]; oc[j] = best_c[j]; }
        }
    } else {
        // threshold + warp-aggregated stream compaction: one atomic per warp and image
        unsigned ballots[VEC];
        int total = 0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            ballots[j] = __ballot_sync(0xffffffffu, active && (score[j] >= P.thr));
            total += __popc(ballots[j]);
        }
        if (total == 0) return;
        int base = 0;
        if (lane == 0) base = atomicAdd(P.cand_count + b, total);
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned lt = lanemask_lt();
        int running = base;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            if ((ballots[j] >> lane) & 1u) {
                const int slot = running + __popc(ballots[j] & lt);
                if (slot < P.capacity) {
                    const long long row = (long long)b * P.capacity + slot;
                    float* ob = P.cand_box + row * n_par;
                    if (n_par == 4) {
                        *reinterpret_cast<float4*>(ob) = make_float4(box[j][0], box[j][1], box[j][2], box[j][3]);
                    } else {
#pragma unroll
                        for (int p = 0; p < 5; ++p)
                            if (p < n_par) ob[p] = box[j][p];
                    }
                    P.cand_score[row] = score[j];
                    P.cand_cls[row] = best_c[j];
                    P.cand_idx[row] = (int)(cand0 + j);
                }
            }
            running += __popc(ballots[j]);
        }
    }
}

template <int KIND, bool COMPACT>
__global__ void __launch_bounds__(kDecodeThreads, MYDET_DECODE_MINBLOCKS)
decode_kernel(const __grid_constant__ DecodeParams P) {
    // a dependent launched programmatically behind this grid (the post-process of mydet_detect) may be scheduled once
    // every CTA of this grid has started; it waits for this grid's completion before reading.  No-op otherwise.
    asm volatile("griddepcontrol.launch_dependents;");
    // CTA -> level (levels are laid out back to back in CTA index space)
    int l = 0;
#pragma unroll 1
    while (l + 1 < P.n_levels && (int)blockIdx.x >= P.lv[l + 1].first_block) ++l;
    const LevelDev& L = P.lv[l];
    const int unit = ((int)blockIdx.x - L.first_block) * kDecodeWarps + (int)(threadIdx.x >> 5);
    if (L.vec == 4) decode_unit<KIND, COMPACT, 4>(P, L, unit);
    else decode_unit<KIND, COMPACT, 1>(P, L, unit);
}

// --------------------------------------------------------------------------------------- host
static bool plane_contiguous(const mydet_level_t& s) {
    const int64_t hw = (int64_t)s.n_h * s.n_w;
    auto ok = [&](const float* p, int64_t sb, int64_t sa, int64_t sh, int64_t sw, int64_t sc) {
        if (!p) return true;
        if (sw != 1 || sh != s.n_w) return false;
        if (reinterpret_cast<uintptr_t>(p) & 15) return false;
        return (sb % 4 == 0) && (sa % 4 == 0) && (sc % 4 == 0);
    };
    if (hw % 4 != 0) return false;
    return ok(s.bbox, s.bbox_stride[0], s.bbox_stride[1], s.bbox_stride[2], s.bbox_stride[3], s.bbox_stride[4]) &&
           ok(s.conf, s.conf_stride[0], s.conf_stride[1], s.conf_stride[2], s.conf_stride[3], 0) &&
           ok(s.cls, s.cls_stride[0], s.cls_stride[1], s.cls_stride[2], s.cls_stride[3], s.cls_stride[4]);
}

int build_decode_params(DecodeParams& P, int kind, const mydet_level_t* levels, int n_levels, int batch,
                        int n_cls, int n_param, float img_h, float img_w, int* n_blocks, int64_t* n_total) {
    MYDET_REQUIRE(kind >= MYDET_KIND_YOLO && kind <= MYDET_KIND_UV5, "unknown decode kind %d", kind);
    MYDET_REQUIRE(levels && n_levels >= 1 && n_levels <= MYDET_MAX_LEVELS, "n_levels must be in [1,%d]", MYDET_MAX_LEVELS);
    MYDET_REQUIRE(batch >= 0 && n_cls >= 0 && n_cls <= MYDET_MAX_CLASS_ID + 1, "bad batch/n_cls");
    MYDET_REQUIRE(n_param == 4 || n_param == 5, "n_param must be 4 or 5");
    if (kind == MYDET_KIND_RAPID) MYDET_REQUIRE(n_param == 5, "RAPiD decode needs n_param == 5");
    if (kind == MYDET_KIND_YOLO || kind == MYDET_KIND_FCOS || kind == MYDET_KIND_UV5)
        MYDET_REQUIRE(n_param == 4, "this decode kind needs n_param == 4");
    if (kind == MYDET_KIND_FCOS || kind == MYDET_KIND_RETINA || kind == MYDET_KIND_UV5)
        MYDET_REQUIRE(n_cls > 0, "this decode kind needs n_cls > 0 (the reference crashes on 0, SURVEY 0.1)");
    memset(&P, 0, sizeof(P));
    P.n_levels = n_levels; P.batch = batch; P.n_cls = n_cls; P.n_param = n_param;
    P.img_h = img_h; P.img_w = img_w; P.img_max = img_h > img_w ? img_h : img_w;
    int blocks = 0;
    int64_t offset = 0;
    for (int i = 0; i < n_levels; ++i) {
        const mydet_level_t& s = levels[i];
        LevelDev& d = P.lv[i];
        MYDET_REQUIRE(s.bbox, "level %d: bbox pointer is NULL", i);
        MYDET_REQUIRE(kind == MYDET_KIND_RETINA || s.conf, "level %d: conf pointer is NULL", i);
        MYDET_REQUIRE(n_cls == 0 || s.cls, "level %d: cls pointer is NULL", i);
        MYDET_REQUIRE(s.n_anchor >= 1 && s.n_anchor <= MYDET_MAX_ANCHORS && s.n_h >= 1 && s.n_w >= 1,
                      "level %d: bad n_anchor/n_h/n_w", i);
        d.bbox = s.bbox; d.conf = s.conf; d.cls = s.cls;
        d.bs_b = s.bbox_stride[0]; d.bs_a = s.bbox_stride[1]; d.bs_h = s.bbox_stride[2]; d.bs_w = s.bbox_stride[3]; d.bs_p = s.bbox_stride[4];
        d.cs_b = s.conf_stride[0]; d.cs_a = s.conf_stride[1]; d.cs_h = s.conf_stride[2]; d.cs_w = s.conf_stride[3];
        d.ks_b = s.cls_stride[0]; d.ks_a = s.cls_stride[1]; d.ks_h = s.cls_stride[2]; d.ks_w = s.cls_stride[3]; d.ks_c = s.cls_stride[4];
        d.n_a = s.n_anchor; d.n_h = s.n_h; d.n_w = s.n_w; d.n_hw = s.n_h * s.n_w;
        d.vec = plane_contiguous(s) ? 4 : 1;
        const int groups = (d.n_hw + d.vec - 1) / d.vec;
        d.chunks = (groups + 31) / 32;
        d.first_block = blocks;
        d.out_offset = (int)offset;
        d.stride = s.stride;
        for (int a = 0; a < MYDET_MAX_ANCHORS; ++a) { d.aw[a] = s.anchor_w[a]; d.ah[a] = s.anchor_h[a]; }
        const long long units = (long long)batch * d.n_a * d.chunks;
        blocks += (int)((units + kDecodeWarps - 1) / kDecodeWarps);
        offset += (int64_t)d.n_a * d.n_hw;
    }
    MYDET_REQUIRE(offset <= MYDET_MAX_CANDIDATES, "more than %d candidates per image", MYDET_MAX_CANDIDATES);
    *n_blocks = blocks;
    *n_total = offset;
    return 0;
}

template <bool COMPACT>
static int launch_decode(int kind, const DecodeParams& P, int blocks, cudaStream_t st) {
    if (blocks == 0) return 0;
    switch (kind) {
        case MYDET_KIND_YOLO:   decode_kernel<MYDET_KIND_YOLO, COMPACT><<<blocks, kDecodeThreads, 0, st>>>(P); break;
        case MYDET_KIND_FCOS:   decode_kernel<MYDET_KIND_FCOS, COMPACT><<<blocks, kDecodeThreads, 0, st>>>(P); break;
        case MYDET_KIND_RAPID:  decode_kernel<MYDET_KIND_RAPID, COMPACT><<<blocks, kDecodeThreads, 0, st>>>(P); break;
        case MYDET_KIND_RETINA: decode_kernel<MYDET_KIND_RETINA, COMPACT><<<blocks, kDecodeThreads, 0, st>>>(P); break;
        default:                decode_kernel<MYDET_KIND_UV5, COMPACT><<<blocks, kDecodeThreads, 0, st>>>(P); break;
    }
    return launch_status("decode_kernel");
}

int decode_compact_impl(int kind, const mydet_level_t* levels, int n_levels, int batch, int n_cls, int n_param,
                        float img_h, float img_w, float conf_thres, float* cand_box, float* cand_score,
                        int32_t* cand_cls, int32_t* cand_idx, int32_t* cand_count, int32_t capacity,
                        int state_clean, int64_t* n_total_out, cudaStream_t st) {
    DecodeParams P;
    int blocks = 0;
    int64_t n_total = 0;
    int rc = build_decode_params(P, kind, levels, n_levels, batch, n_cls, n_param, img_h, img_w, &blocks, &n_total);
    if (rc) return rc;
    MYDET_REQUIRE(cand_box && cand_score && cand_cls && cand_idx && cand_count && capacity > 0,
                  "compact decode: NULL output or capacity <= 0");
    MYDET_REQUIRE(n_param != 4 || (reinterpret_cast<uintptr_t>(cand_box) & 15) == 0,
                  "compact decode: cand_box must be 16-byte aligned (4-parameter boxes are stored as 128-bit vectors)");
    P.thr = conf_thres;
    P.cand_box = cand_box; P.cand_score = cand_score; P.cand_cls = cand_cls; P.cand_idx = cand_idx;
    P.cand_count = cand_count; P.capacity = capacity;
    if (n_total_out) *n_total_out = n_total;
    if (batch == 0) return 0;
    if (!state_clean) MYDET_CUDA(cudaMemsetAsync(cand_count, 0, sizeof(int32_t) * (size_t)batch, st));
    return launch_decode<true>(kind, P, blocks, st);
}

}  // namespace mydet

using namespace mydet;

MYDET_API int mydet_decode_dense(int kind, const mydet_level_t* levels, int n_levels, int batch, int n_cls,
                                 int n_param, float img_h, float img_w, float* out_box, int64_t* out_cls,
                                 float* out_score, int64_t n_total, void* stream) {
    DecodeParams P;
    int blocks = 0;
    int64_t total = 0;
    int rc = build_decode_params(P, kind, levels, n_levels, batch, n_cls, n_param, img_h, img_w, &blocks, &total);
    if (rc) return rc;
    MYDET_REQUIRE(out_box && out_cls && out_score, "dense decode: NULL output");
    MYDET_REQUIRE(n_total == total, "n_total is %lld but the levels hold %lld candidates per image",
                  (long long)n_total, (long long)total);
    P.out_box = out_box; P.out_cls = reinterpret_cast<long long*>(out_cls); P.out_score = out_score;
    P.n_total = n_total;
    return launch_decode<false>(kind, P, blocks, (cudaStream_t)stream);
}

MYDET_API int mydet_decode_compact(int kind, const mydet_level_t* levels, int n_levels, int batch, int n_cls,
                                   int n_param, float img_h, float img_w, float conf_thres, float* cand_box,
                                   float* cand_score, int32_t* cand_cls, int32_t* cand_idx, int32_t* cand_count,
                                   int32_t capacity, int state_clean, void* stream) {
    return decode_compact_impl(kind, levels, n_levels, batch, n_cls, n_param, img_h, img_w, conf_thres, cand_box,
                               cand_score, cand_cls, cand_idx, cand_count, capacity, state_clean, nullptr,
                               (cudaStream_t)stream);
}
