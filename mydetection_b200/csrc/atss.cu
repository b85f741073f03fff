// ATSS anchor-to-GT assignment for one pyramid level.
//
// Replaces the target construction of FCOS_ATSS_Layer.forward (models/detlayers/fcos2.py:253-341)
// and _get_atss_threshold (:385-405).  Three kernels, everything stays on the device (the
// reference moves predictions to the CPU, loops over images and GTs in Python and copies the
// dense targets back):
//   prepare   : per image, GTs ordered by area descending (stable)                     (:299-303)
//   threshold : per (image, GT): the k nearest anchor centres of EVERY level (one warp per
//               level, exhaustive (d2, index)-ordered selection = torch.topk(largest=False)),
//               IoU of those L*k square anchors with the GT, thr = mean + unbiased std  (:385-405)
//   assign    : per (image, cell) of this level: ignore mask = max_GT IoU(pred, GT) > t (:306-308),
//               then the GTs in area order; positive iff IoU(anchor, GT) > thr and the cell lies
//               strictly inside the GT; the last (= smallest) positive GT owns TargetLTRB, classes
//               accumulate multi-hot                                                    (:312-341)
// IoU arithmetic is bboxes_iou's (utils/bbox_ops.py:38-49), bit-exact.
#include "internal.cuh"

namespace mydet {

constexpr int kAtssMaxK = 16;

struct AtssWs {
    float4* gt_sorted;   // B*max_gt   cxcywh, area-descending
    int* cls_sorted;     // B*max_gt
    int* src_sorted;     // B*max_gt   original GT index
    float* thr;          // B*max_gt   in sorted order
};

static size_t carve_atss(AtssWs& w, void* base, int batch, int max_gt) {
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
    const size_t bg = (size_t)batch * (size_t)(max_gt > 0 ? max_gt : 1);
    const size_t o0 = take(bg * 16), o1 = take(bg * 4), o2 = take(bg * 4), o3 = take(bg * 4);
    if (base) {
        char* p = static_cast<char*>(base);
        w.gt_sorted = (float4*)(p + o0); w.cls_sorted = (int*)(p + o1); w.src_sorted = (int*)(p + o2); w.thr = (float*)(p + o3);
    }
    return off;
}

// bboxes_iou(a, b, xyxy=False) for one pair, cxcywh inputs.
__device__ __forceinline__ float iou_cxcywh(float acx, float acy, float aw, float ah, float bcx, float bcy, float bw, float bh) {
    const float ahw = __fmul_rn(aw, 0.5f), ahh = __fmul_rn(ah, 0.5f), bhw = __fmul_rn(bw, 0.5f), bhh = __fmul_rn(bh, 0.5f);
    const float tlx = fmax_nan(__fsub_rn(acx, ahw), __fsub_rn(bcx, bhw)), tly = fmax_nan(__fsub_rn(acy, ahh), __fsub_rn(bcy, bhh));
    const float brx = fmin_nan(__fadd_rn(acx, ahw), __fadd_rn(bcx, bhw)), bry = fmin_nan(__fadd_rn(acy, ahh), __fadd_rn(bcy, bhh));
    const float en = (tlx < brx && tly < bry) ? 1.0f : 0.0f;
    const float inter = __fmul_rn(__fmul_rn(__fsub_rn(brx, tlx), __fsub_rn(bry, tly)), en);
    return iou_from_parts(inter, __fadd_rn(__fmul_rn(aw, ah), __fmul_rn(bw, bh)));
}

// The same, for callers that only compare the result with a NON-NEGATIVE threshold (or take a maximum that is then so
// compared): boxes that do not overlap give +-0 (or NaN for degenerate boxes) whatever the union is -- never above such a
// threshold -- so the multiply / divide tail is skipped for them (the vast majority of anchor-GT and cell-GT pairs).
__device__ __forceinline__ float iou_cxcywh_or_zero(float acx, float acy, float aw, float ah, float bcx, float bcy, float bw, float bh) {
    const float ahw = __fmul_rn(aw, 0.5f), ahh = __fmul_rn(ah, 0.5f), bhw = __fmul_rn(bw, 0.5f), bhh = __fmul_rn(bh, 0.5f);
    const float tlx = fmax_nan(__fsub_rn(acx, ahw), __fsub_rn(bcx, bhw)), tly = fmax_nan(__fsub_rn(acy, ahh), __fsub_rn(bcy, bhh));
    const float brx = fmin_nan(__fadd_rn(acx, ahw), __fadd_rn(bcx, bhw)), bry = fmin_nan(__fadd_rn(acy, ahh), __fadd_rn(bcy, bhh));
    if (!(tlx < brx && tly < bry)) return 0.0f;
    const float inter = __fmul_rn(__fsub_rn(brx, tlx), __fsub_rn(bry, tly));         // en == 1
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(__fmul_rn(aw, ah), __fmul_rn(bw, bh)), inter));
}

__global__ void atss_prepare_kernel(const float* gt_box, const long long* gt_cls, const int* gt_count, int max_gt, AtssWs w) {
    extern __shared__ float s_area[];
    const int b = blockIdx.x;
    const int n = min(max(gt_count[b], 0), max_gt);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float* g = gt_box + ((long long)b * max_gt + i) * 4;
        s_area[i] = __fmul_rn(g[2], g[3]);                              // fcos2.py:299
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float ai = s_area[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += (s_area[j] > ai || (s_area[j] == ai && j < i)) ? 1 : 0;
        const float* g = gt_box + ((long long)b * max_gt + i) * 4;
        const long long o = (long long)b * max_gt + rank;
        w.gt_sorted[o] = make_float4(g[0], g[1], g[2], g[3]);
        w.cls_sorted[o] = (int)gt_cls[(long long)b * max_gt + i];
        w.src_sorted[o] = i;
    }
}

struct AtssGeom {
    int n_levels, k, img_h, img_w;
    int stride[MYDET_MAX_LEVELS];
    float side[MYDET_MAX_LEVELS];
};

// grid (max_gt, B); one warp per level.
__global__ void atss_threshold_kernel(AtssGeom G, const int* gt_count, int max_gt, AtssWs w, float* thr_user) {
    __shared__ float s_iou[MYDET_MAX_LEVELS * kAtssMaxK];
    const int b = blockIdx.y, g = blockIdx.x;
    const int n_gt = min(max(gt_count[b], 0), max_gt);
    if (g >= n_gt) return;
    const int level = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4 gt = w.gt_sorted[(long long)b * max_gt + g];
    if (level < G.n_levels) {
        const int s = G.stride[level];
        const int n_w = G.img_w / s, n_h = G.img_h / s;
        const float fs = (float)s, half = __fmul_rn(0.5f, fs);
        // Candidates.  When the GT centre lies inside the image, the k (<= 16) nearest cell centres of a regular
        // grid all lie within 5 cells of the cell that holds the centre, so only that (clipped) 11x11 window is
        // searched -- with the same (distance, index) order, hence the same picks as the exhaustive scan over all n
        // anchors (9 rounds x 6 400 anchors per GT at stride 8), which remains the path for centres outside the image
        // and for grids too thin for the argument (tests/test_kernel_claims_cpu.py checks both windows by brute force).
        int r_lo = 0, c_lo = 0, w_rows = n_h, w_cols = n_w;
        // the window argument needs an m x m block of cells (m = ceil(sqrt(k)) <= 4) next to the centre's cell INSIDE the
        // grid: its centres are within (m - 0.5) sqrt(2) <= 4.95 cells, cells outside the window are >= 5.5 away.  A
        // grid thinner than m (a 1 x 20 level: the 9 nearest of a corner reach 8.5 cells) takes the exhaustive scan.
        const int m_blk = G.k <= 1 ? 1 : (G.k <= 4 ? 2 : (G.k <= 9 ? 3 : 4));
        if (gt.x >= 0.f && gt.x <= (float)G.img_w && gt.y >= 0.f && gt.y <= (float)G.img_h && n_h >= m_blk && n_w >= m_blk) {
            const int col0 = min(max((int)floorf(gt.x / fs), 0), n_w - 1), row0 = min(max((int)floorf(gt.y / fs), 0), n_h - 1);
            // k <= 9 and the 3x3 block around the centre's cell lies inside the grid: those 9 centres are within
            // 1.5 sqrt(2) = 2.13 cells, every cell outside a 5x5 window is >= 2.5 cells away -- 25 candidates, one per lane
            const bool small = G.k <= 9 && col0 >= 1 && col0 + 1 <= n_w - 1 && row0 >= 1 && row0 + 1 <= n_h - 1;
            const int R = small ? 2 : 5;
            c_lo = max(col0 - R, 0); r_lo = max(row0 - R, 0);
            w_cols = min(col0 + R, n_w - 1) - c_lo + 1; w_rows = min(row0 + R, n_h - 1) - r_lo + 1;
        }
        const int n_cand = w_rows * w_cols;
        auto cand = [&](int q, float& d, int& i) {             // q-th cell of the window: squared distance of its centre, flat index
            const int wr = q / w_cols;
            const int row = r_lo + wr, col = c_lo + (q - wr * w_cols);
            i = row * n_w + col;
            const float ax = __fadd_rn(__fmul_rn((float)col, fs), half);
            const float ay = __fadd_rn(__fmul_rn((float)row, fs), half);
            const float dx = __fsub_rn(gt.x, ax), dy = __fsub_rn(gt.y, ay);
            d = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));                            // :396
        };
        auto anchor_iou = [&](int i) {                         // IoU of the square anchor at flat index i with the GT, :401
            const int row = i / n_w, col = i - row * n_w;
            const float ax = __fadd_rn(__fmul_rn((float)col, fs), half);
            const float ay = __fadd_rn(__fmul_rn((float)row, fs), half);
            return iou_cxcywh(gt.x, gt.y, gt.z, gt.w, ax, ay, G.side[level], G.side[level]);
        };
        if (n_cand <= 32) {
            // one candidate per lane: its position in the (distance, index) order = the number of candidates before it;
            // the first k positions are torch.topk(largest=False)'s picks, in its order
            float d = INFINITY;
            int i = 0x7fffffff;
            if (lane < n_cand) cand(lane, d, i);
            int rank = 0;
            for (int j = 0; j < n_cand; ++j) {
                const float od = __shfl_sync(0xffffffffu, d, j);
                const int oi = __shfl_sync(0xffffffffu, i, j);
                rank += (od < d || (od == d && oi < i)) ? 1 : 0;
            }
            if (lane < n_cand && rank < G.k) s_iou[level * G.k + rank] = anchor_iou(i);
        } else {
            // the 11 x 11 window is at most 4 cells per lane: their (distance, index) pairs are computed ONCE and the k
            // selection rounds run on registers; the exhaustive scan (centre outside the image) recomputes them per round
            constexpr int kPerLane = 4;
            const bool in_regs = n_cand <= 32 * kPerLane;
            float cd[kPerLane];
            int ci[kPerLane];
#pragma unroll
            for (int u = 0; u < kPerLane; ++u) {
                cd[u] = INFINITY; ci[u] = 0x7fffffff;
                if (in_regs && lane + 32 * u < n_cand) cand(lane + 32 * u, cd[u], ci[u]);
            }
            float last_d = -1.0f;
            int last_i = -1, my_i = 0;
            for (int round = 0; round < G.k; ++round) {
                float best_d = INFINITY;
                int best_i = 0x7fffffff;
                if (in_regs) {
#pragma unroll
                    for (int u = 0; u < kPerLane; ++u) {
                        const float d = cd[u];
                        const int i = ci[u];
                        const bool after_last = d > last_d || (d == last_d && i > last_i);
                        if (after_last && (d < best_d || (d == best_d && i < best_i))) { best_d = d; best_i = i; }
                    }
                } else {
                    for (int q = lane; q < n_cand; q += 32) {
                        float d;
                        int i;
                        cand(q, d, i);
                        const bool after_last = d > last_d || (d == last_d && i > last_i);
                        if (after_last && (d < best_d || (d == best_d && i < best_i))) { best_d = d; best_i = i; }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float od = __shfl_xor_sync(0xffffffffu, best_d, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
                    if (od < best_d || (od == best_d && oi < best_i)) { best_d = od; best_i = oi; }
                }
                last_d = best_d; last_i = best_i;
                if (lane == round) my_i = best_i;              // every lane holds the pick; lane `round` keeps it
            }
            if (lane < G.k) s_iou[level * G.k + lane] = anchor_iou(my_i);      // the k IoUs side by side, not one per round on lane 0
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int cnt = G.n_levels * G.k;
        double sum = 0.0;
        for (int i = 0; i < cnt; ++i) sum += (double)s_iou[i];
        const double mean = sum / cnt;
        double ss = 0.0;
        for (int i = 0; i < cnt; ++i) { const double d = (double)s_iou[i] - mean; ss += d * d; }
        const float fmean = (float)mean;
        const float fstd = (float)sqrt(ss / (cnt - 1));                                    // unbiased, :403
        const float thr = __fadd_rn(fmean, fstd);                                          // :404
        w.thr[(long long)b * max_gt + g] = thr;
        if (thr_user) thr_user[(long long)b * max_gt + w.src_sorted[(long long)b * max_gt + g]] = thr;
    }
}

// thresholds computed by an earlier level's call (caller's GT order) -> area-sorted order
__global__ void atss_load_thr_kernel(const int* gt_count, int max_gt, AtssWs w, const float* thr_user) {
    const int b = blockIdx.y, g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= min(max(gt_count[b], 0), max_gt)) return;
    const long long o = (long long)b * max_gt + g;
    w.thr[o] = thr_user[(long long)b * max_gt + w.src_sorted[o]];
}

struct AssignParams {
    const float* t; long long ts_b, ts_h, ts_w, ts_p;
    int n_h, n_w, n_cls, max_gt;
    float stride, side, ignore_thres;
    float center_region, anch_min, anch_max;     // FCOS mode (FCOSLayer, fcos2.py:113-133)
    const int* gt_count;
    unsigned char* positive; unsigned char* ignored;
    float* target_ltrb; float* target_conf; float* target_cls;
};

constexpr int kAssignThreads = 128;

// ATSS: positive iff IoU(anchor, GT) > adaptive threshold and the cell lies inside the GT (fcos2.py:321-331).
// !ATSS: FCOSLayer's rule (fcos2.py:113-133): the cell centre lies strictly inside the GT's central region
// (the GT shrunk by center_region) and anch_min < max(l,t,r,b) < anch_max.  Everything else is shared.
template <bool ATSS>
__device__ __forceinline__ void assign_body(const AssignParams& P, const AtssWs& w, int block_in_level) {
    extern __shared__ float4 s_gt[];                       // max_gt boxes, then thr, cls, and the list of GTs that can matter here
    float* s_thr = reinterpret_cast<float*>(s_gt + P.max_gt);
    int* s_cls = reinterpret_cast<int*>(s_thr + P.max_gt);
    int* s_list = s_cls + P.max_gt;
    __shared__ float s_hull[kAssignThreads / 32][4];
    __shared__ int s_nlist;
    const int b = blockIdx.y;
    const int n_hw = P.n_h * P.n_w;
    const int cell0 = block_in_level * kAssignThreads;
    const int n_gt = min(max(P.gt_count[b], 0), P.max_gt);
    for (int i = threadIdx.x; i < n_gt; i += kAssignThreads) {
        s_gt[i] = w.gt_sorted[(long long)b * P.max_gt + i];
        s_thr[i] = ATSS ? w.thr[(long long)b * P.max_gt + i] : 0.f;
        s_cls[i] = w.cls_sorted[(long long)b * P.max_gt + i];
    }
    // zero this CTA's slab of the class target (coalesced), ones are scattered after the barrier
    {
        const int cells = min(kAssignThreads, n_hw - cell0);
        float* slab = P.target_cls + ((long long)b * n_hw + cell0) * P.n_cls;
        const long long total = (long long)cells * P.n_cls;
        for (long long i = threadIdx.x; i < total; i += kAssignThreads) slab[i] = 0.0f;
    }
    const int cell = cell0 + threadIdx.x;
    const bool valid = cell < n_hw;
    const int row = cell / P.n_w, col = cell - row * P.n_w;
    const float half = __fmul_rn(P.stride, 0.5f);
    // fcos2.py:256-259  linspace(0,img,n+1)[:-1] + 0.5*stride  == col*stride + stride/2 for integer strides
    const float gx = __fadd_rn(__fmul_rn((float)col, P.stride), half);
    const float gy = __fadd_rn(__fmul_rn((float)row, P.stride), half);
    // un-clamped predicted box of this cell (fcos2.py:42, :253, :444-450)
    float pcx = 0.f, pcy = 0.f, pw = 0.f, ph = 0.f;
    if (valid) {
        const float* t = P.t + b * P.ts_b + row * P.ts_h + col * P.ts_w;
        const float l = __fmul_rn(expf(t[0]), P.stride), tp = __fmul_rn(expf(t[P.ts_p]), P.stride);
        const float r = __fmul_rn(expf(t[2 * P.ts_p]), P.stride), bt = __fmul_rn(expf(t[3 * P.ts_p]), P.stride);
        pcx = __fadd_rn(gx, __fmul_rn(__fsub_rn(r, l), 0.5f)); pcy = __fadd_rn(gy, __fmul_rn(__fsub_rn(bt, tp), 0.5f));
        pw = __fadd_rn(l, r); ph = __fadd_rn(tp, bt);
    }
    // GT cull.  A GT can only matter to a cell of this CTA if it meets the hull of the CTA's predicted boxes and cell
    // centres: the ignore test needs IoU(pred, GT) > a non-negative threshold, i.e. an overlap, and a positive cell lies
    // inside its GT.  The hull is built from the SAME float corner expressions the pair tests use (min / max are
    // monotonic), compared non-strictly, so no GT that could pass a test is dropped; the area order is kept.  A
    // 128-cell CTA of the finest level is a strip two rows high: ~80 % of the GTs go.
    const bool skip_disjoint = P.ignore_thres >= 0.f;      // else a disjoint pair's IoU of 0 could exceed the threshold: keep every GT and the full formula
    {
        float hx1 = INFINITY, hy1 = INFINITY, hx2 = -INFINITY, hy2 = -INFINITY;
        if (valid) {
            const float phw = __fmul_rn(pw, 0.5f), phh = __fmul_rn(ph, 0.5f);
            hx1 = fminf(gx, __fsub_rn(pcx, phw)); hy1 = fminf(gy, __fsub_rn(pcy, phh));      // fminf / fmaxf drop a NaN corner:
            hx2 = fmaxf(gx, __fadd_rn(pcx, phw)); hy2 = fmaxf(gy, __fadd_rn(pcy, phh));      // such a box overlaps nothing anyway
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            hx1 = fminf(hx1, __shfl_xor_sync(0xffffffffu, hx1, o)); hy1 = fminf(hy1, __shfl_xor_sync(0xffffffffu, hy1, o));
            hx2 = fmaxf(hx2, __shfl_xor_sync(0xffffffffu, hx2, o)); hy2 = fmaxf(hy2, __shfl_xor_sync(0xffffffffu, hy2, o));
        }
        if ((threadIdx.x & 31) == 0) {
            float* h = s_hull[threadIdx.x >> 5];
            h[0] = hx1; h[1] = hy1; h[2] = hx2; h[3] = hy2;
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        float hx1 = s_hull[0][0], hy1 = s_hull[0][1], hx2 = s_hull[0][2], hy2 = s_hull[0][3];
#pragma unroll
        for (int q = 1; q < kAssignThreads / 32; ++q) {
            hx1 = fminf(hx1, s_hull[q][0]); hy1 = fminf(hy1, s_hull[q][1]);
            hx2 = fmaxf(hx2, s_hull[q][2]); hy2 = fmaxf(hy2, s_hull[q][3]);
        }
        int count = 0;
        for (int g0 = 0; g0 < n_gt; g0 += 32) {
            const int g = g0 + (int)threadIdx.x;
            bool keep = false;
            if (g < n_gt) {
                const float4 gt = s_gt[g];
                float hw = __fmul_rn(gt.z, 0.5f), hh = __fmul_rn(gt.w, 0.5f);
                if (!ATSS) {     // the centre region is larger than the GT when center_region > 1: cull against the union
                    hw = fmaxf(hw, __fmul_rn(__fmul_rn(gt.z, P.center_region), 0.5f));
                    hh = fmaxf(hh, __fmul_rn(__fmul_rn(gt.w, P.center_region), 0.5f));
                }
                keep = !skip_disjoint ||
                       (fmaxf(hx1, __fsub_rn(gt.x, hw)) <= fminf(hx2, __fadd_rn(gt.x, hw)) &&
                        fmaxf(hy1, __fsub_rn(gt.y, hh)) <= fminf(hy2, __fadd_rn(gt.y, hh)));
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (keep) s_list[count + __popc(bal & ((1u << threadIdx.x) - 1u))] = g;
            count += __popc(bal);
        }
        if (threadIdx.x == 0) s_nlist = count;
    }
    __syncthreads();
    if (!valid) return;
    const int n_list = s_nlist;

    float best_iou = -INFINITY;
    bool positive = false;
    float4 ltrb = make_float4(0.f, 0.f, 0.f, 0.f);
    float* cls_row = P.target_cls + ((long long)b * n_hw + cell) * P.n_cls;
    for (int q = 0; q < n_list; ++q) {
        const int g = s_list[q];
        const float4 gt = s_gt[g];
        best_iou = fmax_nan(best_iou, skip_disjoint ? iou_cxcywh_or_zero(pcx, pcy, pw, ph, gt.x, gt.y, gt.z, gt.w)
                                                 : iou_cxcywh(pcx, pcy, pw, ph, gt.x, gt.y, gt.z, gt.w));   // :306-307
        const float hw = __fmul_rn(gt.z, 0.5f), hh = __fmul_rn(gt.w, 0.5f);                  // :408-414, cr = 1
        const float tl = __fsub_rn(gx, __fsub_rn(gt.x, hw)), tt = __fsub_rn(gy, __fsub_rn(gt.y, hh));
        const float tr = __fsub_rn(__fadd_rn(gt.x, hw), gx), tb = __fsub_rn(__fadd_rn(gt.y, hh), gy);
        bool pos;
        if (ATSS) {
            const bool inside = tl > 0.f && tt > 0.f && tr > 0.f && tb > 0.f;                // :321
            // :329-331: the anchor IoU only where the cell lies inside the GT (the result is ANDed with `inside`)
            pos = inside && iou_cxcywh(gx, gy, P.side, P.side, gt.x, gt.y, gt.z, gt.w) > s_thr[g];
        } else {
            // _xywh_to_xyxy(bb, cr): c -/+ (w * cr) / 2                                        :408-414
            const float chw = __fmul_rn(__fmul_rn(gt.z, P.center_region), 0.5f), chh = __fmul_rn(__fmul_rn(gt.w, P.center_region), 0.5f);
            const bool centre = gx > __fsub_rn(gt.x, chw) && gx < __fadd_rn(gt.x, chw) &&
                                gy > __fsub_rn(gt.y, chh) && gy < __fadd_rn(gt.y, chh);      // :123-124
            const float mx = fmax_nan(fmax_nan(tl, tt), fmax_nan(tr, tb));                            // :126
            pos = centre && P.anch_min < mx && mx < P.anch_max;                              // :127-129
        }
        if (pos) {
            positive = true;
            ltrb = make_float4(tl, tt, tr, tb);                                              // :335, last writer wins
            const int c = s_cls[g];
            if (c >= 0 && c < P.n_cls) cls_row[c] = 1.0f;                                    // :339-340
        }
    }
    const long long o = (long long)b * n_hw + cell;
    P.positive[o] = positive ? 1 : 0;
    P.ignored[o] = (n_gt > 0 && best_iou > P.ignore_thres) ? 1 : 0;                          // :308
    reinterpret_cast<float4*>(P.target_ltrb)[o] = ltrb;
    P.target_conf[o] = positive ? 1.0f : 0.0f;                                               // :337
}

template <bool ATSS>
__global__ void __launch_bounds__(kAssignThreads) assign_kernel(const __grid_constant__ AssignParams P, AtssWs w) {
    assign_body<ATSS>(P, w, (int)blockIdx.x);
}

// All pyramid levels in one grid: the CTAs of the levels are laid out back to back (finest level first), so the few
// CTAs of the coarse levels fill the tail of the big one instead of running as four under-filled launches.
struct AssignLevels {
    AssignParams lv[MYDET_MAX_LEVELS];
    int first_block[MYDET_MAX_LEVELS + 1];
    int n_levels;
};
__global__ void __launch_bounds__(kAssignThreads) assign_levels_kernel(const __grid_constant__ AssignLevels A, AtssWs w) {
    int l = 0;
#pragma unroll 1
    while (l + 1 < A.n_levels && (int)blockIdx.x >= A.first_block[l + 1]) ++l;
    assign_body<true>(A.lv[l], w, (int)blockIdx.x - A.first_block[l]);
}

}  // namespace mydet

using namespace mydet;

MYDET_API size_t mydet_atss_workspace_bytes(int batch, int max_gt) {
    AtssWs w;
    return carve_atss(w, nullptr, batch, max_gt);
}

MYDET_API int mydet_atss_assign(const float* t_ltrb, const int64_t t_stride[4], int batch, int level, int n_levels,
                                const int32_t* strides, const float* anchor_sides, int img_h, int img_w,
                                const float* gt_box, const int64_t* gt_cls, const int32_t* gt_count, int max_gt,
                                int topk, float ignore_thres, int n_cls, uint8_t* positive, uint8_t* ignored,
                                float* target_ltrb, float* target_conf, float* target_cls, float* thr_out,
                                int thr_is_input, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MYDET_REQUIRE(n_levels >= 1 && n_levels <= MYDET_MAX_LEVELS && level >= 0 && level < n_levels, "bad level / n_levels");
    MYDET_REQUIRE(strides && anchor_sides && t_stride, "NULL host array");
    MYDET_REQUIRE(topk >= 1 && topk <= kAtssMaxK, "topk must be in [1,%d]", kAtssMaxK);
    MYDET_REQUIRE(n_levels * topk >= 2, "need at least two candidate anchors for the std");
    MYDET_REQUIRE(batch >= 0 && max_gt >= 0 && n_cls > 0, "bad batch / max_gt / n_cls");
    MYDET_REQUIRE(max_gt <= 2048, "more than 2048 GT boxes per image");
    AtssGeom G;
    G.n_levels = n_levels; G.k = topk; G.img_h = img_h; G.img_w = img_w;
    for (int i = 0; i < n_levels; ++i) {
        MYDET_REQUIRE(strides[i] > 0 && img_h % strides[i] == 0 && img_w % strides[i] == 0,
                      "image size must be divisible by every stride (fcos2.py:266)");
        MYDET_REQUIRE((img_h / strides[i]) * (img_w / strides[i]) >= topk,
                      "level %d has fewer than k anchors (torch.topk would raise, fcos2.py:397)", i);
        G.stride[i] = strides[i]; G.side[i] = anchor_sides[i];
    }
    if (batch == 0) return 0;
    MYDET_REQUIRE(t_ltrb && gt_count && positive && ignored && target_ltrb && target_conf && target_cls, "NULL tensor pointer");
    MYDET_REQUIRE((reinterpret_cast<uintptr_t>(target_ltrb) & 15) == 0, "target_ltrb must be 16-byte aligned (written as 128-bit vectors)");
    MYDET_REQUIRE(max_gt == 0 || (gt_box && gt_cls), "NULL GT pointer");
    AtssWs w;
    const size_t need = carve_atss(w, workspace, batch, max_gt);
    if (!workspace || need > workspace_bytes) {
        set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        return MYDET_ERR_WORKSPACE;
    }
    if (max_gt > 0) {
        atss_prepare_kernel<<<batch, 128, sizeof(float) * max_gt, st>>>(gt_box, reinterpret_cast<const long long*>(gt_cls), gt_count, max_gt, w);
        if (thr_is_input) {
            MYDET_REQUIRE(thr_out, "thr_is_input needs the thresholds in thr_out");
            atss_load_thr_kernel<<<dim3((max_gt + 127) / 128, batch), 128, 0, st>>>(gt_count, max_gt, w, thr_out);
        } else {
            atss_threshold_kernel<<<dim3(max_gt, batch), 32 * n_levels, 0, st>>>(G, gt_count, max_gt, w, thr_out);
        }
    }
    AssignParams P;
    P.t = t_ltrb; P.ts_b = t_stride[0]; P.ts_h = t_stride[1]; P.ts_w = t_stride[2]; P.ts_p = t_stride[3];
    P.n_h = img_h / strides[level]; P.n_w = img_w / strides[level]; P.n_cls = n_cls; P.max_gt = max_gt;
    P.stride = (float)strides[level]; P.side = anchor_sides[level]; P.ignore_thres = ignore_thres;
    P.gt_count = gt_count; P.positive = positive; P.ignored = ignored;
    P.target_ltrb = target_ltrb; P.target_conf = target_conf; P.target_cls = target_cls;
    const int n_hw = P.n_h * P.n_w;
    const size_t smem = (size_t)max_gt * (16 + 4 + 4 + 4);
    P.center_region = 0.f; P.anch_min = 0.f; P.anch_max = 0.f;
    assign_kernel<true><<<dim3((n_hw + kAssignThreads - 1) / kAssignThreads, batch), kAssignThreads, smem, st>>>(P, w);
    return launch_status("atss kernels");
}

MYDET_API int mydet_fcos_assign(const float* t_ltrb, const int64_t t_stride[4], int batch, int stride, int img_h, int img_w,
                                const float* gt_box, const int64_t* gt_cls, const int32_t* gt_count, int max_gt,
                                float center_region, float anch_min, float anch_max, float ignore_thres, int n_cls,
                                uint8_t* positive, uint8_t* ignored, float* target_ltrb, float* target_conf,
                                float* target_cls, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MYDET_REQUIRE(t_stride && stride > 0, "NULL stride array / bad stride");
    MYDET_REQUIRE(batch >= 0 && max_gt >= 0 && n_cls > 0, "bad batch / max_gt / n_cls");
    MYDET_REQUIRE(max_gt <= 2048, "more than 2048 GT boxes per image");
    if (batch == 0) return 0;
    MYDET_REQUIRE(t_ltrb && gt_count && positive && ignored && target_ltrb && target_conf && target_cls, "NULL tensor pointer");
    MYDET_REQUIRE((reinterpret_cast<uintptr_t>(target_ltrb) & 15) == 0, "target_ltrb must be 16-byte aligned (written as 128-bit vectors)");
    MYDET_REQUIRE(max_gt == 0 || (gt_box && gt_cls), "NULL GT pointer");
    AtssWs w;
    const size_t need = carve_atss(w, workspace, batch, max_gt);
    if (!workspace || need > workspace_bytes) {
        set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        return MYDET_ERR_WORKSPACE;
    }
    if (max_gt > 0)
        atss_prepare_kernel<<<batch, 128, sizeof(float) * max_gt, st>>>(gt_box, reinterpret_cast<const long long*>(gt_cls), gt_count, max_gt, w);
    AssignParams P;
    P.t = t_ltrb; P.ts_b = t_stride[0]; P.ts_h = t_stride[1]; P.ts_w = t_stride[2]; P.ts_p = t_stride[3];
    P.n_h = img_h / stride; P.n_w = img_w / stride; P.n_cls = n_cls; P.max_gt = max_gt;   // int(img / stride), fcos2.py:27
    P.stride = (float)stride; P.side = 0.f; P.ignore_thres = ignore_thres;
    P.center_region = center_region; P.anch_min = anch_min; P.anch_max = anch_max;
    P.gt_count = gt_count; P.positive = positive; P.ignored = ignored;
    P.target_ltrb = target_ltrb; P.target_conf = target_conf; P.target_cls = target_cls;
    const int n_hw = P.n_h * P.n_w;
    const size_t smem = (size_t)max_gt * (16 + 4 + 4 + 4);
    assign_kernel<false><<<dim3((n_hw + kAssignThreads - 1) / kAssignThreads, batch), kAssignThreads, smem, st>>>(P, w);
    return launch_status("fcos assign kernels");
}

MYDET_API int mydet_atss_assign_levels(const mydet_atss_level_t* levels, int n_levels, const int32_t* strides,
                                       const float* anchor_sides, int batch, int img_h, int img_w, const float* gt_box,
                                       const int64_t* gt_cls, const int32_t* gt_count, int max_gt, int topk,
                                       float ignore_thres, int n_cls, float* thr_out, void* workspace,
                                       size_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MYDET_REQUIRE(levels && n_levels >= 1 && n_levels <= MYDET_MAX_LEVELS, "n_levels must be in [1,%d]", MYDET_MAX_LEVELS);
    MYDET_REQUIRE(strides && anchor_sides, "NULL host array");
    MYDET_REQUIRE(topk >= 1 && topk <= kAtssMaxK, "topk must be in [1,%d]", kAtssMaxK);
    MYDET_REQUIRE(n_levels * topk >= 2, "need at least two candidate anchors for the std");
    MYDET_REQUIRE(batch >= 0 && max_gt >= 0 && n_cls > 0, "bad batch / max_gt / n_cls");
    MYDET_REQUIRE(max_gt <= 2048, "more than 2048 GT boxes per image");
    AtssGeom G;
    G.n_levels = n_levels; G.k = topk; G.img_h = img_h; G.img_w = img_w;
    for (int i = 0; i < n_levels; ++i) {
        MYDET_REQUIRE(strides[i] > 0 && img_h % strides[i] == 0 && img_w % strides[i] == 0,
                      "image size must be divisible by every stride (fcos2.py:266)");
        MYDET_REQUIRE((img_h / strides[i]) * (img_w / strides[i]) >= topk,
                      "level %d has fewer than k anchors (torch.topk would raise, fcos2.py:397)", i);
        G.stride[i] = strides[i]; G.side[i] = anchor_sides[i];
    }
    if (batch == 0) return 0;
    MYDET_REQUIRE(gt_count, "NULL gt_count");
    MYDET_REQUIRE(max_gt == 0 || (gt_box && gt_cls), "NULL GT pointer");
    AtssWs w;
    const size_t need = carve_atss(w, workspace, batch, max_gt);
    if (!workspace || need > workspace_bytes) {
        set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        return MYDET_ERR_WORKSPACE;
    }
    AssignLevels A;
    A.n_levels = n_levels;
    int blocks = 0;
    for (int i = 0; i < n_levels; ++i) {
        const mydet_atss_level_t& L = levels[i];
        MYDET_REQUIRE(L.t_ltrb && L.positive && L.ignored && L.target_ltrb && L.target_conf && L.target_cls, "NULL tensor pointer in level %d", i);
        MYDET_REQUIRE((reinterpret_cast<uintptr_t>(L.target_ltrb) & 15) == 0, "target_ltrb of level %d must be 16-byte aligned (written as 128-bit vectors)", i);
        AssignParams& P = A.lv[i];
        P.t = L.t_ltrb; P.ts_b = L.t_stride[0]; P.ts_h = L.t_stride[1]; P.ts_w = L.t_stride[2]; P.ts_p = L.t_stride[3];
        P.n_h = img_h / strides[i]; P.n_w = img_w / strides[i]; P.n_cls = n_cls; P.max_gt = max_gt;
        P.stride = (float)strides[i]; P.side = anchor_sides[i]; P.ignore_thres = ignore_thres;
        P.center_region = 0.f; P.anch_min = 0.f; P.anch_max = 0.f;
        P.gt_count = gt_count; P.positive = L.positive; P.ignored = L.ignored;
        P.target_ltrb = L.target_ltrb; P.target_conf = L.target_conf; P.target_cls = L.target_cls;
        A.first_block[i] = blocks;
        blocks += (P.n_h * P.n_w + kAssignThreads - 1) / kAssignThreads;
    }
    A.first_block[n_levels] = blocks;
    if (max_gt > 0) {
        atss_prepare_kernel<<<batch, 128, sizeof(float) * max_gt, st>>>(gt_box, reinterpret_cast<const long long*>(gt_cls), gt_count, max_gt, w);
        atss_threshold_kernel<<<dim3(max_gt, batch), 32 * n_levels, 0, st>>>(G, gt_count, max_gt, w, thr_out);
    }
    const size_t smem = (size_t)max_gt * (16 + 4 + 4 + 4);
    assign_levels_kernel<<<dim3(blocks, batch), kAssignThreads, smem, st>>>(A, w);
    return launch_status("atss kernels (all levels)");
}
