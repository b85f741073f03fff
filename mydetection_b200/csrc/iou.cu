// Pairwise IoU matrices.
//   mydet_iou_aabb_pairwise replaces bboxes_iou (utils/bbox_ops.py:6-49), bit-exact: the same
//   float32 operations in the same order (c -/+ wh/2, max/min, (br-tl).x*(br-tl).y*en,
//   area_i/(area_a+area_b-area_i)), no FMA contraction.
//   mydet_iou_rot_pairwise replaces iou_rle (utils/bbox_ops.py:52-100) with exact float64 polygon
//   clipping (rotgeom.cuh); like the reference it returns float64.
// Layout: each CTA computes a 32 (rows of a) x 128 (columns of b) tile; the b-tile is staged in
// shared memory once and reused by the 32 rows; stores are coalesced along K.
#include "internal.cuh"
#include "rotgeom.cuh"

namespace mydet {

constexpr int kIouCols = 128;
constexpr int kIouRows = 32;

// V columns per thread: V = 4 writes one 128-bit vector per row (k % 4 == 0 and a 16-byte aligned `out`), and the row
// box read from shared memory serves 4 pairs; V = 1 is the general case.
template <int V>
__global__ void __launch_bounds__(kIouCols) iou_aabb_kernel(const float* __restrict__ a, long long n,
                                                           const float* __restrict__ b, long long k, int xyxy,
                                                           float* __restrict__ out) {
    __shared__ float4 lo_hi_a[kIouRows];   // tl.x tl.y br.x br.y of the row boxes
    __shared__ float area_a[kIouRows];
    const long long col = ((long long)blockIdx.x * kIouCols + threadIdx.x) * V;
    const long long row0 = (long long)blockIdx.y * kIouRows;
    if (threadIdx.x < kIouRows && row0 + threadIdx.x < n) {
        const float4 v = reinterpret_cast<const float4*>(a)[row0 + threadIdx.x];
        if (xyxy) {
            lo_hi_a[threadIdx.x] = v;
            area_a[threadIdx.x] = __fmul_rn(__fsub_rn(v.z, v.x), __fsub_rn(v.w, v.y));   // :36 prod(hi-lo)
        } else {
            const float hw = __fmul_rn(v.z, 0.5f), hh = __fmul_rn(v.w, 0.5f);
            lo_hi_a[threadIdx.x] = make_float4(__fsub_rn(v.x, hw), __fsub_rn(v.y, hh), __fadd_rn(v.x, hw), __fadd_rn(v.y, hh));
            area_a[threadIdx.x] = __fmul_rn(v.z, v.w);                                    // :45 prod(wh)
        }
    }
    __syncthreads();
    if (col >= k) return;                  // V == 4: k % 4 == 0, so a thread has all of its columns or none
    float4 cb[V];
    float area_b[V];
#pragma unroll
    for (int c = 0; c < V; ++c) {
        const float4 vb = reinterpret_cast<const float4*>(b)[col + c];
        if (xyxy) {
            cb[c] = vb;
            area_b[c] = __fmul_rn(__fsub_rn(vb.z, vb.x), __fsub_rn(vb.w, vb.y));
        } else {
            const float hw = __fmul_rn(vb.z, 0.5f), hh = __fmul_rn(vb.w, 0.5f);
            cb[c] = make_float4(__fsub_rn(vb.x, hw), __fsub_rn(vb.y, hh), __fadd_rn(vb.x, hw), __fadd_rn(vb.y, hh));
            area_b[c] = __fmul_rn(vb.z, vb.w);
        }
    }
    const int rows = (int)min((long long)kIouRows, n - row0);
#pragma unroll 2
    for (int r = 0; r < rows; ++r) {
        const float4 ca = lo_hi_a[r];
        const float aa = area_a[r];
        float res[V];
#pragma unroll
        for (int c = 0; c < V; ++c) {
            const float tlx = fmax_nan(ca.x, cb[c].x), tly = fmax_nan(ca.y, cb[c].y);
            const float brx = fmin_nan(ca.z, cb[c].z), bry = fmin_nan(ca.w, cb[c].w);
            const float en = (tlx < brx && tly < bry) ? 1.0f : 0.0f;                            // :47
            const float inter = __fmul_rn(__fmul_rn(__fsub_rn(brx, tlx), __fsub_rn(bry, tly)), en);  // :48
            res[c] = iou_from_parts(inter, __fadd_rn(aa, area_b[c]));                           // :49
        }
        if (V == 4) *reinterpret_cast<float4*>(out + (row0 + r) * k + col) = make_float4(res[0], res[1], res[2], res[3 % V]);
        else out[(row0 + r) * k + col] = res[0];
    }
}

__global__ void __launch_bounds__(kIouCols) iou_rot_kernel(const float* __restrict__ a, long long n,
                                                          const float* __restrict__ b, long long k,
                                                          double* __restrict__ out) {
    __shared__ float ax[kIouRows][4], ay[kIouRows][4], acx[kIouRows], acy[kIouRows], ar[kIouRows];
    const long long col = (long long)blockIdx.x * kIouCols + threadIdx.x;
    const long long row0 = (long long)blockIdx.y * kIouRows;
    if (threadIdx.x < kIouRows && row0 + threadIdx.x < n) {
        const float* p = a + (row0 + threadIdx.x) * 5;
        float v[5] = {p[0], p[1], p[2], p[3], p[4]};
        float r;
        make_rot_box(v, ax[threadIdx.x], ay[threadIdx.x], r);
        acx[threadIdx.x] = v[0]; acy[threadIdx.x] = v[1]; ar[threadIdx.x] = r;
    }
    __syncthreads();
    if (col >= k) return;
    const float* p = b + col * 5;
    float v[5] = {p[0], p[1], p[2], p[3], p[4]};
    float bx[4], by[4], br;
    make_rot_box(v, bx, by, br);
    const int rows = (int)min((long long)kIouRows, n - row0);
    for (int r = 0; r < rows; ++r) {
        // same cull as the oracle: circumscribed circles apart => IoU is exactly 0
        const double dx = (double)acx[r] - (double)v[0], dy = (double)acy[r] - (double)v[1];
        const double rr = (double)ar[r] + (double)br + 1e-3;
        double iou = 0.0;
        if (dx * dx + dy * dy <= rr * rr) iou = rot_iou_f64(ax[r], ay[r], bx, by);
        out[(row0 + r) * k + col] = iou;
    }
}

// Many small rotated-IoU matrices in ONE launch: the matching IoU of a whole evaluation (CEPDOFeval.computeIoU is
// called once per (image, category), utils/evaluation/cepdof.py:67-99 -- thousands of tiny dt x gt problems).
// Segment s: rows a[seg[s].a0 .. +na), columns b[seg[s].b0 .. +nb), row-major output at out + seg[s].out0.
// One thread per output element; the segment of an element is found by binary search in the out0 prefix.
template <typename C>
__global__ void __launch_bounds__(256) iou_rot_segments_kernel(const C* __restrict__ a, const C* __restrict__ b,
                                                               const long long* __restrict__ seg, int n_seg,
                                                               long long total, double* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    int lo = 0, hi = n_seg - 1;                      // last segment whose out0 <= e (out0 is non-decreasing)
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (seg[(long long)mid * 5 + 4] <= e) lo = mid; else hi = mid - 1;
    }
    const long long* sg = seg + (long long)lo * 5;
    const long long nb = sg[3], local = e - sg[4];
    const long long r = local / nb, c = local - r * nb;
    const C* pa = a + (sg[0] + r) * 5;
    const C* pb = b + (sg[2] + c) * 5;
    C va[5] = {pa[0], pa[1], pa[2], pa[3], pa[4]}, vb[5] = {pb[0], pb[1], pb[2], pb[3], pb[4]};
    C ax[4], ay[4], bx[4], by[4], ra, rb;
    make_rot_box(va, ax, ay, ra);
    make_rot_box(vb, bx, by, rb);
    const double dx = (double)va[0] - (double)vb[0], dy = (double)va[1] - (double)vb[1];
    const double rr = (double)ra + (double)rb + 1e-3;
    out[e] = (dx * dx + dy * dy <= rr * rr) ? rot_iou_f64(ax, ay, bx, by) : 0.0;
}

// Row-wise max / arg-max of the bboxes_iou matrix, batched, WITHOUT materialising it: what every training
// branch does next with that matrix (`bboxes_iou(...).max(dim=1)`: yolov3.py:106-107 and :94-95, fcos2.py:104-106,
// retinanet.py:106-107).  One thread per row box; the image's GT boxes are staged in shared memory (corners
// and areas computed once).  Same float32 arithmetic as iou_aabb_kernel, torch.max semantics: first index of
// the maximum, NaN propagates (the first NaN wins).
constexpr int kRowmaxThreads = 128;
constexpr int kRowmaxStage = 512;          // GT boxes per shared-memory stage
__global__ void __launch_bounds__(kRowmaxThreads) iou_rowmax_kernel(const float* __restrict__ a, long long a_bs, long long a_pitch,
                                                                    long long n, const float* __restrict__ gt,
                                                                    const int* __restrict__ gt_count, int max_gt, int xyxy,
                                                                    float* __restrict__ out_max, long long* __restrict__ out_arg) {
    __shared__ float4 s_box[kRowmaxStage];
    __shared__ float s_area[kRowmaxStage];
    const int b = blockIdx.y;
    const long long row = (long long)blockIdx.x * kRowmaxThreads + threadIdx.x;
    int n_gt = max_gt;
    if (gt_count) n_gt = min(max(gt_count[b], 0), max_gt);
    float4 ca = make_float4(0.f, 0.f, 0.f, 0.f);
    float area_a = 0.f;
    if (row < n) {
        const float* p = a + b * a_bs + row * a_pitch;
        const float v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
        if (xyxy) {
            ca = make_float4(v0, v1, v2, v3);
            area_a = __fmul_rn(__fsub_rn(v2, v0), __fsub_rn(v3, v1));
        } else {
            const float hw = __fmul_rn(v2, 0.5f), hh = __fmul_rn(v3, 0.5f);
            ca = make_float4(__fsub_rn(v0, hw), __fsub_rn(v1, hh), __fadd_rn(v0, hw), __fadd_rn(v1, hh));
            area_a = __fmul_rn(v2, v3);
        }
    }
    float best = -1.0f;
    long long arg = -1;
    bool have = false, is_nan = false;
    for (int g0 = 0; g0 < n_gt; g0 += kRowmaxStage) {
        const int cnt = min(kRowmaxStage, n_gt - g0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += kRowmaxThreads) {
            const float4 v = reinterpret_cast<const float4*>(gt)[(long long)b * max_gt + g0 + i];
            if (xyxy) {
                s_box[i] = v;
                s_area[i] = __fmul_rn(__fsub_rn(v.z, v.x), __fsub_rn(v.w, v.y));
            } else {
                const float hw = __fmul_rn(v.z, 0.5f), hh = __fmul_rn(v.w, 0.5f);
                s_box[i] = make_float4(__fsub_rn(v.x, hw), __fsub_rn(v.y, hh), __fadd_rn(v.x, hw), __fadd_rn(v.y, hh));
                s_area[i] = __fmul_rn(v.z, v.w);
            }
        }
        __syncthreads();
        if (row < n && !is_nan) {
#pragma unroll 4
            for (int i = 0; i < cnt; ++i) {
                const float4 cb = s_box[i];
                const float tlx = fmax_nan(ca.x, cb.x), tly = fmax_nan(ca.y, cb.y);
                const float brx = fmin_nan(ca.z, cb.z), bry = fmin_nan(ca.w, cb.w);
                const float en = (tlx < brx && tly < bry) ? 1.0f : 0.0f;
                const float inter = __fmul_rn(__fmul_rn(__fsub_rn(brx, tlx), __fsub_rn(bry, tly)), en);
                const float iou = iou_from_parts(inter, __fadd_rn(area_a, s_area[i]));
                if (iou != iou) { if (!is_nan) { best = iou; arg = g0 + i; is_nan = true; } }
                else if (!is_nan && (!have || iou > best)) { best = iou; arg = g0 + i; have = true; }
            }
        }
    }
    if (row < n) {
        out_max[(long long)b * n + row] = best;
        if (out_arg) out_arg[(long long)b * n + row] = arg;
    }
}

__global__ void corners_kernel(const float* __restrict__ in, long long n, int n_param, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = in + i * n_param;
    float* o = out + i * n_param;
    const float cx = p[0], cy = p[1], hw = __fmul_rn(p[2], 0.5f), hh = __fmul_rn(p[3], 0.5f);
    o[0] = __fsub_rn(cx, hw); o[1] = __fsub_rn(cy, hh); o[2] = __fadd_rn(cx, hw); o[3] = __fadd_rn(cy, hh);
    for (int k = 4; k < n_param; ++k) o[k] = p[k];
}

// xywha2vertex, bbox_ops.py:137-172 (angle already in radians; float32 sinf/cosf like torch.sin/cos)
__global__ void vertex_kernel(const float* __restrict__ in, long long n, int n_param, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = in + i * n_param;
    const float s = (float)sin((double)p[4]), c = (float)cos((double)p[4]);
    const float hh = __fmul_rn(p[3], 0.5f), hw = __fmul_rn(p[2], 0.5f);
    const float vx = __fmul_rn(hh, s), vy = -__fmul_rn(hh, c), hx = __fmul_rn(hw, c), hy = __fmul_rn(hw, s);
    float* o = out + i * 8;
    o[0] = __fsub_rn(__fadd_rn(p[0], vx), hx); o[1] = __fsub_rn(__fadd_rn(p[1], vy), hy);
    o[2] = __fadd_rn(__fadd_rn(p[0], vx), hx); o[3] = __fadd_rn(__fadd_rn(p[1], vy), hy);
    o[4] = __fadd_rn(__fsub_rn(p[0], vx), hx); o[5] = __fadd_rn(__fsub_rn(p[1], vy), hy);
    o[6] = __fsub_rn(__fsub_rn(p[0], vx), hx); o[7] = __fsub_rn(__fsub_rn(p[1], vy), hy);
}

}  // namespace mydet

using namespace mydet;

MYDET_API int mydet_cxcywh_to_x1y1x2y2(const float* in, int64_t n, int n_param, float* out, void* stream) {
    MYDET_REQUIRE(n >= 0 && n_param >= 4, "need n >= 0 and at least 4 columns");
    if (n == 0) return 0;
    MYDET_REQUIRE(in && out, "NULL tensor pointer");
    corners_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, n, n_param, out);
    return launch_status("corners_kernel");
}

MYDET_API int mydet_xywha2vertex(const float* in, int64_t n, int n_param, float* out, void* stream) {
    MYDET_REQUIRE(n >= 0 && n_param >= 5, "need n >= 0 and at least 5 columns");
    if (n == 0) return 0;
    MYDET_REQUIRE(in && out, "NULL tensor pointer");
    vertex_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, n, n_param, out);
    return launch_status("vertex_kernel");
}

MYDET_API int mydet_iou_aabb_pairwise(const float* a, int64_t n, const float* b, int64_t k, int xyxy, float* out,
                                      void* stream) {
    MYDET_REQUIRE(n >= 0 && k >= 0, "negative size");
    if (n == 0 || k == 0) return 0;
    MYDET_REQUIRE(a && b && out, "NULL tensor pointer");
    MYDET_REQUIRE(((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0, "box arrays must be 16-byte aligned");
    const bool vec = (k % 4 == 0) && ((uintptr_t)out & 15) == 0;
    const long long per_cta = (long long)kIouCols * (vec ? 4 : 1);
    const dim3 grid((unsigned)((k + per_cta - 1) / per_cta), (unsigned)((n + kIouRows - 1) / kIouRows));
    MYDET_REQUIRE(grid.y <= 65535, "too many rows for one launch");
    if (vec) iou_aabb_kernel<4><<<grid, kIouCols, 0, (cudaStream_t)stream>>>(a, n, b, k, xyxy, out);
    else iou_aabb_kernel<1><<<grid, kIouCols, 0, (cudaStream_t)stream>>>(a, n, b, k, xyxy, out);
    return launch_status("iou_aabb_kernel");
}

MYDET_API int mydet_iou_rot_pairwise(const float* a, int64_t n, const float* b, int64_t k, double* out, void* stream) {
    MYDET_REQUIRE(n >= 0 && k >= 0, "negative size");
    if (n == 0 || k == 0) return 0;
    MYDET_REQUIRE(a && b && out, "NULL tensor pointer");
    const dim3 grid((unsigned)((k + kIouCols - 1) / kIouCols), (unsigned)((n + kIouRows - 1) / kIouRows));
    MYDET_REQUIRE(grid.y <= 65535, "too many rows for one launch");
    iou_rot_kernel<<<grid, kIouCols, 0, (cudaStream_t)stream>>>(a, n, b, k, out);
    return launch_status("iou_rot_kernel");
}

MYDET_API int mydet_iou_rot_segments(const void* a, const void* b, int boxes_are_f64, const int64_t* segments, int n_segments,
                                     int64_t total, double* out, void* stream) {
    MYDET_REQUIRE(n_segments >= 0 && total >= 0, "negative size");
    if (n_segments == 0 || total == 0) return 0;
    MYDET_REQUIRE(a && b && segments && out, "NULL tensor pointer");
    MYDET_REQUIRE(total <= 0x7fffffffLL * 256, "too many pairs for one launch");
    const unsigned grid = (unsigned)((total + 255) / 256);
    const long long* sg = reinterpret_cast<const long long*>(segments);
    if (boxes_are_f64)
        iou_rot_segments_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const double*>(a), static_cast<const double*>(b),
                                                                                sg, n_segments, total, out);
    else
        iou_rot_segments_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const float*>(a), static_cast<const float*>(b),
                                                                               sg, n_segments, total, out);
    return launch_status("iou_rot_segments_kernel");
}

MYDET_API int mydet_iou_aabb_rowmax(const float* a, int64_t a_batch_stride, int64_t a_pitch, int64_t n, const float* gt,
                                    const int32_t* gt_count, int max_gt, int batch, int xyxy, float* out_max,
                                    int64_t* out_arg, void* stream) {
    MYDET_REQUIRE(n >= 0 && batch >= 0 && max_gt >= 0 && a_pitch >= 4 && a_batch_stride >= 0, "bad sizes / strides");
    if (n == 0 || batch == 0) return 0;
    MYDET_REQUIRE(a && out_max && (max_gt == 0 || gt), "NULL tensor pointer");
    MYDET_REQUIRE(((uintptr_t)gt & 15) == 0, "GT boxes must be 16-byte aligned");
    MYDET_REQUIRE(batch <= 65535, "batch too large for one launch");
    const dim3 grid((unsigned)((n + kRowmaxThreads - 1) / kRowmaxThreads), (unsigned)batch);
    iou_rowmax_kernel<<<grid, kRowmaxThreads, 0, (cudaStream_t)stream>>>(a, a_batch_stride, a_pitch, n, gt, gt_count, max_gt, xyxy,
                                                                      out_max, reinterpret_cast<long long*>(out_arg));
    return launch_status("iou_rowmax_kernel");
}
