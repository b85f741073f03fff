// Rasterised rotated IoU: `iou_rle` (utils/bbox_ops.py:52-100) the way the reference computes it -- the two polygons
// are turned into binary masks on an img_hw canvas by pycocotools (maskApi.c: rleFrPoly) and the IoU is a pixel count
// (rleIou, iscrowd = 0) -- instead of the exact polygon intersection of mydet_iou_rot_pairwise.
//
// pycocotools is not in this image; the algorithm is restated from its published source in oracle/raster.c, which is
// pinned to five run-length encodings derived by hand (tests/test_oracle_golden.py::test_raster_known_answers) and which
// this file reproduces bit for bit.  What rleFrPoly does, in closed form for a convex quad:
//   * corners are rounded onto a 5x finer integer grid, X = (int)(5 x + .5);
//   * every edge is walked point by point along its major axis, the minor coordinate being (int)(start + slope t + .5);
//   * the mask changes value where the walk steps over the line between sub-columns 5 c + 2 and 5 c + 3 of pixel column c,
//     at row ceil((v + .5) / 5 - .5) for the smaller sub-row v of the two points, clamped to [0, h];
//   * a closed convex walk crosses each such line exactly twice (or not at all): pixel column c of the mask is the run
//     [y_lo, y_hi).
// So instead of walking ~5 perimeter points per pixel and sorting them, kernel 1 computes for every box and every pixel
// column the two crossings directly (one warp per box, lanes over columns; an x-major edge gives the crossing of a column
// in O(1), a y-major edge after a <= 2-step search for the step where its rounded x moves across the line), and kernel 2
// intersects the per-column runs of each pair.
#include "internal.cuh"
#include "rotgeom.cuh"

namespace mydet {

struct RasterHead { int x0, ncols, area, flags; };            // flags bit 0: a column with more than two crossings (never for a convex quad)

__device__ __forceinline__ int raster_row(int v, int h) {     // ceil of ((v + .5) / 5 - .5) clamped to [0, h]
    double yd = ((double)v + .5) / 5.0 - .5;
    if (yd < 0) yd = 0; else if (yd > (double)h) yd = (double)h;
    return (int)ceil(yd);
}

// crossing of the line between sub-columns tgt = 5 c + 2 and tgt + 1 by the walk of one edge; returns false if none
__device__ bool edge_crossing(int xs, int ys, int xe, int ye, int tgt, int h, int& row) {
    const int dx = abs(xe - xs), dy = abs(ys - ye);
    const bool flip = (dx >= dy && xs > xe) || (dx < dy && ys > ye);
    if (flip) { int t = xs; xs = xe; xe = t; t = ys; ys = ye; ye = t; }
    if (dx >= dy) {
        if (dx == 0 || tgt < xs || tgt + 1 > xe) return false;
        const double s = (double)(ye - ys) / dx;
        const int v0 = (int)(ys + s * (tgt - xs) + .5), v1 = (int)(ys + s * (tgt + 1 - xs) + .5);
        row = raster_row(min(v0, v1), h);
        return true;
    }
    const double s = (double)(xe - xs) / dy;
    const int ua = (int)(xs + s * 0 + .5), ub = (int)(xs + s * dy + .5);
    if (tgt < min(ua, ub) || tgt + 1 > max(ua, ub)) return false;
    // the rounded x is monotone in t (|s| < 1): find the step t -> t + 1 that moves it across the line
    int t = (int)floor(((double)tgt + .5 - xs) / s);
    t = max(0, min(dy - 1, t));
    if (s > 0) {
        while (t > 0 && (int)(xs + s * t + .5) > tgt) --t;
        while (t < dy - 1 && (int)(xs + s * (t + 1) + .5) <= tgt) ++t;
    } else {
        while (t > 0 && (int)(xs + s * t + .5) < tgt + 1) --t;
        while (t < dy - 1 && (int)(xs + s * (t + 1) + .5) >= tgt + 1) ++t;
    }
    row = raster_row(ys + t, h);
    return true;
}

// one warp per box: header + one (lo, hi) run per pixel column
__global__ void __launch_bounds__(256) raster_spans_kernel(const float* __restrict__ boxes, long long n, int h, int w, int pitch,
                                                           RasterHead* __restrict__ head, ushort2* __restrict__ spans) {
    const long long i = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const float* p = boxes + i * 5;
    float v[5] = {p[0], p[1], p[2], p[3], p[4]}, fx[4], fy[4], r;
    make_rot_box(v, fx, fy, r);                                  // float32 corners, as xywha2vertex(...).tolist() hands them over
    int X[5], Y[5];
#pragma unroll
    for (int k = 0; k < 4; ++k) { X[k] = (int)(5.0 * (double)fx[k] + .5); Y[k] = (int)(5.0 * (double)fy[k] + .5); }
    X[4] = X[0]; Y[4] = Y[0];
    const int umin = min(min(X[0], X[1]), min(X[2], X[3])), umax = max(max(X[0], X[1]), max(X[2], X[3]));
    // columns whose line 5 c + 2 | 5 c + 3 lies inside [umin, umax], on the canvas
    int c_lo = (umin - 2 >= 0) ? (umin - 2 + 4) / 5 : -((2 - umin) / 5);
    int c_hi = (umax - 3 >= 0) ? (umax - 3) / 5 : -((3 - umax + 4) / 5);
    c_lo = max(c_lo, 0); c_hi = min(c_hi, w - 1);
    const int ncols = max(0, min(c_hi - c_lo + 1, pitch));
    int area = 0, flags = 0;
    ushort2* my = spans + i * pitch;
    for (int c = c_lo + lane; c < c_lo + ncols; c += 32) {
        int rows[4], m = 0, row;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (edge_crossing(X[e], Y[e], X[e + 1], Y[e + 1], 5 * c + 2, h, row)) { if (m < 4) rows[m] = row; ++m; }
        int lo = 0, hi = 0;
        if (m >= 2) {
            lo = min(rows[0], rows[1]); hi = max(rows[0], rows[1]);
            if (m > 2) flags = 1;
        } else if (m == 1) flags = 1;
        my[c - c_lo] = make_ushort2((unsigned short)lo, (unsigned short)hi);
        area += hi - lo;
    }
    area = __reduce_add_sync(0xffffffffu, area);
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0) head[i] = RasterHead{c_lo, ncols, area, flags};
}

constexpr int kRasCols = 128, kRasRows = 32;
__global__ void __launch_bounds__(kRasCols) raster_iou_kernel(const RasterHead* __restrict__ ha, const ushort2* __restrict__ sa, long long n,
                                                              const RasterHead* __restrict__ hb, const ushort2* __restrict__ sb, long long k,
                                                              int pitch, double* __restrict__ out) {
    __shared__ RasterHead s_head[kRasRows];
    const long long col = (long long)blockIdx.x * kRasCols + threadIdx.x;
    const long long row0 = (long long)blockIdx.y * kRasRows;
    if (threadIdx.x < kRasRows && row0 + threadIdx.x < n) s_head[threadIdx.x] = ha[row0 + threadIdx.x];
    __syncthreads();
    if (col >= k) return;
    const RasterHead B = hb[col];
    const ushort2* pb = sb + col * pitch;
    const int rows = (int)min((long long)kRasRows, n - row0);
    for (int r = 0; r < rows; ++r) {
        const RasterHead A = s_head[r];
        const int c0 = max(A.x0, B.x0), c1 = min(A.x0 + A.ncols, B.x0 + B.ncols);
        long long inter = 0;
        const ushort2* pa = sa + (row0 + r) * pitch;
        for (int c = c0; c < c1; ++c) {
            const ushort2 a = pa[c - A.x0], b = pb[c - B.x0];
            const int lo = max((int)a.x, (int)b.x), hi = min((int)a.y, (int)b.y);
            inter += max(0, hi - lo);
        }
        // rleIou: i / u, and 0 when the masks do not meet (maskApi.c: `if(i==0) u=1`)
        double iou = inter > 0 ? (double)inter / (double)((long long)A.area + B.area - inter) : 0.0;
        if ((A.flags | B.flags) & 1) iou = __longlong_as_double(0x7ff8000000000000LL);    // not a convex walk: say so loudly
        out[(row0 + r) * k + col] = iou;
    }
}

static size_t raster_carve(long long n, long long k, int pitch, char* base, RasterHead** ha, ushort2** sa, RasterHead** hb, ushort2** sb) {
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_ha = take(sizeof(RasterHead) * (size_t)n), o_hb = take(sizeof(RasterHead) * (size_t)k);
    const size_t o_sa = take(sizeof(ushort2) * (size_t)n * pitch), o_sb = take(sizeof(ushort2) * (size_t)k * pitch);
    if (base) {
        *ha = reinterpret_cast<RasterHead*>(base + o_ha); *hb = reinterpret_cast<RasterHead*>(base + o_hb);
        *sa = reinterpret_cast<ushort2*>(base + o_sa); *sb = reinterpret_cast<ushort2*>(base + o_sb);
    }
    return off;
}

}  // namespace mydet

using namespace mydet;

MYDET_API size_t mydet_iou_raster_workspace_bytes(int64_t n, int64_t k, int canvas_w) {
    if (n < 0 || k < 0 || canvas_w <= 0) return 0;
    return raster_carve(n, k, canvas_w, nullptr, nullptr, nullptr, nullptr, nullptr);
}

MYDET_API int mydet_iou_raster_pairwise(const float* a, int64_t n, const float* b, int64_t k, int canvas_h, int canvas_w,
                                        double* out, void* workspace, size_t workspace_bytes, void* stream) {
    MYDET_REQUIRE(n >= 0 && k >= 0, "negative size");
    MYDET_REQUIRE(canvas_h > 0 && canvas_w > 0 && canvas_h <= 65535 && canvas_w <= 65535, "canvas must be within 1..65535 pixels");
    if (n == 0 || k == 0) return 0;
    MYDET_REQUIRE(a && b && out && workspace, "NULL tensor pointer");
    MYDET_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
    if (workspace_bytes < mydet_iou_raster_workspace_bytes(n, k, canvas_w)) {
        set_error("raster IoU workspace too small: %zu < %zu bytes", workspace_bytes, mydet_iou_raster_workspace_bytes(n, k, canvas_w));
        return MYDET_ERR_WORKSPACE;
    }
    RasterHead *ha, *hb;
    ushort2 *sa, *sb;
    raster_carve(n, k, canvas_w, static_cast<char*>(workspace), &ha, &sa, &hb, &sb);
    cudaStream_t st = (cudaStream_t)stream;
    raster_spans_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(a, n, canvas_h, canvas_w, canvas_w, ha, sa);
    raster_spans_kernel<<<(unsigned)((k + 7) / 8), 256, 0, st>>>(b, k, canvas_h, canvas_w, canvas_w, hb, sb);
    const dim3 grid((unsigned)((k + kRasCols - 1) / kRasCols), (unsigned)((n + kRasRows - 1) / kRasRows));
    MYDET_REQUIRE(grid.y <= 65535, "too many rows for one launch");
    raster_iou_kernel<<<grid, kRasCols, 0, st>>>(ha, sa, n, hb, sb, k, canvas_w, out);
    return launch_status("raster_iou_kernel");
}
