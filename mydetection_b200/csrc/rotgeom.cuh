// Rotated-rectangle geometry shared by the rotated NMS mask kernel and the pairwise IoU kernel.
//
// Stands in for iou_rle (utils/bbox_ops.py:52-100).  The reference rasterises the two polygons
// with pycocotools, which is unavailable (DESIGN.md "Oracle": parity unpinned); here the area of
// the intersection of the SAME two polygons is computed exactly by Sutherland-Hodgman clipping.
// The polygons are built exactly as the reference builds them (bbox_ops.py:88-89, :137-172):
// float32 corners  c +/- verti -/+ hori  with verti=(h/2)(sin,-cos), hori=(w/2)(cos,sin).
#pragma once
#include "common.cuh"

namespace mydet {

struct __align__(16) RotBox {       // 16 floats, one per box, precomputed once
    float x[4], y[4]; // corners tl,tr,br,bl in float32, as the reference computes them
    float cx, cy;     // centre
    float r;          // circumscribed radius (cull)
    float area2;      // signed 2*area of the corner polygon (shoelace), float64-rounded to float32
    float x0, y0, x1, y1;   // axis-aligned hull of the corners (cull)
};
__device__ __forceinline__ void rot_box_hull(RotBox& q) {
    q.x0 = fminf(fminf(q.x[0], q.x[1]), fminf(q.x[2], q.x[3]));
    q.x1 = fmaxf(fmaxf(q.x[0], q.x[1]), fmaxf(q.x[2], q.x[3]));
    q.y0 = fminf(fminf(q.y[0], q.y[1]), fminf(q.y[2], q.y[3]));
    q.y1 = fmaxf(fmaxf(q.y[0], q.y[1]), fmaxf(q.y[2], q.y[3]));
}

__device__ __forceinline__ void make_rot_box(const float* b, float* x, float* y, float& r) {
    // degrees -> radians: a * pi / 180 in float32 (bbox_ops.py:88-89).  sin/cos are evaluated in
    // float64 and rounded once, which reproduces a correctly rounded float32 sinf/cosf.
    const float rad = __fdiv_rn(__fmul_rn(b[4], 3.14159265358979323846f), 180.0f);
    const float s = (float)sin((double)rad), c = (float)cos((double)rad);
    const float hh = __fmul_rn(b[3], 0.5f), hw = __fmul_rn(b[2], 0.5f);
    const float vx = __fmul_rn(hh, s), vy = -__fmul_rn(hh, c);
    const float hx = __fmul_rn(hw, c), hy = __fmul_rn(hw, s);
    x[0] = __fsub_rn(__fadd_rn(b[0], vx), hx); y[0] = __fsub_rn(__fadd_rn(b[1], vy), hy);
    x[1] = __fadd_rn(__fadd_rn(b[0], vx), hx); y[1] = __fadd_rn(__fadd_rn(b[1], vy), hy);
    x[2] = __fadd_rn(__fsub_rn(b[0], vx), hx); y[2] = __fadd_rn(__fsub_rn(b[1], vy), hy);
    x[3] = __fsub_rn(__fsub_rn(b[0], vx), hx); y[3] = __fsub_rn(__fsub_rn(b[1], vy), hy);
    r = 0.5f * sqrtf(b[2] * b[2] + b[3] * b[3]);
}

// The evaluator's variant (utils/evaluation/cepdof.py:181-207, :232-236): boxes are Python floats, the corners are built
// with numpy in float64 -- degree * pi / 180, then c +/- verti -/+ hori in double.
__device__ __forceinline__ void make_rot_box(const double* b, double* x, double* y, double& r) {
    const double rad = b[4] * 3.14159265358979323846 / 180.0;
    const double s = sin(rad), c = cos(rad);
    const double vx = (b[3] / 2) * s, vy = -(b[3] / 2) * c, hx = (b[2] / 2) * c, hy = (b[2] / 2) * s;
    x[0] = b[0] + vx - hx; y[0] = b[1] + vy - hy;
    x[1] = b[0] + vx + hx; y[1] = b[1] + vy + hy;
    x[2] = b[0] - vx + hx; y[2] = b[1] - vy + hy;
    x[3] = b[0] - vx - hx; y[3] = b[1] - vy - hy;
    r = 0.5 * sqrt(b[2] * b[2] + b[3] * b[3]);
}

template <typename C>
__device__ __forceinline__ double signed_area2_f64(const C* x, const C* y) {
    double a2 = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int n = (k + 1) & 3;
        a2 += (double)x[k] * (double)y[n] - (double)x[n] * (double)y[k];
    }
    return a2;
}

// Intersection area of convex quads A and B (B's orientation given by sgn), arithmetic in T.
// (ox, oy) is subtracted from every corner first: exact in float32 for nearby boxes, and it keeps
// the float32 variant's rounding error relative to the box size instead of the image size.
template <typename T, typename C = float>
__device__ __forceinline__ T clip_area(const C* ax, const C* ay, const C* bx, const C* by,
                                       T sgn, C ox, C oy) {
    T px[8], py[8], qx[8], qy[8];
    int n = 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) { px[k] = (T)(ax[k] - ox); py[k] = (T)(ay[k] - oy); }
#pragma unroll 1
    for (int e = 0; e < 4 && n > 0; ++e) {
        const T ex0 = (T)(bx[e] - ox), ey0 = (T)(by[e] - oy);
        const T ex = (T)(bx[(e + 1) & 3] - ox) - ex0, ey = (T)(by[(e + 1) & 3] - oy) - ey0;
        int m = 0;
        T dp = sgn * (ex * (py[0] - ey0) - ey * (px[0] - ex0));
#pragma unroll 1
        for (int k = 0; k < n; ++k) {
            const int k2 = (k + 1 == n) ? 0 : k + 1;
            const T dq = sgn * (ex * (py[k2] - ey0) - ey * (px[k2] - ex0));
            // a convex quad clipped by 4 half planes has <= 8 vertices in exact arithmetic; rounding can make an
            // intermediate polygon non-convex (near-collinear vertices) and produce a 9th: it is dropped, never
            // written past the arrays (oracle/rotiou.c applies the same rule)
            if (dp >= (T)0 && m < 8) { qx[m] = px[k]; qy[m] = py[k]; ++m; }
            if ((dp >= (T)0) != (dq >= (T)0) && m < 8) {
                const T t = dp / (dp - dq);
                qx[m] = px[k] + t * (px[k2] - px[k]);
                qy[m] = py[k] + t * (py[k2] - py[k]);
                ++m;
            }
            dp = dq;
        }
        n = m;
#pragma unroll 1
        for (int k = 0; k < n; ++k) { px[k] = qx[k]; py[k] = qy[k]; }
    }
    if (n < 3) return (T)0;
    T a2 = (T)0;
#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        const int k2 = (k + 1 == n) ? 0 : k + 1;
        a2 += px[k] * py[k2] - px[k2] * py[k];
    }
    return (T)0.5 * (a2 < (T)0 ? -a2 : a2);
}

// Exact (float64) IoU, same operation order as oracle/rotiou.c::quad_iou.
template <typename C>
__device__ __forceinline__ double rot_iou_f64(const C* ax, const C* ay, const C* bx, const C* by) {
    const double a2A = signed_area2_f64(ax, ay), a2B = signed_area2_f64(bx, by);
    const double areaA = 0.5 * fabs(a2A), areaB = 0.5 * fabs(a2B);
    double inter = 0.0;
    if (areaA > 0.0 && areaB > 0.0) inter = clip_area<double, C>(ax, ay, bx, by, a2B >= 0.0 ? 1.0 : -1.0, (C)0, (C)0);
    const double uni = areaA + areaB - inter;
    return (inter > 0.0 && uni > 0.0) ? inter / uni : 0.0;      // a zero numerator would take the divide's slow path for the same 0
}

// Decision "IoU >= thr" (ge) or "IoU > thr": float32 clipping in box-local coordinates, and an
// exact float64 re-evaluation whenever the float32 value is within 1e-3 of the threshold.
__device__ __forceinline__ bool rot_overlaps(const RotBox& A, const RotBox& B, double thr, bool ge) {
    // (1) circumscribed circles apart => empty intersection (slack covers float32 rounding)
    const float dx = A.cx - B.cx, dy = A.cy - B.cy;
    const float rr = (A.r + B.r) * 1.00001f + 1e-3f;
    if (dx * dx + dy * dy > rr * rr) return ge ? (0.0 >= thr) : (0.0 > thr);
    // (2) IoU <= min(area)/max(area)
    const float aA = 0.5f * fabsf(A.area2), aB = 0.5f * fabsf(B.area2);
    const float lo = fminf(aA, aB), hi = fmaxf(aA, aB);
    if (!(lo > 0.f)) return ge ? (0.0 >= thr) : (0.0 > thr);
    if ((double)lo * 1.0001 < thr * (double)hi) return false;
    // (3) float32 clip around A's centre
    const float inter = clip_area<float>(A.x, A.y, B.x, B.y, B.area2 >= 0.f ? 1.f : -1.f, A.cx, A.cy);
    const float uni = aA + aB - inter;
    const float iou = (inter > 0.f && uni > 0.f) ? inter / uni : 0.f;
    const float d = iou - (float)thr;
    if (fabsf(d) < 1e-3f) {
        const double e = rot_iou_f64(A.x, A.y, B.x, B.y);
        return ge ? (e >= thr) : (e > thr);
    }
    return d > 0.f;
}

}  // namespace mydet
