// Host harness for the rotated-NMS pair decision (TEST INFRASTRUCTURE, not part of libmydet).
//
// Compiles mydetection_b200/csrc/rotgeom.cuh -- make_rot_box, clip_area<float/double>, rot_iou_f64, rot_overlaps:
// the device functions the rotated mask kernel and the pairwise IoU kernel call -- with the host compiler (the
// CUDA round-to-nearest intrinsics become plain float operations, -ffp-contract=off), and restates the per-pair cull
// chain of mask_rot_spatial_kernel (nms_large.cu: circle test + area-ratio bound, hull bound) in front of it, so
// that millions of adversarial pairs can be compared with the oracle's float64 IoU on a machine without a GPU.
//
//   rotgeom_host <pairs.bin> <out.bin> n thr ge
// pairs.bin: n x 10 float32 (box A, box B as cx,cy,w,h,deg); out.bin: n x {int32 decision, int32 stage, float iou32, double iou64}
// stage: 0 = culled by circle/area, 1 = culled by the hull bound, 2 = decided by the float32 clip, 3 = float64 re-check.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <cuda_runtime.h>
#undef __device__
#undef __forceinline__
#define __device__
#define __forceinline__ inline
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
#include "../../mydetection_b200/csrc/rotgeom.cuh"

using namespace mydet;

static RotBox make(const float* v) {                    // spatial_gather_kernel<true>
    RotBox q;
    make_rot_box(v, q.x, q.y, q.r);
    q.cx = v[0]; q.cy = v[1];
    q.area2 = (float)signed_area2_f64(q.x, q.y);
    rot_box_hull(q);
    return q;
}

#pragma pack(push, 1)
struct Out { int32_t decision, stage; float iou32; double iou64; };
#pragma pack(pop)

int main(int argc, char** argv) {
    if (argc != 6) return 2;
    const long n = atol(argv[3]);
    const double thr_d = atof(argv[4]);
    const bool ge = atoi(argv[5]) != 0;
    const float thr_f = (float)thr_d;
    float* in = (float*)malloc(sizeof(float) * 10 * (size_t)n);
    Out* out = (Out*)malloc(sizeof(Out) * (size_t)n);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(in, sizeof(float) * 10, n, f) != (size_t)n) return 3;
    fclose(f);
    for (long i = 0; i < n; ++i) {
        const RotBox A = make(in + 10 * i), B = make(in + 10 * i + 5);
        Out o; o.decision = 0; o.stage = 0;
        o.iou64 = rot_iou_f64(A.x, A.y, B.x, B.y);
        {   // what rot_overlaps' float32 stage sees (reported for the error statistics)
            const float aA = 0.5f * fabsf(A.area2), aB = 0.5f * fabsf(B.area2);
            const float inter = clip_area<float>(A.x, A.y, B.x, B.y, B.area2 >= 0.f ? 1.f : -1.f, A.cx, A.cy);
            const float uni = aA + aB - inter;
            o.iou32 = uni > 0.f ? inter / uni : 0.f;
        }
        // ---- mask_rot_spatial_kernel: "my" box A (row), candidate B (column)
        const float mcx = A.cx, mcy = A.cy, mr = A.r * 1.00001f + 1e-3f, ma = 0.5f * fabsf(A.area2);
        const float oa = 0.5f * fabsf(B.area2);
        const float dx = mcx - B.cx, dy = mcy - B.cy, rr = fmaf(B.r, 1.00001f, mr);
        const bool pass = (fmaf(dx, dx, dy * dy) <= rr * rr) & (fminf(ma, oa) * 1.0001f >= thr_f * fmaxf(ma, oa));
        if (pass) {
            o.stage = 1;
            const float ix = fminf(A.x1, B.x1) - fmaxf(A.x0, B.x0);
            const float iy = fminf(A.y1, B.y1) - fmaxf(A.y0, B.y0);
            bool alive = (ix > -1e-3f && iy > -1e-3f);
            if (alive) {
                const float ub = (ix + 2e-3f) * (iy + 2e-3f);
                if (ub * 1.0001f < thr_f * (ma + oa - ub)) alive = false;
            }
            if (alive) {
                o.decision = rot_overlaps(A, B, thr_d, ge) ? 1 : 0;
                o.stage = fabsf(o.iou32 - (float)thr_d) < 1e-3f ? 3 : 2;
            }
        }
        out[i] = o;
    }
    f = fopen(argv[2], "wb");
    if (!f || fwrite(out, sizeof(Out), n, f) != (size_t)n) return 4;
    fclose(f);
    return 0;
}
