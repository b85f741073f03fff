// Shared device/host helpers for libmydet (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/mydet.h"

#define MYDET_API extern "C" __attribute__((visibility("default")))

namespace mydet {

void set_error(const char* fmt, ...);

#define MYDET_REQUIRE(cond, ...)                       \
    do {                                               \
        if (!(cond)) {                                 \
            ::mydet::set_error(__VA_ARGS__);           \
            return MYDET_ERR_INVALID;                  \
        }                                              \
    } while (0)

#define MYDET_CUDA(expr)                                                              \
    do {                                                                              \
        cudaError_t e__ = (expr);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            ::mydet::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));      \
            return (int)e__;                                                          \
        }                                                                             \
    } while (0)

static inline int launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kNumSMs = 148;  // B200

// ------------------------------------------------------------------------------ device helpers
#ifdef __CUDACC__

// Streaming 128-bit load: read-only path, do not allocate in L1 (every logit is read once).
__device__ __forceinline__ float4 ld_stream_v4(const float* p) {
    float4 r;
#ifdef MYDET_DECODE_EVICT_FIRST
    // the logits are read exactly once: mark their L2 lines evict-first, so that the candidates the decode writes (and
    // the post-process reads back a few microseconds later) are what stays in the 126 MB L2
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(pol));
#else
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
#endif
    return r;
}
__device__ __forceinline__ float ld_stream(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// sigmoid exactly as 1/(1+exp(-x)) with IEEE division: torch's CPU kernel computes the same
// expression; only the exp implementation differs (<= 2 ulp), see DESIGN.md "tolerances".
__device__ __forceinline__ float sigmoid_f(float x) {
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

// Order-preserving map float -> uint32 (larger float => larger key); -0 is folded onto +0.
__device__ __forceinline__ uint32_t float_key(float s) {
    uint32_t u = __float_as_uint(s + 0.0f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// ---- system-scope flag accesses and NVLS multicast stores of the fused detections exchange (exchange.cu, stage E of
// the post-process kernel).  A multimem.st is a store to the multicast mapping of a buffer: the NVSwitch replicates it
// into the copy of every GPU bound to the multicast object.
__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void multimem_st_v4(float4* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st_u32(unsigned* p, unsigned v) {
    asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void multimem_st_release_u32(unsigned* p, unsigned v) {
    asm volatile("multimem.st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
constexpr long long kExchangeSpinCycles = 2000000000LL;   // ~1 s at 1.9 GHz: a flag wait gives up after this, it never hangs

// torch.max / torch.min (and Tensor.max(dim)) propagate NaN; fmaxf / fminf return the other operand.  One FMNMX either way.
__device__ __forceinline__ float fmax_nan(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmin_nan(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// torchvision-compatible float32 IoU of two corner boxes with precomputed areas:
// inter = max(0,xx2-xx1)*max(0,yy2-yy1); ovr = inter/(area_i+area_j-inter).  No FMA contraction.
__device__ __forceinline__ float iou_corners(float ax1, float ay1, float ax2, float ay2, float aarea,
                                             float bx1, float by1, float bx2, float by2, float barea) {
    float xx1 = fmaxf(ax1, bx1), yy1 = fmaxf(ay1, by1);
    float xx2 = fminf(ax2, bx2), yy2 = fminf(ay2, by2);
    float w = fmaxf(0.0f, __fsub_rn(xx2, xx1));
    float h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
    float inter = __fmul_rn(w, h);
    float uni = __fsub_rn(__fadd_rn(aarea, barea), inter);
    return __fdiv_rn(inter, uni);
}

// inter / (area_a + area_b - inter), bboxes_iou :49, without the divide where its result is known: a zero numerator
// (+-0: a disjoint pair) over area_a + area_b > 0 (also +inf) gives that same signed zero.  This is not a micro-optimisation:
// a zero numerator sends the IEEE divide down its slow path (a subroutine of ~50 instructions), and anchor-vs-GT
// matrices are almost entirely disjoint pairs -- 16 384 x 8 192: 284 -> 160 us with a third of the pairs disjoint,
// 282 -> 118 us with nearly all of them (scripts/iou_variants.py).  The test is per lane: a warp-uniform vote that
// lets all lanes divide when one must runs the slow path for the whole warp again (297 us).
__device__ __forceinline__ float iou_from_parts(float inter, float area_sum) {
    if (inter == 0.0f && area_sum > 0.0f) return inter;
    return __fdiv_rn(inter, __fsub_rn(area_sum, inter));
}

// iou_corners(...) > thr without the divide where the answer is known: inter == 0 gives 0, -0 or NaN, none of which is
// above a non-negative threshold.  (Not a micro-optimisation: a zero numerator sends the IEEE divide down its slow path,
// a subroutine of ~50 instructions, and in NMS nearly every pair is disjoint.)
__device__ __forceinline__ bool iou_corners_gt(float ax1, float ay1, float ax2, float ay2, float aarea,
                                               float bx1, float by1, float bx2, float by2, float barea, float thr) {
    float xx1 = fmaxf(ax1, bx1), yy1 = fmaxf(ay1, by1);
    float xx2 = fminf(ax2, bx2), yy2 = fminf(ay2, by2);
    float w = fmaxf(0.0f, __fsub_rn(xx2, xx1));
    float h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
    float inter = __fmul_rn(w, h);
    if (inter == 0.0f && thr >= 0.0f) return false;
    float uni = __fsub_rn(__fadd_rn(aarea, barea), inter);
    return __fdiv_rn(inter, uni) > thr;
}

#endif  // __CUDACC__

// Largest float f with (double)f <= t: then  ((double)x > t)  <=>  (x > f)  for every float x,
// so the kernels can do torchvision's float-vs-double comparison with one FSETP.
static inline float float_at_or_below(double t) {
    float f = (float)t;
    if ((double)f > t) f = nextafterf(f, -INFINITY);
    return f;
}
// Smallest float f with (double)f >= t:  ((double)x >= t) <=> (x >= f).
static inline float float_at_or_above(double t) {
    float f = (float)t;
    if ((double)f < t) f = nextafterf(f, INFINITY);
    return f;
}

}  // namespace mydet
