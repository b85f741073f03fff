// Per-element arithmetic of the image pre-processing kernels (preprocess.cu), written as host/device inline
// functions of ONE output element each, so that the same source can be compiled by the host compiler in a test
// harness (tests/host_harness/preprocess_host.cpp) and compared bit for bit with Pillow / the reference on a
// machine without a GPU.  The kernels in preprocess.cu only map a thread index to an element and call these.
//
// What is restated (reference: api/detection.py:158-162, :177-205; utils/image_ops.py:22-106, :165-188):
//   tvf.resize(PIL image)  = Pillow Image.resize(BILINEAR): libImaging/Resample.c, the two-pass 8 bits-per-channel
//                            path -- triangle filter, support max(scale, 1), coefficients normalised in double and
//                            rounded to 22-bit fixed point, horizontal pass rounded to uint8, then the vertical pass;
//   tvf.pad(fill=0)        on the uint8 image, i.e. BEFORE the normalisation;
//   tvf.to_tensor          uint8 / 255 in float32;
//   format_tensor_img      'RGB_1' | 'RGB_1_norm' ((x - mean) / std) | 'BGR_255_norm' (channel swap, x * 255 - mean).
// Every floating-point operation is a single IEEE operation (the library is built with -fmad=false, the harness
// with -ffp-contract=off): identical bits on both sides.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#if defined(__CUDACC__)
#define MYDET_HD __host__ __device__ __forceinline__
#else
#define MYDET_HD inline
#endif

namespace mydet {
namespace pre {

constexpr int kPrecisionBits = 32 - 8 - 2;      // Resample.c: PRECISION_BITS

enum { kFormatRGB1 = 0, kFormatRGB1Norm = 1, kFormatBGR255Norm = 2 };

// Number of coefficient slots per output coordinate (Resample.c precompute_coeffs: ksize).
MYDET_HD int resample_ksize(int in_size, int out_size) {
    double scale = (double)in_size / out_size;
    double support = scale < 1.0 ? 1.0 : scale;           // bilinear filter: support 1.0 * filterscale
    int c = (int)support;
    if ((double)c < support) ++c;                         // ceil
    return c * 2 + 1;
}

MYDET_HD double triangle(double x) {
    if (x < 0.0) x = -x;
    return x < 1.0 ? 1.0 - x : 0.0;
}

// Coefficients of output coordinate xx: bounds[0] = first tap, bounds[1] = tap count, kk[0 .. ksize) fixed point.
MYDET_HD void resample_coeffs(int in_size, int out_size, int ksize, int xx, int* bounds, int* kk) {
    const double scale = (double)in_size / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = filterscale;
    const double center = 0.0 + (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) ww += triangle((x + xmin - center + 0.5) * ss);
    for (int x = 0; x < ksize; ++x) {
        int v = 0;
        if (x < xmax) {
            double k = triangle((x + xmin - center + 0.5) * ss);
            if (ww != 0.0) k /= ww;
            v = k < 0 ? (int)(-0.5 + k * (double)(1 << kPrecisionBits)) : (int)(0.5 + k * (double)(1 << kPrecisionBits));
        }
        kk[x] = v;
    }
    bounds[0] = xmin;
    bounds[1] = xmax;
}

MYDET_HD uint8_t clip8(int acc) {
    int v = acc >> kPrecisionBits;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// One pixel (3 channels) of the horizontal pass: src_row points at the first byte of the source row (RGB interleaved).
MYDET_HD void resample_h_pixel(const uint8_t* src_row, const int* bounds_h, const int* kk_h, int ksize_h, int xx,
                               uint8_t* out3) {
    const int x0 = bounds_h[2 * xx], n = bounds_h[2 * xx + 1];
    const int* k = kk_h + (long long)xx * ksize_h;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    const uint8_t* p = src_row + 3ll * x0;
#if defined(__CUDA_ARCH__)
#pragma unroll 4                                      // the byte loads of four taps in flight before the first multiply-add
#endif
    for (int t = 0; t < n; ++t) {
        const int w = k[t];
        s0 += (int)p[3 * t + 0] * w;
        s1 += (int)p[3 * t + 1] * w;
        s2 += (int)p[3 * t + 2] * w;
    }
    out3[0] = clip8(s0); out3[1] = clip8(s1); out3[2] = clip8(s2);
}

// One pixel of the vertical pass on the uint8 intermediate (rows of row_pitch bytes, RGB interleaved).
MYDET_HD void resample_v_pixel(const uint8_t* tmp, long long row_pitch, const int* bounds_v, const int* kk_v, int ksize_v,
                               int yy, int x, uint8_t* out3) {
    const int y0 = bounds_v[2 * yy], n = bounds_v[2 * yy + 1];
    const int* k = kk_v + (long long)yy * ksize_v;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    const uint8_t* p = tmp + (long long)y0 * row_pitch + 3ll * x;
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
    for (int t = 0; t < n; ++t) {
        const int w = k[t];
        s0 += (int)p[0] * w;
        s1 += (int)p[1] * w;
        s2 += (int)p[2] * w;
        p += row_pitch;
    }
    out3[0] = clip8(s0); out3[1] = clip8(s1); out3[2] = clip8(s2);
}

// to_tensor + format_tensor_img of one pixel: rgb[3] uint8 -> the three output planes' values.
MYDET_HD void format_pixel(const uint8_t* rgb, int format, float* out3) {
    const float r = (float)rgb[0] / 255.0f, g = (float)rgb[1] / 255.0f, b = (float)rgb[2] / 255.0f;
    if (format == kFormatRGB1) {
        out3[0] = r; out3[1] = g; out3[2] = b;
    } else if (format == kFormatRGB1Norm) {                       // tvf.normalize: sub then div, float32 constants
        out3[0] = (r - 0.485f) / 0.229f;
        out3[1] = (g - 0.456f) / 0.224f;
        out3[2] = (b - 0.406f) / 0.225f;
    } else {                                                      // t[[2,1,0]] * 255, then (x - mean) / 1
        out3[0] = b * 255.0f - 102.9801f;
        out3[1] = g * 255.0f - 115.9465f;
        out3[2] = r * 255.0f - 122.7717f;
    }
}

// The same through a table: plane c of a pixel depends only on ONE source byte (channel c, or 2 - c for the BGR format),
// so lut[c * 256 + v] = format_pixel of the grey pixel (v, v, v), plane c, is bit for bit what format_pixel computes --
// and replaces three IEEE divisions by 255 plus the normalisation per pixel by three table reads.
MYDET_HD void format_lut_entry(int v, int format, float* lut) {
    const uint8_t rgb[3] = {(uint8_t)v, (uint8_t)v, (uint8_t)v};
    float o[3];
    format_pixel(rgb, format, o);
    lut[v] = o[0]; lut[256 + v] = o[1]; lut[512 + v] = o[2];
}
MYDET_HD void format_pixel_lut(const uint8_t* rgb, int format, const float* lut, float* out3) {
    const bool bgr = !(format == kFormatRGB1 || format == kFormatRGB1Norm);
    out3[0] = lut[rgb[bgr ? 2 : 0]];
    out3[1] = lut[256 + rgb[1]];
    out3[2] = lut[512 + rgb[bgr ? 0 : 2]];
}

// Geometry of one call, shared by the kernels and the harness.
struct Geometry {
    int in_h, in_w;            // source image
    int rs_h, rs_w;            // size after the resize (== in_h, in_w when the image is only padded)
    int left, top;             // where the resized image sits in the output
    int out_h, out_w;          // padded output
    int ksize_h, ksize_v;
    int format;
    int direct;                // no resize at all: the output reads the source directly
    int v_first;               // vertical pass first (Pillow's rule for very tall images, see make_plan)
};

// One output pixel (all three planes) of the final pass.  `img` is the uint8 intermediate of this image -- rows of
// 3 * rs_w bytes after a horizontal first pass, rows of 3 * in_w bytes after a vertical one (G.v_first) -- or, when
// G.direct, the source image (rows of src_row_pitch bytes).
MYDET_HD void final_pixel(const Geometry& G, const uint8_t* img, long long row_pitch, const int* bounds_h, const int* kk_h,
                          const int* bounds_v, const int* kk_v, int y, int x, float* out3, const float* lut = nullptr) {
    uint8_t rgb[3] = {0, 0, 0};                                   // the zero padding of tvf.pad(fill=0)
    const int ry = y - G.top, rx = x - G.left;
    if (ry >= 0 && ry < G.rs_h && rx >= 0 && rx < G.rs_w) {
        if (G.direct) {
            const uint8_t* p = img + (long long)ry * row_pitch + 3ll * rx;
            rgb[0] = p[0]; rgb[1] = p[1]; rgb[2] = p[2];
        } else if (G.v_first) {
            resample_h_pixel(img + (long long)ry * row_pitch, bounds_h, kk_h, G.ksize_h, rx, rgb);
        } else {
            resample_v_pixel(img, row_pitch, bounds_v, kk_v, G.ksize_v, ry, rx, rgb);
        }
    }
    if (lut) format_pixel_lut(rgb, G.format, lut, out3);
    else format_pixel(rgb, G.format, out3);
}

// (column, row, image) of linear work item i of a (batch, rows, cols) space, x fastest.  32-bit arithmetic whenever the
// index fits: a 64-bit divide is ~100 instructions, and every item starts with two of them.
MYDET_HD void split_index(long long i, int cols, int rows, int* x, int* y, long long* b) {
    if (i <= 0x7fffffffLL) {
        const unsigned u = (unsigned)i, row = u / (unsigned)cols;
        *x = (int)(u - row * (unsigned)cols);
        const unsigned bb = row / (unsigned)rows;
        *y = (int)(row - bb * (unsigned)rows);
        *b = (long long)bb;
    } else {
        *x = (int)(i % cols);
        const long long row = i / cols;
        *y = (int)(row % rows);
        *b = row / rows;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Work items.  A kernel thread (or one iteration of the harness loop) with linear index i does exactly one of these.

// i in [0, rs_w + rs_h): the filter bank entry of one output column / row.
MYDET_HD void coeff_item(const Geometry& G, int i, int* bounds_h, int* kk_h, int* bounds_v, int* kk_v) {
    if (i < G.rs_w) {
        resample_coeffs(G.in_w, G.rs_w, G.ksize_h, i, bounds_h + 2 * i, kk_h + (long long)i * G.ksize_h);
    } else if (i < G.rs_w + G.rs_h) {
        const int y = i - G.rs_w;
        resample_coeffs(G.in_h, G.rs_h, G.ksize_v, y, bounds_v + 2 * y, kk_v + (long long)y * G.ksize_v);
    }
}

// One pixel of the uint8 intermediate, x fastest -- neighbouring threads read overlapping spans of one source row (or
// neighbouring bytes of the same source rows) and write neighbouring bytes.
//   horizontal first pass: i in [0, batch * in_h * rs_w) = (b, y, xx),  intermediate in_h x rs_w
//   vertical first pass:   i in [0, batch * rs_h * in_w) = (b, yy, x),  intermediate rs_h x in_w
MYDET_HD void first_item(const Geometry& G, long long i, const uint8_t* src, long long src_image_stride,
                         long long src_row_pitch, const int* bounds_h, const int* kk_h, const int* bounds_v,
                         const int* kk_v, uint8_t* tmp) {
    const int cols = G.v_first ? G.in_w : G.rs_w, rows = G.v_first ? G.rs_h : G.in_h;
    int x, y;
    long long b;
    split_index(i, cols, rows, &x, &y, &b);
    const uint8_t* image = src + b * src_image_stride;
    uint8_t px[3];
    if (G.v_first) resample_v_pixel(image, src_row_pitch, bounds_v, kk_v, G.ksize_v, y, x, px);
    else resample_h_pixel(image + (long long)y * src_row_pitch, bounds_h, kk_h, G.ksize_h, x, px);
    uint8_t* o = tmp + i * 3;
    o[0] = px[0]; o[1] = px[1]; o[2] = px[2];
}

// Four consecutive floats as one 16-byte store.  The host build checks the alignment the device store needs and
// writes the same four values, so the harness exercises this branch too.
MYDET_HD void store4(float* p, float a, float b, float c, float d) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
#else
    if (reinterpret_cast<uintptr_t>(p) & 15) abort();
    p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#endif
}

// i in [0, batch * out_h * quads_per_row): 4 consecutive output pixels (b, y, 4q .. 4q+3) in all three planes.
// vec_ok: out_w % 4 == 0 and dst 16-byte aligned, so each plane takes one 16-byte store.
MYDET_HD void final_item(const Geometry& G, long long i, const uint8_t* img, long long image_stride,
                         long long row_pitch, const int* bounds_h, const int* kk_h, const int* bounds_v, const int* kk_v,
                         float* dst, int quads_per_row, int vec_ok, const float* lut = nullptr) {
    int q, y;
    long long b;
    split_index(i, quads_per_row, G.out_h, &q, &y, &b);
    const uint8_t* im = img + b * image_stride;
    const long long plane = (long long)G.out_h * G.out_w;
    float* o = dst + b * 3 * plane + (long long)y * G.out_w;
    const int x0 = q * 4;
    float v[4][3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
        v[k][0] = v[k][1] = v[k][2] = 0.f;
        if (x0 + k < G.out_w) final_pixel(G, im, row_pitch, bounds_h, kk_h, bounds_v, kk_v, y, x0 + k, v[k], lut);
    }
    if (vec_ok && x0 + 3 < G.out_w) {
        for (int c = 0; c < 3; ++c) store4(o + c * plane + x0, v[0][c], v[1][c], v[2][c], v[3][c]);
        return;
    }
    for (int k = 0; k < 4 && x0 + k < G.out_w; ++k)
        for (int c = 0; c < 3; ++c) o[c * plane + x0 + k] = v[k][c];
}

// ------------------------------------------------------------------------------------------------------------------
// Host-side plan of one call: validated geometry, work-item counts and the workspace layout (byte offsets of the
// 256-byte aligned segments).  Used by mydet_preprocess and by the host harness, so the sizes the kernels are
// launched with are the sizes the harness checks under AddressSanitizer.
struct Plan {
    Geometry G;
    int n_coeff_items;                 // rs_w + rs_h
    long long n_first_items;           // pixels of the uint8 intermediate: batch * in_h * rs_w, or batch * rs_h * in_w (v_first)
    int quads_per_row;                 // ceil(out_w / 4)
    long long n_final_items;           // batch * out_h * quads_per_row
    size_t off_bounds_h, off_kk_h, off_bounds_v, off_kk_v, off_tmp, workspace_bytes;
    long long tmp_image_stride, tmp_row_pitch;
};

// Returns NULL on success, else a static message describing the rejected argument.
inline const char* make_plan(int batch, int in_h, int in_w, int rs_h, int rs_w, int left, int top, int out_h, int out_w,
                             int format, Plan* P) {
    if (batch < 0) return "negative batch";
    if (in_h <= 0 || in_w <= 0 || rs_h <= 0 || rs_w <= 0 || out_h <= 0 || out_w <= 0) return "image sizes must be positive";
    if (in_h > 65536 || in_w > 65536 || out_h > 65536 || out_w > 65536) return "image sides above 65536";
    if (left < 0 || top < 0 || (long long)left + rs_w > out_w || (long long)top + rs_h > out_h)
        return "the resized image does not fit inside the output";
    if (format < kFormatRGB1 || format > kFormatBGR255Norm) return "unknown input format";
    Geometry& G = P->G;
    G.in_h = in_h; G.in_w = in_w; G.rs_h = rs_h; G.rs_w = rs_w; G.left = left; G.top = top;
    G.out_h = out_h; G.out_w = out_w; G.format = format;
    G.ksize_h = resample_ksize(in_w, rs_w); G.ksize_v = resample_ksize(in_h, rs_h);
    G.direct = (rs_h == in_h && rs_w == in_w) ? 1 : 0;       // Image.resize returns a copy when the size is unchanged
    // Pillow's Image.resize (PIL/Image.py, as installed: 12.2.0) resizes a very tall image vertically first:
    //   if self.size[1] > self.size[0] * 100 and size[1] < self.size[1]: vertical resize, then horizontal resize
    // -- the rounding of the uint8 intermediate then happens on the other axis, so the rule is part of the bits.
    G.v_first = (!G.direct && (long long)in_h > 100ll * in_w && rs_h < in_h) ? 1 : 0;
    P->n_coeff_items = rs_w + rs_h;
    P->n_first_items = G.v_first ? (long long)batch * rs_h * in_w : (long long)batch * in_h * rs_w;
    P->quads_per_row = (out_w + 3) / 4;
    P->n_final_items = (long long)batch * out_h * P->quads_per_row;
    if (P->n_first_items / 256 >= 0x7fffffffll || P->n_final_items / 256 >= 0x7fffffffll) return "batch too large for one launch";
    size_t off = 0;
    auto take = [&off](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    P->off_bounds_h = take(sizeof(int) * 2 * (size_t)rs_w);
    P->off_kk_h = take(sizeof(int) * (size_t)rs_w * G.ksize_h);
    P->off_bounds_v = take(sizeof(int) * 2 * (size_t)rs_h);
    P->off_kk_v = take(sizeof(int) * (size_t)rs_h * G.ksize_v);
    P->off_tmp = take(G.direct ? 0 : (size_t)P->n_first_items * 3);
    P->workspace_bytes = G.direct ? 256 : off;
    P->tmp_row_pitch = 3ll * (G.v_first ? in_w : rs_w);
    P->tmp_image_stride = P->tmp_row_pitch * (G.v_first ? rs_h : in_h);
    return nullptr;
}

}  // namespace pre
}  // namespace mydet
