// Image pre-processing on the device: uint8 HWC frames -> the float32 NCHW tensor the
// model consumes.  Replaces Detector._preprocess_pil + tvf.to_tensor + format_tensor_img
// (api/detection.py:158-162, :177-205; utils/image_ops.py:22-106, :165-188), which run on the CPU through Pillow.
//
// Three stream-ordered launches per batch of equally sized frames, no host synchronisation, caller-owned workspace:
//   resample_coeff_kernel   the fixed-point filter banks of both axes (one thread per output coordinate)
//   first_pass_kernel       horizontal pass: source rows -> uint8 intermediate (in_h x rs_w), as Pillow rounds it
//                           (vertical pass first for images more than 100x taller than wide: Pillow's own rule)
//   final_kernel            the other pass + zero padding + /255 + normalisation / channel order -> float32 planes,
//                           each thread produces 4 consecutive pixels of a row and writes three 16-byte vectors
// Byte work, bound by instruction issue (ncu: 74-85 % issue-active), not by HBM: no tensor cores; no shared-memory
// staging -- a variant that staged each CTA's source span with coalesced word loads was measured SLOWER (682 vs 486 us
// per 64 x 1080p batch: 207 k small CTAs with two barriers each; the taps of neighbouring threads are served by L1 anyway).
// Per-pixel float work is a 768-entry table (format_lut_entry), index splitting is 32-bit.
// All arithmetic lives in preprocess_core.cuh (host/device), which tests compile for the host and compare with
// Pillow and the reference bit for bit.
#include "internal.cuh"
#include "preprocess_core.cuh"

namespace mydet {
namespace pre {

__global__ void resample_coeff_kernel(Geometry G, int* __restrict__ bounds_h, int* __restrict__ kk_h,
                                      int* __restrict__ bounds_v, int* __restrict__ kk_v) {
    coeff_item(G, (int)(blockIdx.x * blockDim.x + threadIdx.x), bounds_h, kk_h, bounds_v, kk_v);
}

__global__ void first_pass_kernel(Geometry G, const uint8_t* __restrict__ src, long long src_image_stride,
                                  long long src_row_pitch, const int* __restrict__ bounds_h, const int* __restrict__ kk_h,
                                  const int* __restrict__ bounds_v, const int* __restrict__ kk_v,
                                  uint8_t* __restrict__ tmp, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) first_item(G, i, src, src_image_stride, src_row_pitch, bounds_h, kk_h, bounds_v, kk_v, tmp);
}

__global__ void final_kernel(Geometry G, const uint8_t* __restrict__ img, long long image_stride, long long row_pitch,
                             const int* __restrict__ bounds_h, const int* __restrict__ kk_h,
                             const int* __restrict__ bounds_v, const int* __restrict__ kk_v, float* __restrict__ dst,
                             int quads_per_row, long long total_quads, int vec_ok) {
    __shared__ float s_lut[768];                       // to_tensor + format_tensor_img of every byte value, per plane
    for (int v = threadIdx.x; v < 256; v += blockDim.x) format_lut_entry(v, G.format, s_lut);
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total_quads)
        final_item(G, i, img, image_stride, row_pitch, bounds_h, kk_h, bounds_v, kk_v, dst, quads_per_row, vec_ok, s_lut);
}

}  // namespace pre
}  // namespace mydet

using namespace mydet;

MYDET_API size_t mydet_preprocess_workspace_bytes(int batch, int in_h, int in_w, int resized_h, int resized_w) {
    pre::Plan P;
    if (batch <= 0 || pre::make_plan(batch, in_h, in_w, resized_h, resized_w, 0, 0, resized_h, resized_w, 0, &P)) return 256;
    return P.workspace_bytes;
}

MYDET_API int mydet_preprocess(const uint8_t* src, int batch, int64_t src_image_stride, int64_t src_row_pitch, int in_h,
                               int in_w, int resized_h, int resized_w, int left, int top, int out_h, int out_w,
                               int format, float* dst, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    pre::Plan P;
    const char* why = pre::make_plan(batch, in_h, in_w, resized_h, resized_w, left, top, out_h, out_w, format, &P);
    MYDET_REQUIRE(!why, "mydet_preprocess: %s", why);
    MYDET_REQUIRE(src_row_pitch >= 3ll * in_w && (batch <= 1 || src_image_stride >= src_row_pitch * in_h),
                  "source pitch / image stride smaller than the image");
    if (batch == 0) return 0;
    MYDET_REQUIRE(src && dst, "NULL tensor pointer");
    const pre::Geometry& G = P.G;
    const int vec_ok = (out_w % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) ? 1 : 0;
    const unsigned final_blocks = (unsigned)((P.n_final_items + 255) / 256);
    if (G.direct) {
        pre::final_kernel<<<final_blocks, 256, 0, st>>>(G, src, src_image_stride, src_row_pitch, nullptr, nullptr, nullptr,
                                                        nullptr, dst, P.quads_per_row, P.n_final_items, vec_ok);
        return launch_status("final_kernel");
    }
    MYDET_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
    if (P.workspace_bytes > workspace_bytes) {
        set_error("workspace too small: %zu bytes given, %zu needed", workspace_bytes, P.workspace_bytes);
        return MYDET_ERR_WORKSPACE;
    }
    char* ws = static_cast<char*>(workspace);
    int* bounds_h = reinterpret_cast<int*>(ws + P.off_bounds_h);
    int* kk_h = reinterpret_cast<int*>(ws + P.off_kk_h);
    int* bounds_v = reinterpret_cast<int*>(ws + P.off_bounds_v);
    int* kk_v = reinterpret_cast<int*>(ws + P.off_kk_v);
    uint8_t* tmp = reinterpret_cast<uint8_t*>(ws + P.off_tmp);
    pre::resample_coeff_kernel<<<(unsigned)((P.n_coeff_items + 127) / 128), 128, 0, st>>>(G, bounds_h, kk_h, bounds_v, kk_v);
    if (int rc = launch_status("resample_coeff_kernel")) return rc;
    pre::first_pass_kernel<<<(unsigned)((P.n_first_items + 255) / 256), 256, 0, st>>>(
        G, src, src_image_stride, src_row_pitch, bounds_h, kk_h, bounds_v, kk_v, tmp, P.n_first_items);
    if (int rc = launch_status("first_pass_kernel")) return rc;
    pre::final_kernel<<<final_blocks, 256, 0, st>>>(G, tmp, P.tmp_image_stride, P.tmp_row_pitch, bounds_h, kk_h, bounds_v,
                                                    kk_v, dst, P.quads_per_row, P.n_final_items, vec_ok);
    return launch_status("final_kernel");
}
