// Host harness for the image pre-processing kernels (TEST INFRASTRUCTURE, not part of libmydet).
//
// Compiles mydetection_b200/csrc/preprocess_core.cuh -- the very functions the CUDA kernels of preprocess.cu call,
// work-item index mapping and workspace plan included -- with the host compiler, and runs every work item of one call
// in a loop, in the order coefficient bank -> horizontal pass -> final pass.  Built with AddressSanitizer, every buffer
// sized exactly as the plan says, so an index that would leave a buffer on the GPU aborts here.
// tests/test_host_cpu.py compares its output bit for bit with Pillow / the reference fixtures.
//
//   preprocess_host <src.bin> <dst.bin> batch in_h in_w rs_h rs_w left top out_h out_w format row_pad
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../mydetection_b200/csrc/preprocess_core.cuh"

using namespace mydet::pre;

int main(int argc, char** argv) {
    if (argc != 14) { fprintf(stderr, "usage: see the header comment\n"); return 2; }
    int a[11];
    for (int i = 0; i < 11; ++i) a[i] = atoi(argv[3 + i]);
    const int batch = a[0], in_h = a[1], in_w = a[2], rs_h = a[3], rs_w = a[4], left = a[5], top = a[6], out_h = a[7],
              out_w = a[8], format = a[9], row_pad = a[10];
    Plan P;
    if (const char* why = make_plan(batch, in_h, in_w, rs_h, rs_w, left, top, out_h, out_w, format, &P)) {
        fprintf(stderr, "plan rejected: %s\n", why);
        return 3;
    }
    const long long row_pitch = 3ll * in_w + row_pad, image_stride = row_pitch * in_h;
    const size_t src_bytes = (size_t)batch * image_stride, dst_count = (size_t)batch * 3 * out_h * out_w;
    uint8_t* src = (uint8_t*)malloc(src_bytes ? src_bytes : 1);
    float* dst = (float*)aligned_alloc(16, (dst_count * sizeof(float) + 15) / 16 * 16);   // torch allocations are 512-byte aligned
    char* ws = (char*)malloc(P.workspace_bytes);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(src, 1, src_bytes, f) != src_bytes) { fprintf(stderr, "cannot read %s\n", argv[1]); return 4; }
    fclose(f);
    memset(dst, 0xff, dst_count * sizeof(float));            // NaN pattern: every element must be overwritten
    const Geometry& G = P.G;
    const int vec_ok = (out_w % 4 == 0 && ((uintptr_t)dst & 15) == 0) ? 1 : 0;      // as mydet_preprocess decides
    float lut[768];                                           // as final_kernel builds it in shared memory
    for (int v = 0; v < 256; ++v) format_lut_entry(v, G.format, lut);
    if (G.direct) {
        for (long long i = 0; i < P.n_final_items; ++i)
            final_item(G, i, src, image_stride, row_pitch, nullptr, nullptr, nullptr, nullptr, dst, P.quads_per_row, vec_ok, lut);
    } else {
        int* bounds_h = (int*)(ws + P.off_bounds_h);
        int* kk_h = (int*)(ws + P.off_kk_h);
        int* bounds_v = (int*)(ws + P.off_bounds_v);
        int* kk_v = (int*)(ws + P.off_kk_v);
        uint8_t* tmp = (uint8_t*)(ws + P.off_tmp);
        // the kernel is launched with ceil(n / 128) * 128 threads: run the surplus indices too, they must do nothing
        for (int i = 0; i < (P.n_coeff_items + 127) / 128 * 128; ++i) coeff_item(G, i, bounds_h, kk_h, bounds_v, kk_v);
        for (long long i = 0; i < P.n_first_items; ++i)
            first_item(G, i, src, image_stride, row_pitch, bounds_h, kk_h, bounds_v, kk_v, tmp);
        for (long long i = 0; i < P.n_final_items; ++i)
            final_item(G, i, tmp, P.tmp_image_stride, P.tmp_row_pitch, bounds_h, kk_h, bounds_v, kk_v, dst, P.quads_per_row, vec_ok, lut);
    }
    f = fopen(argv[2], "wb");
    if (!f || fwrite(dst, sizeof(float), dst_count, f) != dst_count) { fprintf(stderr, "cannot write %s\n", argv[2]); return 5; }
    fclose(f);
    free(src); free(dst); free(ws);
    return 0;
}
