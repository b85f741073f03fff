// Structures and launchers shared between the translation units of libmydet.
#pragma once
#include "common.cuh"

namespace mydet {

// postprocess_small.cu
struct PPParams {
    const float* boxes;
    const float* scores;
    const void* cls;
    const int* src_idx;
    int* counts;
    long long pitch;
    int n_per_image, n_param, box_format, cls_is_i64;
    float conf_thres;
    int topk;          // effective K (<= kpad)
    float nms_thr_f;   // float_at_or_below(nms_thres)
    int kpad;          // power of two >= K, >= 32
    float* out_box;
    float* out_score;
    long long* out_cls;
    int* out_idx;
    int* out_count;
    int* status;
    int out_cap;
    // fused exchange (mydet_postprocess_scatter): every detection row is also stored, packed as
    // (box, score, class) floats, into the gathered buffer of each peer GPU over NVLink
    float* peer[8];
    int n_peers;
    long long peer_row0;        // first image row of this rank inside the gathered buffer
    long long peer_rows_total;  // images in the gathered buffer (all ranks)
    int peer_vec;               // image blocks of the gathered buffers are 16-byte aligned: coalesced vector stores
    float* peer_mc;             // multicast mapping of the same buffer (NVLS): ONE store reaches every rank; NULL = unicast loop
    int peer_protocol;          // sequence flags + back-pressure (include/mydet.h: exchange protocol)
    int peer_self;              // index of this rank's own buffer in peer[] (its flags are read locally)
    int consume;                // zero counts[b] once it has been read
    int force_scan;             // skip the sampled front end of the select (tests)
};
// pdl: programmatic dependent launch -- the kernel may be scheduled while its predecessor on the stream (the decode) is
// still draining; it waits (griddepcontrol.wait) before its first global read.  Internal flag bit of pp_flags.
constexpr int kPPFlagPdl = 1 << 16;
int launch_postprocess_small(const PPParams& P, int batch, cudaStream_t st, bool pdl = false);

// Gathered-detections buffer of the fused exchange, in 32-bit words (include/mydet.h):
//   float rows[images_total][out_cap][P+2]; int32 counts[images_total];                      (the data)
//   uint32 seq[images_total]   -- how often image i has been PUBLISHED: written, with release semantics at system scope, by
//                                 the publish step that follows the producing kernel on its stream (mydet_exchange_publish
//                                 or the prologue of mydet_exchange_consume_counts) -- NOT by the producing kernel
//                                 itself, whose CTAs would each end in a system-scope drain of their peer stores
//   uint32 ack[8]              -- ack[q], written by consumer rank q into EVERY rank's copy: how many publications q has consumed
//   uint32 want[8]             -- local only: publications this rank's consumer has waited for so far
//   uint32 prod[images_total]  -- local only: how often the producing kernel has WRITTEN image i (its own rows only)
// seq / ack / prod start on 16-byte boundaries.
struct ExchangeLayout { long long counts_off, seq_off, ack_off, want_off, prod_off, total_words; };
__host__ __device__ inline ExchangeLayout exchange_layout(long long images_total, int out_cap, int n_param) {
    ExchangeLayout L;
    L.counts_off = images_total * out_cap * (n_param + 2);
    L.seq_off = (L.counts_off + images_total + 3) & ~3LL;
    L.ack_off = L.seq_off + ((images_total + 3) & ~3LL);
    L.want_off = L.ack_off + 8;
    L.prod_off = L.want_off + 8;
    L.total_words = L.prod_off + ((images_total + 3) & ~3LL);
    return L;
}

// nms_large.cu
struct LargeArgs {
    const float* boxes; const float* scores; const void* cls; int cls_is_i64; const int* src_idx; const int* counts;
    int batch; long long pitch; int n, n_param, box_format; float conf_thres; int topk; double thr; int ge; bool rot;
    float* out_box; float* out_score; long long* out_cls; int* out_idx; int* out_count; int* status; int out_cap;
    long long* keep64;
    int* votes;
    bool ws_clean = false;   // persistent workspace: the bit matrix is clean on entry and is left clean (no wholesale memset)
};
int run_large(const LargeArgs& A, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t large_workspace_bytes(int batch, int n, bool rot);

// decode.cu
int decode_compact_impl(int kind, const mydet_level_t* levels, int n_levels, int batch, int n_cls, int n_param,
                        float img_h, float img_w, float conf_thres, float* cand_box, float* cand_score,
                        int32_t* cand_cls, int32_t* cand_idx, int32_t* cand_count, int32_t capacity,
                        int state_clean, int64_t* n_total_out, cudaStream_t st);

}  // namespace mydet
