// Fused threshold -> top-k select -> sort -> class-aware AABB NMS for images whose post-select
// candidate set fits one CTA's shared memory (<= MYDET_SMALL_K = 1024 boxes; the reference's
// hard cap is 512).  One CTA per image; nothing leaves the SM between the stages.
//
// Replaces ImageObjects.post_process / non_max_suppression (utils/structures.py:92-173) and the
// CPU kernel of torchvision.ops.nms it calls, with identical arithmetic:
//   keep  score >= conf_thres                       (:98, float32 compare)
//   top-k by (score desc, index asc)                (:99-101, torch.topk; tie policy is ours)
//   per class, visiting order = stable score desc:  suppressed iff (double)iou > nms_thres
//   output = class ascending, then score descending (:158-171)
//
// Stages: (A) MSB-first 8-bit radix select on the 64-bit key (score, ~index) with early exit,
// (B) gather + 64-bit bitonic sort on (class, ~score, index), (C) upper-triangular IoU bit matrix
// in shared memory, same-class pairs only, (D) single-warp sweep that jumps from kept box to kept
// box with ffs, (E) ordered compaction of the survivors.
#include "internal.cuh"

namespace mydet {

constexpr int kPPThreads = 1024;

__device__ __forceinline__ int load_cls(const void* cls, int is64, long long i) {
    return is64 ? (int)reinterpret_cast<const long long*>(cls)[i] : reinterpret_cast<const int*>(cls)[i];
}

// key for selection: larger = better.  (score key << 32) | (~index)
__device__ __forceinline__ unsigned long long select_key(float s, int i) {
    return ((unsigned long long)float_key(s) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
}

__global__ void __launch_bounds__(kPPThreads, 1) postprocess_small_kernel(const PPParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int kpad = P.kpad;
    const int W = kpad >> 5;  // mask words per row

    // shared layout
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);           // kpad
    float4* sbox = reinterpret_cast<float4*>(keys + kpad);                                // kpad
    float* sarea = reinterpret_cast<float*>(sbox + kpad);                                 // kpad
    int* scls = reinterpret_cast<int*>(sarea + kpad);                                     // kpad
    unsigned* mask = reinterpret_cast<unsigned*>(scls + kpad);                            // kpad * W
    unsigned* hist = mask + (size_t)kpad * W;                                             // 256
    unsigned* keptw = hist + 256;                                                         // 32
    __shared__ unsigned long long s_prefix;
    __shared__ int s_need, s_done, s_nsel, s_total, s_flags;

    int n = P.n_per_image;
    int flags = 0;
    if (P.counts) {
        int c = P.counts[b];
        if (c > n) flags |= 4; else n = c < 0 ? 0 : c;
    }
    const float* scores = P.scores + (long long)b * P.pitch;
    const float* boxes = P.boxes + (long long)b * P.pitch * P.n_param;
    const long long cls_base = (long long)b * P.pitch;
    const float thr = P.conf_thres;
    const int K = P.topk;

    if (tid == 0) { s_nsel = 0; s_total = 0; s_flags = 0; s_done = 0; s_prefix = 0ull; }
    for (int i = tid; i < kpad * W; i += kPPThreads) mask[i] = 0u;
    __syncthreads();

    // ---- (A0) how many candidates pass the threshold
    {
        int local = 0;
        for (int i = tid; i < n; i += kPPThreads) local += (scores[i] >= thr) ? 1 : 0;
        local = __reduce_add_sync(0xffffffffu, local);
        if ((tid & 31) == 0 && local) atomicAdd(&s_total, local);
    }
    __syncthreads();
    const int total = s_total;
    unsigned long long kth = 0ull;  // select every passing key >= kth
    if (total > K) {
        // ---- (A) radix select of the K-th largest 64-bit key, 8 bits per pass, MSB first
        if (tid == 0) s_need = K;
        for (int pass = 7; pass >= 0; --pass) {
            const int shift = pass * 8;
            for (int i = tid; i < 256; i += kPPThreads) hist[i] = 0u;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            const unsigned long long himask = (pass == 7) ? 0ull : (~0ull << (shift + 8));
            for (int i = tid; i < n; i += kPPThreads) {
                const float s = scores[i];
                if (s >= thr) {
                    const unsigned long long k = select_key(s, i);
                    if ((k & himask) == prefix) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
                }
            }
            __syncthreads();
            if (tid < 32) {
                // lane l owns digits [8l, 8l+8); find the digit that holds the need-th largest
                unsigned h[8], mine = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { h[j] = hist[tid * 8 + j]; mine += h[j]; }
                // above = number of keys in digits greater than this lane's range
                unsigned above = 0;
                for (int l = 31; l >= 0; --l) {
                    unsigned v = __shfl_sync(0xffffffffu, mine, l);
                    if (l > tid) above += v;
                }
                const int need = s_need;
                if ((int)above < need && need <= (int)(above + mine)) {
                    unsigned acc = above;
                    for (int j = 7; j >= 0; --j) {
                        if ((int)acc < need && need <= (int)(acc + h[j])) {
                            s_prefix = prefix | ((unsigned long long)(tid * 8 + j) << shift);
                            s_need = need - (int)acc;
                            if ((int)h[j] == need - (int)acc) s_done = 1;  // whole bucket is taken
                            break;
                        }
                        acc += h[j];
                    }
                }
            }
            __syncthreads();
            if (s_done) break;
        }
        kth = s_prefix;
    }

    // ---- (B) gather the selected candidates as sort keys: class asc, score desc, index asc
    for (int i = tid; i < kpad; i += kPPThreads) keys[i] = ~0ull;
    __syncthreads();
    for (int i = tid; i < n; i += kPPThreads) {
        const float s = scores[i];
        if (s >= thr && select_key(s, i) >= kth) {
            int c = load_cls(P.cls, P.cls_is_i64, cls_base + i);
            if (c < 0 || c > MYDET_MAX_CLASS_ID) { atomicOr(&s_flags, 1); c = c < 0 ? 0 : MYDET_MAX_CLASS_ID; }
            const int slot = atomicAdd(&s_nsel, 1);
            if (slot < kpad)
                keys[slot] = ((unsigned long long)c << 52) |
                             ((unsigned long long)(~float_key(s)) << 20) | (unsigned long long)i;
        }
    }
    __syncthreads();
    const int m = min(s_nsel, kpad);

    // bitonic sort, ascending
    for (int size = 2; size <= kpad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (kpad >> 1); t += kPPThreads) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = keys[lo], c = keys[hi];
                if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
            }
            __syncthreads();
        }
    }

    // ---- corners and areas exactly as structures.py:128-143 + torchvision: c -/+ w/2, (x2-x1)*(y2-y1)
    for (int r = tid; r < kpad; r += kPPThreads) {
        float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
        int c = 0x7fffffff;
        float area = 0.f;
        if (r < m) {
            const unsigned long long k = keys[r];
            const int i = (int)(k & 0xfffffu);
            c = (int)(k >> 52);
            const float* bx = boxes + (long long)i * P.n_param;
            const float v0 = bx[0], v1 = bx[1], v2 = bx[2], v3 = bx[3];
            if (P.box_format == MYDET_BOX_CXCYWH) {
                const float hw = __fmul_rn(v2, 0.5f), hh = __fmul_rn(v3, 0.5f);
                c4 = make_float4(__fsub_rn(v0, hw), __fsub_rn(v1, hh), __fadd_rn(v0, hw), __fadd_rn(v1, hh));
            } else {
                c4 = make_float4(v0, v1, v2, v3);
            }
            area = __fmul_rn(__fsub_rn(c4.z, c4.x), __fsub_rn(c4.w, c4.y));
        }
        sbox[r] = c4; sarea[r] = area; scls[r] = c;
    }
    __syncthreads();

    // ---- (C) IoU bit matrix: mask[r][w] bit j  <=>  box (32w+j) is after r, same class, iou > thr
    {
        const int warp = tid >> 5, lane = tid & 31, nwarps = kPPThreads >> 5;
        const float thr_f = P.nms_thr_f;
        for (int r = warp; r < m; r += nwarps) {
            const float4 a = sbox[r];
            const float aarea = sarea[r];
            const int ac = scls[r];
            for (int w = r >> 5; w < W; ++w) {
                if (scls[w << 5] > ac) break;          // sorted by class: nothing further can match
                const int j = (w << 5) + lane;
                bool hit = false;
                if (j > r && j < m && scls[j] == ac) {
                    const float4 c4 = sbox[j];
                    hit = iou_corners(a.x, a.y, a.z, a.w, aarea, c4.x, c4.y, c4.z, c4.w, sarea[j]) > thr_f;
                }
                const unsigned bits = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) mask[r * W + w] = bits;
            }
        }
    }
    __syncthreads();

    // ---- (D) sweep: lane l owns word l of the "removed" vector; visit only surviving rows
    if (tid < 32) {
        const int lane = tid;
        unsigned removed = 0u, kept = 0u;
        // rows >= m do not exist
        {
            const int lo = lane << 5;
            if (lo >= m) removed = 0xffffffffu;
            else if (lo + 32 > m) removed = ~0u << (m - lo);
        }
        for (int w = 0; w < W; ++w) {
            unsigned done = 0u;
            while (true) {
                const unsigned cur = __shfl_sync(0xffffffffu, removed, w);
                const unsigned alive = ~cur & ~done;
                if (!alive) break;
                const int bit = __ffs(alive) - 1;
                done |= 1u << bit;
                const int r = (w << 5) + bit;
                if (lane == w) kept |= 1u << bit;
                if (lane >= w && lane < W) removed |= mask[r * W + lane];
            }
        }
        keptw[lane] = (lane < W) ? kept : 0u;
    }
    __syncthreads();

    // ---- (E) ordered output
    {
        int pos = -1;
        int nk = 0;
        for (int w = 0; w < W; ++w) nk += __popc(keptw[w]);
        for (int r = tid; r < m; r += kPPThreads) {
            const unsigned wbits = keptw[r >> 5];
            if ((wbits >> (r & 31)) & 1u) {
                int before = __popc(wbits & ((1u << (r & 31)) - 1u));
                for (int w = 0; w < (r >> 5); ++w) before += __popc(keptw[w]);
                pos = before;
                if (pos < P.out_cap) {
                    const unsigned long long k = keys[r];
                    const int i = (int)(k & 0xfffffu);
                    const long long orow = (long long)b * P.out_cap + pos;
                    const float* bx = boxes + (long long)i * P.n_param;
                    for (int p = 0; p < P.n_param; ++p) P.out_box[orow * P.n_param + p] = bx[p];
                    P.out_score[orow] = scores[i];
                    P.out_cls[orow] = load_cls(P.cls, P.cls_is_i64, cls_base + i);
                    P.out_idx[orow] = P.src_idx ? P.src_idx[cls_base + i] : i;
                }
            }
        }
        if (tid == 0) {
            if (nk > P.out_cap) { nk = P.out_cap; flags |= 2; }
            P.out_count[b] = nk;
            if (P.status) P.status[b] = flags | s_flags;
        }
    }
}

size_t pp_small_smem_bytes(int kpad) {
    const size_t W = (size_t)kpad / 32;
    return (size_t)kpad * (8 + 16 + 4 + 4) + (size_t)kpad * W * 4 + 256 * 4 + 32 * 4;
}

int launch_postprocess_small(const PPParams& P, int batch, cudaStream_t st) {
    const size_t smem = pp_small_smem_bytes(P.kpad);
    MYDET_CUDA(cudaFuncSetAttribute(postprocess_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)pp_small_smem_bytes(MYDET_SMALL_K)));
    postprocess_small_kernel<<<batch, kPPThreads, smem, st>>>(P);
    return launch_status("postprocess_small_kernel");
}

}  // namespace mydet
