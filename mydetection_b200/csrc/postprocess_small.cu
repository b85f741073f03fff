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
// (B) gather + 64-bit bitonic sort on (class, ~score, index), (C) lower-triangular IoU bit matrix
// in shared memory, same-class pairs only, (D) the greedy sweep evaluated as a parallel fixed
// point (a serial sweep by one warp was 50 % of this kernel, profiles/r1), (E) ordered compaction.
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

// number of candidates whose 32-bit score key is cached in shared memory (the rest is re-read from L2)
__host__ __device__ inline int pp_cache_elems(int n_per_image, int kpad) {
    const int budget = (kpad <= 512) ? 40960 : 12288;   // floats; keeps the CTA under 227 KB
    return n_per_image < budget ? n_per_image : budget;
}

__global__ void __launch_bounds__(kPPThreads, 1) postprocess_small_kernel(const PPParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kpad = P.kpad;
    const int W = kpad >> 5;   // mask words per row
    const int Wp = W + 1;      // padded row pitch: conflict-free column walks

    // shared layout
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);           // kpad
    float4* sbox = reinterpret_cast<float4*>(keys + kpad);                                // kpad
    float* sarea = reinterpret_cast<float*>(sbox + kpad);                                 // kpad
    int* scls = reinterpret_cast<int*>(sarea + kpad);                                     // kpad
    unsigned* mask = reinterpret_cast<unsigned*>(scls + kpad);                            // kpad * Wp
    unsigned* hist = mask + (size_t)kpad * Wp;                                            // 256
    unsigned* keptw = hist + 256;                                                         // 2 * 32
    unsigned* ucache = keptw + 64;                                                        // cache_n
    __shared__ unsigned long long s_prefix;
    __shared__ int s_need, s_done, s_nsel, s_total, s_flags;

    int n = P.n_per_image;
    int flags = 0;
    if (P.counts) {
        int c = P.counts[b];
        if (c > n) flags |= 4; else n = c < 0 ? 0 : c;
    }
    const int cache_n = min(n, pp_cache_elems(P.n_per_image, kpad));
    const float* scores = P.scores + (long long)b * P.pitch;
    const float* boxes = P.boxes + (long long)b * P.pitch * P.n_param;
    const long long cls_base = (long long)b * P.pitch;
    const float thr = P.conf_thres;
    const int K = P.topk;

    if (tid == 0) { s_nsel = 0; s_total = 0; s_flags = 0; s_done = 0; s_prefix = 0ull; }
    for (int i = tid; i < kpad * Wp; i += kPPThreads) mask[i] = 0u;
    __syncthreads();

    // 32-bit score key of candidate i, 0 when it fails the threshold (or is NaN)
    auto ukey_global = [&](int i) -> unsigned { const float s = scores[i]; return (s >= thr) ? float_key(s) : 0u; };
    auto ukey = [&](int i) -> unsigned { return i < cache_n ? ucache[i] : ukey_global(i); };

    // ---- (A0) stage the keys, count the candidates that pass the threshold
    {
        int local = 0;
        for (int i = tid; i < n; i += kPPThreads) {
            const unsigned u = ukey_global(i);
            if (i < cache_n) ucache[i] = u;
            local += u ? 1 : 0;
        }
        local = __reduce_add_sync(0xffffffffu, local);
        if (lane == 0 && local) atomicAdd(&s_total, local);
    }
    __syncthreads();
    const int total = s_total;
    unsigned long long kth = 1ull;  // select every key >= kth (key 0.. = failed threshold)
    if (total > K) {
        // ---- (A) radix select of the K-th largest 64-bit key (score key, ~index), 8 bits per pass,
        // MSB first; warp-aggregated histogram (scores cluster in a few digits)
        if (tid == 0) s_need = K;
        for (int pass = 7; pass >= 0; --pass) {
            const int shift = pass * 8;
            for (int i = tid; i < 256; i += kPPThreads) hist[i] = 0u;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            const unsigned long long himask = (pass == 7) ? 0ull : (~0ull << (shift + 8));
            for (int base = 0; base < n; base += kPPThreads) {
                const int i = base + tid;
                unsigned digit = 256u;  // "not a candidate"
                if (i < n) {
                    const unsigned u = ukey(i);
                    if (u) {
                        const unsigned long long k = ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
                        if ((k & himask) == prefix) digit = (unsigned)(k >> shift) & 255u;
                    }
                }
                const unsigned peers = __match_any_sync(0xffffffffu, digit);
                if (digit < 256u && lane == __ffs(peers) - 1) atomicAdd(&hist[digit], (unsigned)__popc(peers));
            }
            __syncthreads();
            if (tid < 32) {
                // lane l owns digits [8l, 8l+8); find the digit that holds the need-th largest
                unsigned h[8], mine = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { h[j] = hist[tid * 8 + j]; mine += h[j]; }
                unsigned above = 0;   // keys in digits above this lane's range
                for (int l = 31; l >= 0; --l) {
                    const unsigned v = __shfl_sync(0xffffffffu, mine, l);
                    if (l > tid) above += v;
                }
                const int need = s_need;
                if ((int)above < need && need <= (int)(above + mine)) {
                    unsigned acc = above;
                    for (int j = 7; j >= 0; --j) {
                        if ((int)acc < need && need <= (int)(acc + h[j])) {
                            s_prefix = prefix | ((unsigned long long)(tid * 8 + j) << shift);
                            s_need = need - (int)acc;
                            if ((int)h[j] == need - (int)acc) s_done = 1;  // the whole bucket is taken
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
        if (kth == 0ull) kth = 1ull;
    }

    // ---- (B) gather the selected candidates as sort keys: class asc, score desc, index asc
    for (int i = tid; i < kpad; i += kPPThreads) keys[i] = ~0ull;
    __syncthreads();
    for (int base = 0; base < n; base += kPPThreads) {
        const int i = base + tid;
        unsigned u = 0u;
        bool take = false;
        if (i < n) {
            u = ukey(i);
            take = u && ((((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (unsigned)i)) >= kth);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        int slot0 = 0;
        if (lane == 0 && bal) slot0 = atomicAdd(&s_nsel, __popc(bal));
        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
        if (take) {
            int c = load_cls(P.cls, P.cls_is_i64, cls_base + i);
            if (c < 0 || c > MYDET_MAX_CLASS_ID) { atomicOr(&s_flags, 1); c = c < 0 ? 0 : MYDET_MAX_CLASS_ID; }
            const int slot = slot0 + __popc(bal & ((1u << lane) - 1u));
            if (slot < kpad)
                keys[slot] = ((unsigned long long)c << 52) | ((unsigned long long)(~u) << 20) | (unsigned long long)i;
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

    // ---- (C) IoU bit matrix, LOWER triangle: mask[r][w] bit j  <=>  box (32w+j) ranks before r,
    // has r's class and iou > thr, i.e. it suppresses r if it is itself kept
    {
        const int nwarps = kPPThreads >> 5;
        const float thr_f = P.nms_thr_f;
        for (int r = warp; r < m; r += nwarps) {
            const float4 a = sbox[r];
            const float aarea = sarea[r];
            const int ac = scls[r];
            for (int w = r >> 5; w >= 0; --w) {
                if (scls[(w << 5) + 31] < ac) break;    // sorted by class: nothing earlier can match
                const int j = (w << 5) + lane;
                bool hit = false;
                if (j < r && scls[j] == ac) {
                    const float4 c4 = sbox[j];
                    // torchvision evaluates the pair from the higher-ranked box: areas (i = j, j = r)
                    hit = iou_corners(c4.x, c4.y, c4.z, c4.w, sarea[j], a.x, a.y, a.z, a.w, aarea) > thr_f;
                }
                const unsigned bits = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) mask[r * Wp + w] = bits;
            }
        }
    }
    // ---- (D) greedy sweep as a fixed point: kept[r] = no kept j < r suppresses r.  The sequential
    // result is the unique fixed point; after t rounds the first t rows are final, in practice a
    // handful of rounds suffice.  Thread r owns row r, warp w's ballot IS word w of the kept vector.
    unsigned* kept_a = keptw;
    unsigned* kept_b = keptw + 32;
    if (tid < 32) {
        const int lo = tid << 5;
        kept_a[tid] = (lo >= m) ? 0u : ((lo + 32 > m) ? ((1u << (m - lo)) - 1u) : 0xffffffffu);
    }
    __syncthreads();
    const int my_row = tid;  // kPPThreads == MYDET_SMALL_K >= kpad
    for (int round = 0; round <= m; ++round) {
        bool alive = false;
        if (my_row < m) {
            unsigned sup = 0u;
            const int ac = scls[my_row];
            for (int w = my_row >> 5; w >= 0; --w) {
                if (scls[(w << 5) + 31] < ac) break;
                sup |= mask[my_row * Wp + w] & kept_a[w];
            }
            alive = (sup == 0u);
        }
        const unsigned word = __ballot_sync(0xffffffffu, alive);
        bool changed = false;
        if (lane == 0 && warp < W) { kept_b[warp] = word; changed = (word != kept_a[warp]); }
        const int any = __syncthreads_or(changed ? 1 : 0);
        unsigned* t = kept_a; kept_a = kept_b; kept_b = t;
        if (!any) break;
    }
    keptw = kept_a;

    // ---- (E) ordered output
    {
        int nk = 0;
        for (int w = 0; w < W; ++w) nk += __popc(keptw[w]);
        for (int r = tid; r < m; r += kPPThreads) {
            const unsigned wbits = keptw[r >> 5];
            if ((wbits >> (r & 31)) & 1u) {
                int pos = __popc(wbits & ((1u << (r & 31)) - 1u));
                for (int w = 0; w < (r >> 5); ++w) pos += __popc(keptw[w]);
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

size_t pp_small_smem_bytes(int kpad, int n_per_image) {
    const size_t W = (size_t)kpad / 32;
    return (size_t)kpad * (8 + 16 + 4 + 4) + (size_t)kpad * (W + 1) * 4 + 256 * 4 + 64 * 4 +
           (size_t)pp_cache_elems(n_per_image, kpad) * 4;
}

int launch_postprocess_small(const PPParams& P, int batch, cudaStream_t st) {
    const size_t smem = pp_small_smem_bytes(P.kpad, P.n_per_image);
    MYDET_CUDA(cudaFuncSetAttribute(postprocess_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    postprocess_small_kernel<<<batch, kPPThreads, smem, st>>>(P);
    return launch_status("postprocess_small_kernel");
}

}  // namespace mydet
