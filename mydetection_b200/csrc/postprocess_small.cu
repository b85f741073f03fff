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
// "index" is the candidate's flat index (src_idx) when the input is a compacted buffer whose slot
// order is arbitrary, so the result does not depend on the compaction order.
//
// Stages (r1 profile notes in profiles/):
//  (A0) stage 64-bit select keys (score key, ~index) in shared memory, count threshold survivors
//  (A)  MSB-first 8-bit radix select of the K-th largest key; per-warp private histograms (scores
//       cluster in a few digits: one shared histogram serialised 1024 threads on 2-3 addresses);
//       early exit as soon as the bucket holding the K-th key is taken whole
//  (B)  gather: one global round trip fetches class, box and source index of every selected
//       candidate into shared memory; 64-bit bitonic sort on (class, ~score, index) + slot payload
//  (C)  lower-triangular IoU bit matrix in shared memory, same-class pairs only
//  (D)  the greedy sweep evaluated as a parallel fixed point (a serial one-warp sweep was 50 % of
//       the first version of this kernel)
//  (E)  ordered compaction of the survivors, from shared memory
#include "internal.cuh"

namespace mydet {

constexpr int kPPThreads = 1024;
constexpr int kPPWarps = kPPThreads / 32;

__device__ __forceinline__ int load_cls(const void* cls, int is64, long long i) {
    return is64 ? (int)reinterpret_cast<const long long*>(cls)[i] : reinterpret_cast<const int*>(cls)[i];
}

// number of candidates whose select key is cached in shared memory (the rest is re-read from L2)
__host__ __device__ inline int pp_cache_elems(int n_per_image, int kpad) {
    const int budget = (kpad <= 512) ? 16384 : 4096;
    return n_per_image < budget ? n_per_image : budget;
}
__host__ __device__ inline size_t pp_mask_bytes(int kpad) {
    const size_t m = (size_t)kpad * (size_t)(kpad / 32 + 1) * 4;
    const size_t h = (size_t)kPPWarps * 256 * 4;   // the per-warp histograms alias the mask
    return m > h ? m : h;
}

__global__ void __launch_bounds__(kPPThreads, 1) postprocess_small_kernel(const PPParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kpad = P.kpad;
    const int W = kpad >> 5;   // mask words per row
    const int Wp = W + 1;      // padded row pitch: conflict-free column walks

    // ---- shared layout (16-byte pieces first)
    unsigned char* sp = smem_raw;
    float4* gbox = reinterpret_cast<float4*>(sp); sp += (size_t)kpad * 16;                 // by gather slot
    float4* sbox = reinterpret_cast<float4*>(sp); sp += (size_t)kpad * 16;                 // by sorted rank: corners
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(sp); sp += (size_t)kpad * 8;
    unsigned* mask = reinterpret_cast<unsigned*>(sp); sp += pp_mask_bytes(kpad);           // kpad * Wp words
    unsigned* whist = mask;                                                                // 32 x 256, aliases mask
    float* sarea = reinterpret_cast<float*>(sp); sp += (size_t)kpad * 4;
    int* scls = reinterpret_cast<int*>(sp); sp += (size_t)kpad * 4;                        // by sorted rank
    int* gsrc = reinterpret_cast<int*>(sp); sp += (size_t)kpad * 4;                        // by gather slot
    float* gang = reinterpret_cast<float*>(sp); sp += (size_t)kpad * 4;                    // 5th box column, by slot
    unsigned* hist = reinterpret_cast<unsigned*>(sp); sp += 256 * 4;
    unsigned* keptw = reinterpret_cast<unsigned*>(sp); sp += 64 * 4;
    unsigned long long* ckey = reinterpret_cast<unsigned long long*>(sp); sp += (size_t)pp_cache_elems(P.n_per_image, kpad) * 8;
    unsigned short* pay = reinterpret_cast<unsigned short*>(sp);                           // slot payload of the sort
    __shared__ unsigned long long s_prefix;
    __shared__ int s_need, s_done, s_nsel, s_total, s_flags;

    int n = P.n_per_image;
    int flags = 0;
    if (P.counts) {
        int c = P.counts[b];
        if (c > n) flags |= 4; else n = c < 0 ? 0 : c;
    }
    const int cache_n = min(n, pp_cache_elems(P.n_per_image, kpad));
    const long long row0 = (long long)b * P.pitch;
    const float* scores = P.scores + row0;
    const float* boxes = P.boxes + row0 * P.n_param;
    const int* src = P.src_idx ? P.src_idx + row0 : nullptr;
    const float thr = P.conf_thres;
    const int K = P.topk;

    if (tid == 0) { s_nsel = 0; s_total = 0; s_flags = 0; s_done = 0; s_prefix = 0ull; }

    // select key of candidate i: (score key << 32) | ~tie index; 0 when it fails the threshold / is NaN
    auto key_global = [&](int i) -> unsigned long long {
        const float s = scores[i];
        if (!(s >= thr)) return 0ull;
        const unsigned tie = src ? (unsigned)src[i] : (unsigned)i;
        return ((unsigned long long)float_key(s) << 32) | (unsigned long long)(0xffffffffu - tie);
    };
    auto key_of = [&](int i) -> unsigned long long { return i < cache_n ? ckey[i] : key_global(i); };

    // ---- (A0) stage the keys, count the candidates that pass the threshold
    {
        int local = 0;
        for (int i = tid; i < n; i += kPPThreads) {
            const unsigned long long k = key_global(i);
            if (i < cache_n) ckey[i] = k;
            local += k ? 1 : 0;
        }
        local = __reduce_add_sync(0xffffffffu, local);
        if (lane == 0 && local) atomicAdd(&s_total, local);
    }
    __syncthreads();
    const int total = s_total;
    unsigned long long kth = 1ull;  // select every key >= kth
    if (total > K) {
        // ---- (A) radix select of the K-th largest key
        if (tid == 0) s_need = K;
        unsigned* myhist = whist + warp * 256;
        for (int pass = 7; pass >= 0; --pass) {
            const int shift = pass * 8;
            for (int i = tid; i < kPPWarps * 256; i += kPPThreads) whist[i] = 0u;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            const unsigned long long himask = (pass == 7) ? 0ull : (~0ull << (shift + 8));
            for (int i = tid; i < n; i += kPPThreads) {
                const unsigned long long k = key_of(i);
                if (k && (k & himask) == prefix) atomicAdd(&myhist[(unsigned)(k >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid < 256) {
                unsigned sum = 0;
#pragma unroll 8
                for (int w = 0; w < kPPWarps; ++w) sum += whist[w * 256 + tid];
                hist[tid] = sum;
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

    // ---- (B) gather the selected candidates: sort key (class asc, score desc, tie index asc) and,
    // in the same global round trip, their box / class / source index into shared memory
    for (int i = tid; i < kpad; i += kPPThreads) { keys[i] = ~0ull; pay[i] = (unsigned short)i; }
    for (int i = tid; i < kpad * Wp; i += kPPThreads) mask[i] = 0u;   // the histograms are dead now
    __syncthreads();
    for (int base = 0; base < n; base += kPPThreads) {
        const int i = base + tid;
        unsigned long long k = 0ull;
        if (i < n) k = key_of(i);
        const bool take = k >= kth && k != 0ull;
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        int slot0 = 0;
        if (lane == 0 && bal) slot0 = atomicAdd(&s_nsel, __popc(bal));
        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
        if (take) {
            const int slot = slot0 + __popc(bal & ((1u << lane) - 1u));
            if (slot < kpad) {
                int c = load_cls(P.cls, P.cls_is_i64, row0 + i);
                const float* bx = boxes + (long long)i * P.n_param;
                gbox[slot] = make_float4(bx[0], bx[1], bx[2], bx[3]);
                if (P.n_param == 5) gang[slot] = bx[4];
                if (c < 0 || c > MYDET_MAX_CLASS_ID) { atomicOr(&s_flags, 1); c = c < 0 ? 0 : MYDET_MAX_CLASS_ID; }
                const unsigned tie = 0xffffffffu - (unsigned)(k & 0xffffffffull);
                gsrc[slot] = (int)tie;
                if (tie > 0xfffffu) atomicOr(&s_flags, 8);   // tie index does not fit the 20-bit field
                keys[slot] = ((unsigned long long)c << 52) | ((unsigned long long)(~(unsigned)(k >> 32)) << 20) |
                             (unsigned long long)(tie & 0xfffffu);
            }
        }
    }
    __syncthreads();
    const int m = min(s_nsel, kpad);

    // bitonic sort of (key, slot), ascending
    for (int size = 2; size <= kpad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (kpad >> 1); t += kPPThreads) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = keys[lo], c = keys[hi];
                if ((a > c) == up) {
                    keys[lo] = c; keys[hi] = a;
                    const unsigned short pa = pay[lo]; pay[lo] = pay[hi]; pay[hi] = pa;
                }
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
            c = (int)(keys[r] >> 52);
            const float4 v = gbox[pay[r]];
            if (P.box_format == MYDET_BOX_CXCYWH) {
                const float hw = __fmul_rn(v.z, 0.5f), hh = __fmul_rn(v.w, 0.5f);
                c4 = make_float4(__fsub_rn(v.x, hw), __fsub_rn(v.y, hh), __fadd_rn(v.x, hw), __fadd_rn(v.y, hh));
            } else {
                c4 = v;
            }
            area = __fmul_rn(__fsub_rn(c4.z, c4.x), __fsub_rn(c4.w, c4.y));
        }
        sbox[r] = c4; sarea[r] = area; scls[r] = c;
    }
    __syncthreads();

    // ---- (C) IoU bit matrix, LOWER triangle: mask[r][w] bit j  <=>  box (32w+j) ranks before r,
    // has r's class and iou > thr, i.e. it suppresses r if it is itself kept
    {
        const float thr_f = P.nms_thr_f;
        for (int r = warp; r < m; r += kPPWarps) {
            const float4 a = sbox[r];
            const float aarea = sarea[r];
            const int ac = scls[r];
            for (int w = r >> 5; w >= 0; --w) {
                if (scls[(w << 5) + 31] < ac) break;    // sorted by class: nothing earlier can match
                const int j = (w << 5) + lane;
                bool hit = false;
                if (j < r && scls[j] == ac) {
                    const float4 c4 = sbox[j];
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

    // ---- (E) ordered output, everything from shared memory
    {
        int nk = 0;
        for (int w = 0; w < W; ++w) nk += __popc(keptw[w]);
        const int r = tid;
        if (r < m) {
            const unsigned wbits = keptw[r >> 5];
            if ((wbits >> (r & 31)) & 1u) {
                int pos = __popc(wbits & ((1u << (r & 31)) - 1u));
                for (int w = 0; w < (r >> 5); ++w) pos += __popc(keptw[w]);
                if (pos < P.out_cap) {
                    const unsigned long long k = keys[r];
                    const int slot = pay[r];
                    const long long orow = (long long)b * P.out_cap + pos;
                    const float4 v = gbox[slot];
                    if (P.n_param == 4) {
                        reinterpret_cast<float4*>(P.out_box)[orow] = v;
                    } else {
                        float* ob = P.out_box + orow * 5;
                        ob[0] = v.x; ob[1] = v.y; ob[2] = v.z; ob[3] = v.w; ob[4] = gang[slot];
                    }
                    P.out_score[orow] = key_float(~(unsigned)(k >> 20));
                    P.out_cls[orow] = (long long)(k >> 52);
                    P.out_idx[orow] = gsrc[slot];
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
    return (size_t)kpad * (16 + 16 + 8 + 4 + 4 + 4 + 4 + 2) + pp_mask_bytes(kpad) + 256 * 4 + 64 * 4 +
           (size_t)pp_cache_elems(n_per_image, kpad) * 8;
}

int launch_postprocess_small(const PPParams& P, int batch, cudaStream_t st) {
    const size_t smem = pp_small_smem_bytes(P.kpad, P.n_per_image);
    MYDET_CUDA(cudaFuncSetAttribute(postprocess_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    postprocess_small_kernel<<<batch, kPPThreads, smem, st>>>(P);
    return launch_status("postprocess_small_kernel");
}

}  // namespace mydet
