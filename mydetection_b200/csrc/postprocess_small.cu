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
//  (R)  n <= 9 216: scores and tie indices are loaded ONCE into registers; an exact 2048-bin shared-memory
//       histogram of the valid scores gives the bin that holds the K-th score and the exact count above it; the
//       register-resident scores are classified without a second pass: bins above go straight to their slots, the
//       boundary bin (a handful of candidates) to a short list whose best are taken by counting.
//  (S)  larger n: the K-th score is bracketed from a histogram of a SAMPLE of the scores, ONE pass copies the
//       candidates above that edge to a short list and counts them, the count is verified (>= K), radix passes run
//       on the list alone.  Exact whatever the sample says; when the check fails (or a list would not fit, e.g.
//       heavy ties) the scan path (A0)-(B1) runs instead:
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

// developer instrumentation (scripts/pp_phases.py builds with -DMYDET_PP_PROFILE): SM clock at the
// phase boundaries of CTA 0
#ifdef MYDET_PP_PROFILE
__device__ long long g_pp_clock[32];
#define PP_MARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_pp_clock[i] = clock64(); } while (0)
#else
#define PP_MARK(i) do {} while (0)
#endif

__device__ __forceinline__ int load_cls(const void* cls, int is64, long long i) {
    return is64 ? (int)reinterpret_cast<const long long*>(cls)[i] : reinterpret_cast<const int*>(cls)[i];
}

// number of candidates whose select key is cached in shared memory (the rest is re-read from L2)
__host__ __device__ inline int pp_cache_elems(int n_per_image, int kpad) {
    const int budget = (kpad <= 512) ? 16384 : 2048;   // kpad 1024: the 135 KB mask leaves room for 2048 keys (227 KB per CTA)
    return n_per_image < budget ? n_per_image : budget;
}
__host__ __device__ inline size_t pp_mask_bytes(int kpad) {
    const size_t m = (size_t)kpad * (size_t)(kpad / 32 + 1) * 4;
    // the mask region first hosts the 32 padded per-warp histograms of the radix select, then the
    // 4096(+32)+4096 class-sort counters
    const size_t h = (size_t)kPPWarps * 257 * 4;
    return m > h ? m : h;
}

constexpr int kClassBins = MYDET_MAX_CLASS_ID + 1;   // 4096
constexpr int kHistBins = 2 * kPPThreads;            // sample histogram of the front end

// 48 registers per thread (no spills), not the 62 the compiler takes when it may: a 1024-thread CTA then leaves a quarter
// of the SM's register file free, so two CTAs of the bandwidth-bound decode kernel of the NEXT step stay resident on each
// of the 64 SMs a post-process launch occupies.  Measured in the pipelined bench step: 33.3 -> 31.2 us (profiles/r2_history.md).
#ifndef MYDET_PP_MAXNREG
#define MYDET_PP_MAXNREG 48
#endif
__global__ void __maxnreg__(MYDET_PP_MAXNREG) postprocess_small_kernel(const PPParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int kpad = P.kpad;
    const int W = kpad >> 5;   // mask words per row
    const int Wp = W + 1;      // padded row pitch: conflict-free column walks

    // ---- shared layout (16-byte pieces first)
    unsigned char* sp = smem_raw;
    float4* gbox = reinterpret_cast<float4*>(sp); sp += (size_t)kpad * 16;                 // raw box, by gather slot
    float4* sbox = reinterpret_cast<float4*>(sp); sp += (size_t)kpad * 16;                 // corners, by sorted rank
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(sp); sp += (size_t)kpad * 8;   // sort key by rank
    unsigned* mask = reinterpret_cast<unsigned*>(sp); sp += pp_mask_bytes(kpad);           // kpad * Wp words
    float* sarea = reinterpret_cast<float*>(sp); sp += (size_t)kpad * 4;                   // by sorted rank
    int* scls = reinterpret_cast<int*>(sp); sp += (size_t)kpad * 4;                        // by sorted rank
    int* gsrc = reinterpret_cast<int*>(sp); sp += (size_t)kpad * 4;                        // by gather slot
    float* gang = reinterpret_cast<float*>(sp); sp += (size_t)kpad * 4;                    // 5th box column, by slot
    int* gcls = reinterpret_cast<int*>(sp); sp += (size_t)kpad * 4;                        // class, by gather slot
    int* sseg = reinterpret_cast<int*>(sp); sp += (size_t)kpad * 4;                        // first rank of the class, by sorted rank
    int* ppre = reinterpret_cast<int*>(sp); sp += (size_t)kpad * 4;                        // same-class pairs before this rank's row
    unsigned* hist = reinterpret_cast<unsigned*>(sp); sp += 256 * 4;
    unsigned* keptw = reinterpret_cast<unsigned*>(sp); sp += 64 * 4;
    int* wtot = reinterpret_cast<int*>(sp); sp += 64 * 4;
    unsigned long long* ckey = reinterpret_cast<unsigned long long*>(sp); sp += (size_t)pp_cache_elems(P.n_per_image, kpad) * 8;
    unsigned short* pay = reinterpret_cast<unsigned short*>(sp);                           // slot, by sorted rank
    // regions with two lives
    unsigned* whist = mask;                                             // radix select: 32 private histograms
    int* cstart = reinterpret_cast<int*>(mask);                         // class sort: 4096 (+1) segment starts ...
    int* ccur = cstart + kClassBins + 32;                               // ... and 4096 scatter cursors (both die before (C))
    unsigned long long* tkey = reinterpret_cast<unsigned long long*>(sbox);               // class-bucketed keys (die before corners)
    int* sel = reinterpret_cast<int*>(tkey + kpad);                                         // candidate number by slot
    unsigned short* tslot = reinterpret_cast<unsigned short*>(sel + kpad);                  // class-bucketed slots
    unsigned long long* ukey = reinterpret_cast<unsigned long long*>(gbox);                 // undecided bucket of the select: keys ...
    int* uidx = reinterpret_cast<int*>(sarea);                                              // ... and candidate numbers (both die before (B2))
    __shared__ unsigned long long s_prefix;
    __shared__ int s_need, s_done, s_nsel, s_total, s_flags, s_bucket, s_ucount, s_top, s_nvalid;
    __shared__ unsigned s_seq;      // exchange protocol: the sequence number this launch publishes for the image

    // launched as a programmatic dependent of the decode (mydet_detect): everything above overlapped its tail; the
    // decode's stores are visible from here on.  A no-op for an ordinary launch.
    asm volatile("griddepcontrol.wait;" ::: "memory");
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

    if (tid == 0) { s_nsel = 0; s_total = 0; s_flags = 0; s_done = 0; s_prefix = 0ull; s_ucount = 0; s_nvalid = 0; }
    __syncthreads();
    PP_MARK(0);

    // select key of candidate i: (score key << 32) | ~tie index; 0 when it fails the threshold / is NaN
    auto key_global = [&](int i) -> unsigned long long {
        const float s = scores[i];
        if (!(s >= thr)) return 0ull;
        const unsigned tie = src ? (unsigned)src[i] : (unsigned)i;
        return ((unsigned long long)float_key(s) << 32) | (unsigned long long)(0xffffffffu - tie);
    };
    auto key_of = [&](int i) -> unsigned long long { return i < cache_n ? ckey[i] : key_global(i); };
    auto key_list = [&](int i) -> unsigned long long { return ukey[i]; };

    // ---- radix-select machinery, shared by the histogram front end (short list only) and the scan path
    // start of a select: combine the per-warp AND / OR pairs left in hist[0..63] / hist[64..127] into the
    // highest bit in which the keys differ at all (s_top) and their common prefix above it
    auto select_init = [&](int need) {
        __syncthreads();
        if (warp == 0) {
            unsigned a_lo = hist[2 * lane], a_hi = hist[2 * lane + 1], o_lo = hist[64 + 2 * lane], o_hi = hist[64 + 2 * lane + 1];
            a_lo = __reduce_and_sync(0xffffffffu, a_lo); a_hi = __reduce_and_sync(0xffffffffu, a_hi);
            o_lo = __reduce_or_sync(0xffffffffu, o_lo); o_hi = __reduce_or_sync(0xffffffffu, o_hi);
            if (lane == 0) {
                const unsigned long long all_and = ((unsigned long long)a_hi << 32) | a_lo, all_or = ((unsigned long long)o_hi << 32) | o_lo;
                const unsigned long long diff = all_and ^ all_or;           // bits that are not common to every key
                const int top = diff ? 63 - __clzll((long long)diff) : 0;   // highest differing bit
                s_top = top;
                s_need = need; s_prefix = (top >= 63) ? 0ull : (all_and & (~0ull << (top + 1)));
            }
        }
        __syncthreads();
    };
    auto publish_and_or = [&](unsigned long long k_and, unsigned long long k_or) {
        const unsigned a_lo = __reduce_and_sync(0xffffffffu, (unsigned)k_and), a_hi = __reduce_and_sync(0xffffffffu, (unsigned)(k_and >> 32));
        const unsigned o_lo = __reduce_or_sync(0xffffffffu, (unsigned)k_or), o_hi = __reduce_or_sync(0xffffffffu, (unsigned)(k_or >> 32));
        if (lane == 0) {
            hist[2 * warp] = a_lo; hist[2 * warp + 1] = a_hi;
            hist[64 + 2 * warp] = o_lo; hist[64 + 2 * warp + 1] = o_hi;
        }
    };
    // one histogram pass over `cnt` keys delivered by key_at(i); digit = bits [hi_bit-7, hi_bit]
    auto pass = [&](auto key_at, int cnt, int hi_bit, bool priv) -> int {
        // priv: 32 padded per-warp histograms (all n keys: the digits cluster); else one shared one (short list)
        const int shift = hi_bit >= 7 ? hi_bit - 7 : 0;
        const unsigned dmask = (hi_bit >= 7) ? 255u : ((1u << (hi_bit + 1)) - 1u);
        if (priv) { for (int i = tid; i < kPPWarps * 257; i += kPPThreads) whist[i] = 0u; }
        else if (tid < 256) hist[tid] = 0u;
        unsigned* myhist = priv ? whist + warp * 257 : hist;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const unsigned long long himask = (hi_bit >= 63) ? 0ull : (~0ull << (hi_bit + 1));
#pragma unroll 2
        for (int base = 0; base < cnt; base += kPPThreads) {
            const int i = base + tid;
            unsigned digit = 256u;                              // not a candidate of this pass
            if (i < cnt) {
                const unsigned long long k = key_at(i);
                if (k && (k & himask) == prefix) digit = (unsigned)(k >> shift) & dmask;
            }
            // scores cluster: when the whole warp agrees, one lane adds 32
            const unsigned first = __shfl_sync(0xffffffffu, digit, 0);
            if (__all_sync(0xffffffffu, digit == first)) {
                if (lane == 0 && first < 256u) atomicAdd(&myhist[first], 32u);
            } else if (digit < 256u) {
                atomicAdd(&myhist[digit], 1u);
            }
        }
        __syncthreads();
        if (priv) {
            if (tid < 256) {
                unsigned sum = 0;
#pragma unroll 8
                for (int w = 0; w < kPPWarps; ++w) sum += whist[w * 257 + tid];
                hist[tid] = sum;
            }
            __syncthreads();
        }
        if (tid < 32) {
            // lane l owns digits [8l, 8l+8); find the digit that holds the need-th largest
            unsigned h[8], mine = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { h[j] = hist[tid * 8 + j]; mine += h[j]; }
            unsigned incl = mine;                               // suffix sum over lanes >= tid
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_down_sync(0xffffffffu, incl, o);
                if (tid + o < 32) incl += v;
            }
            const unsigned above = incl - mine;                 // keys in digits above this lane's range
            const int need = s_need;
            if ((int)above < need && need <= (int)(above + mine)) {
                unsigned acc = above;
                bool found = false;
#pragma unroll
                for (int j = 7; j >= 0; --j) {                         // unrolled: h[] stays in registers
                    if (!found && (int)acc < need && need <= (int)(acc + h[j])) {
                        s_prefix = prefix | ((unsigned long long)(tid * 8 + j) << shift);
                        s_need = need - (int)acc;
                        s_bucket = (int)h[j];
                        if ((int)h[j] == need - (int)acc) s_done = 1;  // the whole bucket is taken
                        found = true;
                    }
                    acc += h[j];
                }
            }
        }
        __syncthreads();
        return shift;
    };
    // the undecided keys at or above the K-th key take the slots after the sure ones
    auto append_list = [&](unsigned long long kth_key) {
        const int ucount = min(s_ucount, kpad);
        for (int base = 0; base < ucount; base += kPPThreads) {
            const int u = base + tid;
            const bool take = u < ucount && ukey[u] >= kth_key;
            const unsigned bal = __ballot_sync(0xffffffffu, take);
            if (bal) {
                int slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(&s_nsel, __popc(bal));
                slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                if (take) {
                    const int slot = slot0 + __popc(bal & lt_mask);
                    if (slot < kpad) { sel[slot] = uidx[u]; keys[slot] = ukey[u]; }
                }
            }
        }
        __syncthreads();
    };

    // list entries (ukey / lidx) at or above kth_key -> selection slots
    int* lidx = reinterpret_cast<int*>(sbox);                // list: candidate numbers (2 kpad ints; tkey's space)
    auto take_from_list = [&](int c, unsigned long long kth_key) {
        for (int base = 0; base < c; base += kPPThreads) {
            const int u = base + tid;
            const unsigned long long k = (u < c) ? ukey[u] : 0ull;
            const bool take = u < c && k >= kth_key;
            const unsigned bal = __ballot_sync(0xffffffffu, take);
            if (bal) {
                int slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(&s_nsel, __popc(bal));
                slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                if (take) {
                    const int slot = slot0 + __popc(bal & lt_mask);
                    if (slot < kpad) { sel[slot] = lidx[u]; keys[slot] = k; }
                }
            }
        }
        __syncthreads();
    };
    // K-th largest key of the list by 8-bit radix passes (the list has more than `need` entries)
    auto list_kth = [&](int c, int need) -> unsigned long long {
        unsigned long long k_and = ~0ull, k_or = 0ull;
        for (int u = tid; u < c; u += kPPThreads) { const unsigned long long k = ukey[u]; k_and &= k; k_or |= k; }
        publish_and_or(k_and, k_or);
        select_init(need);
        int shift = pass(key_list, c, s_top, false);
        while (!(s_done || shift == 0)) shift = pass(key_list, c, shift - 1, false);
        return s_prefix;
    };
    const int list_cap = 2 * kpad;
    bool fast_done = false;                      // block-uniform

    // ---- (R) register-resident front end (n <= 9 216, e.g. the 8 525 candidates of D1 @640).  Scores and tie
    // indices are loaded ONCE into registers; an EXACT 2048-bin histogram of all valid scores (shared-memory atomics,
    // bins linear over the score range) gives the bin T that holds the K-th score and the exact count above it; the
    // register-resident scores are then classified without a second pass over memory: bins above T go straight to
    // their slots, bin T (a handful of candidates) to the short list, of which the best are taken by counting.
    constexpr int UR = 9;
    const bool reg_path = !P.force_scan && n <= UR * kPPThreads;
    if (reg_path) {
        unsigned* sh = mask;                     // 2048 bins (the mask region is free until (B2))
        static_assert(kHistBins == 2 * kPPThreads, "thread t owns bins 2t and 2t+1");
        reinterpret_cast<uint2*>(sh)[tid] = make_uint2(0u, 0u);
        float sv[UR];
        unsigned tv[UR];
#pragma unroll
        for (int u = 0; u < UR; ++u) {
            const int i = u * kPPThreads + tid;
            sv[u] = (i < n) ? scores[i] : __int_as_float(0x7fc00000);   // NaN pads the tail: fails every compare
            tv[u] = (i < n && src) ? (unsigned)src[i] : (unsigned)i;
        }
        unsigned kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
        for (int u = 0; u < UR; ++u)
            if (sv[u] >= thr) { const unsigned k = float_key(sv[u]); kmin = min(kmin, k); kmax = max(kmax, k); }
        kmin = __reduce_min_sync(0xffffffffu, kmin); kmax = __reduce_max_sync(0xffffffffu, kmax);
        if (lane == 0) { wtot[warp] = (int)kmin; wtot[32 + warp] = (int)kmax; }
        __syncthreads();
        const float smin = key_float(__reduce_min_sync(0xffffffffu, (unsigned)wtot[lane]));
        const float smax = key_float(__reduce_max_sync(0xffffffffu, (unsigned)wtot[32 + lane]));
        const float scale = (smax > smin) ? __fdiv_rn((float)(kHistBins - 1), __fsub_rn(smax, smin)) : 0.0f;
        auto bin_of = [&](float q) -> int {                  // the SAME function builds and reads the histogram
            return min(kHistBins - 1, max(0, (int)__fmul_rn(__fsub_rn(q, smin), scale)));
        };
#pragma unroll
        for (int u = 0; u < UR; ++u)
            if (sv[u] >= thr) atomicAdd(&sh[bin_of(sv[u])], 1u);
        __syncthreads();
        // suffix sums over the bins: `above` = candidates in bins above this thread's pair
        const uint2 h = reinterpret_cast<const uint2*>(sh)[tid];
        const int mine = (int)(h.x + h.y);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_down_sync(0xffffffffu, incl, o); if (lane + o < 32) incl += v; }
        if (lane == 0) wtot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int v = wtot[lane];
            int wi = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_down_sync(0xffffffffu, wi, o); if (lane + o < 32) wi += t; }
            wtot[32 + lane] = wi - v;
            if (lane == 0) s_total = wi;
        }
        __syncthreads();
        {
            const int total_valid = s_total;
            const int above1 = wtot[32 + warp] + incl - mine;   // above bin 2*tid+1
            const int above0 = above1 + (int)h.y;               // above bin 2*tid
            if (total_valid > K) {
                // the one bin T with  above(T) < K <= above(T) + hist[T]  holds the K-th score
                if (above1 < K && K <= above1 + (int)h.y) { s_top = 2 * tid + 1; s_bucket = (int)h.y; s_need = K - above1; }
                else if (above0 < K && K <= above0 + (int)h.x) { s_top = 2 * tid; s_bucket = (int)h.x; s_need = K - above0; }
            } else if (tid == 0) {
                s_top = -1; s_bucket = 0; s_need = 0;           // everything valid is selected
            }
        }
        __syncthreads();
        PP_MARK(16);
        for (int i = tid; i < kClassBins + 32; i += kPPThreads) { cstart[i] = 0; }   // class histogram (the bins are dead)
        for (int i = tid; i < kClassBins; i += kPPThreads) { ccur[i] = 0; }
        if (s_bucket <= list_cap) {
            const int T = s_top, need = s_need;
            unsigned sure_bits = 0u, und_bits = 0u;
#pragma unroll
            for (int u = 0; u < UR; ++u) {
                const bool valid = sv[u] >= thr;
                const int bq = bin_of(sv[u]);
                sure_bits |= ((valid && bq > T) ? 1u : 0u) << u;
                und_bits |= ((valid && bq == T) ? 1u : 0u) << u;
            }
            // slots: exclusive warp scan of (sure count | undecided count << 16), one shared atomic per warp
            const int packed = __popc(sure_bits) | (__popc(und_bits) << 16);
            int pin = packed;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pin, o); if (lane >= o) pin += t; }
            const int wtotal = __shfl_sync(0xffffffffu, pin, 31);
            int bs = 0, bu = 0;
            if (lane == 0) {
                if (wtotal & 0xffff) bs = atomicAdd(&s_nsel, wtotal & 0xffff);
                if (wtotal >> 16) bu = atomicAdd(&s_ucount, wtotal >> 16);
            }
            bs = __shfl_sync(0xffffffffu, bs, 0) + ((pin - packed) & 0xffff);
            bu = __shfl_sync(0xffffffffu, bu, 0) + ((pin - packed) >> 16);
            if (sure_bits | und_bits) {
#pragma unroll
                for (int u = 0; u < UR; ++u) {
                    const bool sure = (sure_bits >> u) & 1u, und = (und_bits >> u) & 1u;
                    if (sure || und) {
                        const int i = u * kPPThreads + tid;
                        const unsigned long long k = ((unsigned long long)float_key(sv[u]) << 32) | (unsigned long long)(0xffffffffu - tv[u]);
                        if (sure) {
                            if (bs < kpad) { sel[bs] = i; keys[bs] = k; }
                            ++bs;
                        } else {
                            if (bu < list_cap) { ukey[bu] = k; lidx[bu] = i; }
                            ++bu;
                        }
                    }
                }
            }
            __syncthreads();
            PP_MARK(17);
            const int ucount = min(s_ucount, list_cap);
            if (T >= 0 && ucount > 0) {
                if (need >= ucount) {
                    take_from_list(ucount, 0ull);
                } else if (ucount <= kPPThreads / 8) {
                    // rank by counting, 8 threads per key (keys are unique: the tie index is)
                    const int e = tid >> 3, part = tid & 7;
                    const unsigned long long ke = (e < ucount) ? ukey[e] : 0ull;
                    int rank = 0;
                    if (e < ucount)
                        for (int v = part; v < ucount; v += 8) rank += (ukey[v] > ke) ? 1 : 0;
                    rank += __shfl_xor_sync(0xffffffffu, rank, 1);
                    rank += __shfl_xor_sync(0xffffffffu, rank, 2);
                    rank += __shfl_xor_sync(0xffffffffu, rank, 4);
                    const bool take = e < ucount && part == 0 && rank < need;
                    const unsigned bal = __ballot_sync(0xffffffffu, take);
                    if (bal) {
                        int slot0 = 0;
                        if (lane == 0) slot0 = atomicAdd(&s_nsel, __popc(bal));
                        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                        if (take) {
                            const int slot = slot0 + __popc(bal & lt_mask);
                            if (slot < kpad) { sel[slot] = lidx[e]; keys[slot] = ke; }
                        }
                    }
                    __syncthreads();
                } else {
                    take_from_list(ucount, list_kth(ucount, need));
                }
            }
            fast_done = true;
        } else {                                 // a boundary bin too full for the list (heavy ties): scan path
            if (tid == 0) { s_total = 0; s_ucount = 0; s_nsel = 0; s_done = 0; s_prefix = 0ull; }
            __syncthreads();
        }
    }
    PP_MARK(18);

    // ---- (S) sampled front end.  The K-th score is bracketed from a SAMPLE of the scores (every s-th
    // candidate, <= 2048 of them, one or two loads per thread) binned into a 2048-bin shared-memory histogram:
    // the bin below which  K/s + 4 sqrt(K/s) + 4  samples lie gives a float edge that, with high probability,
    // has between K and ~1.5 K candidates at or above it.  ONE pass over all scores then copies exactly those
    // candidates to a short list and counts them; if the list holds >= K entries (or every valid candidate)
    // the top K of the list are the top K of the image -- exact, whatever the sample looked like -- and are
    // picked by radix passes over the list alone.  Otherwise (estimate too tight, list overflow, huge n) the
    // scan path below runs.  Replaces staging + full radix pass + slot assignment: 23 k -> ~9 k cycles.
    if (!P.force_scan && !reg_path) {
        unsigned* sh = mask;                     // 2048 sample bins (the mask region is free until (B2))
        const int stride = max(1, (n + 2 * kPPThreads - 1) / (2 * kPPThreads));
        static_assert(kHistBins == 2 * kPPThreads, "thread t owns bins 2t and 2t+1");
        reinterpret_cast<uint2*>(sh)[tid] = make_uint2(0u, 0u);
        const int i0 = tid * stride + (stride >> 1), i1 = (tid + kPPThreads) * stride + (stride >> 1);
        const float q0 = (i0 < n) ? scores[i0] : __int_as_float(0x7fc00000);
        const float q1 = (i1 < n) ? scores[i1] : __int_as_float(0x7fc00000);
        const bool ok0 = q0 >= thr, ok1 = q1 >= thr;            // NaN fails
        // range of the valid samples: the bins are linear over [smin, smax], whatever the score scale is
        {
            unsigned kmin = 0xffffffffu, kmax = 0u;
            if (ok0) { const unsigned k = float_key(q0); kmin = min(kmin, k); kmax = max(kmax, k); }
            if (ok1) { const unsigned k = float_key(q1); kmin = min(kmin, k); kmax = max(kmax, k); }
            kmin = __reduce_min_sync(0xffffffffu, kmin); kmax = __reduce_max_sync(0xffffffffu, kmax);
            if (lane == 0) { wtot[warp] = (int)kmin; wtot[32 + warp] = (int)kmax; }
        }
        __syncthreads();
        const float smin = key_float(__reduce_min_sync(0xffffffffu, (unsigned)wtot[lane]));
        const float smax = key_float(__reduce_max_sync(0xffffffffu, (unsigned)wtot[32 + lane]));
        const float scale = (smax > smin) ? __fdiv_rn((float)(kHistBins - 1), __fsub_rn(smax, smin)) : 0.0f;
        auto sample_bin = [&](float q) -> int {
            const float f = __fmul_rn(__fsub_rn(q, smin), scale);
            return min(kHistBins - 1, max(0, (int)f));          // the verification below makes rounding harmless
        };
        if (ok0) atomicAdd(&sh[sample_bin(q0)], 1u);
        if (ok1) atomicAdd(&sh[sample_bin(q1)], 1u);
        __syncthreads();
        // suffix sums over the bins: `above` = samples in bins above this thread's pair
        const uint2 h = reinterpret_cast<const uint2*>(sh)[tid];
        const int mine = (int)(h.x + h.y);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_down_sync(0xffffffffu, incl, o); if (lane + o < 32) incl += v; }
        if (lane == 0) wtot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int v = wtot[lane];
            int wi = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_down_sync(0xffffffffu, wi, o); if (lane + o < 32) wi += t; }
            wtot[32 + lane] = wi - v;
            if (lane == 0) s_total = wi;
        }
        __syncthreads();
        const int stotal = s_total;              // valid samples
        const float ks = (float)K / (float)stride;
        const int target = (stride == 1) ? K : (int)(ks + 4.0f * sqrtf(ks) + 4.0f);
        {
            const int above1 = wtot[32 + warp] + incl - mine;   // above bin 2*tid+1
            const int above0 = above1 + (int)h.y;               // above bin 2*tid
            if (stotal >= target) {
                // the one bin T with  above(T) < target <= above(T) + hist[T]
                if (above1 < target && target <= above1 + (int)h.y) { s_top = 2 * tid + 1; s_bucket = above1 + (int)h.y; }
                else if (above0 < target && target <= above0 + (int)h.x) { s_top = 2 * tid; s_bucket = above0 + (int)h.x; }
            } else if (tid == 0) {
                s_top = 0; s_bucket = stotal;                   // few valid candidates: all of them go to the list
            }
        }
        __syncthreads();
        PP_MARK(16);
        for (int i = tid; i < kClassBins + 32; i += kPPThreads) { cstart[i] = 0; }   // class histogram (the sample bins are dead)
        for (int i = tid; i < kClassBins; i += kPPThreads) { ccur[i] = 0; }
        // expected list length; leave the fast path to the scan when it would not fit
        if ((long long)s_bucket * stride <= (long long)list_cap - list_cap / 8) {
            const int T = s_top;
            // lower edge of bin T (approximately: exactness comes from the count check below, not from here)
            float edge = thr;
            if (T > 0 && scale > 0.0f) edge = fmaxf(thr, __fadd_rn(smin, __fdiv_rn((float)T, scale)));
            if (!(edge == edge)) edge = thr;                    // NaN threshold: nothing is valid either way
            // (S2) one pass: scores AND tie indices are loaded up front as one batch of independent loads
            constexpr int U = 9;                 // 9 x 1024 >= 8 525: one round on the D1 @640 geometry
            int nvalid = 0;
#pragma unroll 1
            for (int base = 0; base < n; base += U * kPPThreads) {
                float sv[U];
                unsigned tv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = base + u * kPPThreads + tid;
                    sv[u] = (i < n) ? scores[i] : __int_as_float(0x7fc00000);   // NaN pads the tail: fails the compares
                    tv[u] = (i < n && src) ? (unsigned)src[i] : (unsigned)i;
                }
                unsigned bits = 0u;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    bits |= ((sv[u] >= edge) ? 1u : 0u) << u;
                    nvalid += (sv[u] >= thr) ? 1 : 0;
                }
                // list positions: exclusive warp scan of the per-thread counts, one shared atomic per warp
                const int cnt = __popc(bits);
                int incl2 = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl2, o); if (lane >= o) incl2 += t; }
                const int wtotal = __shfl_sync(0xffffffffu, incl2, 31);
                int pos = 0;
                if (lane == 0 && wtotal) pos = atomicAdd(&s_ucount, wtotal);
                pos = __shfl_sync(0xffffffffu, pos, 0) + incl2 - cnt;
                if (bits) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if ((bits >> u) & 1u) {
                            if (pos < list_cap) {
                                ukey[pos] = ((unsigned long long)float_key(sv[u]) << 32) | (unsigned long long)(0xffffffffu - tv[u]);
                                lidx[pos] = base + u * kPPThreads + tid;
                            }
                            ++pos;
                        }
                    }
                }
            }
            nvalid = __reduce_add_sync(0xffffffffu, nvalid);
            if (lane == 0 && nvalid) atomicAdd(&s_nvalid, nvalid);
            __syncthreads();
            PP_MARK(17);
            const int c = s_ucount, v = s_nvalid;
            if (c <= list_cap && (c >= K || c == v)) {
                // (S3) the K best of the list (all of it when it has no more than K entries)
                take_from_list(c, (c > K) ? list_kth(c, K) : 0ull);
                fast_done = true;
            }
        }
        if (!fast_done) {                        // the scan path below starts from scratch
            if (tid == 0) { s_total = 0; s_ucount = 0; s_nsel = 0; s_done = 0; s_prefix = 0ull; }
            __syncthreads();
        }
    }
    PP_MARK(18);

    if (!fast_done) {
    // ---- (A0) stage the keys, count the threshold survivors, AND / OR of all keys (common prefix)
    unsigned long long k_and = ~0ull, k_or = 0ull;
    {
        int local = 0;
        constexpr int U = 8;   // independent global loads in flight per thread
#pragma unroll 1
        for (int base = 0; base < n; base += U * kPPThreads) {
            float sv[U];
            unsigned tv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * kPPThreads + tid;
                sv[u] = (i < n) ? scores[i] : __int_as_float(0x7fc00000);   // NaN fails the threshold
                tv[u] = (i < n && src) ? (unsigned)src[i] : (unsigned)i;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * kPPThreads + tid;
                unsigned long long k = 0ull;
                if (sv[u] >= thr) {
                    k = ((unsigned long long)float_key(sv[u]) << 32) | (unsigned long long)(0xffffffffu - tv[u]);
                    k_and &= k; k_or |= k; ++local;
                }
                if (i < cache_n) ckey[i] = k;
            }
        }
        local = __reduce_add_sync(0xffffffffu, local);
        if (lane == 0 && local) atomicAdd(&s_total, local);
        publish_and_or(k_and, k_or);                                    // hist is free until the first pass
    }
    __syncthreads();
    const int total = s_total;
    PP_MARK(1);
    unsigned long long kth = 1ull;  // select every key >= kth
    bool use_list = false;                       // block-uniform
    unsigned long long list_bucket = 0ull, list_mask = 0ull;
    if (total > K) {
        // ---- (A) radix select of the K-th largest key, 8 bits per pass starting at the first bit in which
        // the keys differ at all.  Only the FIRST pass scans all n keys: the keys of the bucket that holds the
        // K-th key ("undecided", usually a few dozen) are then copied to a short list and the remaining passes
        // run on that list.  (Four full passes + two slot-assignment passes were 21 k of the kernel's 42 k cycles.)
        select_init(K);
        if (tid == 0) s_ucount = 0;
        int shift = pass(key_of, n, s_top, true);                           // the only pass over all n keys
        bool done = s_done || shift == 0;
        if (!done && s_bucket <= kpad) {
            // copy the undecided bucket (key + candidate number) to the short list; count, per warp, the
            // keys ABOVE the bucket: they are selected whatever happens next
            const unsigned long long bucket = s_prefix, bmask = ~0ull << shift;
            int sure = 0;
#pragma unroll 1
            for (int base = 0; base < n; base += kPPThreads) {
                const int i = base + tid;
                const unsigned long long k = (i < n) ? key_of(i) : 0ull;
                const unsigned long long kb = k & bmask;
                sure += __popc(__ballot_sync(0xffffffffu, k != 0ull && kb > bucket));
                const bool inb = k != 0ull && kb == bucket;
                const unsigned bal = __ballot_sync(0xffffffffu, inb);
                if (bal) {
                    int base_u = 0;
                    if (lane == 0) base_u = atomicAdd(&s_ucount, __popc(bal));
                    base_u = __shfl_sync(0xffffffffu, base_u, 0);
                    if (inb) { const int u = base_u + __popc(bal & lt_mask); ukey[u] = k; uidx[u] = i; }
                }
            }
            if (lane == 0) wtot[warp] = sure;
            __syncthreads();
            use_list = true;
            list_bucket = bucket; list_mask = bmask;
            const int ucount = s_ucount;
            while (!done) {
                shift = pass(key_list, ucount, shift - 1, false);
                done = s_done || shift == 0;
            }
        } else {
            while (!done) {                                         // heavy ties: keep scanning all keys
                shift = pass(key_of, n, shift - 1, true);
                done = s_done || shift == 0;
            }
        }
        kth = s_prefix;
        if (kth == 0ull) kth = 1ull;
    }
    PP_MARK(2);

    // ---- (B1) slots for the selected candidates without contended atomics: per-warp counts, scan, assign
    for (int i = tid; i < kClassBins + 32; i += kPPThreads) { cstart[i] = 0; }   // class histogram (whist is dead)
    for (int i = tid; i < kClassBins; i += kPPThreads) { ccur[i] = 0; }
    if (!use_list) {
        int mine = 0;
#pragma unroll 1
        for (int base = 0; base < n; base += kPPThreads) {
            const int i = base + tid;
            const unsigned long long k = (i < n) ? key_of(i) : 0ull;
            mine += __popc(__ballot_sync(0xffffffffu, k >= kth && k != 0ull));
        }
        if (lane == 0) wtot[warp] = mine;
    }
    __syncthreads();
    if (warp == 0) {
        const int v = wtot[lane];
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        wtot[32 + lane] = incl - v;
        if (lane == 31) s_nsel = incl;
    }
    __syncthreads();
    {
        int running = wtot[32 + warp];
#pragma unroll 1
        for (int base = 0; base < n; base += kPPThreads) {
            const int i = base + tid;
            const unsigned long long k = (i < n) ? key_of(i) : 0ull;
            const bool take = use_list ? (k != 0ull && (k & list_mask) > list_bucket) : (k >= kth && k != 0ull);
            const unsigned bal = __ballot_sync(0xffffffffu, take);
            if (take) {
                const int slot = running + __popc(bal & lt_mask);
                if (slot < kpad) { sel[slot] = i; keys[slot] = k; }
            }
            running += __popc(bal);
        }
    }
    __syncthreads();
    if (use_list) append_list(kth);
    }   // !fast_done
    const int m = min(s_nsel, kpad);
    PP_MARK(3);

    // ---- (B2) ONE batch of global loads: class, box, source index of the selected candidates;
    // histogram of their classes
#pragma unroll 1
    for (int slot = tid; slot < m; slot += kPPThreads) {
        const int i = sel[slot];
        const unsigned long long k = keys[slot];
        int c = load_cls(P.cls, P.cls_is_i64, row0 + i);
        const float* bx = boxes + (long long)i * P.n_param;
        if (P.n_param == 4) {
            gbox[slot] = *reinterpret_cast<const float4*>(bx);
        } else {
            gbox[slot] = make_float4(bx[0], bx[1], bx[2], bx[3]);
            gang[slot] = bx[4];
        }
        if (c < 0 || c > MYDET_MAX_CLASS_ID) { atomicOr(&s_flags, 1); c = c < 0 ? 0 : MYDET_MAX_CLASS_ID; }
        const unsigned tie = 0xffffffffu - (unsigned)(k & 0xffffffffull);
        gsrc[slot] = (int)tie;
        if (tie > 0xfffffu) atomicOr(&s_flags, 8);   // tie index does not fit the 20-bit field
        gcls[slot] = c;
        // in-class order: score descending, tie index ascending
        keys[slot] = ((unsigned long long)(~(unsigned)(k >> 32)) << 20) | (unsigned long long)(tie & 0xfffffu);
        atomicAdd(&cstart[c], 1);
    }
    __syncthreads();
    PP_MARK(4);

    // ---- (B3) sort by (class, key) = counting sort on the class + rank-by-counting inside the class.
    // (A 45-stage bitonic network is a serial chain of ~300 cycles per stage for one CTA.)
    {
        // exclusive scan of the 4096 class counts: 4 bins per thread
        const int c0 = tid * 4;
        const int v0 = cstart[c0], v1 = cstart[c0 + 1], v2 = cstart[c0 + 2], v3 = cstart[c0 + 3];
        const int mine = v0 + v1 + v2 + v3;
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int v = wtot[lane];
            int wi = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            wtot[32 + lane] = wi - v;
        }
        __syncthreads();
        const int excl = wtot[32 + warp] + incl - mine;
        cstart[c0] = excl; cstart[c0 + 1] = excl + v0; cstart[c0 + 2] = excl + v0 + v1; cstart[c0 + 3] = excl + v0 + v1 + v2;
        if (tid == 0) cstart[kClassBins] = m;
        __syncthreads();
        // scatter into class buckets (order inside a bucket is arbitrary, fixed by the ranking below)
#pragma unroll 1
        for (int slot = tid; slot < m; slot += kPPThreads) {
            const int c = gcls[slot];
            const int pos = cstart[c] + atomicAdd(&ccur[c], 1);
            tkey[pos] = keys[slot];
            tslot[pos] = (unsigned short)slot;
        }
        __syncthreads();
        // rank inside the bucket; keys are unique, so ranks are a permutation
#pragma unroll 1
        for (int p = tid; p < m; p += kPPThreads) {
            const int slot = tslot[p];
            const int c = gcls[slot];
            const int s0 = cstart[c], s1 = cstart[c + 1];
            const unsigned long long mykey = tkey[p];
            int rank = 0;
#pragma unroll 4
            for (int q = s0; q < s1; ++q) rank += (tkey[q] < mykey) ? 1 : 0;
            const int r = s0 + rank;
            keys[r] = ((unsigned long long)c << 52) | mykey;
            pay[r] = (unsigned short)slot;
        }
        __syncthreads();
    }
    PP_MARK(5);

    // ---- corners and areas exactly as structures.py:128-143 + torchvision: c -/+ w/2, (x2-x1)*(y2-y1)
    // (cstart is copied out of the mask region first: (C) overwrites it)
    int seg0 = 0;   // first sorted rank of this thread's class
    {
        const int r = tid;
        float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
        int c = 0x7fffffff;
        float area = 0.f;
        if (r < m) {
            c = (int)(keys[r] >> 52);
            seg0 = cstart[c];
            const float4 v = gbox[pay[r]];
            if (P.box_format == MYDET_BOX_CXCYWH) {
                const float hw = __fmul_rn(v.z, 0.5f), hh = __fmul_rn(v.w, 0.5f);
                c4 = make_float4(__fsub_rn(v.x, hw), __fsub_rn(v.y, hh), __fadd_rn(v.x, hw), __fadd_rn(v.y, hh));
            } else {
                c4 = v;
            }
            area = __fmul_rn(__fsub_rn(c4.z, c4.x), __fsub_rn(c4.w, c4.y));
        }
        __syncthreads();          // every thread has read tkey/sel (alias sbox) and cstart (alias mask)
        if (r < kpad) { sbox[r] = c4; sarea[r] = area; scls[r] = c; sseg[r] = seg0; }
    }
    __syncthreads();
    PP_MARK(6);

    // ---- (C) IoU bit matrix, LOWER triangle: mask[r][w] bit j  <=>  box (32w+j) ranks before r,
    // has r's class and iou > thr, i.e. it suppresses r if it is itself kept.  Row r only needs the
    // words [seg0>>5, r>>5].
    // Few pairs (the multi-class case: ~1 650 same-class pairs among 512 boxes of 80 classes): the pairs are
    // enumerated and dealt evenly to the threads -- row = binary search in the prefix sums of the row lengths,
    // one IoU per pair, the rare hit sets its bit with a shared atomic.  (A thread or an 8-lane group per ROW runs
    // every warp for the longest row in it: 7 k cycles at 16 % lane utilisation.)
    // Many pairs (few classes): a group of 8 lanes owns a row and takes 8 predecessors per step; rows are dealt
    // round-robin, so the groups stay balanced.
    bool pairs_done = false;                     // block-uniform
    {
        const int p_r = (tid < m) ? tid - seg0 : 0;                     // predecessors of row tid inside its class
        int incl = p_r;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int v = wtot[lane];
            int wi = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            wtot[32 + lane] = wi - v;
            if (lane == 31) s_total = wi;
        }
        __syncthreads();
        const int n_pairs = s_total;
        if (n_pairs <= 8 * kPPThreads) {
            if (tid < kpad) ppre[tid] = wtot[32 + warp] + incl - p_r;
            for (int i = tid; i < m * Wp; i += kPPThreads) mask[i] = 0u;
            __syncthreads();
            const float thr_f = P.nms_thr_f;
#pragma unroll 1
            for (int q = tid; q < n_pairs; q += kPPThreads) {
                int lo = 0, hi = m - 1;                                  // last row r with ppre[r] <= q
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (ppre[mid] <= q) lo = mid; else hi = mid - 1;
                }
                const int r = lo, j = sseg[r] + (q - ppre[r]);
                const float4 a = sbox[r], c4 = sbox[j];
                if (iou_corners_gt(c4.x, c4.y, c4.z, c4.w, sarea[j], a.x, a.y, a.z, a.w, sarea[r], thr_f))
                    atomicOr(&mask[r * Wp + (j >> 5)], 1u << (j & 31));
            }
            pairs_done = true;
        }
    }
    if (!pairs_done)
    {
        const float thr_f = P.nms_thr_f;
        const int grp = tid >> 3, part = tid & 7, gshift = lane & 24;   // group's byte inside the warp ballot
        constexpr int kGroups = kPPThreads / 8;
        const int rounds = (m + kGroups - 1) / kGroups;                 // block-uniform trip count: ballots stay full
#pragma unroll 1
        for (int t = 0; t < rounds; ++t) {
            const int r = t * kGroups + grp;
            const bool live = r < m;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            float aarea = 0.f;
            int s0 = 0;
            if (live) { a = sbox[r]; aarea = sarea[r]; s0 = sseg[r]; }
            // chunks of 8 predecessors, aligned to 8 so that a chunk never straddles a mask word; the warp walks
            // the longest row of its 4 groups.  The chunk holding r itself is included, so that every word
            // of [seg0>>5, r>>5] the sweep reads is written.
            const int j_first = s0 & ~7;
            const int steps = live ? (((r - j_first) >> 3) + 1) : 0;
            int wsteps = steps;
            wsteps = max(wsteps, __shfl_xor_sync(0xffffffffu, wsteps, 8));
            wsteps = max(wsteps, __shfl_xor_sync(0xffffffffu, wsteps, 16));
            unsigned acc = 0u;                                          // the mask word being assembled (leader)
#pragma unroll 1
            for (int c = 0; c < wsteps; ++c) {
                const int j0 = j_first + (c << 3), j = j0 + part;
                bool hit = false;
                if (c < steps && j >= s0 && j < r) {
                    const float4 c4 = sbox[j];
                    hit = iou_corners_gt(c4.x, c4.y, c4.z, c4.w, sarea[j], a.x, a.y, a.z, a.w, aarea, thr_f);
                }
                const unsigned byte = (__ballot_sync(0xffffffffu, hit) >> gshift) & 0xffu;
                if (part == 0 && c < steps) {
                    acc |= byte << (j0 & 24);
                    if ((j0 & 24) == 24 || c == steps - 1) { mask[r * Wp + (j0 >> 5)] = acc; acc = 0u; }
                }
            }
        }
    }
    __syncthreads();
    PP_MARK(7);

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
    {
        const int r = tid;  // kPPThreads == MYDET_SMALL_K >= kpad
        const int w_lo = seg0 >> 5, w_hi = r >> 5;
#pragma unroll 1
        for (int round = 0; round <= m; ++round) {
            bool alive = false;
            if (r < m) {
                unsigned sup = 0u;
#pragma unroll 1
                for (int w = w_hi; w >= w_lo; --w) sup |= mask[r * Wp + w] & kept_a[w];
                alive = (sup == 0u);
            }
            const unsigned word = __ballot_sync(0xffffffffu, alive);
            bool changed = false;
            if (lane == 0 && warp < W) { kept_b[warp] = word; changed = (word != kept_a[warp]); }
            const int any = __syncthreads_or(changed ? 1 : 0);
            unsigned* t = kept_a; kept_a = kept_b; kept_b = t;
            if (!any) break;
        }
    }
    keptw = kept_a;
    PP_MARK(8);

    // ---- (E) ordered output, everything from shared memory
    {
        // exclusive prefix of the kept-word popcounts, once (every thread looping over the words was 10 % of the
        // kernel's instructions): wtot[w] = kept rows before word w, wtot[32] = all kept rows
        if (warp == 0) {
            const int v = (lane < W) ? __popc(keptw[lane]) : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            wtot[lane] = incl - v;
            if (lane == 31) wtot[32] = incl;
        }
        // exchange protocol (back-pressure): the image's rows may only be overwritten once every consumer has acknowledged
        // the previous publication of this image slot.  One lane of the last warp polls the LOCAL ack words while warp 0
        // builds the prefix; the acks have normally arrived long ago.  Bounded: a consumer that never acknowledges costs
        // ~1 s once and raises status bit 16 instead of hanging the GPU.
        if (P.n_peers > 0 && P.peer_protocol && tid == kPPThreads - 32) {
            const ExchangeLayout XL = exchange_layout(P.peer_rows_total, P.out_cap, P.n_param);
            const unsigned* own = reinterpret_cast<const unsigned*>(P.peer[P.peer_self]);
            const unsigned prev = own[XL.prod_off + P.peer_row0 + b];      // written by this image's CTA of the previous launch
            const long long t0 = clock64();
            bool late = false;
            for (int q = 0; q < P.n_peers && !late; ++q)
                while ((int)(ld_acquire_sys(own + XL.ack_off + q) - prev) < 0) {
                    if (clock64() - t0 > kExchangeSpinCycles) { late = true; break; }
                    __nanosleep(64);
                }
            s_seq = prev + 1u;
            if (late) atomicOr(&s_flags, 16);
        }
        __syncthreads();
        int nk = wtot[32];
        float* stage = reinterpret_cast<float*>(mask);      // packed rows for the vector peer stores (mask is dead)
        const int r = tid;
        if (r < m) {
            const unsigned wbits = keptw[r >> 5];
            if ((wbits >> (r & 31)) & 1u) {
                const int pos = wtot[r >> 5] + __popc(wbits & ((1u << (r & 31)) - 1u));
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
                    const float sc = key_float(~(unsigned)(k >> 20));
                    const long long cl = (long long)(k >> 52);
                    P.out_score[orow] = sc;
                    P.out_cls[orow] = cl;
                    P.out_idx[orow] = gsrc[slot];
                    if (P.n_peers > 0) {
                        // the path's only exchange, fused: this row goes into every rank's gathered buffer (peer
                        // stores over NVLink / NVSwitch; own copy included).  Vector mode: the packed rows of the image
                        // are first assembled in shared memory (the dead mask region) and sent as coalesced 16-byte
                        // stores below -- per-row scalar stores cost ~6 us PER PEER (partial-sector NVLink writes).
                        const int np2 = P.n_param + 2;
                        const float ang = (P.n_param == 5) ? gang[slot] : 0.f;
                        if (P.peer_vec) {
                            float* o = stage + pos * np2;
                            o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
                            if (P.n_param == 5) { o[4] = ang; o[5] = sc; o[6] = (float)cl; }
                            else { o[4] = sc; o[5] = (float)cl; }
                        } else {
                            const long long grow = ((P.peer_row0 + b) * P.out_cap + pos) * np2;
                            const int nq = P.peer_mc ? 1 : P.n_peers;          // multicast mapping: one store reaches every rank
#pragma unroll 1
                            for (int q = 0; q < nq; ++q) {
                                float* o = (P.peer_mc ? P.peer_mc : P.peer[q]) + grow;
                                o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
                                if (P.n_param == 5) { o[4] = ang; o[5] = sc; o[6] = (float)cl; }
                                else { o[4] = sc; o[5] = (float)cl; }
                            }
                        }
                    }
                }
            }
        }
        if (P.n_peers > 0 && P.peer_vec) {
            const int np2 = P.n_param + 2;
            const int nrow = min(nk, P.out_cap);
            const int nfl = nrow * np2, nvec = (nfl + 3) >> 2;
            if (tid < 4 && nfl + tid < 4 * nvec) stage[nfl + tid] = 0.f;       // pad the last vector
            __syncthreads();
            const long long img0 = (P.peer_row0 + b) * (long long)P.out_cap * np2;   // 16-byte aligned (host check)
            const float4* sv4 = reinterpret_cast<const float4*>(stage);
#pragma unroll 1
            if (P.peer_mc) {
                // NVLS: the buffer's multicast address -- the switch replicates ONE 16-byte store into every rank's copy
                // (own copy included), instead of n_peers unicast stores per vector (~1.8 us per peer and step)
                for (int i = tid; i < nvec; i += kPPThreads) multimem_st_v4(reinterpret_cast<float4*>(P.peer_mc + img0) + i, sv4[i]);
            } else {
#pragma unroll 1
                for (int i = tid; i < nvec; i += kPPThreads) {
                    const float4 val = sv4[i];
#pragma unroll 1
                    for (int q = 0; q < P.n_peers; ++q) reinterpret_cast<float4*>(P.peer[q] + img0)[i] = val;
                }
            }
        }
        if (tid == 0) {
            if (nk > P.out_cap) { nk = P.out_cap; flags |= 2; }
            P.out_count[b] = nk;
            if (P.status) P.status[b] = flags | s_flags;
            if (P.consume && P.counts) P.counts[b] = 0;   // every thread read it before the first barrier
            if (P.n_peers > 0) {
                const ExchangeLayout XL = exchange_layout(P.peer_rows_total, P.out_cap, P.n_param);
                const long long row = P.peer_row0 + b;
                const int nq = P.peer_mc ? 1 : P.n_peers;
                for (int q = 0; q < nq; ++q) {
                    int* base = reinterpret_cast<int*>(P.peer_mc ? P.peer_mc : P.peer[q]);
                    if (P.peer_mc) multimem_st_u32(reinterpret_cast<unsigned*>(base + XL.counts_off + row), (unsigned)nk);
                    else base[XL.counts_off + row] = nk;
                }
                if (P.peer_protocol) {
                    // the image has been written once more.  Its PUBLICATION (seq, release at system scope) is left to the
                    // publish step that follows this kernel on the stream: a release here would end every CTA in a
                    // system-scope drain of its peer stores (~1 us per CTA with the SM's registers and shared memory held)
                    reinterpret_cast<unsigned*>(P.peer[P.peer_self])[XL.prod_off + row] = s_seq;
                }
            }
        }
    }
    __syncthreads();
    PP_MARK(9);
}

#ifdef MYDET_PP_PROFILE
extern "C" __attribute__((visibility("default"))) int mydet_debug_pp_clocks(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_pp_clock, sizeof(long long) * 32);
}
#endif

size_t pp_small_smem_bytes(int kpad, int n_per_image) {
    return (size_t)kpad * (16 + 16 + 8 + 4 + 4 + 4 + 4 + 4 + 4 + 4 + 2) + pp_mask_bytes(kpad) + 256 * 4 + 64 * 4 + 64 * 4 +
           (size_t)pp_cache_elems(n_per_image, kpad) * 8;
}

int launch_postprocess_small(const PPParams& P, int batch, cudaStream_t st, bool pdl) {
    const size_t smem = pp_small_smem_bytes(P.kpad, P.n_per_image);
    MYDET_CUDA(cudaFuncSetAttribute(postprocess_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!pdl) {
        postprocess_small_kernel<<<batch, kPPThreads, smem, st>>>(P);
        return launch_status("postprocess_small_kernel");
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)batch); cfg.blockDim = dim3(kPPThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    MYDET_CUDA(cudaLaunchKernelEx(&cfg, postprocess_small_kernel, P));
    return launch_status("postprocess_small_kernel");
}

}  // namespace mydet
