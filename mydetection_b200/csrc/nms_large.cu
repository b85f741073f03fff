// NMS for candidate sets that do not fit one CTA (> MYDET_SMALL_K boxes per image): the dense
// scene sweep (10k-50k boxes, BASELINE.json configs[4]) and rotated NMS at 10k boxes (configs[2]).
//
// Replaces ImageObjects.non_max_suppression (utils/structures.py:111-173, via torchvision.ops.nms)
// and nms_rotbb (utils/bbox_ops.py:250-306).  Pipeline, all stream-ordered, no host round trip:
//   keys   : 64-bit sort key per candidate (class asc, score desc, index asc); failed threshold = ~0
//   sort   : bitonic sort of (key, slot) in shared memory (<= 16 384 keys per image) or split over CTAs -> order
//   gather : sorted corner boxes / rotated quads
//   mask   : upper-triangular 64x64-tile IoU bit matrix, column tile staged in shared memory,
//            one row per thread; rotated boxes use the cull + polygon clipping of rotgeom.cuh
//   sweep  : one CTA per image walks the 64-row blocks; the diagonal word chain is resolved in
//            registers by one thread, the kept rows are OR-ed into the removed vector by all
//   emit   : ordered compaction of the survivors
#include "internal.cuh"
#include "rotgeom.cuh"
#include "sweep_fixpoint.cuh"
#include <stdlib.h>

namespace mydet {

constexpr int kTile = 64;
constexpr int kAdjMax = 16;            // adjacency words per tile: tiles <= 1024, i.e. n <= 65 536 in the spatial path
constexpr int kSpatialMaxN = kAdjMax * 64 * kTile;

struct LargeWs {          // carved out of the caller's workspace
    unsigned long long* keys;   // B*n
    int* order;                 // B*n   sorted position -> candidate slot
    int* m;                     // B     number of valid (thresholded / selected) candidates
    float4* box;                // B*n   sorted corners (AABB)
    float* area;                // B*n
    int* cls;                   // B*n
    RotBox* rbox;               // B*n   sorted rotated quads (ROT)
    unsigned long long* mask;   // B*n*words
    unsigned long long* kept;   // B*words  survivors, bit per sorted row
    int* rowpos;                // B*n      sorted row -> position in the emitted list (kept rows)
    // spatially ordered variant (rotated path, n <= 16384): boxes live in Morton order of their centres
    unsigned long long* bk;     // B*npad   big sort (n > 16384): keys ...
    int* bp;                    // B*npad   ... and payload
    unsigned long long* skeys;  // B*n      (class << 52 | morton << 20 | tie), indexed by candidate slot
    int* slot_of_spos;          // B*n      spatial position -> candidate slot (payload of the Morton sort)
    int* rank_of_slot;          // B*n      candidate slot -> score rank (inverse of `order`)
    int* rank_of_spos;          // B*n      spatial position -> score rank
    int* spos_of_rank;          // B*n      score rank -> spatial position
    float4* tile_hull;          // B*tiles  hull of the 64 boxes of a spatial tile
    unsigned long long* diag_all; // B*tiles*64  per score block, TRANSPOSED: entry c, bit a = "a suppresses c" (both of the block)
    unsigned long long* adj_blk;  // B*tiles*64*aw  per score block row: the tile_adj words of that row's spatial tile
    unsigned long long* tile_adj; // B*tiles*aw  bit j of tile i: some row of tile i has a nonzero mask word j
    int2* tile_cls;               // B*tiles     class range of a spatial tile (axis-aligned path)
    int aw;                       // adjacency words per tile = ceil(tiles / 64)
    int words;                  // ceil(n/64)
    // appended last, so that every field above keeps its offset (and the kernels that take this struct their code)
    int fx_cap;                 // entries per image of fx_list
    fx::Entry* fx_list;         // B*fx_cap  non-empty mask words of an image (fixed-point sweep)
    // rotated path, broad phase / narrow phase split (rot_broad_kernel, rot_narrow_kernel)
    float4* cull4;              // B*n      (cx, cy, circumscribed radius, area) in spatial order
    float4* hull4;              // B*n      axis-aligned hull (x0, y0, x1, y1) of the corners, spatial order
    float4* axes4;              // B*n      half-axis vectors (Hx, Hy, Vx, Vy): H = (tr - tl) / 2, V = (tl - bl) / 2
    float4* hull16;             // B*n16    hull of 16 consecutive spatial positions (n16 = ceil(n/16))
    float4* hull32;             // B*n32    hull of 32 consecutive spatial positions
    unsigned* pairs;            // B*pair_cap  candidate pairs (lower position << 16 | higher position) that reach the clip
    int* pair_count;            // B        entries appended (may exceed pair_cap: the image then takes the tile kernel)
    int pair_cap;
    int n16, n32;
    // entry list of the fixed-point sweep, appended to by the mask kernels themselves: the thread whose atomicOr turns a
    // 32-bit mask half-word non-zero records (row, half-word index) -- exactly one entry per non-empty half-word, so
    // the sweep no longer walks the adjacency map of every row to find them (75 of its 100 us at 10 000 boxes)
    int* fx_count;              // B
    // lazy narrow phase of the rotated path (rot_filter_kernel / rot_clip_kernel)
    unsigned* pairs2;           // B*pair_cap  the listed pairs that survive the oriented-extent bound
    int* pair2_count;           // B
    unsigned char* lz_high;     // B*n  by spatial position: some surviving pair has a higher-scored partner for this box
    unsigned char* lz_dead;     // B*n  by spatial position: overlapped (>= thr) by a ROOT (a box with lz_high == 0)
};
// record the first bit of a mask half-word in the image's entry list (entries beyond the capacity are dropped: the
// count still says so, and the sweep then falls back to the adjacency walk)
__device__ __forceinline__ void fx_note(const LargeWs& w, int b, int row, int half, unsigned old) {
    if (old != 0u || w.fx_cap == 0) return;
    const int at = atomicAdd(&w.fx_count[b], 1);
    if (at < w.fx_cap) { fx::Entry& e = w.fx_list[(long long)b * w.fx_cap + at]; e.row = row; e.word = half; }
}

static size_t carve(LargeWs& w, void* base, int batch, int n, bool rot) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t bn = (size_t)batch * n;
    w.words = (n + 63) / 64;
    const size_t o_keys = take(bn * 8), o_order = take(bn * 4), o_m = take((size_t)batch * 4);
    const size_t o_box = take(rot ? 0 : bn * 16), o_area = take(rot ? 0 : bn * 4), o_cls = take(rot ? 0 : bn * 4);
    const size_t o_rbox = take(rot ? bn * sizeof(RotBox) : 0);
    const size_t o_mask = take(bn * (size_t)w.words * 8), o_kept = take((size_t)batch * w.words * 8);
    const size_t o_rowpos = take(rot ? bn * 4 : 0);
    size_t npad_big = 0;
    if (n > 16384) { npad_big = 32768; while (npad_big < (size_t)n) npad_big <<= 1; }
    const size_t o_bk = take((size_t)batch * npad_big * 8), o_bp = take((size_t)batch * npad_big * 4);
    const bool sp = n <= kSpatialMaxN;            // buffers of the spatially ordered path
    w.aw = (w.words + 63) / 64;
    const size_t o_skeys = take(sp ? bn * 8 : 0), o_ros = take(sp ? bn * 4 : 0), o_sor = take(sp ? bn * 4 : 0);
    const size_t o_sos = take(sp ? bn * 4 : 0), o_rosl = take(sp ? bn * 4 : 0);
    const size_t o_hull = take(sp ? (size_t)batch * w.words * 16 : 0);
    const size_t o_tcls = take(sp ? (size_t)batch * w.words * 8 : 0);
    const size_t o_adj = take(sp ? (size_t)batch * w.words * w.aw * 8 : 0);
    const size_t o_diag = take(sp ? (size_t)batch * w.words * kTile * 8 : 0);
    const size_t o_adjb = take(sp ? (size_t)batch * w.words * kTile * w.aw * 8 : 0);
    w.fx_cap = sp ? 16 * n : 0;                     // entries (16 B each) per image: 16 non-empty half-words per row on average
    const size_t o_fx = take((size_t)batch * w.fx_cap * sizeof(fx::Entry));
    const bool bp = sp && rot;                      // broad / narrow phase buffers of the rotated path
    w.n16 = (n + 15) / 16; w.n32 = (n + 31) / 32;
    {
        const long long all = (long long)n * (n - 1) / 2, want = 256ll * n;     // 256 listed partners per box before the tile kernel takes over
        w.pair_cap = bp ? (int)(all < want ? all : want) : 0;
        if (w.pair_cap < 1) w.pair_cap = bp ? 1 : 0;
    }
    const size_t o_cull = take(bp ? bn * 16 : 0), o_h4 = take(bp ? bn * 16 : 0), o_ax = take(bp ? bn * 16 : 0);
    const size_t o_h16 = take(bp ? (size_t)batch * w.n16 * 16 : 0), o_h32 = take(bp ? (size_t)batch * w.n32 * 16 : 0);
    const size_t o_pairs = take((size_t)batch * w.pair_cap * 4), o_pcnt = take(bp ? (size_t)batch * 4 : 0);
    const size_t o_fxc = take(sp ? (size_t)batch * 4 : 0);
    const size_t o_p2 = take((size_t)batch * w.pair_cap * 4), o_p2c = take(bp ? (size_t)batch * 4 : 0);
    const size_t o_lzh = take(bp ? bn : 0), o_lzd = take(bp ? bn : 0);
    if (base) {
        char* p = static_cast<char*>(base);
        w.keys = (unsigned long long*)(p + o_keys); w.order = (int*)(p + o_order); w.m = (int*)(p + o_m);
        w.box = (float4*)(p + o_box); w.area = (float*)(p + o_area); w.cls = (int*)(p + o_cls);
        w.rbox = (RotBox*)(p + o_rbox);
        w.mask = (unsigned long long*)(p + o_mask); w.kept = (unsigned long long*)(p + o_kept);
        w.rowpos = (int*)(p + o_rowpos);
        w.bk = (unsigned long long*)(p + o_bk); w.bp = (int*)(p + o_bp);
        w.skeys = (unsigned long long*)(p + o_skeys); w.rank_of_spos = (int*)(p + o_ros); w.spos_of_rank = (int*)(p + o_sor);
        w.slot_of_spos = (int*)(p + o_sos); w.rank_of_slot = (int*)(p + o_rosl);
        w.tile_hull = (float4*)(p + o_hull);
        w.tile_cls = (int2*)(p + o_tcls);
        w.tile_adj = (unsigned long long*)(p + o_adj);
        w.diag_all = (unsigned long long*)(p + o_diag);
        w.adj_blk = (unsigned long long*)(p + o_adjb);
        w.fx_list = (fx::Entry*)(p + o_fx);
        w.cull4 = (float4*)(p + o_cull); w.hull4 = (float4*)(p + o_h4); w.axes4 = (float4*)(p + o_ax);
        w.hull16 = (float4*)(p + o_h16); w.hull32 = (float4*)(p + o_h32);
        w.pairs = (unsigned*)(p + o_pairs); w.pair_count = (int*)(p + o_pcnt);
        w.fx_count = (int*)(p + o_fxc);
        w.pairs2 = (unsigned*)(p + o_p2); w.pair2_count = (int*)(p + o_p2c);
        w.lz_high = (unsigned char*)(p + o_lzh); w.lz_dead = (unsigned char*)(p + o_lzd);
    }
    return off;
}

size_t large_workspace_bytes(int batch, int n, bool rot) {
    LargeWs w;
    return carve(w, nullptr, batch, n, rot);
}

// ---------------------------------------------------------------------------- keys
struct KeyParams {
    const float* scores; const void* cls; const int* counts; const int* src_idx;
    long long pitch; int n, cls_is_i64, use_cls; float thr;
    int* status;
    // spatially ordered path: Morton keys of the box centres, written next to the score keys so that both
    // sorts are independent of each other and run in ONE launch
    const float* boxes; int n_param, box_format;
    unsigned long long* skeys;
};

__device__ __forceinline__ unsigned part1by1(unsigned v) {      // spread the low 16 bits to the even positions
    v &= 0x0000ffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

__global__ void keys_kernel(KeyParams P, unsigned long long* keys, int* m) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = P.n;
    if (P.counts) { const int c = P.counts[b]; n = c < n ? (c < 0 ? 0 : c) : n; }
    bool valid = false;
    unsigned long long key = ~0ull, skey = ~0ull;
    if (i < n) {
        const float s = P.scores[(long long)b * P.pitch + i];
        if (s >= P.thr) {
            int c = 0;
            if (P.use_cls) {
                const long long ci = (long long)b * P.pitch + i;
                c = P.cls_is_i64 ? (int)reinterpret_cast<const long long*>(P.cls)[ci] : reinterpret_cast<const int*>(P.cls)[ci];
                if (c < 0 || c > MYDET_MAX_CLASS_ID) { if (P.status) atomicOr(P.status + b, 1); c = c < 0 ? 0 : MYDET_MAX_CLASS_ID; }
            }
            // equal scores rank by the candidate's flat index when the input is a compacted buffer
            // (slot order is arbitrary), else by the slot itself
            const unsigned tie = P.src_idx ? (unsigned)P.src_idx[(long long)b * P.pitch + i] : (unsigned)i;
            if (tie > 0xfffffu && P.status) atomicOr(P.status + b, 8);
            key = ((unsigned long long)c << 52) | ((unsigned long long)(~float_key(s)) << 20) | (unsigned long long)(tie & 0xfffffu);
            valid = true;
            if (P.skeys) {
                const float* bx = P.boxes + ((long long)b * P.pitch + i) * P.n_param;
                float cx = bx[0], cy = bx[1];
                if (P.box_format == MYDET_BOX_X1Y1X2Y2) { cx = 0.5f * (bx[0] + bx[2]); cy = 0.5f * (bx[1] + bx[3]); }
                const unsigned qx = (unsigned)fminf(fmaxf(cx, 0.f), 65535.f), qy = (unsigned)fminf(fmaxf(cy, 0.f), 65535.f);
                const unsigned morton = part1by1(qx) | (part1by1(qy) << 1);
                // class first (boxes of different classes never interact), then the Z-order of the centre;
                // equal codes in slot order (any order is valid: mask bits are directed by score rank).
                // Rotated boxes have no class: the field groups them by SIZE instead, so that a tile of 32 neighbours is
                // not inflated by the one 250-px box among 30-px ones (RAPiD candidates: 5 % of the boxes come from the
                // coarsest level, and 80 % of the tiles held one -- the broad phase then tested 60 % of ALL pairs)
                int group = c;
                if (!P.use_cls && P.n_param == 5) {
                    const float ext = fmaxf(bx[2], bx[3]);
                    group = ext < 96.f ? 0 : (ext < 256.f ? 1 : 2);
                }
                skey = ((unsigned long long)group << 52) | ((unsigned long long)morton << 20) | (unsigned long long)((unsigned)i & 0xfffffu);
            }
        }
    }
    if (i < P.n) {
        keys[(long long)b * P.n + i] = key;
        if (P.skeys) P.skeys[(long long)b * P.n + i] = skey;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(m + b, __popc(bal));
}

// ---------------------------------------------------------------------------- sort
// Sort for images of at most 16 384 candidates: one CTA per image, bitonic network over (key, slot) held
// entirely in shared memory (192 KB).  32 images sort concurrently on 32 SMs in ~230 us, where an
// all-pairs rank kernel (rank = number of smaller keys) needed 545 us for the same batch.
constexpr int kSortThreads = 1024;
constexpr int kSortMaxN = 16384;
// grid = B CTAs, or 2 B when a second, independent key array is sorted by the same launch (score keys + Morton
// keys: 64 CTAs of a 32-image batch instead of twice 32).  inv (first array only, optional): inverse permutation.
__global__ void __launch_bounds__(kSortThreads, 1) sort_smem_kernel(const unsigned long long* keys, int* order, int* inv,
                                                                    const unsigned long long* keys2, int* order2,
                                                                    int n, int npad, int B) {
    extern __shared__ unsigned long long skeys[];
    int* spay = reinterpret_cast<int*>(skeys + npad);
    const bool second = (int)blockIdx.x >= B;
    const int b = second ? blockIdx.x - B : blockIdx.x, tid = threadIdx.x;
    const unsigned long long* kb = (second ? keys2 : keys) + (long long)b * n;
    int* ob = (second ? order2 : order) + (long long)b * n;
    for (int i = tid; i < npad; i += kSortThreads) { skeys[i] = (i < n) ? kb[i] : ~0ull; spay[i] = i; }
    __syncthreads();
#pragma unroll 1
    for (int size = 2; size <= npad; size <<= 1) {
#pragma unroll 1
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll 4
            for (int t = tid; t < (npad >> 1); t += kSortThreads) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = skeys[lo], c = skeys[hi];
                if ((a > c) == up) {
                    skeys[lo] = c; skeys[hi] = a;
                    const int pa = spay[lo]; spay[lo] = spay[hi]; spay[hi] = pa;
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += kSortThreads)
        if (skeys[i] != ~0ull) {
            ob[i] = spay[i];
            if (inv && !second) inv[(long long)b * n + spay[i]] = i;
        }
}

// The same network for npad == 16 384 with the data in REGISTERS: thread t owns the 16 consecutive elements
// 16t .. 16t+15.  Strides 1..8 are compare-exchanges inside a thread, strides 16..256 are warp shuffles with lane
// t ^ (stride/16), and only strides >= 512 (15 of the 105 steps) go through shared memory.  The all-shared-memory
// kernel above is bound by shared-memory bandwidth (2 x 12 B per element and step): 244 us -> ~90 us per launch.
constexpr int kRegSortN = 16384, kRegSortE = 16;
__device__ __forceinline__ int sort_pad(int g) { return g + (g >> 5); }       // 1 pad word per 32: blocked access is 2-way at worst
// PIK ("payload in key"): the low 20 bits of the key ARE the element's slot (Morton keys always; score keys when no
// src_idx remaps the tie index), so only the 16 keys are carried (32 registers instead of 48: no spills at 1024 threads).
template <int S, bool PIK>
__device__ __forceinline__ void sort_reg_step(unsigned long long (&k)[kRegSortE], int (&p)[kRegSortE], int g0, int size) {
#pragma unroll
    for (int e = 0; e < kRegSortE; ++e) {
        if ((e & S) == 0) {
            const bool up = ((g0 + e) & size) == 0;
            const unsigned long long a = k[e], c = k[e + S];
            if ((a > c) == up) {
                k[e] = c; k[e + S] = a;
                if (!PIK) { const int t = p[e]; p[e] = p[e + S]; p[e + S] = t; }
            }
        }
    }
}
template <bool PIK>
__device__ __forceinline__ void sort_reg16k_body(const unsigned long long* kb, int* ob, int* invb, int n,
                                                 unsigned long long* skeys, int* spay) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int g0 = tid * kRegSortE;
    unsigned long long k[kRegSortE];
    int p[kRegSortE];
#pragma unroll
    for (int e = 0; e < kRegSortE; ++e) { const int g = g0 + e; k[e] = (g < n) ? kb[g] : ~0ull; p[e] = g; }
#pragma unroll 1
    for (int size = 2; size <= kRegSortN; size <<= 1) {
        if (size >= 1024) {
            // strides size/2 .. 512 in shared memory (8 pairs per thread), then back to the registers
            __syncthreads();                                           // everybody has read the previous contents
#pragma unroll
            for (int e = 0; e < kRegSortE; ++e) { skeys[sort_pad(g0 + e)] = k[e]; if (!PIK) spay[sort_pad(g0 + e)] = p[e]; }
            __syncthreads();
#pragma unroll 1
            for (int stride = size >> 1; stride >= 512; stride >>= 1) {
#pragma unroll 4
                for (int t = tid; t < (kRegSortN >> 1); t += kSortThreads) {
                    const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                    const bool up = (lo & size) == 0;
                    const int il = sort_pad(lo), ih = sort_pad(hi);
                    const unsigned long long a = skeys[il], c = skeys[ih];
                    if ((a > c) == up) {
                        skeys[il] = c; skeys[ih] = a;
                        if (!PIK) { const int pa = spay[il]; spay[il] = spay[ih]; spay[ih] = pa; }
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int e = 0; e < kRegSortE; ++e) { k[e] = skeys[sort_pad(g0 + e)]; if (!PIK) p[e] = spay[sort_pad(g0 + e)]; }
        }
        // strides 256 .. 16: the partner element lives in lane ^ d, same slot
#pragma unroll 1
        for (int d = min(size >> 5, 16); d >= 1; d >>= 1) {
            const bool lower = (lane & d) == 0;
#pragma unroll
            for (int e = 0; e < kRegSortE; ++e) {
                const bool up = ((g0 + e) & size) == 0;
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, k[e], d);
                int op = 0;
                if (!PIK) op = __shfl_xor_sync(0xffffffffu, p[e], d);
                const bool take = (lower == up) ? (other < k[e]) : (other > k[e]);     // keep the min / the max
                if (take) { k[e] = other; if (!PIK) p[e] = op; }
            }
        }
        if (size >= 16) sort_reg_step<8, PIK>(k, p, g0, size);
        if (size >= 8) sort_reg_step<4, PIK>(k, p, g0, size);
        if (size >= 4) sort_reg_step<2, PIK>(k, p, g0, size);
        sort_reg_step<1, PIK>(k, p, g0, size);
    }
#pragma unroll
    for (int e = 0; e < kRegSortE; ++e) {
        const int g = g0 + e;
        if (g < n && k[e] != ~0ull) {
            const int slot = PIK ? (int)(k[e] & 0xfffffull) : p[e];
            ob[g] = slot;
            if (invb) invb[slot] = g;
        }
    }
}
// grid = B CTAs, or 2 B (score keys, then Morton keys: always payload-in-key); pik1: the first array is too
__global__ void __launch_bounds__(kSortThreads, 1) sort_reg16k_kernel(const unsigned long long* keys, int* order, int* inv,
                                                                      const unsigned long long* keys2, int* order2, int n, int B,
                                                                      int pik1) {
    extern __shared__ unsigned long long skeys[];                      // sort_pad(16384) keys, then as many payloads
    int* spay = reinterpret_cast<int*>(skeys + sort_pad(kRegSortN));
    const bool second = (int)blockIdx.x >= B;
    const int b = second ? blockIdx.x - B : blockIdx.x;
    const unsigned long long* kb = (second ? keys2 : keys) + (long long)b * n;
    int* ob = (second ? order2 : order) + (long long)b * n;
    int* invb = (inv && !second) ? inv + (long long)b * n : nullptr;
    if (second || pik1) sort_reg16k_body<true>(kb, ob, invb, n, skeys, spay);
    else sort_reg16k_body<false>(kb, ob, invb, n, skeys, spay);
}

// LSD radix sort, one CTA per image (n <= 16 384), for keys whose low 20 bits ARE the element's slot and whose
// initial order is slot order ("payload in key": Morton keys always, score keys when no src_idx remaps the tie
// index).  A stable sort then only has to look at the bits above the tie field: 4 passes of 8 bits over the 32-bit score
// (or Morton) field, plus 2 passes of 6 bits over the class field when any key has a class -- 4 to 6 passes with three
// barriers each, where the bitonic network above needs 105 compare-exchange stages (153 us at 16 384 keys).
// Keys live in registers in WARP-STRIPED order (round r of warp w, lane l: element 512 w + 32 r + l -- 1024 E per
// round for E < 16), so that within a warp, round after round, lane after lane is index order: the rank of a key among
// the equal digits before it = the warp's running count of that digit (one shared-memory word per warp and digit,
// updated by the leader of each __match_any_sync group) + its position inside the group; an exclusive scan of the
// 256 x 32 counters in (digit, warp) order turns that into the global rank, and the keys are scattered through shared
// memory.  Invalid keys (~0) have the largest digit in every pass and stay at the end.
constexpr int kRadixPitch = 257;                 // counters [warp][digit], padded: conflict-free for both access patterns
template <int E>
__global__ void __launch_bounds__(kSortThreads, 1) sort_radix_kernel(const unsigned long long* keys, int* order, int* inv,
                                                                     const unsigned long long* keys2, int* order2, int n, int B) {
    extern __shared__ unsigned long long skeys[];                      // 1024 E keys, then the counters
    unsigned* cnt = reinterpret_cast<unsigned*>(skeys + 1024 * E);     // 32 x 257
    __shared__ unsigned s_warp_tot[32];
    __shared__ unsigned s_cls_or;
    const bool second = (int)blockIdx.x >= B;
    const int b = second ? blockIdx.x - B : blockIdx.x;
    const unsigned long long* kb = (second ? keys2 : keys) + (long long)b * n;
    int* ob = (second ? order2 : order) + (long long)b * n;
    int* invb = (inv && !second) ? inv + (long long)b * n : nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int base = warp * (32 * E) + lane;                           // element of round r: base + 32 r
    unsigned long long k[E];
    unsigned cls_or = 0;
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int g = base + 32 * r;
        k[r] = (g < n) ? kb[g] : ~0ull;
        if (k[r] != ~0ull) cls_or |= (unsigned)(k[r] >> 52);
    }
    if (tid == 0) s_cls_or = 0;
    __syncthreads();
    cls_or = __reduce_or_sync(0xffffffffu, cls_or);
    if (lane == 0 && cls_or) atomicOr(&s_cls_or, cls_or);
    __syncthreads();
    const int n_pass = s_cls_or ? (s_cls_or < 64u ? 5 : 6) : 4;      // class ids below 64 (size groups, few classes): one 6-bit pass
#pragma unroll 1
    for (int pass = 0; pass < n_pass; ++pass) {
        const int shift = pass < 4 ? 20 + 8 * pass : 52 + 6 * (pass - 4);
        const unsigned dmask = pass < 4 ? 255u : 63u;
        for (int i = tid; i < 32 * kRadixPitch; i += kSortThreads) cnt[i] = 0u;
        __syncthreads();
        unsigned pre2[E / 2];                                          // 16-bit ranks, two per register (64-register budget)
        unsigned* mycnt = cnt + warp * kRadixPitch;
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const unsigned d = (unsigned)(k[r] >> shift) & dmask;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const int leader = __ffs(peers) - 1;
            unsigned old = 0;
            if (lane == leader) { old = mycnt[d]; mycnt[d] = old + __popc(peers); }
            old = __shfl_sync(0xffffffffu, old, leader);
            const unsigned pr = old + __popc(peers & lt);
            if (r & 1) pre2[r >> 1] |= pr << 16; else pre2[r >> 1] = pr;
            __syncwarp();
        }
        __syncthreads();
        // exclusive scan over the counters in (digit major, warp minor) order: thread t owns digit t/4, warps 8 (t%4) .. +8
        {
            const int d = tid >> 2, w0 = (tid & 3) * 8;
            unsigned v[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { v[j] = cnt[(w0 + j) * kRadixPitch + d]; sum += v[j]; }
            unsigned incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (lane == 31) s_warp_tot[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const unsigned t0 = s_warp_tot[lane];
                unsigned wi = t0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
                s_warp_tot[lane] = wi - t0;
            }
            __syncthreads();
            unsigned run = s_warp_tot[warp] + incl - sum;
#pragma unroll
            for (int j = 0; j < 8; ++j) { cnt[(w0 + j) * kRadixPitch + d] = run; run += v[j]; }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const unsigned d = (unsigned)(k[r] >> shift) & dmask;
            skeys[mycnt[d] + ((pre2[r >> 1] >> (16 * (r & 1))) & 0xffffu)] = k[r];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < E; ++r) k[r] = skeys[base + 32 * r];
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int g = base + 32 * r;
        if (g < n && k[r] != ~0ull) {
            const int slot = (int)(k[r] & 0xfffffull);
            ob[g] = slot;
            if (invb) invb[slot] = g;
        }
    }
}
template <int E>
static int launch_sort_radix(const unsigned long long* keys, int* order, int* inv, const unsigned long long* keys2, int* order2,
                             int n, int B, cudaStream_t st) {
    const size_t smem = (size_t)1024 * E * 8 + (size_t)32 * kRadixPitch * 4;
    MYDET_CUDA(cudaFuncSetAttribute(sort_radix_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sort_radix_kernel<E><<<keys2 ? 2 * B : B, kSortThreads, smem, st>>>(keys, order, inv, keys2, order2, n, B);
    return launch_status("sort_radix_kernel");
}

// Sort for larger images (16 384 < n <= 2^20): the same bitonic network split over CTAs.  16 384-key chunks
// are sorted / merged in shared memory (strides < 16 384), the few stages with longer strides are
// compare-exchanges in global memory.  Replaces the O(n^2) all-pairs rank kernel (2.3e9 compares at 48 k).
constexpr int kChunk = kSortMaxN;
__global__ void sort_big_load_kernel(const unsigned long long* keys, unsigned long long* bk, int* bp, int n, int npad) {
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    bk[(long long)b * npad + i] = (i < n) ? keys[(long long)b * n + i] : ~0ull;
    bp[(long long)b * npad + i] = i;
}
// first != 0: full network for sizes 2..kChunk; else: the strides kChunk/2..1 of the merge step `size`
__global__ void __launch_bounds__(kSortThreads, 1) sort_big_chunk_kernel(unsigned long long* bk, int* bp, int npad, int size_arg, int first) {
    extern __shared__ unsigned long long skeys[];
    int* spay = reinterpret_cast<int*>(skeys + kChunk);
    const int c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    unsigned long long* gk = bk + (long long)b * npad + (long long)c * kChunk;
    int* gp = bp + (long long)b * npad + (long long)c * kChunk;
    for (int i = tid; i < kChunk; i += kSortThreads) { skeys[i] = gk[i]; spay[i] = gp[i]; }
    __syncthreads();
    const int gbase = c * kChunk;
#pragma unroll 1
    for (int size = first ? 2 : size_arg; size <= (first ? kChunk : size_arg); size <<= 1) {
#pragma unroll 1
        for (int stride = min(size >> 1, kChunk >> 1); stride > 0; stride >>= 1) {
#pragma unroll 4
            for (int t = tid; t < (kChunk >> 1); t += kSortThreads) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (((first == 2 ? 0 : gbase) + lo) & size) == 0;     // first == 2: every chunk ascending (rank merge)
                const unsigned long long a = skeys[lo], d = skeys[hi];
                if ((a > d) == up) {
                    skeys[lo] = d; skeys[hi] = a;
                    const int pa = spay[lo]; spay[lo] = spay[hi]; spay[hi] = pa;
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < kChunk; i += kSortThreads) { gk[i] = skeys[i]; gp[i] = spay[i]; }
}
__global__ void sort_big_global_kernel(unsigned long long* bk, int* bp, int npad, int size, int stride) {
    const int b = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (npad >> 1)) return;
    const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
    const bool up = (lo & size) == 0;
    unsigned long long* gk = bk + (long long)b * npad;
    int* gp = bp + (long long)b * npad;
    const unsigned long long a = gk[lo], d = gk[hi];
    if ((a > d) == up) {
        gk[lo] = d; gk[hi] = a;
        const int pa = gp[lo]; gp[lo] = gp[hi]; gp[hi] = pa;
    }
}
__global__ void sort_big_store_kernel(const unsigned long long* bk, const int* bp, int* order, int n, int npad) {
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && bk[(long long)b * npad + i] != ~0ull) order[(long long)b * n + i] = bp[(long long)b * npad + i];
}

__global__ void invert_kernel(const int* order, const int* m, int* inv, int n) {
    const int b = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < m[b]) inv[(long long)b * n + order[(long long)b * n + r]] = r;
}

// Merge of the sorted 16 384-key chunks of an image BY RANKING: keys are unique (their tie field is), so the final
// position of an element is its index inside its own chunk plus, for every other chunk, the number of keys below it --
// one binary search per other chunk (the chunks of an image are 128 KB each and sit in L2).  Replaces the log2(chunks)
// bitonic merge steps, each of which was a full pass of the shared-memory chunk kernel plus the long-stride global
// stages (48 384 boxes: 3 chunk-kernel launches of ~130 us per key array -> 1 + this kernel).
__global__ void __launch_bounds__(256) sort_big_rank_kernel(const unsigned long long* __restrict__ bk, const int* __restrict__ bp,
                                                            int* __restrict__ order, int n, int npad) {
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    const unsigned long long* gk = bk + (long long)b * npad;
    const unsigned long long key = gk[i];
    if (key == ~0ull) return;                                   // padding / invalid: left out of the order
    const int chunks = npad / kChunk, mine = i / kChunk;
    int rank = i - mine * kChunk;
    for (int c = 0; c < chunks; ++c) {
        if (c == mine) continue;
        const unsigned long long* ck = gk + (long long)c * kChunk;
        int lo = 0, hi = kChunk;                                // first index with ck[idx] >= key  (= keys below `key`)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (ck[mid] < key) lo = mid + 1; else hi = mid;
        }
        rank += lo;
    }
    if (rank < n) order[(long long)b * n + rank] = bp[(long long)b * npad + i];
}

static int sort_big(const unsigned long long* keys, int* order, int n, int B, LargeWs& w, cudaStream_t st) {
    int npad = 2 * kChunk;
    while (npad < n) npad <<= 1;
    const int chunks = npad / kChunk;
    MYDET_CUDA(cudaFuncSetAttribute(sort_big_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kChunk * 12)));
    sort_big_load_kernel<<<dim3((npad + 255) / 256, B), 256, 0, st>>>(keys, w.bk, w.bp, n, npad);
    const char* menv = getenv("MYDET_SORT_BIG_MERGE");          // developer A/B switch: 1 = the bitonic merge steps
    const bool bitonic_merge = menv && menv[0] == '1';
    // first = 1: chunks alternate ascending / descending (a bitonic sequence for the merge steps); 2: all ascending
    sort_big_chunk_kernel<<<dim3(chunks, B), kSortThreads, (size_t)kChunk * 12, st>>>(w.bk, w.bp, npad, 0, bitonic_merge ? 1 : 2);
    if (bitonic_merge) {
        for (int size = 2 * kChunk; size <= npad; size <<= 1) {
            for (int stride = size >> 1; stride >= kChunk; stride >>= 1)
                sort_big_global_kernel<<<dim3((npad / 2 + 255) / 256, B), 256, 0, st>>>(w.bk, w.bp, npad, size, stride);
            sort_big_chunk_kernel<<<dim3(chunks, B), kSortThreads, (size_t)kChunk * 12, st>>>(w.bk, w.bp, npad, size, 0);
        }
        sort_big_store_kernel<<<dim3((n + 255) / 256, B), 256, 0, st>>>(w.bk, w.bp, order, n, npad);
    } else {
        sort_big_rank_kernel<<<dim3((npad + 255) / 256, B), 256, 0, st>>>(w.bk, w.bp, order, n, npad);
    }
    return launch_status("sort_big kernels");
}

static bool sort_force_bitonic() {            // developer A/B switch (tests compare both sorts)
    const char* e = getenv("MYDET_SORT_BITONIC");
    return e && e[0] == '1';
}
// order[b][rank] = index of the rank-th smallest key of image b (invalid keys ~0 are left out); inv (optional)
// = its inverse.  keys2 / order2 (optional): a second key array of the same shape, sorted alongside.
static int sort_keys(const unsigned long long* keys, int* order, int* inv, const unsigned long long* keys2, int* order2,
                     int n, int B, LargeWs& w, cudaStream_t st, bool pik1 = false) {
    if (n <= kSortMaxN && n > 1024 && pik1 && !sort_force_bitonic()) {
        if (n <= 2048) return launch_sort_radix<2>(keys, order, inv, keys2, order2, n, B, st);
        if (n <= 4096) return launch_sort_radix<4>(keys, order, inv, keys2, order2, n, B, st);
        if (n <= 8192) return launch_sort_radix<8>(keys, order, inv, keys2, order2, n, B, st);
        return launch_sort_radix<16>(keys, order, inv, keys2, order2, n, B, st);
    }
    if (n <= kSortMaxN) {
        int npad = 64;
        while (npad < n) npad <<= 1;
        if (npad == kRegSortN) {
            const size_t smem = (size_t)(kRegSortN + kRegSortN / 32) * 12;
            MYDET_CUDA(cudaFuncSetAttribute(sort_reg16k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sort_reg16k_kernel<<<keys2 ? 2 * B : B, kSortThreads, smem, st>>>(keys, order, inv, keys2, order2, n, B, pik1 ? 1 : 0);
            return launch_status("sort_reg16k_kernel");
        }
        MYDET_CUDA(cudaFuncSetAttribute(sort_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSortMaxN * 12)));
        sort_smem_kernel<<<keys2 ? 2 * B : B, kSortThreads, (size_t)npad * 12, st>>>(keys, order, inv, keys2, order2, n, npad, B);
        return launch_status("sort_smem_kernel");
    }
    int rc = sort_big(keys, order, n, B, w, st);
    if (rc) return rc;
    if (inv) invert_kernel<<<dim3((n + 255) / 256, B), 256, 0, st>>>(order, w.m, inv, n);
    if (keys2) rc = sort_big(keys2, order2, n, B, w, st);
    return rc;
}

// ---------------------------------------------------------------------------- gather
struct GatherParams {
    const float* boxes; long long pitch; int n, n_param, box_format;
};
template <bool ROT>
__global__ void gather_kernel(GatherParams P, const unsigned long long* keys, const int* order, const int* m, LargeWs w) {
    const int b = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m[b]) return;
    const long long row = (long long)b * P.n + r;
    const int i = order[row];
    const float* bx = P.boxes + ((long long)b * P.pitch + i) * P.n_param;
    if (ROT) {
        float v[5] = {bx[0], bx[1], bx[2], bx[3], bx[4]};
        RotBox q;
        make_rot_box(v, q.x, q.y, q.r);
        q.cx = v[0]; q.cy = v[1];
        q.area2 = (float)signed_area2_f64(q.x, q.y);
        rot_box_hull(q);
        w.rbox[row] = q;
    } else {
        const float v0 = bx[0], v1 = bx[1], v2 = bx[2], v3 = bx[3];
        float4 c4;
        if (P.box_format == MYDET_BOX_CXCYWH) {
            const float hw = __fmul_rn(v2, 0.5f), hh = __fmul_rn(v3, 0.5f);
            c4 = make_float4(__fsub_rn(v0, hw), __fsub_rn(v1, hh), __fadd_rn(v0, hw), __fadd_rn(v1, hh));
        } else {
            c4 = make_float4(v0, v1, v2, v3);
        }
        w.box[row] = c4;
        w.area[row] = __fmul_rn(__fsub_rn(c4.z, c4.x), __fsub_rn(c4.w, c4.y));
        w.cls[row] = (int)(keys[(long long)b * P.n + i] >> 52);
    }
}

// ---------------------------------------------------------------------------- mask
// Axis-aligned mask.  grid (upper-triangular tile pair, image); 64 threads; thread t owns row row_tile*64+t.
__global__ void __launch_bounds__(kTile) mask_kernel(LargeWs w, const int* m, int n, float thr_f) {
    // blockIdx.x enumerates the upper-triangular tile pairs (rt <= ct) row by row
    const int T = (n + kTile - 1) / kTile;
    const int b = blockIdx.y;
    int rt = (int)((2.0f * T + 1.0f - sqrtf((2.0f * T + 1.0f) * (2.0f * T + 1.0f) - 8.0f * (float)blockIdx.x)) * 0.5f);
    rt = max(0, min(rt, T - 1));
    while (rt > 0 && (long long)rt * T - (long long)rt * (rt - 1) / 2 > (long long)blockIdx.x) --rt;
    while ((long long)(rt + 1) * T - (long long)(rt + 1) * rt / 2 <= (long long)blockIdx.x) ++rt;
    const int ct = rt + (int)((long long)blockIdx.x - ((long long)rt * T - (long long)rt * (rt - 1) / 2));
    const int mb = m[b];
    if (rt * kTile >= mb || ct * kTile >= mb) return;
    const long long base = (long long)b * n;
    const int t = threadIdx.x;
    const int r = rt * kTile + t;
    const int c0 = ct * kTile;
    unsigned long long bits = 0ull;
    {
        __shared__ float4 cbox[kTile];
        __shared__ float carea[kTile];
        __shared__ int ccls[kTile];
        if (c0 + t < mb) { cbox[t] = w.box[base + c0 + t]; carea[t] = w.area[base + c0 + t]; ccls[t] = w.cls[base + c0 + t]; }
        __syncthreads();
        if (r < mb) {
            const float4 a = w.box[base + r];
            const float aarea = w.area[base + r];
            const int ac = w.cls[base + r];
            const int lim = min(kTile, mb - c0);
            // sorted by class: the tile can only match if its class range reaches ac
            if (ccls[0] <= ac && ccls[lim - 1] >= ac) {
#pragma unroll 4
                for (int j = 0; j < lim; ++j) {
                    const float4 c4 = cbox[j];
                    if (c0 + j > r && ccls[j] == ac && iou_corners_gt(a.x, a.y, a.z, a.w, aarea, c4.x, c4.y, c4.z, c4.w, carea[j], thr_f))
                        bits |= 1ull << j;
                }
            }
        }
    }
    if (r < mb) w.mask[(base + r) * w.words + ct] = bits;
}

// Rotated mask: one CTA owns a row tile and walks kColChunk consecutive column tiles.
//  (1) per column tile: one coalesced 1 KB load of cull records, then every thread runs 64 branch-free
//      circumscribed-circle tests for its row; the few survivors pass the area-ratio and hull bounds and
//      are appended to a queue in shared memory;
//  (2) the queue is drained once it holds at least one entry per thread (or at the end of the chunk):
//      one pair per thread, polygon clipping with every lane busy.
// History (profiles/r1_history.md): clipping inside the per-row loop ran with ~1 of 32 lanes active
// (585 us per 10 k-box image); one CTA per 64x64 tile spent its time on set-up and barriers (209 us);
// draining after every tile clipped ~10 pairs with 64 threads (136 us).
constexpr int kColChunk = 8;
constexpr int kQueueCap = kTile * kTile + kTile;
__global__ void __launch_bounds__(kTile) mask_rot_kernel(LargeWs w, const int* m, int n, double thr_d, int ge) {
    const int rt = blockIdx.x, cc = blockIdx.y, b = blockIdx.z;
    const int T = (n + kTile - 1) / kTile;
    const int ct_lo = max(rt, cc * kColChunk), ct_hi = min(T, (cc + 1) * kColChunk);
    if (ct_lo >= ct_hi) return;
    const int mb = m[b];
    if (rt * kTile >= mb || ct_lo * kTile >= mb) return;
    const long long base = (long long)b * n;
    const int t = threadIdx.x;
    const int r = rt * kTile + t;
    __shared__ float4 ccull[2][kTile];
    __shared__ unsigned short queue[kQueueCap];          // (tile in chunk) << 12 | row << 6 | col
    __shared__ unsigned bits32[kColChunk][kTile][2];
    __shared__ int qn;
    const bool ge_mode = ge != 0;
    const float thr_f = (float)thr_d;
    float mcx = 3.0e18f, mcy = 3.0e18f, mr = 0.f, ma = 0.f, mx0 = 0.f, my0 = 0.f, mx1 = 0.f, my1 = 0.f;
    if (r < mb) {
        const RotBox q = w.rbox[base + r];
        mcx = q.cx; mcy = q.cy; mr = q.r * 1.00001f + 1e-3f; ma = 0.5f * fabsf(q.area2);
        mx0 = q.x0; my0 = q.y0; mx1 = q.x1; my1 = q.y1;
    }
#pragma unroll
    for (int k = 0; k < kColChunk; ++k) { bits32[k][t][0] = 0u; bits32[k][t][1] = 0u; }
    if (t == 0) qn = 0;

    auto drain = [&](int total) {
        for (int e = t; e < total; e += kTile) {
            const int code = queue[e], k = code >> 12, rt_ = (code >> 6) & 63, j = code & 63;
            const RotBox A = w.rbox[base + rt * kTile + rt_];
            const RotBox B = w.rbox[base + (ct_lo + k) * kTile + j];
            if (rot_overlaps(A, B, thr_d, ge_mode)) atomicOr(&bits32[k][rt_][j >> 5], 1u << (j & 31));
        }
    };

    int buf = 0;
#pragma unroll 1
    for (int ct = ct_lo; ct < ct_hi; ++ct, buf ^= 1) {
        const int c0 = ct * kTile;
        if (c0 >= mb) break;
        {
            float4 c = make_float4(-3.0e18f, -3.0e18f, 0.f, 0.f);              // never passes the circle test
            if (c0 + t < mb) {
                const RotBox& q = w.rbox[base + c0 + t];
                c = make_float4(q.cx, q.cy, q.r, 0.5f * fabsf(q.area2));
            }
            ccull[buf][t] = c;
        }
        __syncthreads();                       // cull records visible; the previous tile's pushes are complete
        const int pending = qn;
        if (pending >= kTile) {                // uniform
            drain(pending);
            __syncthreads();
            if (t == 0) qn = 0;
            __syncthreads();
        }
        if (r < mb) {
            unsigned cand_lo = 0u, cand_hi = 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float4 c = ccull[buf][j];
                const float dx = mcx - c.x, dy = mcy - c.y, rr = fmaf(c.z, 1.00001f, mr);
                if (fmaf(dx, dx, dy * dy) <= rr * rr) cand_lo |= 1u << j;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float4 c = ccull[buf][32 + j];
                const float dx = mcx - c.x, dy = mcy - c.y, rr = fmaf(c.z, 1.00001f, mr);
                if (fmaf(dx, dx, dy * dy) <= rr * rr) cand_hi |= 1u << j;
            }
            unsigned long long cand = ((unsigned long long)cand_hi << 32) | cand_lo;
            const int first = r + 1 - c0;                                       // only columns ranked after the row
            if (first >= kTile) cand = 0ull; else if (first > 0) cand &= ~0ull << first;
            while (cand) {
                const int j = __ffsll((long long)cand) - 1;
                cand &= cand - 1;
                const float oa = ccull[buf][j].w;
                const float lo = fminf(ma, oa), hi = fmaxf(ma, oa);
                if (lo * 1.0001f < thr_f * hi) continue;                         // IoU <= lo/hi < thr
                const RotBox& q = w.rbox[base + c0 + j];
                const float ix = fminf(mx1, q.x1) - fmaxf(mx0, q.x0);
                const float iy = fminf(my1, q.y1) - fmaxf(my0, q.y0);
                if (!(ix > -1e-3f && iy > -1e-3f)) continue;                     // hulls apart: IoU == 0
                const float ub = (ix + 2e-3f) * (iy + 2e-3f);
                if (ub * 1.0001f < thr_f * (ma + oa - ub)) continue;             // IoU <= ub/(a+b-ub) < thr
                queue[atomicAdd(&qn, 1)] = (unsigned short)(((ct - ct_lo) << 12) | (t << 6) | j);
            }
        }
    }
    __syncthreads();
    drain(qn);
    __syncthreads();
    if (r < mb) {
#pragma unroll 1
        for (int ct = ct_lo; ct < ct_hi; ++ct) {
            const int c0 = ct * kTile;
            if (c0 >= mb) break;
            unsigned long long bits = ((unsigned long long)bits32[ct - ct_lo][t][1] << 32) | bits32[ct - ct_lo][t][0];
            if (ge_mode ? (thr_d <= 0.0) : (thr_d < 0.0)) {   // degenerate threshold: every later box is suppressed
                bits = 0ull;
                const int lim = min(kTile, mb - c0);
                for (int j = 0; j < lim; ++j) if (c0 + j > r) bits |= 1ull << j;
            }
            w.mask[(base + r) * w.words + ct] = bits;
        }
    }
}

// ---------------------------------------------------------------------------- spatially ordered rotated NMS
// With boxes in score order a 64-box tile is spread over the whole image, so every tile pair has to be
// culled pair by pair: 5e7 circle tests per 10 k-box image, 2.9 ms per 32-image batch even at 9 instructions
// per test.  Here the MASK lives in Morton order of the box centres: a tile is a compact patch, tile pairs
// whose hulls are disjoint are skipped outright (~90 % of them), and the bit for an overlapping pair is set
// in the row of the higher-scored box (rank comparison) instead of being implied by the tile position.
// The sweep still visits boxes in score order and gathers the bits it needs through the rank<->position maps.
// after both sorts: slot_of_spos[spos] = candidate slot (payload of the Morton sort), rank_of_slot = inverse of the
// score order.  Build the quads in spatial order, both rank <-> position maps, and the hull of every 64-box tile
// (one warp pair per tile).
template <bool ROT>
__global__ void __launch_bounds__(kTile) spatial_gather_kernel(GatherParams P, const unsigned long long* keys, const int* order,
                                                               const int* m, LargeWs w) {
    const int tile = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const int mb = m[b];
    const int spos = tile * kTile + t;
    __shared__ float4 part[2];
    __shared__ int2 cpart[2];
    float x0 = 3.0e18f, y0 = 3.0e18f, x1 = -3.0e18f, y1 = -3.0e18f;
    int c_lo = 0x7fffffff, c_hi = -1;
    if (spos < mb) {
        const long long row = (long long)b * P.n + spos;
        const int i = w.slot_of_spos[row];
        const int r = w.rank_of_slot[(long long)b * P.n + i];
        w.rank_of_spos[row] = r;
        const float* bx = P.boxes + ((long long)b * P.pitch + i) * P.n_param;
        if (ROT) {
            float v[5] = {bx[0], bx[1], bx[2], bx[3], bx[4]};
            RotBox q;
            make_rot_box(v, q.x, q.y, q.r);
            q.cx = v[0]; q.cy = v[1];
            q.area2 = (float)signed_area2_f64(q.x, q.y);
            rot_box_hull(q);
            w.rbox[row] = q;
            x0 = q.x0; y0 = q.y0; x1 = q.x1; y1 = q.y1;
            w.cull4[row] = make_float4(q.cx, q.cy, q.r, 0.5f * fabsf(q.area2));
            w.hull4[row] = make_float4(q.x0, q.y0, q.x1, q.y1);
            w.axes4[row] = make_float4(0.5f * (q.x[1] - q.x[0]), 0.5f * (q.y[1] - q.y[0]), 0.5f * (q.x[0] - q.x[3]), 0.5f * (q.y[0] - q.y[3]));
        } else {
            const float v0 = bx[0], v1 = bx[1], v2 = bx[2], v3 = bx[3];
            float4 c4;
            if (P.box_format == MYDET_BOX_CXCYWH) {
                const float hw = __fmul_rn(v2, 0.5f), hh = __fmul_rn(v3, 0.5f);
                c4 = make_float4(__fsub_rn(v0, hw), __fsub_rn(v1, hh), __fadd_rn(v0, hw), __fadd_rn(v1, hh));
            } else {
                c4 = make_float4(v0, v1, v2, v3);
            }
            w.box[row] = c4;
            w.area[row] = __fmul_rn(__fsub_rn(c4.z, c4.x), __fsub_rn(c4.w, c4.y));
            const int c = (int)(keys[(long long)b * P.n + i] >> 52);
            w.cls[row] = c;
            c_lo = c_hi = c;
            x0 = fminf(c4.x, c4.z); y0 = fminf(c4.y, c4.w); x1 = fmaxf(c4.x, c4.z); y1 = fmaxf(c4.y, c4.w);
        }
        w.spos_of_rank[(long long)b * P.n + r] = spos;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o));
        x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
        c_lo = min(c_lo, __shfl_xor_sync(0xffffffffu, c_lo, o)); c_hi = max(c_hi, __shfl_xor_sync(0xffffffffu, c_hi, o));
        if (ROT && o == 8 && (t & 15) == 0 && tile * kTile + (t & ~15) < mb)      // groups of 16 positions are complete here
            w.hull16[(long long)b * w.n16 + tile * 4 + (t >> 4)] = make_float4(x0 - 1e-2f, y0 - 1e-2f, x1 + 1e-2f, y1 + 1e-2f);
    }
    if (ROT && (t & 31) == 0 && tile * kTile + t < mb)
        w.hull32[(long long)b * w.n32 + tile * 2 + (t >> 5)] = make_float4(x0 - 1e-2f, y0 - 1e-2f, x1 + 1e-2f, y1 + 1e-2f);
    if ((t & 31) == 0) { part[t >> 5] = make_float4(x0, y0, x1, y1); cpart[t >> 5] = make_int2(c_lo, c_hi); }
    __syncthreads();
    if (t == 0 && tile * kTile < mb) {
        w.tile_hull[(long long)b * w.words + tile] = make_float4(fminf(part[0].x, part[1].x) - 1e-2f, fminf(part[0].y, part[1].y) - 1e-2f,
                                                                 fmaxf(part[0].z, part[1].z) + 1e-2f, fmaxf(part[0].w, part[1].w) + 1e-2f);
        w.tile_cls[(long long)b * w.words + tile] = make_int2(min(cpart[0].x, cpart[1].x), max(cpart[0].y, cpart[1].y));
    }
}

// Axis-aligned mask in spatial order.  grid (row tile ti, chunk of column tiles, image); unordered tile pairs
// ti <= tj once; pairs of tiles with disjoint hulls or class ranges are skipped.  The IoU is torchvision's
// arithmetic (iou_corners); a hit sets the bit in the row of the higher-ranked box with a global atomicOr.
__global__ void __launch_bounds__(kTile) mask_aabb_spatial_kernel(LargeWs w, const int* m, int n, float thr_f) {
    const int ti = blockIdx.x, cc = blockIdx.y, b = blockIdx.z;
    const int T = (n + kTile - 1) / kTile;
    const int tj_lo = max(ti, cc * kColChunk), tj_hi = min(T, (cc + 1) * kColChunk);
    if (tj_lo >= tj_hi) return;
    const int mb = m[b];
    if (ti * kTile >= mb || tj_lo * kTile >= mb) return;
    const long long base = (long long)b * n;
    const int t = threadIdx.x;
    const int r = ti * kTile + t;
    __shared__ float4 cbox[2][kTile];
    __shared__ float carea[2][kTile];
    __shared__ int ccls[2][kTile];
    const float4 my_hull = w.tile_hull[(long long)b * w.words + ti];
    const int2 my_cls = w.tile_cls[(long long)b * w.words + ti];
    float4 a = make_float4(3.0e18f, 3.0e18f, 3.0e18f, 3.0e18f);
    float aarea = 0.f;
    int ac = -2, arank = 0;
    if (r < mb) { a = w.box[base + r]; aarea = w.area[base + r]; ac = w.cls[base + r]; arank = w.rank_of_spos[base + r]; }
    unsigned* mask32 = reinterpret_cast<unsigned*>(w.mask);
    int buf = 0;
#pragma unroll 1
    for (int tj = tj_lo; tj < tj_hi; ++tj) {
        const int c0 = tj * kTile;
        if (c0 >= mb) break;
        const float4 oh = w.tile_hull[(long long)b * w.words + tj];
        const int2 oc = w.tile_cls[(long long)b * w.words + tj];
        if (oh.x > my_hull.z || my_hull.x > oh.z || oh.y > my_hull.w || my_hull.y > oh.w) continue;   // patches apart
        if (oc.x > my_cls.y || my_cls.x > oc.y) continue;                                             // no common class
        if (c0 + t < mb) { cbox[buf][t] = w.box[base + c0 + t]; carea[buf][t] = w.area[base + c0 + t]; ccls[buf][t] = w.cls[base + c0 + t]; }
        else { cbox[buf][t] = make_float4(-3.0e18f, -3.0e18f, -3.0e18f, -3.0e18f); carea[buf][t] = 0.f; ccls[buf][t] = -1; }
        __syncthreads();
        if (r < mb) {
            const int j0 = (tj == ti) ? t + 1 : 0;                           // each unordered pair once
#pragma unroll 4
            for (int j = j0; j < kTile; ++j) {
                const float4 c4 = cbox[buf][j];
                // cheap reject: different class or no overlap at all (the common case)
                if (ccls[buf][j] != ac || !(fminf(a.z, c4.z) > fmaxf(a.x, c4.x)) || !(fminf(a.w, c4.w) > fmaxf(a.y, c4.y))) continue;
                const int pb = c0 + j;
                const bool a_first = arank < w.rank_of_spos[base + pb];
                // torchvision evaluates the pair from the higher-ranked box
                const float ovr = a_first ? iou_corners(a.x, a.y, a.z, a.w, aarea, c4.x, c4.y, c4.z, c4.w, carea[buf][j])
                                          : iou_corners(c4.x, c4.y, c4.z, c4.w, carea[buf][j], a.x, a.y, a.z, a.w, aarea);
                if (ovr > thr_f) {
                    const int row = a_first ? r : pb, col = a_first ? pb : r;
                    fx_note(w, b, row, col >> 5, atomicOr(&mask32[((base + row) * w.words + (col >> 6)) * 2 + ((col >> 5) & 1)], 1u << (col & 31)));
                    const int trow = row >> 6, tcol = col >> 6;
                    atomicOr(&w.tile_adj[((long long)b * w.words + trow) * w.aw + (tcol >> 6)], 1ull << (tcol & 63));
                }
            }
        }
        buf ^= 1;
    }
}

// grid (row tile ti, chunk of column tiles, image).  Unordered tile pairs ti <= tj are visited once; a hit
// sets the bit in the row of the higher-scored box with a global atomicOr (hits are rare); the mask is
// zeroed beforehand.  Cull / queue / drain exactly as mask_rot_kernel.
__device__ __forceinline__ void mask_rot_tile(const LargeWs& w, int mb, int n, int b, int ti, int cc, double thr_d, int ge) {
    const int T = (n + kTile - 1) / kTile;
    const int tj_lo = max(ti, cc * kColChunk), tj_hi = min(T, (cc + 1) * kColChunk);
    if (tj_lo >= tj_hi) return;
    if (ti * kTile >= mb || tj_lo * kTile >= mb) return;
    const long long base = (long long)b * n;
    const int t = threadIdx.x;
    const int r = ti * kTile + t;
    __shared__ float4 ccull[2][kTile];
    __shared__ unsigned short queue[kQueueCap];
    __shared__ int qn;
    const bool ge_mode = ge != 0;
    const float thr_f = (float)thr_d;
    const float4 my_hull = w.tile_hull[(long long)b * w.words + ti];
    float mcx = 3.0e18f, mcy = 3.0e18f, mr = 0.f, ma = 0.f, mx0 = 0.f, my0 = 0.f, mx1 = 0.f, my1 = 0.f;
    if (r < mb) {
        const RotBox q = w.rbox[base + r];
        mcx = q.cx; mcy = q.cy; mr = q.r * 1.00001f + 1e-3f; ma = 0.5f * fabsf(q.area2);
        mx0 = q.x0; my0 = q.y0; mx1 = q.x1; my1 = q.y1;
    }
    __syncthreads();                                  // a CTA may run several work items (overflow kernel): previous drain done
    if (t == 0) qn = 0;
    unsigned* mask32 = reinterpret_cast<unsigned*>(w.mask);

    auto drain = [&](int total) {
        for (int e = t; e < total; e += kTile) {
            const int code = queue[e], k = code >> 12, rt_ = (code >> 6) & 63, j = code & 63;
            const int pa = ti * kTile + rt_, pb = (tj_lo + k) * kTile + j;            // spatial positions
            const RotBox A = w.rbox[base + pa];
            const RotBox B = w.rbox[base + pb];
            if (rot_overlaps(A, B, thr_d, ge_mode)) {
                const bool a_first = w.rank_of_spos[base + pa] < w.rank_of_spos[base + pb];
                const int row = a_first ? pa : pb, col = a_first ? pb : pa;           // the higher-scored box suppresses
                fx_note(w, b, row, col >> 5, atomicOr(&mask32[((base + row) * w.words + (col >> 6)) * 2 + ((col >> 5) & 1)], 1u << (col & 31)));
                const int trow = row >> 6, tcol = col >> 6;
                atomicOr(&w.tile_adj[((long long)b * w.words + trow) * w.aw + (tcol >> 6)], 1ull << (tcol & 63));
            }
        }
    };

    int buf = 0;
#pragma unroll 1
    for (int tj = tj_lo; tj < tj_hi; ++tj) {
        const int c0 = tj * kTile;
        if (c0 >= mb) break;
        const float4 oh = w.tile_hull[(long long)b * w.words + tj];
        if (oh.x > my_hull.z || my_hull.x > oh.z || oh.y > my_hull.w || my_hull.y > oh.w) continue;   // patches apart
        {
            float4 c = make_float4(-3.0e18f, -3.0e18f, 0.f, 0.f);
            if (c0 + t < mb) {
                const RotBox& q = w.rbox[base + c0 + t];
                c = make_float4(q.cx, q.cy, q.r, 0.5f * fabsf(q.area2));
            }
            ccull[buf][t] = c;
        }
        __syncthreads();
        const int pending = qn;
        if (pending >= kTile) {
            drain(pending);
            __syncthreads();
            if (t == 0) qn = 0;
            __syncthreads();
        }
        if (r < mb) {
            unsigned cand_lo = 0u, cand_hi = 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float4 c = ccull[buf][j];
                const float dx = mcx - c.x, dy = mcy - c.y, rr = fmaf(c.z, 1.00001f, mr);
                // circle test AND area-ratio bound (IoU <= min / max of the two areas), both branch-free: the serial
                // per-candidate loop below was a third of the kernel's samples, the area bound halves its trips
                {
                    const bool pass = (fmaf(dx, dx, dy * dy) <= rr * rr) & (fminf(ma, c.w) * 1.0001f >= thr_f * fmaxf(ma, c.w));
                    cand_lo |= (pass ? 1u : 0u) << j;            // non-short-circuit: no branch per test
                }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float4 c = ccull[buf][32 + j];
                const float dx = mcx - c.x, dy = mcy - c.y, rr = fmaf(c.z, 1.00001f, mr);
                {
                    const bool pass = (fmaf(dx, dx, dy * dy) <= rr * rr) & (fminf(ma, c.w) * 1.0001f >= thr_f * fmaxf(ma, c.w));
                    cand_hi |= (pass ? 1u : 0u) << j;            // non-short-circuit: no branch per test
                }
            }
            unsigned long long cand = ((unsigned long long)cand_hi << 32) | cand_lo;
            if (tj == ti) cand = (t >= kTile - 1) ? 0ull : (cand & (~0ull << (t + 1)));   // each unordered pair once
            while (cand) {
                const int j = __ffsll((long long)cand) - 1;
                cand &= cand - 1;
                const float oa = ccull[buf][j].w;
                const RotBox& q = w.rbox[base + c0 + j];
                const float ix = fminf(mx1, q.x1) - fmaxf(mx0, q.x0);
                const float iy = fminf(my1, q.y1) - fmaxf(my0, q.y0);
                if (!(ix > -1e-3f && iy > -1e-3f)) continue;
                const float ub = (ix + 2e-3f) * (iy + 2e-3f);
                if (ub * 1.0001f < thr_f * (ma + oa - ub)) continue;
                queue[atomicAdd(&qn, 1)] = (unsigned short)(((tj - tj_lo) << 12) | (t << 6) | j);
            }
        }
        buf ^= 1;
    }
    __syncthreads();
    drain(qn);
}

__global__ void __launch_bounds__(kTile) mask_rot_spatial_kernel(LargeWs w, const int* m, int n, double thr_d, int ge) {
    mask_rot_tile(w, m[blockIdx.z], n, blockIdx.z, blockIdx.x, blockIdx.y, thr_d, ge);
}

// Backs up the broad / narrow phase kernels below: a fixed, small grid walks the (row tile, column chunk) work items of the
// images whose pair list overflowed (pair_count > pair_cap) -- normally none, and the launch costs a few microseconds.
__global__ void __launch_bounds__(kTile) mask_rot_overflow_kernel(LargeWs w, const int* m, int n, int batch, double thr_d, int ge) {
    const int T = (n + kTile - 1) / kTile, chunks = (T + kColChunk - 1) / kColChunk;
    for (int b = 0; b < batch; ++b) {
        if (w.pair_count[b] <= w.pair_cap) continue;
        const int mb = m[b];
        for (int item = blockIdx.x; item < T * chunks; item += gridDim.x)
            mask_rot_tile(w, mb, n, b, item % T, item / T, thr_d, ge);
    }
}

// ---------------------------------------------------------------------------- rotated mask: broad phase + narrow phase
// The tile kernel above spends its time waiting, not computing (ncu, round 1: 35 % of the warp slots active, 11 % of the
// samples on the CTA barrier in front of the queue drain, 15 % in the serial walk over the circle survivors that fetches
// every candidate's hull from L2).  Split:
//   broad  : ONE WARP per row tile of 32 spatial positions, no CTA barrier anywhere.  The lanes test the hulls of 32
//            column sub-tiles (16 positions each) at a time against the row tile's hull; for every overlapping sub-tile
//            the cull data (centre, radius, area) AND the hulls of its 16 boxes are staged in the warp's shared memory with
//            one 16-byte load per lane, every lane runs 16 branch-free circle + area-ratio tests for its row, and the
//            survivors go through the hull-overlap bound (shared memory, no L2 round trip) into a warp-private queue kept
//            with ballots (no atomics).  The queue is appended to the image's pair list with one global atomic per ~512 pairs.
//   narrow : one thread per listed pair: polygon clip (float32 in box-local coordinates, float64 re-check within 1e-3 of
//            the threshold -- rot_overlaps, unchanged) and the bit for the row of the higher-ranked box.  Every lane busy,
//            no queue, no barrier.
// 32 x 16 tiles cut the circle tests per 10 000-box image from 22 M (64 x 64 tiles) to 13 M (scripts/rot_mask_workload.py).
// A pair list that overflows (more than 256 listed partners per box on average) marks its image for the tile kernel.
constexpr int kBroadWarps = 4;
constexpr int kBroadQueue = 1024;          // entries per warp; flushed above kBroadQueue - 512 (a sub-tile adds at most 512)

__global__ void __launch_bounds__(kBroadWarps * 32) rot_broad_kernel(LargeWs w, const int* m, int n, float thr_f) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.y;
    const int R = blockIdx.x * kBroadWarps + warp;
    const int mb = m[b];
    if (R * 32 >= mb) return;                       // warps are independent: no CTA-wide barrier below
    __shared__ float4 s_col[kBroadWarps][2][32];    // [0,16): cull data, [16,32): hulls of the column sub-tile
    __shared__ float4 s_rowh[kBroadWarps][32];      // hulls of the 32 row boxes (refined sub-tile test below)
    __shared__ unsigned s_queue[kBroadWarps][kBroadQueue];
    const long long base = (long long)b * n;
    const int n16 = (mb + 15) >> 4;
    const int p = R * 32 + lane;
    const bool have = p < mb;
    float4 mc = make_float4(3.0e18f, 3.0e18f, 0.f, 0.f), mh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (have) { mc = w.cull4[base + p]; mh = w.hull4[base + p]; }
    const float mr = mc.z * 1.00001f + 1e-3f, ma = mc.w;
    const float4 rh = w.hull32[(long long)b * w.n32 + R];
    s_rowh[warp][lane] = have ? mh : make_float4(3.0e18f, 3.0e18f, -3.0e18f, -3.0e18f);     // an empty hull overlaps nothing
    __syncwarp();
    unsigned* queue = s_queue[warp];
    unsigned* pairs = w.pairs + (long long)b * w.pair_cap;
    int qn = 0;                                     // warp-uniform
    int buf = 0;

    auto flush = [&]() {
        int at = 0;
        if (lane == 0) at = atomicAdd(&w.pair_count[b], qn);
        at = __shfl_sync(0xffffffffu, at, 0);
        for (int e = lane; e < qn; e += 32)
            if (at + e < w.pair_cap) pairs[at + e] = queue[e];
        __syncwarp();
        qn = 0;
    };

    auto load_hull16 = [&](int cbase) {
        const int Cl = cbase + lane;
        return (Cl < n16) ? w.hull16[(long long)b * w.n16 + Cl] : make_float4(3.0e18f, 3.0e18f, -3.0e18f, -3.0e18f);
    };
    auto load_col = [&](int C) {                   // lanes 0-15: cull data, lanes 16-31: hulls of the 16 boxes of sub-tile C
        const int q = C * 16 + (lane & 15);
        float4 v = (lane < 16) ? make_float4(-3.0e18f, -3.0e18f, 0.f, 0.f) : make_float4(3.0e18f, 3.0e18f, -3.0e18f, -3.0e18f);
        if (q < mb) v = (lane < 16) ? w.cull4[base + q] : w.hull4[base + q];
        return v;
    };
    float4 h_next = load_hull16(2 * R);
#pragma unroll 1
    for (int cbase = 2 * R; cbase < n16; cbase += 32) {
        const float4 h = h_next;
        if (cbase + 32 < n16) h_next = load_hull16(cbase + 32);          // in flight while this group of 32 sub-tiles is processed
        bool ov = !(h.x > rh.z || rh.x > h.z || h.y > rh.w || rh.y > h.w);
        if (ov) {
            // the tile's hull is the bounding box of 32 box hulls: a few large boxes inflate it, and it then meets sub-tiles
            // that none of its boxes meets (RAPiD candidates: 2.6x the sub-tile visits of equally sized boxes).  Refine:
            // this lane's sub-tile against each of the 32 row boxes -- 32 cheap tests instead of 16 circle tests per lane,
            // a staged load and a queue pass for every falsely admitted sub-tile
            ov = false;
#pragma unroll 8
            for (int j = 0; j < 32; ++j) {
                const float4 q = s_rowh[warp][j];
                ov |= !(h.x > q.z || q.x > h.z || h.y > q.w || q.y > h.w);
            }
        }
        unsigned tiles = __ballot_sync(0xffffffffu, ov);
        float4 v_next = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tiles) v_next = load_col(cbase + __ffs(tiles) - 1);
#pragma unroll 1
        while (tiles) {
            const int C = cbase + __ffs(tiles) - 1;
            tiles &= tiles - 1;
            s_col[warp][buf][lane] = v_next;
            if (tiles) v_next = load_col(cbase + __ffs(tiles) - 1);       // the next sub-tile's data travels during the tests below
            __syncwarp();
            unsigned pass = 0u;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 c = s_col[warp][buf][j];
                const float dx = mc.x - c.x, dy = mc.y - c.y, rr = fmaf(c.z, 1.00001f, mr);
                // circle test AND area-ratio bound (IoU <= min / max of the two areas), branch-free, as in the tile kernel
                const bool ok = (fmaf(dx, dx, dy * dy) <= rr * rr) & (fminf(ma, c.w) * 1.0001f >= thr_f * fmaxf(ma, c.w));
                pass |= (ok ? 1u : 0u) << j;
            }
            {   // each unordered pair once: only partners at a higher spatial position
                const int d = p - C * 16;           // partner j qualifies iff j > d
                if (!have || d >= 15) pass = 0u;
                else if (d >= 0) pass &= ~0u << (d + 1);
            }
#pragma unroll 1
            while (__any_sync(0xffffffffu, pass != 0u)) {
                bool push = false;
                unsigned code = 0u;
                if (pass) {
                    const int j = __ffs(pass) - 1;
                    pass &= pass - 1;
                    const float4 qh = s_col[warp][buf][16 + j];
                    const float oa = s_col[warp][buf][j].w;
                    const float ix = fminf(mh.z, qh.z) - fmaxf(mh.x, qh.x);
                    const float iy = fminf(mh.w, qh.w) - fmaxf(mh.y, qh.y);
                    if (ix > -1e-3f && iy > -1e-3f) {
                        const float ub = (ix + 2e-3f) * (iy + 2e-3f);
                        push = !(ub * 1.0001f < thr_f * (ma + oa - ub));
                    }
                    code = ((unsigned)p << 16) | (unsigned)(C * 16 + j);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, push);
                if (push) queue[qn + __popc(bal & ((1u << lane) - 1u))] = code;
                qn += __popc(bal);
            }
            buf ^= 1;
            if (qn > kBroadQueue - 512) { __syncwarp(); flush(); }
        }
    }
    __syncwarp();
    if (qn) flush();
}

// Upper bound of the intersection area of two rotated rectangles from their ORIENTED extents: the intersection lies inside
// A and inside the axis-aligned (in A's frame) bounding box of B, and the same with the roles swapped.  Everything is kept
// scaled by the half lengths (no square root): with H, V the half-axis vectors of A and d the centre offset, the overlap
// of the two intervals along H is  min(|H|^2, d.H + E) - max(-|H|^2, d.H - E),  E = |H_B.H| + |V_B.H|,  in units of |H|.
// On the bench workload this bound removes 78 % of the pairs the axis-aligned hull bound lets through (138 k -> 30 k per
// 10 000-box image; that none of the removed pairs reaches the threshold is checked by brute force in tests/test_kernel_claims_cpu.py).
__device__ __forceinline__ float oriented_overlap_bound(const float4 ca, const float4 xa, const float4 cb, const float4 xb) {
    const float dx = cb.x - ca.x, dy = cb.y - ca.y;
    const float hh_ = fabsf(xb.x * xa.x + xb.y * xa.y), vh_ = fabsf(xb.z * xa.x + xb.w * xa.y);   // |H_B.H_A|, |V_B.H_A|
    const float hv_ = fabsf(xb.x * xa.z + xb.y * xa.w), vv_ = fabsf(xb.z * xa.z + xb.w * xa.w);   // |H_B.V_A|, |V_B.V_A|
    auto frame = [](float h2, float v2, float ph, float pv, float eh, float ev, float area) {
        const float ox = fminf(h2, ph + eh) - fmaxf(-h2, ph - eh);
        const float oy = fminf(v2, pv + ev) - fmaxf(-v2, pv - ev);
        return fmaxf(ox, 0.f) * fmaxf(oy, 0.f) * 4.f / area;         // (ox / |H|) (oy / |V|), |H||V| = area / 4
    };
    const float ua = frame(xa.x * xa.x + xa.y * xa.y, xa.z * xa.z + xa.w * xa.w, dx * xa.x + dy * xa.y, dx * xa.z + dy * xa.w,
                           hh_ + vh_, hv_ + vv_, ca.w);
    const float ub = frame(xb.x * xb.x + xb.y * xb.y, xb.z * xb.z + xb.w * xb.w, dx * xb.x + dy * xb.y, dx * xb.z + dy * xb.w,
                           hh_ + hv_, vh_ + vv_, cb.w);
    return fminf(ua, ub);
}

constexpr int kNarrowThreads = 256;
constexpr int kNarrowPer = 4;                       // listed pairs per thread and round

// An image takes the lazy (root-first) narrow phase below when its boxes have many listed partners -- clustered detector
// output -- and the one-pass kernel otherwise (14 partners per box on the bench workload: the two extra passes over the
// list would cost more than the skipped clips save).  Decided on the device from the broad phase's pair count.
constexpr int kLazyPartners = 32;
__device__ __forceinline__ bool rot_image_is_lazy(const LargeWs& w, const int* m, int b, int lazy_mode) {
    return lazy_mode == 2 || (lazy_mode == 1 && w.pair_count[b] > kLazyPartners * m[b]);
}

__device__ __forceinline__ void rot_narrow_body(const LargeWs& w, const int* m, int n, double thr_d, int ge) {
    const int b = blockIdx.y, tid = threadIdx.x;
    const int cnt = w.pair_count[b];
    if (cnt > w.pair_cap) return;                   // overflowed: the tile kernel redoes this image
    const long long base = (long long)b * n;
    const unsigned* pairs = w.pairs + (long long)b * w.pair_cap;
    unsigned* mask32 = reinterpret_cast<unsigned*>(w.mask);
    const bool ge_mode = ge != 0;
    const float thr_f = (float)thr_d;
    __shared__ unsigned s_q[kNarrowThreads * kNarrowPer];
    __shared__ int s_n;
    constexpr int kChunk = kNarrowThreads * kNarrowPer;
    for (int c0 = blockIdx.x * kChunk; c0 < cnt; c0 += gridDim.x * kChunk) {       // block-uniform
        __syncthreads();
        if (tid == 0) s_n = 0;
        __syncthreads();
        // (1) oriented bound on kNarrowPer pairs per thread (independent loads), survivors compacted into shared memory
        unsigned code[kNarrowPer];
        float4 ca[kNarrowPer], cb[kNarrowPer], xa[kNarrowPer], xb[kNarrowPer];
#pragma unroll
        for (int u = 0; u < kNarrowPer; ++u) {
            const int e = c0 + u * kNarrowThreads + tid;
            code[u] = (e < cnt) ? pairs[e] : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < kNarrowPer; ++u) {
            const bool live = code[u] != 0xffffffffu;
            const int pa = live ? (int)(code[u] >> 16) : 0, pb = live ? (int)(code[u] & 0xffffu) : 0;
            ca[u] = w.cull4[base + pa]; cb[u] = w.cull4[base + pb];
            xa[u] = w.axes4[base + pa]; xb[u] = w.axes4[base + pb];
        }
#pragma unroll
        for (int u = 0; u < kNarrowPer; ++u) {
            const float ub = oriented_overlap_bound(ca[u], xa[u], cb[u], xb[u]);
            // keep unless the bound (with 1e-3 relative + 1e-2 absolute slack for float32 rounding) rules the threshold out;
            // degenerate boxes (NaN / zero area) stay in: rot_overlaps decides them as before
            const bool drop = ub * 1.001f + 1e-2f < thr_f * (ca[u].w + cb[u].w - ub);
            const bool keep = code[u] != 0xffffffffu && !drop;
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            int at = 0;
            if ((tid & 31) == 0 && bal) at = atomicAdd(&s_n, __popc(bal));
            at = __shfl_sync(0xffffffffu, at, 0);
            if (keep) s_q[at + __popc(bal & ((1u << (tid & 31)) - 1u))] = code[u];
        }
        __syncthreads();
        // (2) polygon clip on the survivors, every lane busy
        const int total = s_n;
        for (int e = tid; e < total; e += kNarrowThreads) {
            const unsigned cd = s_q[e];
            const int pa = (int)(cd >> 16), pb = (int)(cd & 0xffffu);             // spatial positions, pa < pb
            const RotBox A = w.rbox[base + pa];
            const RotBox B = w.rbox[base + pb];
            if (rot_overlaps(A, B, thr_d, ge_mode)) {
                const bool a_first = w.rank_of_spos[base + pa] < w.rank_of_spos[base + pb];
                const int row = a_first ? pa : pb, col = a_first ? pb : pa;           // the higher-scored box suppresses
                fx_note(w, b, row, col >> 5, atomicOr(&mask32[((base + row) * w.words + (col >> 6)) * 2 + ((col >> 5) & 1)], 1u << (col & 31)));
                const int trow = row >> 6, tcol = col >> 6;
                atomicOr(&w.tile_adj[((long long)b * w.words + trow) * w.aw + (tcol >> 6)], 1ull << (tcol & 63));
            }
        }
    }
}

// ---- lazy narrow phase.  Detector output is clustered: an object draws hundreds of mutually overlapping boxes, of which
// greedy NMS keeps one or two -- yet the kernel above clips EVERY listed pair (1.25 M polygon clips per image for 40
// objects x 250 boxes, 73 % of the whole rotated NMS).  Most of those pairs cannot matter: a box that a certainly-kept box
// suppresses is never kept itself, so its row of the suppression matrix is never read, and its column is already decided.
//   filter : the oriented-extent bound on every listed pair (stage 1 of the kernel above); survivors are compacted into a
//            second list and mark their LOWER-scored box "has a higher-scored partner".  Boxes without the mark are ROOTS:
//            nothing that could suppress them exists, greedy NMS keeps them.
//   clip<0>: pairs whose higher-scored box is a root: clip, set the bit, mark the lower box DEAD on overlap.
//   clip<1>: the remaining pairs, except those with a dead member: a dead higher box never suppresses (it is not kept), a
//            dead lower box is already removed by its root's bit.  The fixed point of the reduced matrix is the greedy set.
// 40 x 250 clustered boxes: 1.25 M -> ~0.1 M clips per image.
__device__ __forceinline__ void rot_filter_body(const LargeWs& w, const int* m, int n, double thr_d) {
    const int b = blockIdx.y, tid = threadIdx.x;
    const int cnt = w.pair_count[b];
    if (cnt > w.pair_cap) return;                   // overflowed: the tile kernel redoes this image
    const long long base = (long long)b * n;
    const unsigned* pairs = w.pairs + (long long)b * w.pair_cap;
    unsigned* out = w.pairs2 + (long long)b * w.pair_cap;
    const float thr_f = (float)thr_d;
    __shared__ unsigned s_q[kNarrowThreads * kNarrowPer];
    __shared__ int s_n, s_at;
    constexpr int kChunk = kNarrowThreads * kNarrowPer;
    for (int c0 = blockIdx.x * kChunk; c0 < cnt; c0 += gridDim.x * kChunk) {       // block-uniform
        __syncthreads();
        if (tid == 0) s_n = 0;
        __syncthreads();
        unsigned code[kNarrowPer];
        float4 ca[kNarrowPer], cb[kNarrowPer], xa[kNarrowPer], xb[kNarrowPer];
#pragma unroll
        for (int u = 0; u < kNarrowPer; ++u) {
            const int e = c0 + u * kNarrowThreads + tid;
            code[u] = (e < cnt) ? pairs[e] : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < kNarrowPer; ++u) {
            const bool live = code[u] != 0xffffffffu;
            const int pa = live ? (int)(code[u] >> 16) : 0, pb = live ? (int)(code[u] & 0xffffu) : 0;
            ca[u] = w.cull4[base + pa]; cb[u] = w.cull4[base + pb];
            xa[u] = w.axes4[base + pa]; xb[u] = w.axes4[base + pb];
        }
#pragma unroll
        for (int u = 0; u < kNarrowPer; ++u) {
            const float ub = oriented_overlap_bound(ca[u], xa[u], cb[u], xb[u]);
            const bool drop = ub * 1.001f + 1e-2f < thr_f * (ca[u].w + cb[u].w - ub);     // same slack as rot_narrow_kernel
            const bool keep = code[u] != 0xffffffffu && !drop;
            if (keep) {
                const int pa = (int)(code[u] >> 16), pb = (int)(code[u] & 0xffffu);
                const bool a_first = w.rank_of_spos[base + pa] < w.rank_of_spos[base + pb];
                w.lz_high[base + (a_first ? pb : pa)] = 1;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            int at = 0;
            if ((tid & 31) == 0 && bal) at = atomicAdd(&s_n, __popc(bal));
            at = __shfl_sync(0xffffffffu, at, 0);
            if (keep) s_q[at + __popc(bal & ((1u << (tid & 31)) - 1u))] = code[u];
        }
        __syncthreads();
        const int total = s_n;
        if (tid == 0 && total) s_at = atomicAdd(&w.pair2_count[b], total);
        __syncthreads();
        for (int e = tid; e < total; e += kNarrowThreads) out[s_at + e] = s_q[e];
    }
}

// lazy_mode: 0 = one-pass narrow phase for every image, 1 = per image by its pair count, 2 = lazy for every image
__global__ void __launch_bounds__(kNarrowThreads) rot_narrow_kernel(LargeWs w, const int* m, int n, double thr_d, int ge, int lazy_mode) {
    if (rot_image_is_lazy(w, m, blockIdx.y, lazy_mode)) rot_filter_body(w, m, n, thr_d);
    else rot_narrow_body(w, m, n, thr_d, ge);
}

template <int PHASE>
__global__ void __launch_bounds__(kNarrowThreads) rot_clip_kernel(LargeWs w, const int* m, int n, double thr_d, int ge, int lazy_mode) {
    const int b = blockIdx.y, tid = threadIdx.x;
    if (w.pair_count[b] > w.pair_cap || !rot_image_is_lazy(w, m, b, lazy_mode)) return;
    const int cnt = w.pair2_count[b];
    const long long base = (long long)b * n;
    const unsigned* pairs = w.pairs2 + (long long)b * w.pair_cap;
    unsigned* mask32 = reinterpret_cast<unsigned*>(w.mask);
    const bool ge_mode = ge != 0;
    __shared__ unsigned s_q[kNarrowThreads * kNarrowPer];
    __shared__ int s_n;
    constexpr int kChunk = kNarrowThreads * kNarrowPer;
    for (int c0 = blockIdx.x * kChunk; c0 < cnt; c0 += gridDim.x * kChunk) {       // block-uniform
        __syncthreads();
        if (tid == 0) s_n = 0;
        __syncthreads();
        // (1) which pairs does this phase own?  code in the queue: higher-scored position << 16 | lower-scored position
        unsigned code[kNarrowPer];
        int ra[kNarrowPer], rb[kNarrowPer];
#pragma unroll
        for (int u = 0; u < kNarrowPer; ++u) {
            const int e = c0 + u * kNarrowThreads + tid;
            code[u] = (e < cnt) ? pairs[e] : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < kNarrowPer; ++u) {
            const bool live = code[u] != 0xffffffffu;
            ra[u] = live ? w.rank_of_spos[base + (code[u] >> 16)] : 0;
            rb[u] = live ? w.rank_of_spos[base + (code[u] & 0xffffu)] : 0;
        }
#pragma unroll
        for (int u = 0; u < kNarrowPer; ++u) {
            bool own = false;
            unsigned ordered = 0u;
            if (code[u] != 0xffffffffu) {
                const int pa = (int)(code[u] >> 16), pb = (int)(code[u] & 0xffffu);
                const int hi = ra[u] < rb[u] ? pa : pb, lo = ra[u] < rb[u] ? pb : pa;
                const bool root = w.lz_high[base + hi] == 0;
                own = PHASE == 0 ? root : (!root && !w.lz_dead[base + hi] && !w.lz_dead[base + lo]);
                ordered = ((unsigned)hi << 16) | (unsigned)lo;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, own);
            int at = 0;
            if ((tid & 31) == 0 && bal) at = atomicAdd(&s_n, __popc(bal));
            at = __shfl_sync(0xffffffffu, at, 0);
            if (own) s_q[at + __popc(bal & ((1u << (tid & 31)) - 1u))] = ordered;
        }
        __syncthreads();
        // (2) polygon clip, every lane busy
        const int total = s_n;
        for (int e = tid; e < total; e += kNarrowThreads) {
            const unsigned cd = s_q[e];
            const int row = (int)(cd >> 16), col = (int)(cd & 0xffffu);           // the higher-scored box suppresses
            const RotBox A = w.rbox[base + min(row, col)];
            const RotBox B = w.rbox[base + max(row, col)];                        // operand order of rot_narrow_kernel: lower position first
            if (rot_overlaps(A, B, thr_d, ge_mode)) {
                if (PHASE == 0) w.lz_dead[base + col] = 1;
                fx_note(w, b, row, col >> 5, atomicOr(&mask32[((base + row) * w.words + (col >> 6)) * 2 + ((col >> 5) & 1)], 1u << (col & 31)));
                const int trow = row >> 6, tcol = col >> 6;
                atomicOr(&w.tile_adj[((long long)b * w.words + trow) * w.aw + (tcol >> 6)], 1ull << (tcol & 63));
            }
        }
    }
}

// The 64 boxes of a score block sit at arbitrary spatial positions: gather their mutual suppression bits
// (row of a, bit of c) for EVERY block in parallel, so that the serial sweep only reads 64 words per block.
__global__ void __launch_bounds__(512) spatial_diag_kernel(LargeWs w, const int* m, int n) {
    const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int mb = m[b];
    const int r0 = t * kTile;
    if (r0 >= mb) return;
    const int rows = min(kTile, mb - r0);
    const long long base_n = (long long)b * n;
    const unsigned long long* mask = w.mask + base_n * w.words;
    __shared__ int s_sp[kTile];
    __shared__ unsigned long long s_adj[kTile][kAdjMax];
    __shared__ unsigned long long s_d[kTile];
    const int aw = w.aw;
    if (tid < kTile) { s_sp[tid] = (tid < rows) ? w.spos_of_rank[base_n + r0 + tid] : 0; s_d[tid] = 0ull; }
    __syncthreads();
    for (int e = tid; e < kTile * aw; e += 512) {
        const int a = e / aw, q = e - a * aw;
        const unsigned long long v = (a < rows) ? w.tile_adj[((long long)b * w.words + (s_sp[a] >> 6)) * aw + q] : 0ull;
        s_adj[a][q] = v;
        w.adj_blk[((long long)b * w.words + t) * (kTile * aw) + e] = v;      // staged per block for the sweep
    }
    __syncthreads();
    const int a = tid & (kTile - 1), cg = tid >> 6;              // 64 rows x 8 column groups
    if (a < rows) {
        const unsigned long long* rowp = mask + (long long)s_sp[a] * w.words;
        unsigned long long wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = cg * 8 + u, wc = s_sp[c] >> 6;
            wv[u] = (c < rows && ((s_adj[a][wc >> 6] >> (wc & 63)) & 1ull)) ? rowp[wc] : 0ull;   // provably empty words are skipped
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = cg * 8 + u;
            if ((wv[u] >> (s_sp[c] & 63)) & 1ull) atomicOr(&s_d[c], 1ull << a);       // transposed: column c collects its suppressors
        }
    }
    __syncthreads();
    if (tid < kTile) w.diag_all[((long long)b * w.words + t) * kTile + tid] = s_d[tid];
}

// ---------------------------------------------------------------------------- sweep + emit
constexpr int kSweepThreads = 512;

struct EmitParams {
    const float* boxes; const float* scores; const void* cls; const int* src_idx;
    long long pitch; int n, n_param, cls_is_i64;
    float* out_box; float* out_score; long long* out_cls; int* out_idx; int* out_count; int* status; int out_cap;
    long long* keep64;   // rotated API: kept candidate indices as int64 (B, pitch)
};

// MODE 0: score-ordered mask, block sweep.  MODE 1: spatially ordered mask, block sweep (the default).
// MODE 2: spatially ordered mask, parallel fixed-point sweep of sweep_fixpoint.cuh (experimental, MYDET_SWEEP_FIXPOINT=1).
template <int MODE>
__global__ void __launch_bounds__(kSweepThreads) sweep_kernel(LargeWs w, const int* m, int n, EmitParams E) {
    constexpr bool SPATIAL = MODE == 1;
    extern __shared__ unsigned long long sm[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int mb = m[b];
    const int words = (mb + 63) / 64;
    unsigned long long* removed = sm;                    // w.words   (indexed by rank, or by spatial position)
    unsigned long long* keptw = sm + w.words;            // w.words   (indexed by rank)
    unsigned long long* diag = keptw + w.words;          // 64
    __shared__ unsigned long long s_kept;
    __shared__ int klist[kTile];
    __shared__ int kslot[kTile];                     // block-local row number of the kept-list entries
    __shared__ int s_prefix_total;
    const long long base_n = (long long)b * n;
    const unsigned long long* mask = w.mask + base_n * w.words;

    for (int i = tid; i < w.words; i += kSweepThreads) { removed[i] = 0ull; keptw[i] = 0ull; }
    __syncthreads();

    if (MODE == 2) {
        // Jacobi rounds until nothing changes; every round is parallel over the rows of the image
        unsigned long long* keep = diag + kTile;             // w.words more (the launch sizes the buffer for it)
        const fx::View V{mask, w.tile_adj + (long long)b * w.words * w.aw, w.spos_of_rank + base_n, mb, w.words, w.aw};
        fx::Entry* list = w.fx_list + (long long)b * w.fx_cap;
#ifdef MYDET_SWEEP_PROFILE
        long long fxt[5]; int fx_rounds = 0;
        fxt[0] = clock64();
#endif
        fx::phase_init(V, keep, removed, keptw, w.words, tid, kSweepThreads);
        // the mask kernels recorded every non-empty 32-bit half-word of the image (fx_note); fetch their bits once
        const int n_entries = w.fx_count[b];
        const bool use_list = n_entries <= w.fx_cap;            // else: walk the adjacency map every round
        // up to 16 entries per thread stay in registers for all rounds; longer lists are re-read from memory every round
        const bool in_regs = use_list && n_entries <= fx::kRegEntries * kSweepThreads;
        unsigned long long ent[fx::kRegEntries];
        if (in_regs) {
#pragma unroll
            for (int k = 0; k < fx::kRegEntries; ++k) ent[k] = 0ull;
            fx::phase_load_entries32(reinterpret_cast<const unsigned*>(mask), w.words, list, n_entries, ent, tid, kSweepThreads);
        } else if (use_list) {
            fx::phase_fill_list32(reinterpret_cast<const unsigned*>(mask), w.words, list, n_entries, tid, kSweepThreads);
        }
        __syncthreads();
#ifdef MYDET_SWEEP_PROFILE
        fxt[1] = clock64();
#endif
        for (;;) {
            if (in_regs) fx::phase_scatter_entries32(ent, keep, removed);
            else if (use_list) fx::phase_scatter_list32(list, n_entries, keep, removed, tid, kSweepThreads);
            else fx::phase_scatter(V, keep, removed, tid, kSweepThreads);
            __syncthreads();
#ifdef MYDET_SWEEP_PROFILE
            ++fx_rounds;
#endif
            if (!__syncthreads_or(fx::phase_update(V, keep, removed, w.words, tid, kSweepThreads))) break;
        }
#ifdef MYDET_SWEEP_PROFILE
        fxt[2] = clock64();
#endif
        fx::phase_to_rank(V, keep, keptw, tid, kSweepThreads);
        __syncthreads();
        for (int i = tid; i < words; i += kSweepThreads) w.kept[(long long)b * w.words + i] = keptw[i];
        __syncthreads();
#ifdef MYDET_SWEEP_PROFILE
        fxt[3] = clock64();
        if (tid == 0 && b == 0) printf("fixpoint sweep (cycles): init+fill %lld  rounds(%d, %d entries) %lld  to_rank %lld\n",
                                       fxt[1] - fxt[0], fx_rounds, n_entries, fxt[2] - fxt[1], fxt[3] - fxt[2]);
#endif
    } else if (SPATIAL) {
        // Software-pipelined: the data of block t+1 (positions, adjacency rows, gathered diagonal words) does
        // not depend on the removed vector, so it is loaded while thread 0 resolves block t.
        __shared__ int sp2[2][kTile];
        __shared__ unsigned long long adj2[2][kTile][kAdjMax];
        const int aw = w.aw;
        constexpr int kAdjPerThread = (kTile * kAdjMax + kSweepThreads - 1) / kSweepThreads;   // 2
        __shared__ unsigned long long dg2[2][kTile];
        auto fetch = [&](int t, int& v_sp, unsigned long long& v_dg, unsigned long long (&v_adj)[kAdjPerThread]) {
            v_sp = 0; v_dg = 0ull;
#pragma unroll
            for (int q = 0; q < kAdjPerThread; ++q) v_adj[q] = 0ull;
            if (t >= words) return;
            const int r0 = t * kTile, rows = min(kTile, mb - r0);
            if (tid < kTile && tid < rows) {
                v_sp = w.spos_of_rank[base_n + r0 + tid];
                v_dg = w.diag_all[((long long)b * w.words + t) * kTile + tid];
            }
#pragma unroll
            for (int q = 0; q < kAdjPerThread; ++q) {
                const int e = tid + q * kSweepThreads;
                if (e < kTile * aw) v_adj[q] = w.adj_blk[((long long)b * w.words + t) * (kTile * aw) + e];
            }
        };
        auto stash = [&](int buf, int v_sp, unsigned long long v_dg, const unsigned long long (&v_adj)[kAdjPerThread]) {
            if (tid < kTile) { sp2[buf][tid] = v_sp; dg2[buf][tid] = v_dg; }
#pragma unroll
            for (int q = 0; q < kAdjPerThread; ++q) {
                const int e = tid + q * kSweepThreads;
                if (e < kTile * aw) adj2[buf][e / aw][e % aw] = v_adj[q];
            }
        };
        int v_sp; unsigned long long v_dg, v_adj[kAdjPerThread];
        fetch(0, v_sp, v_dg, v_adj);
        stash(0, v_sp, v_dg, v_adj);
        fetch(1, v_sp, v_dg, v_adj);              // two blocks of look-ahead: a full iteration to land
        int buf = 0;
#ifdef MYDET_SWEEP_PROFILE
        long long c_a = 0, c_b = 0, c_c = 0, c_d = 0, tk;
#define SW_T(x) do { if (tid == 0) { long long now = clock64(); x += now - tk; tk = now; } } while (0)
        tk = clock64();
#else
#define SW_T(x) do {} while (0)
#endif
        for (int t = 0; t < words; ++t, buf ^= 1) {
            const int r0 = t * kTile;
            const int rows = min(kTile, mb - r0);
            __syncthreads();                       // block t staged, removed vector up to date
            SW_T(c_d);
            if (tid < 32) {
                // Resolve the 64 boxes of the block with ONE warp, as a fixed point: lane l owns boxes l and
                // l+32; a box is kept iff it is not removed yet and no KEPT box of the block suppresses it
                // (dg2 holds, per box, the set of block-mates that would).  The greedy result is the unique
                // fixed point; every round settles at least the next undecided box, typically 2-4 rounds.
                // (A 64-step serial chain in one thread cost 3.6 k cycles per block.)
                const int c0 = tid, c1 = tid + 32;
                bool a0 = false, a1 = false;
                if (c0 < rows) { const int sa = sp2[buf][c0]; a0 = !((removed[sa >> 6] >> (sa & 63)) & 1ull); }
                if (c1 < rows) { const int sa = sp2[buf][c1]; a1 = !((removed[sa >> 6] >> (sa & 63)) & 1ull); }
                const unsigned long long s0 = dg2[buf][c0], s1 = dg2[buf][c1];
                unsigned long long kept = ((unsigned long long)__ballot_sync(0xffffffffu, a1) << 32) | __ballot_sync(0xffffffffu, a0);
#pragma unroll 1
                for (int round = 0; round < kTile; ++round) {
                    const bool k0 = a0 && !(s0 & kept), k1 = a1 && !(s1 & kept);
                    const unsigned long long next = ((unsigned long long)__ballot_sync(0xffffffffu, k1) << 32) | __ballot_sync(0xffffffffu, k0);
                    if (next == kept) break;
                    kept = next;
                }
                if (tid == 0) {
                    s_kept = kept;
                    keptw[t] = kept;
                    w.kept[(long long)b * w.words + t] = kept;
                }
            }
            __syncthreads();
            SW_T(c_b);
            const unsigned long long kept = s_kept;
            if (tid < kTile && ((kept >> tid) & 1ull)) {
                const int kpos = __popcll(kept & ((1ull << tid) - 1ull));
                klist[kpos] = sp2[buf][tid];
                kslot[kpos] = tid;
            }
            stash(buf ^ 1, v_sp, v_dg, v_adj);     // block t+1, fetched one iteration ago
            fetch(t + 2, v_sp, v_dg, v_adj);
            __syncthreads();
            SW_T(c_c);
            {
                // OR the kept rows into the removed vector: one warp per kept row, lanes over its words, and only
                // the words the adjacency map marks non-empty are loaded
                const int nk = __popcll(kept);
                unsigned* removed32 = reinterpret_cast<unsigned*>(removed);
                const int warp = tid >> 5, lane = tid & 31;
                constexpr int kWarps = kSweepThreads / 32;          // 16: a block has at most 64 kept rows = 4 per warp
                constexpr int kRowsPerWarp = kTile / kWarps;        // all of a warp's rows are loaded in ONE round trip
                                                                    // (row after row was 3-4 dependent L2 latencies per block)
#pragma unroll 1
                for (int w0 = 0; w0 < words; w0 += 256) {               // 256 words per round, 8 loads per row in flight
                    unsigned long long v[kRowsPerWarp][8];
#pragma unroll
                    for (int j = 0; j < kRowsPerWarp; ++j) {
                        const int kr = warp + j * kWarps;
                        const bool live = kr < nk;
                        const unsigned long long* rowp = mask + (long long)(live ? klist[kr] : 0) * w.words;
                        const unsigned long long* adj = adj2[buf][live ? kslot[kr] : 0];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int wd = w0 + lane + 32 * u;
                            v[j][u] = (live && wd < words && ((adj[wd >> 6] >> (wd & 63)) & 1ull)) ? rowp[wd] : 0ull;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int wd = w0 + lane + 32 * u;
                        unsigned long long acc = 0ull;
#pragma unroll
                        for (int j = 0; j < kRowsPerWarp; ++j) acc |= v[j][u];
                        if ((unsigned)acc) atomicOr(&removed32[2 * wd], (unsigned)acc);
                        if ((unsigned)(acc >> 32)) atomicOr(&removed32[2 * wd + 1], (unsigned)(acc >> 32));
                    }
                }
            }
        }
        __syncthreads();
#ifdef MYDET_SWEEP_PROFILE
        if (tid == 0 && b == 0) printf("sweep phases (cycles, %d blocks): rem-bits %lld  chain %lld  klist+stash %lld  OR %lld\n", words, c_a, c_b, c_c, c_d);
#endif
    } else {
    for (int t = 0; t < words; ++t) {
        const int r0 = t * kTile;
        const int rows = min(kTile, mb - r0);
        if (tid < kTile) diag[tid] = (tid < rows) ? mask[(long long)(r0 + tid) * w.words + t] : 0ull;
        __syncthreads();
        if (tid == 0) {
            unsigned long long cur = removed[t];
            if (rows < kTile) cur |= ~0ull << rows;
            unsigned long long kept = 0ull;
#pragma unroll
            for (int j = 0; j < kTile; ++j) {
                const bool alive = !((cur >> j) & 1ull);
                if (alive) { kept |= 1ull << j; cur |= diag[j]; }
            }
            s_kept = kept;
            keptw[t] = kept;
            w.kept[(long long)b * w.words + t] = kept;
        }
        __syncthreads();
        const unsigned long long kept = s_kept;
        // OR the kept rows of this block into the removed words that are still ahead.  (kept row, word)
        // pairs are spread over all threads, four independent loads in flight each: a per-word loop over
        // the kept rows issued one dependent load after the other, 24 us per block.
        if (tid < kTile && ((kept >> tid) & 1ull)) klist[__popcll(kept & ((1ull << tid) - 1ull))] = r0 + tid;
        __syncthreads();
        {
            const int nk = __popcll(kept), nw = words - (t + 1);
            const int work = nk * nw;
            unsigned* removed32 = reinterpret_cast<unsigned*>(removed);
            for (int base_i = tid; base_i < work; base_i += 4 * kSweepThreads) {
                unsigned long long v[4];
                int wdv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = base_i + u * kSweepThreads;
                    v[u] = 0ull; wdv[u] = 0;
                    if (idx < work) {
                        const int kr = idx / nw, wd = t + 1 + (idx - kr * nw);
                        wdv[u] = wd;
                        v[u] = mask[(long long)klist[kr] * w.words + wd];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if ((unsigned)v[u]) atomicOr(&removed32[2 * wdv[u]], (unsigned)v[u]);
                    if ((unsigned)(v[u] >> 32)) atomicOr(&removed32[2 * wdv[u] + 1], (unsigned)(v[u] >> 32));
                }
            }
        }
        __syncthreads();
    }
    }

    // ---- emit survivors in sorted order
    // exclusive prefix of popcounts over keptw (serial per 256-word chunk is fine: words <= 16384)
    __shared__ int chunk_sum[kSweepThreads];
    const int per = (words + kSweepThreads - 1) / kSweepThreads;
    int local = 0;
    for (int k = 0; k < per; ++k) { const int wd = tid * per + k; if (wd < words) local += __popcll(keptw[wd]); }
    chunk_sum[tid] = local;
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int k = 0; k < kSweepThreads; ++k) { const int v = chunk_sum[k]; chunk_sum[k] = acc; acc += v; }
        s_prefix_total = acc;
    }
    __syncthreads();
    // one thread per RANK (not per 64-rank word: that was <= 64 dependent load-store pairs in a row per thread, ~30 us
    // of latency at 10 000 boxes): output position = kept ranks before this one
    const int* order = w.order + (long long)b * n;
    constexpr int kEmitBatch = 4;                          // independent order[] loads in flight per thread
    for (int r0 = tid; r0 < mb; r0 += kSweepThreads * kEmitBatch) {
        int idx[kEmitBatch], posv[kEmitBatch];
#pragma unroll
        for (int u = 0; u < kEmitBatch; ++u) {
            const int r = r0 + u * kSweepThreads;
            idx[u] = -1; posv[u] = 0;
            if (r >= mb) continue;
            const int wd = r >> 6, j = r & 63;
            const unsigned long long kw = keptw[wd];
            if (!((kw >> j) & 1ull)) continue;
            const int c = wd / per;
            int pos = chunk_sum[c] + __popcll(kw & ((1ull << j) - 1ull));
            for (int k = c * per; k < wd; ++k) pos += __popcll(keptw[k]);
            posv[u] = pos;
            idx[u] = order[r];
        }
#pragma unroll
        for (int u = 0; u < kEmitBatch; ++u) {
            const int i = idx[u], pos = posv[u], r = r0 + u * kSweepThreads;
            if (i < 0) continue;
            if (E.keep64) {
                E.keep64[(long long)b * E.pitch + pos] = i;
                w.rowpos[(long long)b * n + r] = pos;
            } else if (pos < E.out_cap) {
                const long long orow = (long long)b * E.out_cap + pos;
                const long long irow = (long long)b * E.pitch + i;
                for (int p = 0; p < E.n_param; ++p) E.out_box[orow * E.n_param + p] = E.boxes[irow * E.n_param + p];
                E.out_score[orow] = E.scores[irow];
                E.out_cls[orow] = E.cls ? (E.cls_is_i64 ? reinterpret_cast<const long long*>(E.cls)[irow]
                                                        : (long long)reinterpret_cast<const int*>(E.cls)[irow]) : 0ll;
                E.out_idx[orow] = E.src_idx ? E.src_idx[irow] : i;
            }
        }
    }
#ifdef MYDET_SWEEP_PROFILE
    if (tid == 0 && b == 0) printf("sweep emit done at clock %lld\n", clock64());
#endif
    if (tid == 0) {
        int total = s_prefix_total;
        if (!E.keep64 && total > E.out_cap) { total = E.out_cap; if (E.status) atomicOr(E.status + b, 2); }
        E.out_count[b] = total;
    }
}

// ---------------------------------------------------------------------------- majority votes
// nms_rotbb's vote bookkeeping (utils/bbox_ops.py:291-306): every dropped box votes for the valid
// box it overlaps most (first maximum over the boxes valid at that time = kept boxes ranked
// before it); every kept box starts with its own vote.
__global__ void votes_kernel(LargeWs w, const int* m, int n, int* votes, long long pitch, int spatial) {
    const int b = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m[b]) return;
    const long long base = (long long)b * n;
    const unsigned long long* kept = w.kept + (long long)b * w.words;
    if ((kept[r >> 6] >> (r & 63)) & 1ull) {
        atomicAdd(votes + (long long)b * pitch + w.rowpos[base + r], 1);
        return;
    }
    // the quads are stored by score rank, or by spatial position (then reached through spos_of_rank)
    const RotBox me = w.rbox[base + (spatial ? w.spos_of_rank[base + r] : r)];
    double best = -1.0;
    int best_q = -1;
    for (int wd = 0; wd <= (r >> 6); ++wd) {
        unsigned long long kw = kept[wd];
        if (wd == (r >> 6)) kw &= (1ull << (r & 63)) - 1ull;
        while (kw) {
            const int j = __ffsll((long long)kw) - 1;
            kw &= kw - 1;
            const int q = wd * kTile + j;
            const RotBox o = w.rbox[base + (spatial ? w.spos_of_rank[base + q] : q)];
            const double dx = (double)me.cx - (double)o.cx, dy = (double)me.cy - (double)o.cy;
            const double rr = (double)me.r + (double)o.r + 1e-3;
            double iou = 0.0;
            if (dx * dx + dy * dy <= rr * rr) iou = rot_iou_f64(me.x, me.y, o.x, o.y);
            if (iou > best) { best = iou; best_q = q; }
        }
    }
    if (best_q >= 0) atomicAdd(votes + (long long)b * pitch + w.rowpos[base + best_q], 1);
}

// Leaves the bit matrix of every image all-zero again (persistent workspaces, LargeArgs::ws_clean): the half-words the mask
// kernels recorded, or -- when the list overflowed -- the whole matrix of that image.
__global__ void __launch_bounds__(256) mask_cleanup_kernel(LargeWs w, int n) {
    const int b = blockIdx.y;
    const int cnt = w.fx_count[b];
    unsigned* mask32 = reinterpret_cast<unsigned*>(w.mask + (long long)b * n * w.words);
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cnt <= w.fx_cap) {
        const fx::Entry* list = w.fx_list + (long long)b * w.fx_cap;
        for (long long e = t0; e < cnt; e += stride) mask32[(long long)list[e].row * (2 * w.words) + list[e].word] = 0u;
    } else {
        uint4* m4 = reinterpret_cast<uint4*>(mask32);                  // n * words * 8 bytes: a multiple of 16 (256-byte aligned base)
        const long long total = (long long)n * w.words / 2;
        for (long long i = t0; i < total; i += stride) m4[i] = make_uint4(0u, 0u, 0u, 0u);
        if ((((long long)n * w.words) & 1) && t0 == 0) w.mask[(long long)b * n * w.words + (long long)n * w.words - 1] = 0ull;
    }
}

// ---------------------------------------------------------------------------- host orchestration
int run_large(const LargeArgs& A, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    LargeWs w;
    const size_t need = carve(w, workspace, A.batch, A.n, A.rot);
    if (need > workspace_bytes || !workspace) {
        set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        return MYDET_ERR_WORKSPACE;
    }
    MYDET_REQUIRE(A.topk <= 0 || A.topk >= A.n, "top-k above %d with more candidates than k is not supported yet "
                  "(topk=%d, n=%d)", MYDET_SMALL_K, A.topk, A.n);
    MYDET_REQUIRE(A.batch <= 65535, "batch too large for one launch (<= 65535)");
    const int B = A.batch, n = A.n;
    MYDET_CUDA(cudaMemsetAsync(w.m, 0, sizeof(int) * (size_t)B, st));
    if (A.status) MYDET_CUDA(cudaMemsetAsync(A.status, 0, sizeof(int) * (size_t)B, st));
    // Morton-ordered mask (n <= 65 536).  A degenerate threshold that a DISJOINT pair passes ("IoU >= 0" of the rotated
    // API, a negative threshold with torchvision's "IoU > thr") suppresses boxes that patch culling never pairs up:
    // such calls keep the score-ordered all-pairs kernels.
    const bool disjoint_suppresses = (A.rot && A.ge) ? (A.thr <= 0.0) : (A.thr < 0.0);
    const bool spatial = n <= kSpatialMaxN && !disjoint_suppresses;
    KeyParams K{A.scores, A.cls, A.counts, A.src_idx, A.pitch, n, A.cls_is_i64, (A.cls && !A.rot) ? 1 : 0, A.conf_thres, A.status,
                A.boxes, A.n_param, A.box_format, spatial ? w.skeys : nullptr};
    keys_kernel<<<dim3((n + 255) / 256, B), 256, 0, st>>>(K, w.keys, w.m);
    {
        // score order and Morton order do not depend on each other: one launch sorts both
        const bool pik1 = A.src_idx == nullptr;      // the tie index of the score keys is the slot itself
        const int rc = spatial ? sort_keys(w.keys, w.order, w.rank_of_slot, w.skeys, w.slot_of_spos, n, B, w, st, pik1)
                               : sort_keys(w.keys, w.order, nullptr, nullptr, nullptr, n, B, w, st, pik1);
        if (rc) return rc;
    }
    GatherParams G{A.boxes, A.pitch, n, A.n_param, A.box_format};
    const int tiles = (n + kTile - 1) / kTile;
    EmitParams E{A.boxes, A.scores, A.cls, A.src_idx, A.pitch, n, A.n_param, A.cls_is_i64,
                 A.out_box, A.out_score, A.out_cls, A.out_idx, A.out_count, A.status, A.out_cap, A.keep64};
    const size_t smem = ((size_t)w.words * 2 + kTile) * sizeof(unsigned long long);
    MYDET_REQUIRE(smem <= 200 * 1024, "too many candidates per image for the sweep kernel");
    const int smem_attr = (int)smem > 48 * 1024 ? (int)smem : 48 * 1024;
    if (spatial) {
        // the bit matrix (8 n ceil(n/64) bytes per image: 12.6 MB at 10 000 boxes, 293 MB at 48 384) is sparse -- a few
        // thousand non-zero words -- and clearing it costs more than any kernel of the path.  A caller that keeps the
        // workspace between calls (ws_clean) gets it back clean: mask_cleanup_kernel, last in the pipeline, zeroes exactly
        // the half-words this call set (the entry list again), so no call after the first clears the matrix wholesale.
        if (!A.ws_clean) MYDET_CUDA(cudaMemsetAsync(w.mask, 0, (size_t)B * n * w.words * sizeof(unsigned long long), st));
        MYDET_CUDA(cudaMemsetAsync(w.tile_adj, 0, (size_t)B * w.words * w.aw * 8, st));
        MYDET_CUDA(cudaMemsetAsync(w.fx_count, 0, sizeof(int) * (size_t)B, st));
        const dim3 mgrid(tiles, (tiles + kColChunk - 1) / kColChunk, B);
        if (A.rot) {
            spatial_gather_kernel<true><<<dim3(tiles, B), kTile, 0, st>>>(G, w.keys, w.order, w.m, w);
            const char* tenv = getenv("MYDET_ROT_MASK_TILES");        // =1: the single tile kernel (A/B tests, profiling)
            if (tenv && tenv[0] == '1') {
                mask_rot_spatial_kernel<<<mgrid, kTile, 0, st>>>(w, w.m, n, A.thr, A.ge);
            } else {
                MYDET_CUDA(cudaMemsetAsync(w.pair_count, 0, sizeof(int) * (size_t)B, st));
                rot_broad_kernel<<<dim3((w.n32 + kBroadWarps - 1) / kBroadWarps, B), kBroadWarps * 32, 0, st>>>(w, w.m, n, (float)A.thr);
                // a CTA takes 1024 listed pairs per round; 148 CTAs per image cover the typical list (~14 pairs per box) in one
                const int nb = (int)(((long long)w.pair_cap + kNarrowThreads * kNarrowPer - 1) / (kNarrowThreads * kNarrowPer));
                const dim3 ngrid(nb < 1 ? 1 : (nb > 148 ? 148 : nb), B);
                // developer A/B switch MYDET_ROT_LAZY: 0 = clip every listed pair, 2 = lazy for every image; default 1 = per image
                const char* lzenv = getenv("MYDET_ROT_LAZY");
                const int lazy_mode = (lzenv && lzenv[0] >= '0' && lzenv[0] <= '2') ? lzenv[0] - '0' : 1;
                if (lazy_mode == 0) {
                    rot_narrow_kernel<<<ngrid, kNarrowThreads, 0, st>>>(w, w.m, n, A.thr, A.ge, 0);
                } else {
                    // pair2_count, lz_high, lz_dead are carved one after the other: one memset
                    MYDET_CUDA(cudaMemsetAsync(w.pair2_count, 0, (size_t)((char*)w.lz_dead - (char*)w.pair2_count) + (size_t)B * n, st));
                    rot_narrow_kernel<<<ngrid, kNarrowThreads, 0, st>>>(w, w.m, n, A.thr, A.ge, lazy_mode);
                    // the two clip passes are empty launches for every image that is not lazy: a grid of ~8 CTAs per SM over the
                    // batch (grid-stride inside) keeps those cheap and still fills the GPU for the images that are
                    int cg = 1184 / (B < 1 ? 1 : B);
                    cg = cg < 8 ? 8 : (cg > (int)ngrid.x ? (int)ngrid.x : cg);
                    rot_clip_kernel<0><<<dim3(cg, B), kNarrowThreads, 0, st>>>(w, w.m, n, A.thr, A.ge, lazy_mode);
                    rot_clip_kernel<1><<<dim3(cg, B), kNarrowThreads, 0, st>>>(w, w.m, n, A.thr, A.ge, lazy_mode);
                }
                mask_rot_overflow_kernel<<<148 * 16, kTile, 0, st>>>(w, w.m, n, B, A.thr, A.ge);
            }
        } else {
            spatial_gather_kernel<false><<<dim3(tiles, B), kTile, 0, st>>>(G, w.keys, w.order, w.m, w);
            mask_aabb_spatial_kernel<<<mgrid, kTile, 0, st>>>(w, w.m, n, float_at_or_below(A.thr));
        }
        // parallel fixed-point sweep (sweep_fixpoint.cuh) by default since round 2 (rotated 10 k boxes: 63.8 -> 48.8 us per
        // image, dense scene @1536: 1030 -> 301 us); MYDET_SWEEP_FIXPOINT=0 selects the serial block sweep (A/B tests)
        const char* fxenv = getenv("MYDET_SWEEP_FIXPOINT");
        if (!(fxenv && fxenv[0] == '0')) {
            const size_t smem2 = smem + (size_t)w.words * sizeof(unsigned long long);
            const int attr2 = (int)smem2 > 48 * 1024 ? (int)smem2 : 48 * 1024;
            MYDET_CUDA(cudaFuncSetAttribute(sweep_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, attr2));
            sweep_kernel<2><<<B, kSweepThreads, smem2, st>>>(w, w.m, n, E);
        } else {
            spatial_diag_kernel<<<dim3(w.words, B), 512, 0, st>>>(w, w.m, n);
            MYDET_CUDA(cudaFuncSetAttribute(sweep_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr));
            sweep_kernel<1><<<B, kSweepThreads, smem, st>>>(w, w.m, n, E);
        }
    } else {
        if (A.rot) {
            gather_kernel<true><<<dim3((n + 255) / 256, B), 256, 0, st>>>(G, w.keys, w.order, w.m, w);
            mask_rot_kernel<<<dim3(tiles, (tiles + kColChunk - 1) / kColChunk, B), kTile, 0, st>>>(w, w.m, n, A.thr, A.ge);
        } else {
            gather_kernel<false><<<dim3((n + 255) / 256, B), 256, 0, st>>>(G, w.keys, w.order, w.m, w);
            mask_kernel<<<dim3(tiles * (tiles + 1) / 2, B), kTile, 0, st>>>(w, w.m, n, float_at_or_below(A.thr));
        }
        MYDET_CUDA(cudaFuncSetAttribute(sweep_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr));
        sweep_kernel<0><<<B, kSweepThreads, smem, st>>>(w, w.m, n, E);
    }
    if (A.rot && A.votes) {
        MYDET_CUDA(cudaMemsetAsync(A.votes, 0, sizeof(int) * (size_t)B * (size_t)A.pitch, st));
        votes_kernel<<<dim3((n + 127) / 128, B), 128, 0, st>>>(w, w.m, n, A.votes, A.pitch, spatial ? 1 : 0);
    }
    if (spatial && A.ws_clean) mask_cleanup_kernel<<<dim3(32, B), 256, 0, st>>>(w, n);
    // the all-pairs kernels write whole mask rows; a persistent workspace is handed back clean all the same (a later call
    // of the same geometry with an ordinary threshold takes the spatial path and relies on it)
    if (!spatial && A.ws_clean && n <= kSpatialMaxN)
        MYDET_CUDA(cudaMemsetAsync(w.mask, 0, (size_t)B * n * w.words * sizeof(unsigned long long), st));
    return launch_status("large NMS pipeline");
}

}  // namespace mydet
