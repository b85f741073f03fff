// Parallel fixed-point sweep for the large-N NMS (the default since round 2; MYDET_SWEEP_FIXPOINT=0 selects the serial block
// sweep for A/B tests).
//
// The greedy NMS result is the unique fixed point of
//     keep[i] = valid[i] and not any(keep[j] and M[j, i] for j ranked above i)
// (M[j, i]: "j suppresses i"; the mask kernels only ever set the bit in the row of the higher-ranked box), and the
// Jacobi iteration from keep = valid reaches it in (longest decisive suppression chain + 1) rounds: after round t the
// decisions of every box whose chain is shorter than t are final.  scripts/fixpoint_depth.py measures 5-8 rounds on
// every BASELINE workload, where the block sweep of nms_large.cu walks 157-757 blocks of 64 boxes one after the other.
//
// One CTA per image.  Each round:  scatter -- every still-kept row ORs its non-empty mask words (the per-tile adjacency
// map says which words can be non-empty) into a shared `removed` vector;  update -- keep = valid & ~removed.
// The phases are host/device functions of (thread id, thread count) with the barriers BETWEEN them, so that
// tests/host_harness/sweep_fixpoint_host.cpp can run the same code on the CPU, one "thread" after the other, and
// compare it with the serial greedy sweep.  Rows, words and the keep / removed vectors are indexed by SPATIAL position
// (Morton order of the box centres), kept_by_rank by score rank, exactly as the spatial path of nms_large.cu stores them.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FX_HD __host__ __device__ __forceinline__
#else
#define FX_HD inline
#endif

namespace mydet {
namespace fx {

struct View {                            // one image
    const unsigned long long* mask;      // [position][words_total]: bit c of row p = "p suppresses c" (p ranked above c)
    const unsigned long long* tile_adj;  // [tile][aw]: bit j = some row of the tile has a non-zero mask word j
    const int* spos_of_rank;             // score rank -> spatial position
    int mb;                              // valid boxes of the image
    int words_total;                     // row pitch of the mask in 64-bit words
    int aw;                              // adjacency words per tile
};

FX_HD int count_trailing_zeros(unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}

// vec[word] |= bits.  Device: shared-memory atomics (two 32-bit halves, as the block sweep does); the host harness
// runs the "threads" one after the other, so a plain OR is the same thing.
FX_HD void or_bits(unsigned long long* vec, int word, unsigned long long bits) {
#if defined(__CUDA_ARCH__)
    unsigned* v32 = reinterpret_cast<unsigned*>(vec);
    if ((unsigned)bits) atomicOr(&v32[2 * word], (unsigned)bits);
    if ((unsigned)(bits >> 32)) atomicOr(&v32[2 * word + 1], (unsigned)(bits >> 32));
#else
    vec[word] |= bits;
#endif
}

FX_HD unsigned long long valid_word(int mb, int wd) {             // positions [64 wd, 64 wd + 64) that are < mb
    const int lo = wd * 64;
    if (lo >= mb) return 0ull;
    return (mb - lo >= 64) ? ~0ull : ((1ull << (mb - lo)) - 1ull);
}

// keep = valid, removed = 0, kept_by_rank = 0.   (barrier after)
FX_HD void phase_init(const View& V, unsigned long long* keep, unsigned long long* removed, unsigned long long* kept_by_rank,
                      int n_words, int tid, int nt) {
    for (int wd = tid; wd < n_words; wd += nt) {
        keep[wd] = valid_word(V.mb, wd);
        removed[wd] = 0ull;
        kept_by_rank[wd] = 0ull;
    }
}

// Every still-kept row ORs its non-empty words into `removed`, reading the words through the adjacency map.
// (barrier before and after)  Used when the entry list below overflowed.
FX_HD void phase_scatter(const View& V, const unsigned long long* keep, unsigned long long* removed, int tid, int nt) {
    for (int p = tid; p < V.mb; p += nt) {
        if (!((keep[p >> 6] >> (p & 63)) & 1ull)) continue;
        const unsigned long long* row = V.mask + (long long)p * V.words_total;
        const unsigned long long* adj = V.tile_adj + (long long)(p >> 6) * V.aw;
        for (int q = 0; q < V.aw; ++q) {
            unsigned long long a = adj[q];
            while (a) {
                const int wc = q * 64 + count_trailing_zeros(a);
                a &= a - 1ull;
                const unsigned long long v = row[wc];
                if (v) or_bits(removed, wc, v);
            }
        }
    }
}

// The non-empty mask words of the image as a list: the rounds then touch a few thousand entries instead of walking
// the adjacency map of every row again.
struct Entry { int row; int word; unsigned long long bits; };     // 16 bytes

FX_HD int counter_add(int* counter, int v) {                        // returns the old value
#if defined(__CUDA_ARCH__)
    return atomicAdd(counter, v);
#else
    const int old = *counter;
    *counter = old + v;
    return old;
#endif
}

// Appends every non-empty word the adjacency map points at.  *count may end above `cap`: the list has overflowed and the
// caller falls back to phase_scatter.  Up to kBatch independent loads are in flight per thread.  (barrier before: *count = 0;
// barrier after)
constexpr int kBatch = 8;
FX_HD void phase_build_list(const View& V, Entry* list, int cap, int* count, int tid, int nt) {
    for (int p = tid; p < V.mb; p += nt) {
        const unsigned long long* row = V.mask + (long long)p * V.words_total;
        const unsigned long long* adj = V.tile_adj + (long long)(p >> 6) * V.aw;
        for (int q = 0; q < V.aw; ++q) {
            unsigned long long a = adj[q];
            while (a) {
                int wc[kBatch];
                unsigned long long v[kBatch];
                int k = 0;
                for (; k < kBatch && a; ++k) {
                    wc[k] = q * 64 + count_trailing_zeros(a);
                    a &= a - 1ull;
                }
                for (int j = 0; j < kBatch; ++j) v[j] = (j < k) ? row[wc[j]] : 0ull;
                for (int j = 0; j < kBatch; ++j) {
                    if (v[j]) {
                        const int at = counter_add(count, 1);
                        if (at < cap) { list[at].row = p; list[at].word = wc[j]; list[at].bits = v[j]; }
                    }
                }
            }
        }
    }
}

// One round over the list.   (barrier before and after)
FX_HD void phase_scatter_list(const Entry* list, int n_entries, const unsigned long long* keep, unsigned long long* removed,
                              int tid, int nt) {
    for (int e = tid; e < n_entries; e += nt) {
        const Entry x = list[e];
        if ((keep[x.row >> 6] >> (x.row & 63)) & 1ull) or_bits(removed, x.word, x.bits);
    }
}

// Entries recorded by the mask kernels: `word` is a 32-bit HALF-word index of the row (column >> 5).  Fetch the bits
// once (the mask is final when the sweep starts).   (barrier after)
FX_HD void phase_fill_list32(const unsigned* mask32, int words_total, Entry* list, int n_entries, int tid, int nt) {
    for (int e = tid; e < n_entries; e += nt)
        list[e].bits = mask32[(long long)list[e].row * (2 * words_total) + list[e].word];
}
// One round over that list.   (barrier before and after)
FX_HD void phase_scatter_list32(const Entry* list, int n_entries, const unsigned long long* keep, unsigned long long* removed,
                                int tid, int nt) {
    for (int e = tid; e < n_entries; e += nt) {
        const Entry x = list[e];
        if ((keep[x.row >> 6] >> (x.row & 63)) & 1ull)
            or_bits(removed, x.word >> 1, (x.word & 1) ? (x.bits << 32) : x.bits);
    }
}

// keep = valid & ~removed; removed is cleared for the next round.  Returns whether this thread changed a word.
// (barrier before; the caller ORs the return values over the CTA, which is also the barrier after)
FX_HD int phase_update(const View& V, unsigned long long* keep, unsigned long long* removed, int n_words, int tid, int nt) {
    int changed = 0;
    for (int wd = tid; wd < n_words; wd += nt) {
        const unsigned long long nk = valid_word(V.mb, wd) & ~removed[wd];
        changed |= (nk != keep[wd]) ? 1 : 0;
        keep[wd] = nk;
        removed[wd] = 0ull;
    }
    return changed;
}

// Survivors by score rank (the order the emit stage walks).   (barrier before and after)
// The positions of kBatch ranks are loaded before any of them is used: one CTA works on the image, so a loop of
// dependent load -> test -> atomic would pay one memory latency per rank.
FX_HD void phase_to_rank(const View& V, const unsigned long long* keep, unsigned long long* kept_by_rank, int tid, int nt) {
    for (int r0 = tid; r0 < V.mb; r0 += nt * kBatch) {
        int p[kBatch];
        for (int j = 0; j < kBatch; ++j) { const int r = r0 + j * nt; p[j] = (r < V.mb) ? V.spos_of_rank[r] : -1; }
        for (int j = 0; j < kBatch; ++j) {
            const int r = r0 + j * nt;
            if (p[j] >= 0 && ((keep[p[j] >> 6] >> (p[j] & 63)) & 1ull)) or_bits(kept_by_rank, r >> 6, 1ull << (r & 63));
        }
    }
}

// Register-resident entry list of the device sweep: (row : 16 | half-word : 16 | bits : 32) per entry, kRegEntries per
// thread, loaded ONCE (entry coordinates first, then the mask bits, each as a batch of independent loads) -- the rounds
// then touch shared memory only.  Rows < 65 536 on this path (spatial ordering is used up to 65 536 boxes).
constexpr int kRegEntries = 16;
FX_HD void phase_load_entries32(const unsigned* mask32, int words_total, const Entry* list, int n_entries,
                                unsigned long long (&ent)[kRegEntries], int tid, int nt) {
    int row[kRegEntries], half[kRegEntries];
    for (int k = 0; k < kRegEntries; ++k) {
        const int e = tid + k * nt;
        row[k] = (e < n_entries) ? list[e].row : -1;
        half[k] = (e < n_entries) ? list[e].word : 0;
    }
    for (int k = 0; k < kRegEntries; ++k) {
        unsigned bits = 0u;
        if (row[k] >= 0) bits = mask32[(long long)row[k] * (2 * words_total) + half[k]];
        ent[k] = (row[k] >= 0 && bits) ? (((unsigned long long)row[k] << 48) | ((unsigned long long)half[k] << 32) | bits) : 0ull;
    }
}
FX_HD void phase_scatter_entries32(const unsigned long long (&ent)[kRegEntries], const unsigned long long* keep,
                                   unsigned long long* removed) {
    for (int k = 0; k < kRegEntries; ++k) {
        const unsigned long long x = ent[k];
        if (!x) continue;
        const int row = (int)(x >> 48), half = (int)((x >> 32) & 0xffffu);
        if ((keep[row >> 6] >> (row & 63)) & 1ull)
            or_bits(removed, half >> 1, (half & 1) ? ((x & 0xffffffffull) << 32) : (x & 0xffffffffull));
    }
}

}  // namespace fx
}  // namespace mydet
