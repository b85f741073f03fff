// Host harness for the experimental fixed-point sweep (TEST INFRASTRUCTURE, not part of libmydet).
//
// Runs the phases of mydetection_b200/csrc/sweep_fixpoint.cuh -- the code sweep_kernel<2> of nms_large.cu executes --
// on the CPU: every phase is called for tid = 0 .. nt-1 in turn, with the kernel's barriers between the phases, on
// buffers sized exactly as the kernel sizes them (AddressSanitizer build).  tests/test_sweep_fixpoint_host.py feeds it
// suppression matrices in the spatial layout of nms_large.cu and compares the survivors with the serial greedy sweep.
//
//   sweep_fixpoint_host <in.bin> <out.bin> nt list_cap      (list_cap < 0: 4 * mb, as the library sizes it)
// in.bin : int32 mb, words_total, aw; then mask[mb * words_total] u64, tile_adj[ceil(mb/64) * aw] u64, spos_of_rank[mb] i32
// out.bin: int32 rounds, entries, used_list; kept_by_rank[words_total] u64
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../mydetection_b200/csrc/sweep_fixpoint.cuh"

using namespace mydet::fx;

int main(int argc, char** argv) {
    if (argc != 5) return 2;
    const int nt = atoi(argv[3]);
    FILE* f = fopen(argv[1], "rb");
    int hdr[3];
    if (!f || fread(hdr, sizeof(int), 3, f) != 3) return 3;
    const int mb = hdr[0], words_total = hdr[1], aw = hdr[2], tiles = (mb + 63) / 64;
    unsigned long long* mask = (unsigned long long*)malloc(sizeof(unsigned long long) * ((size_t)mb * words_total + 1));
    unsigned long long* adj = (unsigned long long*)malloc(sizeof(unsigned long long) * ((size_t)tiles * aw + 1));
    int* spos = (int*)malloc(sizeof(int) * ((size_t)mb + 1));
    if (fread(mask, 8, (size_t)mb * words_total, f) != (size_t)mb * words_total) return 4;
    if (fread(adj, 8, (size_t)tiles * aw, f) != (size_t)tiles * aw) return 4;
    if (fread(spos, 4, mb, f) != (size_t)mb) return 4;
    fclose(f);
    // the kernel's shared-memory vectors: w.words entries each
    unsigned long long* keep = (unsigned long long*)malloc(8 * (size_t)words_total);
    unsigned long long* removed = (unsigned long long*)malloc(8 * (size_t)words_total);
    unsigned long long* keptw = (unsigned long long*)malloc(8 * (size_t)words_total);
    const View V{mask, adj, spos, mb, words_total, aw};
    const int cap = atoi(argv[4]) < 0 ? 4 * mb : atoi(argv[4]);
    Entry* list = (Entry*)malloc(sizeof(Entry) * (size_t)(cap > 0 ? cap : 1));
    int entries = 0;                                     // the kernel's shared counter
    for (int t = 0; t < nt; ++t) phase_init(V, keep, removed, keptw, words_total, t, nt);
    for (int t = 0; t < nt; ++t) phase_build_list(V, list, cap, &entries, t, nt);
    const bool use_list = entries <= cap;
    int rounds = 0;
    for (;;) {
        for (int t = 0; t < nt; ++t) {
            if (use_list) phase_scatter_list(list, entries, keep, removed, t, nt);
            else phase_scatter(V, keep, removed, t, nt);
        }
        int changed = 0;
        for (int t = 0; t < nt; ++t) changed |= phase_update(V, keep, removed, words_total, t, nt);   // __syncthreads_or
        ++rounds;
        if (!changed) break;
    }
    for (int t = 0; t < nt; ++t) phase_to_rank(V, keep, keptw, t, nt);
    // The path the library takes since the mask kernels record their own entries (fx_note): one entry per non-empty 32-bit
    // half-word, (row, half-word index), here collected by scanning the matrix (in reverse, as "any order").  Both ways of
    // consuming it -- entries filled in memory, entries packed into per-thread registers -- must give the survivors above.
    {
        const unsigned* mask32 = (const unsigned*)mask;
        std::vector<Entry> l32;
        for (long long i = (long long)mb * words_total * 2 - 1; i >= 0; --i)
            if (mask32[i]) { Entry e; e.row = (int)(i / (2 * words_total)); e.word = (int)(i % (2 * words_total)); e.bits = 0; l32.push_back(e); }
        const int n32 = (int)l32.size();
        Entry* lp = n32 ? l32.data() : list;
        std::vector<unsigned long long> keep2(words_total), removed2(words_total), kept2(words_total);
        for (int variant = 0; variant < 2; ++variant) {
            const bool in_regs = variant == 1;
            if (in_regs && n32 > kRegEntries * nt) break;
            for (int t = 0; t < nt; ++t) phase_init(V, keep2.data(), removed2.data(), kept2.data(), words_total, t, nt);
            std::vector<unsigned long long> regs((size_t)nt * kRegEntries, 0ull);
            for (int t = 0; t < nt; ++t) {
                if (in_regs) phase_load_entries32(mask32, words_total, lp, n32, *(unsigned long long (*)[kRegEntries])&regs[(size_t)t * kRegEntries], t, nt);
                else phase_fill_list32(mask32, words_total, lp, n32, t, nt);
            }
            for (;;) {
                for (int t = 0; t < nt; ++t) {
                    if (in_regs) phase_scatter_entries32(*(unsigned long long (*)[kRegEntries])&regs[(size_t)t * kRegEntries], keep2.data(), removed2.data());
                    else phase_scatter_list32(lp, n32, keep2.data(), removed2.data(), t, nt);
                }
                int changed = 0;
                for (int t = 0; t < nt; ++t) changed |= phase_update(V, keep2.data(), removed2.data(), words_total, t, nt);
                if (!changed) break;
            }
            for (int t = 0; t < nt; ++t) phase_to_rank(V, keep2.data(), kept2.data(), t, nt);
            for (int i = 0; i < words_total; ++i)
                if (kept2[i] != keptw[i]) { fprintf(stderr, "entry-list variant %d differs at word %d\n", variant, i); return 6; }
        }
    }
    f = fopen(argv[2], "wb");
    const int head[3] = {rounds, entries, use_list ? 1 : 0};
    if (!f || fwrite(head, sizeof(int), 3, f) != 3 || fwrite(keptw, 8, words_total, f) != (size_t)words_total) return 5;
    fclose(f);
    free(mask); free(adj); free(spos); free(keep); free(removed); free(keptw); free(list);
    return 0;
}
