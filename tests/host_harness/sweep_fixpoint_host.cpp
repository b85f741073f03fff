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
    f = fopen(argv[2], "wb");
    const int head[3] = {rounds, entries, use_list ? 1 : 0};
    if (!f || fwrite(head, sizeof(int), 3, f) != 3 || fwrite(keptw, 8, words_total, f) != (size_t)words_total) return 5;
    fclose(f);
    free(mask); free(adj); free(spos); free(keep); free(removed); free(keptw); free(list);
    return 0;
}
