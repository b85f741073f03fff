/* Oracle: greedy NMS on the CPU.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * oracle_nms_aabb restates the CPU kernel behind torchvision.ops.nms, the
 * third-party routine the reference calls at utils/structures.py:133,136,143,162
 * (torchvision is unpinned in the reference; 0.26.0 is installed here).  Behaviour
 * restated from its published algorithm and pinned in tests/test_oracle_golden.py
 * against the installed library:
 *   - boxes are x1,y1,x2,y2 float32; areas = (x2-x1)*(y2-y1) in float32;
 *   - candidates are visited in STABLE descending score order;
 *   - inter = max(0,xx2-xx1)*max(0,yy2-yy1);  ovr = inter/(area_i+area_j-inter), float32;
 *   - j is suppressed iff (double)ovr > thr  (strict; thr is a double);
 *   - 0/0 = NaN compares false, so degenerate boxes are kept;
 *   - the result is the kept indices in descending score order.
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction: the generic x86-64
 * torchvision wheel has none either).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Stable descending argsort by merge sort.  Ties keep ascending input index,
 * i.e. the policy "score desc, index asc" that DESIGN.md declares for F5. */
static void merge_sort_desc(const float* key, int64_t* idx, int64_t* tmp, int64_t n) {
    for (int64_t width = 1; width < n; width *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * width) {
            int64_t mid = lo + width < n ? lo + width : n;
            int64_t hi = lo + 2 * width < n ? lo + 2 * width : n;
            int64_t a = lo, b = mid, o = lo;
            while (a < mid && b < hi) {
                /* take from the right run only if strictly greater: stability */
                if (key[idx[b]] > key[idx[a]]) tmp[o++] = idx[b++];
                else tmp[o++] = idx[a++];
            }
            while (a < mid) tmp[o++] = idx[a++];
            while (b < hi) tmp[o++] = idx[b++];
        }
        memcpy(idx, tmp, (size_t)n * sizeof(int64_t));
    }
}

void oracle_argsort_desc_stable(const float* key, int64_t n, int64_t* order) {
    int64_t* tmp = (int64_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) order[i] = i;
    merge_sort_desc(key, order, tmp, n);
    free(tmp);
}

/* returns the number of kept boxes; keep[] receives their indices (score desc). */
int64_t oracle_nms_aabb(const float* xyxy, const float* scores, int64_t n, double thr,
                        int64_t* keep) {
    if (n <= 0) return 0;
    int64_t* order = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    float* area = (float*)malloc((size_t)n * sizeof(float));
    unsigned char* dead = (unsigned char*)calloc((size_t)n, 1);
    oracle_argsort_desc_stable(scores, n, order);
    for (int64_t i = 0; i < n; ++i) {
        const float* b = xyxy + 4 * i;
        float w = b[2] - b[0], h = b[3] - b[1];
        area[i] = w * h;
    }
    int64_t kept = 0;
    for (int64_t oi = 0; oi < n; ++oi) {
        int64_t i = order[oi];
        if (dead[i]) continue;
        keep[kept++] = i;
        const float* bi = xyxy + 4 * i;
        for (int64_t oj = oi + 1; oj < n; ++oj) {
            int64_t j = order[oj];
            if (dead[j]) continue;
            const float* bj = xyxy + 4 * j;
            /* std::max(a,b) = (a<b)?b:a ; std::min(a,b) = (b<a)?b:a */
            float xx1 = bi[0] < bj[0] ? bj[0] : bi[0];
            float yy1 = bi[1] < bj[1] ? bj[1] : bi[1];
            float xx2 = bj[2] < bi[2] ? bj[2] : bi[2];
            float yy2 = bj[3] < bi[3] ? bj[3] : bi[3];
            float w = xx2 - xx1; w = (0.0f < w) ? w : 0.0f;
            float h = yy2 - yy1; h = (0.0f < h) ? h : 0.0f;
            float inter = w * h;
            float uni = area[i] + area[j];
            uni = uni - inter;
            float ovr = inter / uni;
            if ((double)ovr > thr) dead[j] = 1;
        }
    }
    free(order); free(area); free(dead);
    return kept;
}
