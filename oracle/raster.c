/* Oracle: rasterised polygon IoU in the manner of pycocotools' maskApi (rleFrPoly + rleIou, iscrowd = 0).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * PARITY UNPINNED.  The reference's iou_rle (utils/bbox_ops.py:84-96) hands the box corners to
 * pycocotools.mask.frPyObjects / iou.  pycocotools is a third-party C extension that is neither vendored by the
 * reference nor installed in this image (no network), so this file restates its PUBLISHED algorithm from the
 * description of maskApi.c -- 5x super-sampled integer boundary walk, column-crossing points, column-major run
 * lengths, run-length intersection -- and could not be checked against the real library.  It is used for ONE
 * thing: to REPORT how far a raster IoU on the reference's 2048 x 2048 canvas lies from the exact polygon IoU
 * that the kernels compute (tests/test_oracle_golden.py::test_raster_vs_exact_gap).
 * Nothing compares kernel results against it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static int cmp_u32(const void* a, const void* b) {
    const uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return x > y ? 1 : (x < y ? -1 : 0);
}

/* polygon (k points, xy interleaved) -> run lengths (alternating 0-runs / 1-runs, column major); returns the number
 * of runs written to *out (malloc'ed). */
static long rle_from_poly(const double* xy, long k, long h, long w, uint32_t** out) {
    const double scale = 5.0;
    long j, m = 0;
    int* x = (int*)calloc((size_t)k + 1, sizeof(int));
    int* y = (int*)calloc((size_t)k + 1, sizeof(int));
    for (j = 0; j < k; j++) x[j] = (int)(scale * xy[j * 2 + 0] + .5);
    x[k] = x[0];
    for (j = 0; j < k; j++) y[j] = (int)(scale * xy[j * 2 + 1] + .5);
    y[k] = y[0];
    for (j = 0; j < k; j++) {
        const int dx = abs(x[j] - x[j + 1]), dy = abs(y[j] - y[j + 1]);
        m += (dx > dy ? dx : dy) + 1;
    }
    int* u = (int*)malloc(sizeof(int) * (m + 1));
    int* v = (int*)malloc(sizeof(int) * (m + 1));
    m = 0;
    for (j = 0; j < k; j++) {                      /* dense integer points along every edge */
        int xs = x[j], xe = x[j + 1], ys = y[j], ye = y[j + 1], t, d;
        const int dx = abs(xe - xs), dy = abs(ys - ye);
        const int flip = (dx >= dy && xs > xe) || (dx < dy && ys > ye);
        if (flip) { t = xs; xs = xe; xe = t; t = ys; ys = ye; ye = t; }
        const double s = dx >= dy ? (dx ? (double)(ye - ys) / dx : 0.0) : (double)(xe - xs) / dy;
        if (dx >= dy) for (d = 0; d <= dx; d++) { t = flip ? dx - d : d; u[m] = t + xs; v[m] = (int)(ys + s * t + .5); m++; }
        else for (d = 0; d <= dy; d++) { t = flip ? dy - d : d; v[m] = t + ys; u[m] = (int)(xs + s * t + .5); m++; }
    }
    free(x); free(y);
    const long kk = m;
    x = (int*)malloc(sizeof(int) * (kk + 1));
    y = (int*)malloc(sizeof(int) * (kk + 1));
    m = 0;
    for (j = 1; j < kk; j++) if (u[j] != u[j - 1]) {   /* points where the boundary crosses a pixel column */
        double xd = (double)(u[j] < u[j - 1] ? u[j] : u[j] - 1);
        xd = (xd + .5) / scale - .5;
        if (floor(xd) != xd || xd < 0 || xd > w - 1) continue;
        double yd = (double)(v[j] < v[j - 1] ? v[j] : v[j - 1]);
        yd = (yd + .5) / scale - .5;
        if (yd < 0) yd = 0; else if (yd > h) yd = h;
        yd = ceil(yd);
        x[m] = (int)xd; y[m] = (int)yd; m++;
    }
    free(u); free(v);
    long n = m;
    uint32_t* a = (uint32_t*)malloc(sizeof(uint32_t) * (n + 1));
    for (j = 0; j < n; j++) a[j] = (uint32_t)(x[j] * (int)h + y[j]);
    a[n++] = (uint32_t)(h * w);
    free(x); free(y);
    qsort(a, n, sizeof(uint32_t), cmp_u32);
    uint32_t p = 0;
    for (j = 0; j < n; j++) { const uint32_t t = a[j]; a[j] -= p; p = t; }
    uint32_t* b = (uint32_t*)malloc(sizeof(uint32_t) * n);
    j = 0; m = 0;
    b[m++] = a[j++];
    while (j < n) {
        if (a[j] > 0) b[m++] = a[j++];
        else { j++; if (j < n) b[m - 1] += a[j++]; }
    }
    free(a);
    *out = b;
    return m;
}

static double rle_iou(const uint32_t* A, long ka, const uint32_t* B, long kb) {
    uint32_t ca = A[0], cb = B[0], c, ct = 1;
    int va = 0, vb = 0;
    long a = 1, b = 1;
    double inter = 0, uni = 0;
    while (ct > 0) {
        c = ca < cb ? ca : cb;
        if (va || vb) { uni += c; if (va && vb) inter += c; }
        ct = 0;
        ca -= c; if (!ca && a < ka) { ca = A[a++]; va = !va; } ct += ca;
        cb -= c; if (!cb && b < kb) { cb = B[b++]; vb = !vb; } ct += cb;
    }
    if (inter == 0) return 0.0;
    return inter / uni;
}

/* run lengths of one polygon (k points), for the known-answer tests; returns the number of runs (written up to cap) */
int64_t oracle_raster_rle(const double* xy, int64_t k, int64_t h, int64_t w, uint32_t* out, int64_t cap) {
    uint32_t* r;
    const long m = rle_from_poly(xy, (long)k, (long)h, (long)w, &r);
    for (long j = 0; j < m && j < cap; j++) out[j] = r[j];
    free(r);
    return m;
}

/* corners1 (n, 8), corners2 (m, 8): x,y of the 4 vertices; out (n, m) */
void oracle_raster_iou_pairwise(const double* c1, int64_t n, const double* c2, int64_t m, int64_t h, int64_t w, double* out) {
    uint32_t** r2 = (uint32_t**)malloc(sizeof(uint32_t*) * (m > 0 ? m : 1));
    long* k2 = (long*)malloc(sizeof(long) * (m > 0 ? m : 1));
    for (int64_t j = 0; j < m; j++) k2[j] = rle_from_poly(c2 + j * 8, 4, h, w, &r2[j]);
    for (int64_t i = 0; i < n; i++) {
        uint32_t* r1;
        const long k1 = rle_from_poly(c1 + i * 8, 4, h, w, &r1);
        for (int64_t j = 0; j < m; j++) out[i * m + j] = rle_iou(r1, k1, r2[j], k2[j]);
        free(r1);
    }
    for (int64_t j = 0; j < m; j++) free(r2[j]);
    free(r2); free(k2);
}
