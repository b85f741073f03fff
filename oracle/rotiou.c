/* Oracle: rotated-box IoU and rotated greedy NMS on the CPU.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED for the IoU
 * values: the reference rasterises the two polygons on a 2048x2048 canvas with
 * pycocotools (utils/bbox_ops.py:84-96), a third-party C extension that is not
 * vendored, not pinned, not installed and not fetchable here.  This file instead
 * computes the EXACT area of the intersection of the same two polygons:
 *   - corners are built exactly as the reference builds them, in float32:
 *     degrees -> radians as  a * pi / 180   (bbox_ops.py:88-89), then
 *     verti = (h/2)(sin, -cos), hori = (w/2)(cos, sin),
 *     tl,tr,br,bl = c+verti-hori, c+verti+hori, c-verti+hori, c-verti-hori
 *     (xywha2vertex, bbox_ops.py:137-172);
 *   - the corners are widened to double (the reference's .tolist(), :91-92) and
 *     polygon A is clipped against the four edges of polygon B
 *     (Sutherland-Hodgman) in float64; areas by the shoelace formula;
 *   - IoU = inter / (area_A + area_B - inter), 0 when the union is not positive.
 * oracle_nms_rot wraps it in nms_rotbb's control flow (bbox_ops.py:276-306):
 * descending score order (ties: lower index first -- the declared policy, the
 * reference's argsort is unstable), first box always valid, box i dropped iff
 * any IoU with an already-valid box is >= thr, optional majority vote.
 * Build: gcc -O2 -ffp-contract=off -lm.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

void oracle_argsort_desc_stable(const float* key, int64_t n, int64_t* order);

static const float PI_F = 3.14159265358979323846f; /* float32(math.pi), as torch casts the scalar */

typedef struct { double x[4], y[4]; double cx, cy, r, area; } rquad_t;

static void make_quad(const float* b, rquad_t* q) {
    float rad = b[4] * PI_F / 180.0f;
    float s = sinf(rad), c = cosf(rad);
    float hh = b[3] / 2, hw = b[2] / 2;
    float vx = hh * s, vy = -hh * c;
    float hx = hw * c, hy = hw * s;
    float px[4], py[4];
    px[0] = b[0] + vx - hx; py[0] = b[1] + vy - hy;   /* tl */
    px[1] = b[0] + vx + hx; py[1] = b[1] + vy + hy;   /* tr */
    px[2] = b[0] - vx + hx; py[2] = b[1] - vy + hy;   /* br */
    px[3] = b[0] - vx - hx; py[3] = b[1] - vy - hy;   /* bl */
    double a2 = 0.0;
    for (int k = 0; k < 4; ++k) { q->x[k] = px[k]; q->y[k] = py[k]; }
    for (int k = 0; k < 4; ++k) {
        int n = (k + 1) & 3;
        a2 += q->x[k] * q->y[n] - q->x[n] * q->y[k];
    }
    q->area = 0.5 * a2;                     /* signed */
    q->cx = b[0]; q->cy = b[1];
    q->r = 0.5 * sqrt((double)b[2] * b[2] + (double)b[3] * b[3]);
}

static double quad_iou(const rquad_t* A, const rquad_t* B) {
    double areaA = fabs(A->area), areaB = fabs(B->area);
    /* exact cull: circumscribed circles do not touch => empty intersection.
       The 1e-3 slack covers the float32 rounding of the corners. */
    double dx = A->cx - B->cx, dy = A->cy - B->cy, rr = A->r + B->r + 1e-3;
    double inter = 0.0;
    if (dx * dx + dy * dy <= rr * rr && areaA > 0.0 && areaB > 0.0) {
        double px[16], py[16], qx[16], qy[16];
        int n = 4;
        for (int k = 0; k < 4; ++k) { px[k] = A->x[k]; py[k] = A->y[k]; }
        double sgn = B->area >= 0.0 ? 1.0 : -1.0;
        for (int e = 0; e < 4 && n > 0; ++e) {
            double ax = B->x[e], ay = B->y[e];
            double ex = B->x[(e + 1) & 3] - ax, ey = B->y[(e + 1) & 3] - ay;
            int m = 0;
            for (int k = 0; k < n; ++k) {
                int k2 = (k + 1 == n) ? 0 : k + 1;
                double dp = sgn * (ex * (py[k] - ay) - ey * (px[k] - ax));
                double dq = sgn * (ex * (py[k2] - ay) - ey * (px[k2] - ax));
                /* <= 8 vertices in exact arithmetic; a 9th born of rounding is dropped (as rotgeom.cuh does) */
                if (dp >= 0.0 && m < 8) { qx[m] = px[k]; qy[m] = py[k]; ++m; }
                if ((dp >= 0.0) != (dq >= 0.0) && m < 8) {
                    double t = dp / (dp - dq);
                    qx[m] = px[k] + t * (px[k2] - px[k]);
                    qy[m] = py[k] + t * (py[k2] - py[k]);
                    ++m;
                }
            }
            n = m;
            for (int k = 0; k < n; ++k) { px[k] = qx[k]; py[k] = qy[k]; }
        }
        if (n >= 3) {
            double a2 = 0.0;
            for (int k = 0; k < n; ++k) {
                int k2 = (k + 1 == n) ? 0 : k + 1;
                a2 += px[k] * py[k2] - px[k2] * py[k];
            }
            inter = 0.5 * fabs(a2);
        }
    }
    double uni = areaA + areaB - inter;
    return uni > 0.0 ? inter / uni : 0.0;
}

/* out[i*m + j] = IoU(b1[i], b2[j]); boxes are (cx,cy,w,h,degrees) float32. */
void oracle_rot_iou_pairwise(const float* b1, int64_t n, const float* b2, int64_t m, double* out) {
    rquad_t* Q2 = (rquad_t*)malloc((size_t)(m > 0 ? m : 1) * sizeof(rquad_t));
    for (int64_t j = 0; j < m; ++j) make_quad(b2 + 5 * j, &Q2[j]);
    for (int64_t i = 0; i < n; ++i) {
        rquad_t A; make_quad(b1 + 5 * i, &A);
        for (int64_t j = 0; j < m; ++j) out[i * m + j] = quad_iou(&A, &Q2[j]);
    }
    free(Q2);
}

/* out[i*m + j] = IoU of two quadrilaterals given by their corners (x0,y0,...,x3,y3) in double, the form the
 * reference hands to pycocotools (utils/bbox_ops.py:91-96).  Used by oracle/refload.py's pycocotools stand-in. */
static void quad_from_corners(const double* v, rquad_t* q) {
    double a2 = 0.0, sx = 0.0, sy = 0.0, r2 = 0.0;
    for (int k = 0; k < 4; ++k) { q->x[k] = v[2 * k]; q->y[k] = v[2 * k + 1]; sx += v[2 * k]; sy += v[2 * k + 1]; }
    for (int k = 0; k < 4; ++k) {
        int n = (k + 1) & 3;
        a2 += q->x[k] * q->y[n] - q->x[n] * q->y[k];
    }
    q->area = 0.5 * a2; q->cx = 0.25 * sx; q->cy = 0.25 * sy;
    for (int k = 0; k < 4; ++k) {
        double dx = q->x[k] - q->cx, dy = q->y[k] - q->cy;
        if (dx * dx + dy * dy > r2) r2 = dx * dx + dy * dy;
    }
    q->r = sqrt(r2);
}

void oracle_quad_iou_pairwise(const double* q1, int64_t n, const double* q2, int64_t m, double* out) {
    rquad_t* Q2 = (rquad_t*)malloc((size_t)(m > 0 ? m : 1) * sizeof(rquad_t));
    for (int64_t j = 0; j < m; ++j) quad_from_corners(q2 + 8 * j, &Q2[j]);
    for (int64_t i = 0; i < n; ++i) {
        rquad_t A; quad_from_corners(q1 + 8 * i, &A);
        for (int64_t j = 0; j < m; ++j) out[i * m + j] = quad_iou(&A, &Q2[j]);
    }
    free(Q2);
}

/* nms_rotbb control flow.  majority <= 0 means None.  Returns the number kept. */
int64_t oracle_nms_rot(const float* boxes, const float* scores, int64_t n, double thr,
                       int64_t majority, int64_t* keep) {
    if (n <= 0) return 0;
    int64_t* order = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    rquad_t* Q = (rquad_t*)malloc((size_t)n * sizeof(rquad_t));
    int64_t* valid = (int64_t*)malloc((size_t)n * sizeof(int64_t)); /* sorted positions */
    int64_t* votes = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    oracle_argsort_desc_stable(scores, n, order);
    for (int64_t k = 0; k < n; ++k) make_quad(boxes + 5 * order[k], &Q[k]);
    int64_t nv = 0;
    valid[nv] = 0; votes[nv] = 1; ++nv;
    for (int64_t i = 1; i < n; ++i) {
        int hit = 0; double best = -1.0; int64_t best_v = 0;
        for (int64_t v = 0; v < nv; ++v) {
            double iou = quad_iou(&Q[i], &Q[valid[v]]);
            if (iou >= thr) hit = 1;
            if (iou > best) { best = iou; best_v = v; }   /* argmax: first maximum */
            if (hit && majority <= 0) break;
        }
        if (hit) { if (majority > 0) votes[best_v] += 1; continue; }
        valid[nv] = i; votes[nv] = 1; ++nv;
    }
    int64_t kept = 0;
    for (int64_t v = 0; v < nv; ++v)
        if (majority <= 0 || votes[v] >= majority) keep[kept++] = order[valid[v]];
    free(order); free(Q); free(valid); free(votes);
    return kept;
}
