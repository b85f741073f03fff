// Batched tracklet state machine for rotated boxes: the reference keeps one 10-state
// Kalman filter PER OBJECT in numpy (utils/kalman_filter.py:77-142 RotBBoxKalmanFilter, driven by
// utils/structures.py:447-529 KFTracklet) and steps them in Python loops; here all tracklets of a frame advance in one
// launch, one thread per tracklet, float64 like the reference.  State per tracklet: x[10] = (cx, cy, w, h, angle,
// velocities), P[10][10], score, predictions since the last update.
//   predict:    Q = diag(q^2), rows of the four size-like components (and their velocities) scaled by w*h;
//               x = F x, P = F P F^T + Q with F = [[I, I], [0, I]]; angle state wrapped into [0, 180); score decays
//               by the momentum from the second prediction in a row on
//   update:     measurement angle wrapped into [0, 180) then moved by +-180 towards the angle state; R scaled like Q;
//               S = P[:5,:5] + R, K = P[:, :5] S^-1, x += K (z - x[:5]), P -= K P[:5, :]; score momentum
//   likelihood: N(cand; x[:5], P[:5,:5]) for M candidate boxes per tracklet
// 5x5 inverses: Gauss-Jordan with partial pivoting (numpy uses LAPACK's LU; results agree to ~1e-12 relative).
#include "common.cuh"

namespace mydet {

struct KfNoise { double p0[10], q[10], r[5], momentum; };

__device__ __forceinline__ double wrap180(double a) {          // numpy's a % 180 for floats
    double m = fmod(a, 180.0);
    if (m != 0.0 && m < 0.0) m += 180.0;
    return m;
}
__device__ __forceinline__ bool size_like(int i) { return (i % 5) != 4; }    // kalman_filter.py:94 _xywh_mask

// inverse of a 5x5 matrix (row-major, overwritten by the identity), returns the determinant
__device__ double inv5(double a[5][5], double inv[5][5]) {
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) inv[i][j] = (i == j) ? 1.0 : 0.0;
    double det = 1.0;
#pragma unroll 1
    for (int c = 0; c < 5; ++c) {
        int piv = c;
        double best = fabs(a[c][c]);
        for (int r = c + 1; r < 5; ++r) if (fabs(a[r][c]) > best) { best = fabs(a[r][c]); piv = r; }
        if (piv != c) {
            for (int j = 0; j < 5; ++j) {
                double t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t;
                t = inv[c][j]; inv[c][j] = inv[piv][j]; inv[piv][j] = t;
            }
            det = -det;
        }
        const double d = a[c][c];
        det *= d;
        const double s = 1.0 / d;
        for (int j = 0; j < 5; ++j) { a[c][j] *= s; inv[c][j] *= s; }
        for (int r = 0; r < 5; ++r) {
            if (r == c) continue;
            const double f = a[r][c];
            if (f != 0.0)
                for (int j = 0; j < 5; ++j) { a[r][j] -= f * a[c][j]; inv[r][j] -= f * inv[c][j]; }
        }
    }
    return det;
}

__global__ void kf_initiate_kernel(const double* __restrict__ boxes, int n, double* __restrict__ x, double* __restrict__ P,
                                   int* __restrict__ pred_count, const KfNoise nz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double b[5];
    for (int k = 0; k < 5; ++k) b[k] = boxes[(long long)i * 5 + k];
    b[4] = wrap180(b[4]);                                                        // structures.py:463
    double* xi = x + (long long)i * 10;
    double* Pi = P + (long long)i * 100;
    for (int k = 0; k < 5; ++k) { xi[k] = b[k]; xi[5 + k] = 0.0; }
    const double wh = b[2] * b[3];
    for (int r = 0; r < 10; ++r)
        for (int c = 0; c < 10; ++c) Pi[r * 10 + c] = (r == c) ? nz.p0[r] * nz.p0[r] * (size_like(r) ? wh : 1.0) : 0.0;
    if (pred_count) pred_count[i] = 0;
}

__global__ void kf_predict_kernel(double* __restrict__ x, double* __restrict__ P, double* __restrict__ score,
                                  int* __restrict__ pred_count, int n, double* __restrict__ out, const KfNoise nz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* xi = x + (long long)i * 10;
    double* Pi = P + (long long)i * 100;
    const double wh = xi[2] * xi[3];                                              // kalman_filter.py:108: the state BEFORE the step
    // P' = F P F^T + Q,  F = [[I, I], [0, I]]:  (F P)[r][c] = P[r][c] + P[r+5][c] for r < 5;  (M F^T)[r][c] = M[r][c] + M[r][c+5] for c < 5
    for (int r = 0; r < 5; ++r)
        for (int c = 0; c < 10; ++c) Pi[r * 10 + c] += Pi[(r + 5) * 10 + c];
    for (int r = 0; r < 10; ++r)
        for (int c = 0; c < 5; ++c) Pi[r * 10 + c] += Pi[r * 10 + c + 5];
    for (int d = 0; d < 10; ++d) Pi[d * 10 + d] += nz.q[d] * nz.q[d] * (size_like(d) ? wh : 1.0);
    for (int k = 0; k < 5; ++k) {
        const double v = xi[k] + xi[5 + k];
        xi[k] = v;
        if (out) out[(long long)i * 5 + k] = v;                                   // the returned box keeps the un-wrapped angle
    }
    xi[4] = wrap180(xi[4]);                                                       // structures.py:478
    if (pred_count) {
        if (score && pred_count[i] >= 1) score[i] = nz.momentum * score[i];       // :482-483
        pred_count[i] += 1;
    }
}

__global__ void kf_update_kernel(double* __restrict__ x, double* __restrict__ P, double* __restrict__ score,
                                 int* __restrict__ pred_count, const double* __restrict__ meas,
                                 const double* __restrict__ meas_score, const unsigned char* __restrict__ has, int n,
                                 double* __restrict__ out, const KfNoise nz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (has && !has[i]) {
        if (out) for (int k = 0; k < 5; ++k) out[(long long)i * 5 + k] = 0.0;
        return;
    }
    double* xi = x + (long long)i * 10;
    double* Pi = P + (long long)i * 100;
    double z[5];
    for (int k = 0; k < 5; ++k) z[k] = meas[(long long)i * 5 + k];
    {   // structures.py:494-497: the measured angle, modulo 180, on the side of the state's angle
        const double a = wrap180(z[4]), st = xi[4];
        double best = a;
        if (fabs(a - 180.0 - st) < fabs(best - st)) best = a - 180.0;
        if (fabs(a + 180.0 - st) < fabs(best - st)) best = a + 180.0;
        z[4] = best;
    }
    const double wh = xi[2] * xi[3];
    double S[5][5], Si[5][5], y[5];
    for (int r = 0; r < 5; ++r) {
        y[r] = z[r] - xi[r];
        for (int c = 0; c < 5; ++c) S[r][c] = Pi[r * 10 + c] + ((r == c) ? nz.r[r] * nz.r[r] * (size_like(r) ? wh : 1.0) : 0.0);
    }
    inv5(S, Si);
    double K[10][5];
    for (int r = 0; r < 10; ++r)
        for (int c = 0; c < 5; ++c) {
            double acc = 0.0;
            for (int k = 0; k < 5; ++k) acc += Pi[r * 10 + k] * Si[k][c];
            K[r][c] = acc;
        }
    double top[5][10];                                                            // P[:5, :] before it changes
    for (int r = 0; r < 5; ++r)
        for (int c = 0; c < 10; ++c) top[r][c] = Pi[r * 10 + c];
    for (int r = 0; r < 10; ++r) {
        double dx = 0.0;
        for (int k = 0; k < 5; ++k) dx += K[r][k] * y[k];
        xi[r] += dx;
        for (int c = 0; c < 10; ++c) {
            double acc = 0.0;
            for (int k = 0; k < 5; ++k) acc += K[r][k] * top[k][c];
            Pi[r * 10 + c] -= acc;
        }
    }
    if (out) for (int k = 0; k < 5; ++k) out[(long long)i * 5 + k] = xi[k];
    xi[4] = wrap180(xi[4]);                                                       // :499
    if (score && meas_score) score[i] = nz.momentum * score[i] + (1.0 - nz.momentum) * meas_score[i];   // :501
    if (pred_count) pred_count[i] = 0;
}

// one CTA per tracklet: the 5x5 inverse and determinant once, then the candidates in parallel
__global__ void __launch_bounds__(128) kf_likelihood_kernel(const double* __restrict__ x, const double* __restrict__ P, int n,
                                                            const double* __restrict__ cand, int m, double* __restrict__ out) {
    __shared__ double s_inv[5][5], s_mean[5], s_norm;
    const int i = blockIdx.x;
    if (threadIdx.x == 0) {
        double a[5][5], inv[5][5];
        for (int r = 0; r < 5; ++r) {
            s_mean[r] = x[(long long)i * 10 + r];
            for (int c = 0; c < 5; ++c) a[r][c] = P[(long long)i * 100 + r * 10 + c];
        }
        const double det = inv5(a, inv);
        for (int r = 0; r < 5; ++r) for (int c = 0; c < 5; ++c) s_inv[r][c] = inv[r][c];
        const double two_pi = 6.283185307179586;
        s_norm = sqrt(two_pi * two_pi * two_pi * two_pi * two_pi * det);          // structures.py:526
    }
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        double d[5], t[5];
        for (int k = 0; k < 5; ++k) d[k] = cand[(long long)j * 5 + k] - s_mean[k];
        double q = 0.0;
        for (int c = 0; c < 5; ++c) {                                             // (d @ inv) * d summed: :524-525
            t[c] = 0.0;
            for (int k = 0; k < 5; ++k) t[c] += d[k] * s_inv[k][c];
            q += t[c] * d[c];
        }
        out[(long long)i * m + j] = exp(-0.5 * q) / s_norm;
    }
}

static int fill_noise(KfNoise& nz, const double* noise) {
    MYDET_REQUIRE(noise, "NULL noise description (10 initial, 10 process, 5 measurement standard deviations, momentum)");
    for (int k = 0; k < 10; ++k) { nz.p0[k] = noise[k]; nz.q[k] = noise[10 + k]; }
    for (int k = 0; k < 5; ++k) nz.r[k] = noise[20 + k];
    nz.momentum = noise[25];
    return 0;
}

}  // namespace mydet

using namespace mydet;

MYDET_API int mydet_kf_initiate(const double* boxes, int n, const double* noise, double* x, double* P, int32_t* pred_count,
                                void* stream) {
    KfNoise nz;
    if (int rc = fill_noise(nz, noise)) return rc;
    MYDET_REQUIRE(n >= 0, "negative tracklet count");
    if (n == 0) return 0;
    MYDET_REQUIRE(boxes && x && P, "NULL tensor pointer");
    kf_initiate_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(boxes, n, x, P, pred_count, nz);
    return launch_status("kf_initiate_kernel");
}

MYDET_API int mydet_kf_predict(double* x, double* P, double* score, int32_t* pred_count, int n, const double* noise,
                               double* out_boxes, void* stream) {
    KfNoise nz;
    if (int rc = fill_noise(nz, noise)) return rc;
    MYDET_REQUIRE(n >= 0, "negative tracklet count");
    if (n == 0) return 0;
    MYDET_REQUIRE(x && P, "NULL tensor pointer");
    kf_predict_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(x, P, score, pred_count, n, out_boxes, nz);
    return launch_status("kf_predict_kernel");
}

MYDET_API int mydet_kf_update(double* x, double* P, double* score, int32_t* pred_count, const double* meas,
                              const double* meas_score, const uint8_t* has, int n, const double* noise, double* out_boxes,
                              void* stream) {
    KfNoise nz;
    if (int rc = fill_noise(nz, noise)) return rc;
    MYDET_REQUIRE(n >= 0, "negative tracklet count");
    if (n == 0) return 0;
    MYDET_REQUIRE(x && P && meas, "NULL tensor pointer");
    kf_update_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(x, P, score, pred_count, meas, meas_score, has, n, out_boxes, nz);
    return launch_status("kf_update_kernel");
}

MYDET_API int mydet_kf_likelihood(const double* x, const double* P, int n, const double* cand, int m, double* out,
                                  void* stream) {
    MYDET_REQUIRE(n >= 0 && m >= 0, "negative size");
    if (n == 0 || m == 0) return 0;
    MYDET_REQUIRE(x && P && cand && out, "NULL tensor pointer");
    kf_likelihood_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(x, P, n, cand, m, out);
    return launch_status("kf_likelihood_kernel");
}
