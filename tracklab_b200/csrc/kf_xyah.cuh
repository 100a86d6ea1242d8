// 8-d (x, y, a, h, vx, vy, va, vh) constant-velocity Kalman filter, one thread per track, fp64.
//
// Device restatement of the three xyah filters of the reference (only the noise models differ):
//   ByteTrack   /root/reference/plugins/track/byte_track/kalman_filter.py:55-226
//   StrongSORT  /root/reference/plugins/track/strong_sort/sort/kalman_filter.py:47-214
//   BPBReID     /root/reference/plugins/track/bpbreid_strong_sort/sort/kalman_filter.py:47-227
//
// F = [[I, I], [0, I]] and H = [I, 0] only contain 0/1, so F P F^T and H P H^T reduce to sums of two
// entries with a single rounding each: those steps are bit-identical to NumPy/BLAS whatever its
// blocking. The 4x4 Cholesky, the gain solve and K S K^T are textbook order (ulp-level differences
// against LAPACK are expected and covered by the 1e-6 px tolerance of the parity tests).
#pragma once
#include "tk_common.cuh"

namespace tk {

// mean' = mean F^T ; P' = F P F^T + diag(q)
__device__ __forceinline__ void kf8_predict(double* __restrict__ m, double* __restrict__ P, const double* q) {
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = m[i] + m[i + 4];
    // left = F P  (rows 0..3 get row i + row i+4)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) P[i * 8 + j] = P[i * 8 + j] + P[(i + 4) * 8 + j];
    // out = left F^T (cols 0..3 get col j + col j+4)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) P[i * 8 + j] = P[i * 8 + j] + P[i * 8 + j + 4];
#pragma unroll
    for (int i = 0; i < 8; ++i) P[i * 9] = P[i * 9] + q[i];
}

// lower Cholesky of the 4x4 S = P[:4,:4] + diag(r); returns false when S is not positive definite.
// invd[i] = 1 / L[i][i]: the factorisation and the triangular solves multiply by these four reciprocals
// instead of issuing ~80 fp64 divisions per update (each a ~35-instruction subroutine on sm_100a).
__device__ __forceinline__ bool kf8_chol4(const double* __restrict__ P, const double* r, double* L /*16*/, double* S /*16*/,
                                          double* invd /*4*/) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) S[i * 4 + j] = (i == j) ? P[i * 8 + j] + r[i] : P[i * 8 + j];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double d = S[j * 4 + j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[j * 4 + k] * L[j * 4 + k];
        ok = ok && (d > 0.0);
        const double ljj = sqrt(d);
        L[j * 4 + j] = ljj;
        invd[j] = 1.0 / ljj;
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            double s = S[i * 4 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= L[i * 4 + k] * L[j * 4 + k];
            L[i * 4 + j] = s * invd[j];
        }
    }
    return ok;
}

// Measurement update with z[4] and measurement-noise variances r[4] (diagonal).
static __device__ __noinline__ bool kf8_update(double* __restrict__ m, double* __restrict__ P, const double* z, const double* r) {
    double L[16], S[16], invd[4];
    const bool ok = kf8_chol4(P, r, L, S, invd);
    // gain K[8][4]: solve S X = (P H^T)^T column by column, K = X^T
    double K[32];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        double y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double s = P[c * 8 + i];
#pragma unroll
            for (int k = 0; k < i; ++k) s -= L[i * 4 + k] * y[k];
            y[i] = s * invd[i];
        }
#pragma unroll
        for (int i = 3; i >= 0; --i) {
            double s = y[i];
#pragma unroll
            for (int k = i + 1; k < 4; ++k) s -= L[k * 4 + i] * y[k];
            y[i] = s * invd[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) K[c * 4 + i] = y[i];
    }
    double inn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) inn[i] = z[i] - m[i];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) s += inn[i] * K[c * 4 + i];
        m[c] = m[c] + s;
    }
    // P -= K (S K^T)
    double T[32];  // T[i][b] = sum_k S[i][k] K[b][k]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += S[i * 4 + k] * K[b * 4 + k];
            T[i * 8 + b] = s;
        }
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) s += K[a * 4 + i] * T[i * 8 + b];
            P[a * 8 + b] = P[a * 8 + b] - s;
        }
    return ok;
}

// squared Mahalanobis distance of z[4] to (H m, H P H^T + diag(r)) given the Cholesky factor L of S
__device__ __forceinline__ double kf8_maha(const double* m, const double* L, const double* invd, const double* z) {
    double y[4], acc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double s = z[i] - m[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= L[i * 4 + k] * y[k];
        y[i] = s * invd[i];
        acc += y[i] * y[i];
    }
    return acc;
}

}  // namespace tk
