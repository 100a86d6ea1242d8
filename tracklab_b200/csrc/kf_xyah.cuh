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

// ---- octet-cooperative variants: 8 consecutive lanes own one track, lane j holds row j of P and mean[j] ------------
// One thread per track keeps 100+ doubles live (spills to local memory, ~13k cycles of dependent latency per update
// on sm_100a); spreading a track over 8 lanes keeps everything in registers, makes the 512-byte covariance one
// coalesced read per octet, and cuts the critical path to the 4x4 Cholesky (done redundantly by the 8 lanes).
// The arithmetic (formulas and summation order) is identical to kf8_predict / kf8_update above.
// All 32 lanes of the warp must call these; `active` is uniform within an octet.

__device__ __forceinline__ double oct_bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src, 8); }

// q = process-noise variance of this lane's state index j (caller computes the model-specific value)
__device__ __forceinline__ void kf8_octet_predict(double* __restrict__ gm, double* __restrict__ gP, bool active, bool zero_vh, double qj) {
    const int j = threadIdx.x & 7;
    double row[8], m = 0.0;
    if (active) {
        m = gm[j];
#pragma unroll
        for (int c = 0; c < 8; ++c) row[c] = gP[j * 8 + c];
    } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) row[c] = 0.0;
    }
    if (zero_vh && j == 7) m = 0.0;
    const double m_hi = oct_bcast(m, (j + 4) & 7);
    if (j < 4) m = m + m_hi;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const double hi = oct_bcast(row[c], (j + 4) & 7);
        if (j < 4) row[c] = row[c] + hi;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) row[c] = row[c] + row[c + 4];
#pragma unroll
    for (int c = 0; c < 8; ++c) if (c == j) row[c] = row[c] + qj;
    if (active) {
        gm[j] = m;
#pragma unroll
        for (int c = 0; c < 8; ++c) gP[j * 8 + c] = row[c];
    }
}

// z[4]: measurement, r[4]: measurement-noise variances (both uniform within the octet). Returns false if S is not PD.
__device__ __forceinline__ bool kf8_octet_update(double* __restrict__ gm, double* __restrict__ gP, bool active,
                                                 const double* z, const double* r) {
    const int j = threadIdx.x & 7;
    double row[8], m = 0.0;
    if (active) {
        m = gm[j];
#pragma unroll
        for (int c = 0; c < 8; ++c) row[c] = gP[j * 8 + c];
    } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) row[c] = (c == j) ? 1.0 : 0.0;
    }
    double S[16], L[16], invd[4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double v = oct_bcast(row[b], a);
            S[a * 4 + b] = (a == b) ? v + r[a] : v;
        }
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double d = S[c * 4 + c];
#pragma unroll
        for (int k = 0; k < c; ++k) d -= L[c * 4 + k] * L[c * 4 + k];
        ok = ok && (d > 0.0);
        const double lcc = sqrt(d);
        L[c * 4 + c] = lcc;
        invd[c] = 1.0 / lcc;
#pragma unroll
        for (int i = c + 1; i < 4; ++i) {
            double s = S[i * 4 + c];
#pragma unroll
            for (int k = 0; k < c; ++k) s -= L[i * 4 + k] * L[c * 4 + k];
            L[i * 4 + c] = s * invd[c];
        }
    }
    // gain row of this lane: K[j][:] = solve(S, P[j][:4])
    double K[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double s = row[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= L[i * 4 + k] * K[k];
        K[i] = s * invd[i];
    }
#pragma unroll
    for (int i = 3; i >= 0; --i) {
        double s = K[i];
#pragma unroll
        for (int k = i + 1; k < 4; ++k) s -= L[k * 4 + i] * K[k];
        K[i] = s * invd[i];
    }
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double mi = oct_bcast(m, i);
        acc += (z[i] - mi) * K[i];
    }
    m = m + acc;
    // T[:, j] = S K[j]^T, then P[j][b] -= sum_i K[j][i] T[i][b]
    double t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) s += S[i * 4 + k] * K[k];
        t[i] = s;
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) s += K[i] * oct_bcast(t[i], b);
        row[b] = row[b] - s;
    }
    if (active) {
        gm[j] = m;
#pragma unroll
        for (int c = 0; c < 8; ++c) gP[j * 8 + c] = row[c];
    }
    return ok || !active;
}

}  // namespace tk
