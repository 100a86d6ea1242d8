// Stateless, batched cost-matrix kernels and the batched assignment solver (C ABI used by the stress sweep of
// BASELINE.json configs[4] and by callers that want the building blocks without a whole-video tracker).
//
//   tk_iou_matrix     pairwise overlap family, float64, boxes x1y1x2y2
//                     /root/reference/plugins/track/oc_sort/association.py:5-21 (iou), :24-55 (giou), :58-95 (diou), :97-147 (ciou)
//   tk_iou_p1_f32     ByteTrack "+1 pixel" IoU distance in float32  /root/reference/plugins/track/byte_track/matching.py:51-89,182-218
//   tk_cosine_dist    1 - a_hat . b_hat^T on float32 features        /root/reference/plugins/track/strong_sort/sort/nn_matching.py:30-49
//   tk_lap_batched    one assignment problem per CTA (lap.cuh)       byte_track/matching.py:37-48, oc_sort/association.py:187-191,
//                                                                    strong_sort/sort/linear_assignment.py:55
//
// All are HBM/latency-bound: rows are staged once per CTA, outputs are written coalesced; batches fill the 148 SMs.
#include "kf_xyah.cuh"
#include "lap.cuh"
#include "lsap_scipy.cuh"
#include "tk_common.cuh"
#include "trackkern.h"

namespace {

using namespace tk;

__device__ __forceinline__ double d_iou(const double* a, const double* b, double& wh) {
    const double w = fmax(0.0, fmin(a[2], b[2]) - fmax(a[0], b[0]));
    const double h = fmax(0.0, fmin(a[3], b[3]) - fmax(a[1], b[1]));
    wh = __dmul_rn(w, h);
    const double ua = __dsub_rn(__dadd_rn(__dmul_rn(a[2] - a[0], a[3] - a[1]), __dmul_rn(b[2] - b[0], b[3] - b[1])), wh);
    return wh / ua;
}

__device__ double d_overlap(int kind, const double* a, const double* b) {
    double wh;
    const double iou = d_iou(a, b, wh);
    if (kind == TK_ASSO_IOU) return iou;
    const double wc = fmax(a[2], b[2]) - fmin(a[0], b[0]);
    const double hc = fmax(a[3], b[3]) - fmin(a[1], b[1]);
    if (kind == TK_ASSO_GIOU) {
        const double hull = __dmul_rn(wc, hc);
        return __dadd_rn(__dsub_rn(iou, __dsub_rn(hull, wh) / hull), 1.0) / 2.0;
    }
    const double dcx = __dsub_rn((a[0] + a[2]) / 2.0, (b[0] + b[2]) / 2.0);
    const double dcy = __dsub_rn((a[1] + a[3]) / 2.0, (b[1] + b[3]) / 2.0);
    const double inner = __dadd_rn(__dmul_rn(dcx, dcx), __dmul_rn(dcy, dcy));
    const double outer = __dadd_rn(__dmul_rn(wc, wc), __dmul_rn(hc, hc));
    if (kind == TK_ASSO_DIOU) return __dadd_rn(__dsub_rn(iou, inner / outer), 1.0) / 2.0;
    const double w1 = a[2] - a[0], h1 = (a[3] - a[1]) + 1.0, w2 = b[2] - b[0], h2 = (b[3] - b[1]) + 1.0;
    const double at = __dsub_rn(atan(w2 / h2), atan(w1 / h1));
    const double pi = 3.141592653589793;
    const double v = __dmul_rn(4.0 / __dmul_rn(pi, pi), __dmul_rn(at, at));
    const double alpha = v / __dadd_rn(__dsub_rn(1.0, iou), v);
    return __dadd_rn(__dsub_rn(__dsub_rn(iou, inner / outer), __dmul_rn(alpha, v)), 1.0) / 2.0;
}

// a [B, N, 4], b [B, M, 4] -> out [B, N, M]; one CTA per (problem, 32-row tile)
__global__ void __launch_bounds__(256)
iou_matrix_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, int N, int M, int kind) {
    extern __shared__ __align__(16) double sm[];
    const int p = blockIdx.y, r0 = blockIdx.x * 32;
    double* sb = sm;             // [M][4]
    double* sa = sm + 4 * M;     // [32][4]
    const double* ap = a + ((size_t)p * N + r0) * 4;
    const double* bp = b + (size_t)p * M * 4;
    for (int i = threadIdx.x; i < 4 * M; i += blockDim.x) sb[i] = bp[i];
    const int nr = min(32, N - r0);
    for (int i = threadIdx.x; i < 4 * nr; i += blockDim.x) sa[i] = ap[i];
    __syncthreads();
    double* op = out + ((size_t)p * N + r0) * M;
    for (int e = threadIdx.x; e < nr * M; e += blockDim.x) {
        const int i = e / M, j = e - i * M;
        op[e] = d_overlap(kind, sa + 4 * i, sb + 4 * j);
    }
}

__global__ void __launch_bounds__(256)
iou_p1_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int N, int M) {
    extern __shared__ __align__(16) float smf[];
    const int p = blockIdx.y, r0 = blockIdx.x * 32;
    float* sb = smf;
    float* sa = smf + 4 * M;
    const float* ap = a + ((size_t)p * N + r0) * 4;
    const float* bp = b + (size_t)p * M * 4;
    for (int i = threadIdx.x; i < 4 * M; i += blockDim.x) sb[i] = bp[i];
    const int nr = min(32, N - r0);
    for (int i = threadIdx.x; i < 4 * nr; i += blockDim.x) sa[i] = ap[i];
    __syncthreads();
    float* op = out + ((size_t)p * N + r0) * M;
    for (int e = threadIdx.x; e < nr * M; e += blockDim.x) {
        const int i = e / M, j = e - i * M;
        const float* x = sa + 4 * i;
        const float* y = sb + 4 * j;
        float ov = 0.0f;
        const float iw = __fadd_rn(__fsub_rn(fminf(x[2], y[2]), fmaxf(x[0], y[0])), 1.0f);
        if (iw > 0.0f) {
            const float ih = __fadd_rn(__fsub_rn(fminf(x[3], y[3]), fmaxf(x[1], y[1])), 1.0f);
            if (ih > 0.0f) {
                const float ab = __fmul_rn(__fadd_rn(__fsub_rn(y[2], y[0]), 1.0f), __fadd_rn(__fsub_rn(y[3], y[1]), 1.0f));
                const float aa = __fmul_rn(__fadd_rn(__fsub_rn(x[2], x[0]), 1.0f), __fadd_rn(__fsub_rn(x[3], x[1]), 1.0f));
                const float in = __fmul_rn(iw, ih);
                ov = __fdiv_rn(in, __fsub_rn(__fadd_rn(aa, ab), in));
            }
        }
        op[e] = __fsub_rn(1.0f, ov);
    }
}

// L2 norms of the rows of x [R, E] (float32), one warp per row, shuffle reduction
__global__ void __launch_bounds__(256)
row_inv_norm_kernel(const float* __restrict__ x, float* __restrict__ inv_norm, long long R, int E) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= R) return;
    const float* p = x + row * E;
    float s = 0.0f;
    for (int k = threadIdx.x & 31; k < E; k += 32) { const float v = p[k]; s = fmaf(v, v, s); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) inv_norm[row] = sqrtf(s);   // the norm itself; the division happens per element like NumPy
}

// out[p, i, j] = 1 - (a_i / |a_i|) . (b_j / |b_j|)  (nn_matching.py:30-49: float32 normalise, float32 dot).
// Register-tiled: a CTA of 256 threads (16 x 16) owns a (16 TM) x (16 TM) output tile, every thread a TM x TM block of accumulators;
// E is walked in chunks of 32 through shared memory (k-major, so the inner loop reads TM + TM values for TM * TM fused
// multiply-adds). The rows are normalised while they are staged (one division per staged element, like NumPy's a / |a|), not in the
// inner loop. TM = 5 (80 x 80 tiles) wastes 12 % on the 150 x 150 problems of the stress configuration, TM = 4 (64 x 64) 39 %.
template <int TM>
__global__ void __launch_bounds__(256)
cosine_dist_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ na,
                   const float* __restrict__ nb, double* __restrict__ out, int N, int M, int E) {
    constexpr int T = 16 * TM, LD = T + 4;
    __shared__ __align__(16) float ta[32][LD], tb[32][LD];
    const int p = blockIdx.z, i0 = blockIdx.y * T, j0 = blockIdx.x * T;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float* ap = a + (size_t)p * N * E;
    const float* bp = b + (size_t)p * M * E;
    const float* nap = na + (size_t)p * N;
    const float* nbp = nb + (size_t)p * M;
    float acc[TM][TM];
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int c = 0; c < TM; ++c) acc[r][c] = 0.0f;
    const bool vec = (E & 3) == 0;
    for (int k0 = 0; k0 < E; k0 += 32) {
        // stage T rows x 32 columns of both operands, transposed to k-major, divided by the row norm
        for (int q = threadIdx.x; q < T * 8; q += 256) {
            const int row = q >> 3, kq = (q & 7) * 4, k = k0 + kq;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            const int ia = i0 + row, jb = j0 + row;
            if (ia < N) {
                const float* src = ap + (size_t)ia * E + k;
                if (vec && k + 3 < E) va = *reinterpret_cast<const float4*>(src);
                else { if (k < E) va.x = src[0]; if (k + 1 < E) va.y = src[1]; if (k + 2 < E) va.z = src[2]; if (k + 3 < E) va.w = src[3]; }
                const float n = nap[ia];
                va.x = __fdiv_rn(va.x, n); va.y = __fdiv_rn(va.y, n); va.z = __fdiv_rn(va.z, n); va.w = __fdiv_rn(va.w, n);
            }
            if (jb < M) {
                const float* src = bp + (size_t)jb * E + k;
                if (vec && k + 3 < E) vb = *reinterpret_cast<const float4*>(src);
                else { if (k < E) vb.x = src[0]; if (k + 1 < E) vb.y = src[1]; if (k + 2 < E) vb.z = src[2]; if (k + 3 < E) vb.w = src[3]; }
                const float n = nbp[jb];
                vb.x = __fdiv_rn(vb.x, n); vb.y = __fdiv_rn(vb.y, n); vb.z = __fdiv_rn(vb.z, n); vb.w = __fdiv_rn(vb.w, n);
            }
            ta[kq][row] = va.x; ta[kq + 1][row] = va.y; ta[kq + 2][row] = va.z; ta[kq + 3][row] = va.w;
            tb[kq][row] = vb.x; tb[kq + 1][row] = vb.y; tb[kq + 2][row] = vb.z; tb[kq + 3][row] = vb.w;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            float av[TM], bv[TM];
#pragma unroll
            for (int r = 0; r < TM; ++r) av[r] = ta[k][ty * TM + r];
#pragma unroll
            for (int c = 0; c < TM; ++c) bv[c] = tb[k][tx * TM + c];
#pragma unroll
            for (int r = 0; r < TM; ++r)
#pragma unroll
                for (int c = 0; c < TM; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < TM; ++r) {
        const int i = i0 + ty * TM + r;
        if (i >= N) continue;
#pragma unroll
        for (int c = 0; c < TM; ++c) {
            const int j = j0 + tx * TM + c;
            if (j < M) out[((size_t)p * N + i) * M + j] = (double)__fsub_rn(1.0f, acc[r][c]);
        }
    }
}

// one problem per CTA: cost [B, N, M] float64 -> x [B, N], y [B, M]
__global__ void __launch_bounds__(128)
lap_batched_kernel(const double* __restrict__ cost, int N, int M, double cost_limit, int has_limit, int* __restrict__ x_out,
                   int* __restrict__ y_out, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int p = blockIdx.x;
    const bool n_rows = N <= M;
    const int nr = n_rows ? N : M, nc = n_rows ? M : N, ld = lap_pitch(nc);
    double* C = (double*)smraw;
    double* u = C + (size_t)nr * ld;
    int* col4row = (int*)(u + nr);
    int* row4col = col4row + nr;
    int* path = row4col + nc;
    __shared__ int ok_flag;
    const double* cp = cost + (size_t)p * N * M;
    for (int e = threadIdx.x; e < N * M; e += blockDim.x) {
        const int i = e / M, j = e - i * M;
        double c = cp[e];
        if (has_limit) c = fmin(c - cost_limit, 0.0);
        if (n_rows) C[(size_t)i * ld + j] = c; else C[(size_t)j * ld + i] = c;
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) x_out[(size_t)p * N + i] = -1;
    for (int j = threadIdx.x; j < M; j += blockDim.x) y_out[(size_t)p * M + j] = -1;
    __syncthreads();
    if (N == 0 || M == 0) return;
    const bool ok = lap_solve_cta(C, ld, nr, nc, has_limit != 0, u, col4row, row4col, path, &ok_flag);
    if (!ok) { if (threadIdx.x == 0) atomicOr(status, TK_DEV_LAP_INFEASIBLE); return; }
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        const int c = col4row[r];
        if (c < 0) continue;
        if (has_limit && !(C[(size_t)r * ld + c] < 0.0)) continue;
        const int i = n_rows ? r : c, j = n_rows ? c : r;
        x_out[(size_t)p * N + i] = j;
        y_out[(size_t)p * M + j] = i;
    }
}

// ct_dist (association.py:150-171): one CTA per problem — centre distances, block maximum, 1 - d / d.max()
__global__ void __launch_bounds__(256)
ct_dist_matrix_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, int N, int M) {
    __shared__ unsigned long long s_max;
    const int p = blockIdx.x;
    const double* ap = a + (size_t)p * N * 4;
    const double* bp = b + (size_t)p * M * 4;
    double* op = out + (size_t)p * N * M;
    if (threadIdx.x == 0) s_max = 0ull;
    __syncthreads();
    unsigned long long lm = 0ull;
    for (int e = threadIdx.x; e < N * M; e += blockDim.x) {
        const int i = e / M, j = e - i * M;
        const double dx = (ap[4 * i] + ap[4 * i + 2]) / 2.0 - (bp[4 * j] + bp[4 * j + 2]) / 2.0;
        const double dy = (ap[4 * i + 1] + ap[4 * i + 3]) / 2.0 - (bp[4 * j + 1] + bp[4 * j + 3]) / 2.0;
        const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        op[e] = d;
        const unsigned long long k = (unsigned long long)__double_as_longlong(d);   // d >= 0: bit order = value order
        lm = k > lm ? k : lm;
    }
    atomicMax(&s_max, lm);
    __syncthreads();
    const double dmax = __longlong_as_double((long long)s_max);
    for (int e = threadIdx.x; e < N * M; e += blockDim.x) op[e] = 1.0 - op[e] / dmax;
}

// Part-based appearance distance (BPBReID / KPR embeddings): a [B,N,K,E], va [B,N,K], b [B,M,K,E], vb [B,M,K] -> out [B,N,M] float32
//   sum_k w_k * || a_k/|a_k| - b_k/|b_k| || / sum_k w_k / 2,  w_k = va_k * vb_k   (nn_matching.py:99-135 on the restated torchreid
// function, see oracle/bpbreid_np.py). norms: scratch float32 [B*(N+M)*K] of the F.normalize denominators; one warp per pair.
__global__ void __launch_bounds__(256)
part_norm_kernel(const float* __restrict__ x, float* __restrict__ nrm, long long rows, int E) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= rows) return;
    const float* p = x + r * E;
    float s = 0.0f;
    for (int k = threadIdx.x & 31; k < E; k += 32) s = fmaf(p[k], p[k], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) nrm[r] = fmaxf(sqrtf(s), 1e-12f);
}

__global__ void __launch_bounds__(256)
part_dist_kernel(const float* __restrict__ a, const float* __restrict__ va, const float* __restrict__ b, const float* __restrict__ vb,
                 const float* __restrict__ na, const float* __restrict__ nb, float* __restrict__ out, int N, int M, int K, int E) {
    const int p = blockIdx.y, lane = threadIdx.x & 31;
    const long long pair = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (pair >= (long long)N * M) return;
    const int i = (int)(pair / M), j = (int)(pair - (long long)i * M);
    const size_t ra = (size_t)p * N + i, rb = (size_t)p * M + j;
    float num = 0.0f, den = 0.0f;
    for (int k = 0; k < K; ++k) {
        const float* ap = a + (ra * K + k) * E;
        const float* bp = b + (rb * K + k) * E;
        const float an = na[ra * K + k], bn = nb[rb * K + k];
        float acc = 0.0f;
        for (int e = lane; e < E; e += 32) { const float d = __fdiv_rn(ap[e], an) - __fdiv_rn(bp[e], bn); acc = fmaf(d, d, acc); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const float w = __fmul_rn(va[ra * K + k], vb[rb * K + k]);
        num = __fadd_rn(num, __fmul_rn(sqrtf(acc), w));
        den = __fadd_rn(den, w);
    }
    if (lane == 0) out[((size_t)p * N + i) * M + j] = __fdiv_rn(__fdiv_rn(num, den), 2.0f);
}

// Squared Mahalanobis gating distances of the xyah Kalman filters (kalman_filter.py gating_distance of the ByteTrack /
// StrongSORT / BPBReID plugins): mean [T,8], cov [T,8,8], z [D,4] (x, y, a, h) -> out [T,D].
// aspect_const = 1: R = diag((h/20)^2, (h/20)^2, 1e-2, (h/20)^2) (byte_track/kalman_filter.py:126-153,
// strong_sort/sort/kalman_filter.py:114-137 at confidence 0); 0: every term (h/20)^2 (bpbreid_strong_sort/sort/kalman_filter.py:106-136).
__global__ void __launch_bounds__(128)
kf_gate_kernel(const double* __restrict__ mean, const double* __restrict__ cov, const double* __restrict__ z, double* __restrict__ out,
               int T, int D, int aspect_const, int* __restrict__ status) {
    const int t = blockIdx.x;
    __shared__ double c[24];
    if (threadIdx.x == 0) {
        const double* m = mean + (size_t)t * 8;
        const double sp = (1.0 / 20) * m[3];
        const double rr[4] = {sp * sp, sp * sp, aspect_const ? 1e-1 * 1e-1 : sp * sp, sp * sp};
        double L[16], Sm[16], invd[4];
        if (!tk::kf8_chol4(cov + (size_t)t * 64, rr, L, Sm, invd)) atomicOr(status, TK_DEV_BAD_CHOLESKY);
        for (int i = 0; i < 4; ++i) c[i] = m[i];
        for (int i = 0; i < 16; ++i) c[4 + i] = L[i];
        for (int i = 0; i < 4; ++i) c[20 + i] = invd[i];
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) out[(size_t)t * D + d] = tk::kf8_maha(c, c + 4, c + 20, z + (size_t)d * 4);
}

// Stateless Kalman steps (tk_kf_predict / tk_kf_update): one thread per track, float64, the arithmetic of kf_xyah.cuh.
// model 0 = ByteTrack xyah (byte_track/kalman_filter.py:88-124,194-226: std from h, aspect std 1e-2 / 1e-5, R aspect 1e-1),
// model 1 = BoT-SORT xywh (bot_sort/kalman_filter.py:88-124,194-226: std from (w, h)).
__device__ __forceinline__ void kf_noise(int model, const double* m, double* q, double* r) {
    const double wp = 1.0 / 20, wv = 1.0 / 160;
    if (model == 0) {
        const double h = m[3];
        const double sp = wp * h, sv = wv * h;
        const double qq[8] = {sp * sp, sp * sp, 1e-2 * 1e-2, sp * sp, sv * sv, sv * sv, 1e-5 * 1e-5, sv * sv};
        for (int i = 0; i < 8; ++i) q[i] = qq[i];
        r[0] = sp * sp; r[1] = sp * sp; r[2] = 1e-1 * 1e-1; r[3] = sp * sp;
    } else {
        const double sw = wp * m[2], sh = wp * m[3], vw = wv * m[2], vh = wv * m[3];
        const double qq[8] = {sw * sw, sh * sh, sw * sw, sh * sh, vw * vw, vh * vh, vw * vw, vh * vh};
        for (int i = 0; i < 8; ++i) q[i] = qq[i];
        r[0] = sw * sw; r[1] = sh * sh; r[2] = sw * sw; r[3] = sh * sh;
    }
}

__global__ void __launch_bounds__(128)
kf_predict_kernel(double* __restrict__ mean, double* __restrict__ cov, int n, int model) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double m[8], P[64], q[8], r[4];
    for (int i = 0; i < 8; ++i) m[i] = mean[(size_t)t * 8 + i];
    for (int i = 0; i < 64; ++i) P[i] = cov[(size_t)t * 64 + i];
    kf_noise(model, m, q, r);
    tk::kf8_predict(m, P, q);
    for (int i = 0; i < 8; ++i) mean[(size_t)t * 8 + i] = m[i];
    for (int i = 0; i < 64; ++i) cov[(size_t)t * 64 + i] = P[i];
}

__global__ void __launch_bounds__(128)
kf_update_kernel(double* __restrict__ mean, double* __restrict__ cov, const double* __restrict__ z, int n, int model, int* __restrict__ status) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double m[8], P[64], q[8], r[4];
    for (int i = 0; i < 8; ++i) m[i] = mean[(size_t)t * 8 + i];
    for (int i = 0; i < 64; ++i) P[i] = cov[(size_t)t * 64 + i];
    kf_noise(model, m, q, r);
    if (!tk::kf8_update(m, P, z + (size_t)t * 4, r)) atomicOr(status, TK_DEV_BAD_CHOLESKY);
    for (int i = 0; i < 8; ++i) mean[(size_t)t * 8 + i] = m[i];
    for (int i = 0; i < 64; ++i) cov[(size_t)t * 64 + i] = P[i];
}

// OC-SORT velocity-direction-consistency cost (oc_sort/association.py:175-184,246-266): out[d, t] = valid_t * (pi/2 - |acos(clip(
// v_t . dir(prev_t -> det_d)))|) / pi * inertia * weight_d, weight = the column the reference multiplies with (the class column, q1).
__global__ void __launch_bounds__(256)
vdc_cost_kernel(const double* __restrict__ dets, const double* __restrict__ prev, const double* __restrict__ vel, double* __restrict__ out,
                int D, int T, double inertia, int weight_col) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D * T) return;
    const int d = e / T, t = e - d * T;
    const double* db = dets + (size_t)d * 6;
    const double* pv = prev + (size_t)t * 5;
    const double cx1 = (db[0] + db[2]) / 2.0, cy1 = (db[1] + db[3]) / 2.0;
    const double cx2 = (pv[0] + pv[2]) / 2.0, cy2 = (pv[1] + pv[3]) / 2.0;
    const double dx = cx1 - cx2, dy = cy1 - cy2;
    const double norm = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) + 1e-6;
    const double X = dx / norm, Y = dy / norm;
    double c = __dadd_rn(__dmul_rn(vel[(size_t)t * 2 + 1], X), __dmul_rn(vel[(size_t)t * 2], Y));
    c = fmin(fmax(c, -1.0), 1.0);
    const double pi = 3.141592653589793;
    const double ang = (pi / 2.0 - fabs(acos(c))) / pi;
    const double valid = pv[4] < 0 ? 0.0 : 1.0;
    out[e] = __dmul_rn(__dmul_rn(__dmul_rn(valid, ang), inertia), db[weight_col]);
}

// scipy.optimize.linear_sum_assignment, one warp per problem, cost read from global memory (stateless form of lsap_scipy.cuh)
__global__ void __launch_bounds__(32)
lsap_scipy_batched_kernel(const double* __restrict__ cost, int N, int M, int* __restrict__ x_out, int* __restrict__ y_out,
                          int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int p = blockIdx.x;
    const double* c = cost + (size_t)p * N * M;
    const bool transpose = M < N;                       // tall matrix: scipy solves the transposed problem
    const int nr = transpose ? M : N, nc = transpose ? N : M;
    tk::LsapScratch S;
    S.carve(smem_raw, nr, nc);
    bool ok;
    if (transpose) ok = tk::lsap_scipy_warp(nr, nc, [&](int i, int j) { return c[(size_t)j * M + i]; }, S.u, S.v, S.spc, S.path,
                                            S.col4row, S.row4col, S.remaining, S.SR, S.SC);
    else ok = tk::lsap_scipy_warp(nr, nc, [&](int i, int j) { return c[(size_t)i * M + j]; }, S.u, S.v, S.spc, S.path, S.col4row,
                                  S.row4col, S.remaining, S.SR, S.SC);
    __syncwarp();
    int* x = x_out + (size_t)p * N;
    int* y = y_out + (size_t)p * M;
    if (!ok) {
        if (threadIdx.x == 0) atomicOr(status, TK_DEV_LAP_INFEASIBLE);
        for (int i = threadIdx.x; i < N; i += 32) x[i] = -1;
        for (int j = threadIdx.x; j < M; j += 32) y[j] = -1;
        return;
    }
    if (transpose) {   // solver rows are the columns of the caller
        for (int j = threadIdx.x; j < M; j += 32) y[j] = S.col4row[j];
        for (int i = threadIdx.x; i < N; i += 32) x[i] = S.row4col[i];
    } else {
        for (int i = threadIdx.x; i < N; i += 32) x[i] = S.col4row[i];
        for (int j = threadIdx.x; j < M; j += 32) y[j] = S.row4col[j];
    }
}

}  // namespace

extern "C" {

int tk_lsap_scipy_batched(const double* cost, int n_problems, int N, int M, int* x_out, int* y_out, int* status_dev, void* stream) {
    if (!cost || !x_out || !y_out || !status_dev || n_problems <= 0 || N <= 0 || M <= 0) return TK_ERR_ARG;
    const int nr = N <= M ? N : M, nc = N <= M ? M : N;
    const size_t smem = tk::lsap_scipy_scratch_bytes(nr, nc);
    if (smem > 220 * 1024) return TK_ERR_CAPACITY;
    TK_CUDA_TRY(cudaFuncSetAttribute(lsap_scipy_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lsap_scipy_batched_kernel<<<n_problems, 32, smem, (cudaStream_t)stream>>>(cost, N, M, x_out, y_out, status_dev);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_iou_matrix(const double* a, const double* b, double* out, int n_problems, int N, int M, int variant, void* stream) {
    if (!a || !b || !out || n_problems <= 0 || N < 0 || M < 0 || variant < 0 || variant > TK_ASSO_CT_DIST) return TK_ERR_ARG;
    if (N == 0 || M == 0) return TK_OK;
    if (variant == TK_ASSO_CT_DIST) {
        ct_dist_matrix_kernel<<<n_problems, 256, 0, (cudaStream_t)stream>>>(a, b, out, N, M);
        TK_CUDA_TRY(cudaGetLastError());
        return TK_OK;
    }
    const size_t smem = sizeof(double) * 4 * ((size_t)M + 32);
    if (smem > 200 * 1024) return TK_ERR_CAPACITY;
    TK_CUDA_TRY(cudaFuncSetAttribute(iou_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((N + 31) / 32, n_problems);
    iou_matrix_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(a, b, out, N, M, variant);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_iou_p1_f32(const float* a_tlbr, const float* b_tlbr, float* dist_out, int n_problems, int N, int M, void* stream) {
    if (!a_tlbr || !b_tlbr || !dist_out || n_problems <= 0 || N < 0 || M < 0) return TK_ERR_ARG;
    if (N == 0 || M == 0) return TK_OK;
    const size_t smem = sizeof(float) * 4 * ((size_t)M + 32);
    if (smem > 200 * 1024) return TK_ERR_CAPACITY;
    TK_CUDA_TRY(cudaFuncSetAttribute(iou_p1_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((N + 31) / 32, n_problems);
    iou_p1_f32_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(a_tlbr, b_tlbr, dist_out, N, M);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_cosine_dist(const float* a, const float* b, double* out, float* norm_scratch, int n_problems, int N, int M, int E,
                   void* stream) {
    if (!a || !b || !out || !norm_scratch || n_problems <= 0 || N < 0 || M < 0 || E <= 0) return TK_ERR_ARG;
    if (N == 0 || M == 0) return TK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    float* na = norm_scratch;
    float* nb = norm_scratch + (size_t)n_problems * N;
    const long long ra = (long long)n_problems * N, rb = (long long)n_problems * M;
    row_inv_norm_kernel<<<(unsigned)((ra + 7) / 8), 256, 0, st>>>(a, na, ra, E);
    row_inv_norm_kernel<<<(unsigned)((rb + 7) / 8), 256, 0, st>>>(b, nb, rb, E);
    // tile size by covered-area waste: 80 x 80 (TM = 5) or 64 x 64 (TM = 4)
    auto waste = [&](int t) { return (double)(((N + t - 1) / t) * t) * (((M + t - 1) / t) * t) / ((double)N * M); };
    if (waste(80) <= waste(64)) {
        dim3 grid((M + 79) / 80, (N + 79) / 80, n_problems);
        cosine_dist_kernel<5><<<grid, 256, 0, st>>>(a, b, na, nb, out, N, M, E);
    } else {
        dim3 grid((M + 63) / 64, (N + 63) / 64, n_problems);
        cosine_dist_kernel<4><<<grid, 256, 0, st>>>(a, b, na, nb, out, N, M, E);
    }
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_lap_batched(const double* cost, int n_problems, int N, int M, double cost_limit, int has_limit, int* x_out, int* y_out,
                   int* status_dev, void* stream) {
    if (!cost || !x_out || !y_out || !status_dev || n_problems <= 0 || N < 0 || M < 0) return TK_ERR_ARG;
    if (N > tk::LAP_MAX_COLS || M > tk::LAP_MAX_COLS) return TK_ERR_CAPACITY;
    const int nr = N <= M ? N : M, nc = N <= M ? M : N;
    const size_t smem = sizeof(double) * ((size_t)nr * tk::lap_pitch(nc) + nr) + sizeof(int) * ((size_t)nr + 2 * nc) + 64;
    if (smem > 220 * 1024) return TK_ERR_CAPACITY;
    TK_CUDA_TRY(cudaFuncSetAttribute(lap_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lap_batched_kernel<<<n_problems, 128, smem, (cudaStream_t)stream>>>(cost, N, M, cost_limit, has_limit, x_out, y_out, status_dev);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_part_dist(const float* a, const float* va, const float* b, const float* vb, float* out, float* norm_scratch, int n_problems, int N,
                 int M, int K, int E, void* stream) {
    if (!a || !va || !b || !vb || !out || !norm_scratch || n_problems <= 0 || N < 0 || M < 0 || K <= 0 || E <= 0) return TK_ERR_ARG;
    if (N == 0 || M == 0) return TK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    float* na = norm_scratch;
    float* nb = norm_scratch + (size_t)n_problems * N * K;
    const long long rows_a = (long long)n_problems * N * K, rows_b = (long long)n_problems * M * K;
    part_norm_kernel<<<(unsigned)((rows_a + 7) / 8), 256, 0, st>>>(a, na, rows_a, E);
    part_norm_kernel<<<(unsigned)((rows_b + 7) / 8), 256, 0, st>>>(b, nb, rows_b, E);
    dim3 grid((unsigned)(((long long)N * M + 7) / 8), n_problems);
    part_dist_kernel<<<grid, 256, 0, st>>>(a, va, b, vb, na, nb, out, N, M, K, E);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_kf_gate(const double* mean, const double* cov, const double* z, double* out, int n_tracks, int n_dets, int aspect_const,
               int* status_dev, void* stream) {
    if (!mean || !cov || !z || !out || !status_dev || n_tracks < 0 || n_dets < 0) return TK_ERR_ARG;
    if (n_tracks == 0 || n_dets == 0) return TK_OK;
    kf_gate_kernel<<<n_tracks, 128, 0, (cudaStream_t)stream>>>(mean, cov, z, out, n_tracks, n_dets, aspect_const, status_dev);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_kf_predict(double* mean, double* cov, int n_tracks, int model, void* stream) {
    if (!mean || !cov || n_tracks < 0 || model < 0 || model > 1) return TK_ERR_ARG;
    if (n_tracks == 0) return TK_OK;
    kf_predict_kernel<<<(n_tracks + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mean, cov, n_tracks, model);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_kf_update(double* mean, double* cov, const double* z, int n_tracks, int model, int* status_dev, void* stream) {
    if (!mean || !cov || !z || !status_dev || n_tracks < 0 || model < 0 || model > 1) return TK_ERR_ARG;
    if (n_tracks == 0) return TK_OK;
    kf_update_kernel<<<(n_tracks + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mean, cov, z, n_tracks, model, status_dev);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_vdc_cost(const double* dets6, const double* prev_obs5, const double* velocities2, double* out, int n_dets, int n_tracks, double inertia,
                int weight_col, void* stream) {
    if (!dets6 || !prev_obs5 || !velocities2 || !out || n_dets < 0 || n_tracks < 0 || weight_col < 0 || weight_col > 5) return TK_ERR_ARG;
    if (n_dets == 0 || n_tracks == 0) return TK_OK;
    const int n = n_dets * n_tracks;
    vdc_cost_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dets6, prev_obs5, velocities2, out, n_dets, n_tracks, inertia, weight_col);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

}  // extern "C"
