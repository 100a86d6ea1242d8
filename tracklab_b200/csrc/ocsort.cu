// OC-SORT association, whole video per launch, one CTA per video.
//
// Device restatement of
//   /root/reference/plugins/track/oc_sort/ocsort.py:10-54,57-169,203-334      (helpers, KalmanBoxTracker, OCSort.update)
//   /root/reference/plugins/track/oc_sort/association.py:5-171,175-195,242-298 (IoU family, VDC cost, associate)
//   /root/reference/plugins/track/oc_sort/kalmanfilter.py:339-379,383-434,437-526 (predict, ORU freeze/unfreeze, update)
// and of the wrapper filter /root/reference/tracklab/wrappers/track/oc_sort_api.py:50-56.
//
// Same execution shape as bytetrack.cu (one launch walks the frames of a video, one CTA per video,
// assignment on one warp). The reference's deepcopy-based observation-centric re-update is kept as
// "frozen (x, P) + history length + last stored observation", which is all its replay reads.
// Quirks kept: VDC term multiplied by the class column, first round always plain IoU, the
// partial-permutation shortcut that skips the solver, unmatched lists left unsorted unless the OCR/BYTE
// round ran (np.setdiff1d sorts) — that order decides the ids of new tracks.
#include "lap.cuh"
#include "oc_boxes.cuh"
#include "tk_common.cuh"
#include "trackkern.h"

namespace {

using namespace tk;

// Optional per-phase cycle accounting (build with -DTK_PHASE_PROF), read with tk_debug_ocsort_phases().
#ifdef TK_PHASE_PROF
__device__ unsigned long long g_oc_prof[64];
#define PH(k) do { if (threadIdx.x == 0) { const long long _t = clock64(); g_oc_prof[k] += (unsigned long long)(_t - ph_t0); ph_t0 = _t; } } while (0)
#else
#define PH(k) do { } while (0)
#endif

constexpr int OC_THREADS = 128;
constexpr int RING = 8;  // observations kept per track: ages age-1 .. age-RING (delta_t <= RING)

struct OcParams {
    double det_thresh, iou_threshold, inertia, min_conf;
    int max_age, min_hits, delta_t, asso, use_byte;
};

struct OcDev {
    int* hdr;  // 0 frame_count, 1 next uid, 2 n_trk, 3 -, 4 status, 5 n_free
    double *x, *P, *fx, *fP, *last_obs, *vel, *last_z, *ring_obs, *conf, *cls, *det_id;
    int *ring_age, *tsu, *uid, *hits, *streak, *age, *hist_len, *frozen_n, *last_z_idx, *list, *free_list;
    unsigned char *has_vel, *observed, *frozen;
};

__host__ __device__ inline size_t oc_al(size_t x) { return (x + 15) & ~(size_t)15; }

#define OC_FIELDS(X)                                                                                   \
    X(x, double, 7) X(P, double, 49) X(fx, double, 7) X(fP, double, 49) X(last_obs, double, 5) X(vel, double, 2) \
    X(last_z, double, 4) X(ring_obs, double, RING * 5) X(conf, double, 1) X(cls, double, 1) X(det_id, double, 1) \
    X(ring_age, int, RING) X(tsu, int, 1) X(uid, int, 1) X(hits, int, 1) X(streak, int, 1) X(age, int, 1)        \
    X(hist_len, int, 1) X(frozen_n, int, 1) X(last_z_idx, int, 1) X(list, int, 1) X(free_list, int, 1)          \
    X(has_vel, unsigned char, 1) X(observed, unsigned char, 1) X(frozen, unsigned char, 1)

__host__ __device__ inline size_t oc_state_bytes(int cap) {
    size_t s = oc_al(8 * sizeof(int));
#define X(name, type, n) s += oc_al((size_t)cap * (n) * sizeof(type));
    OC_FIELDS(X)
#undef X
    return s;
}

__host__ __device__ inline OcDev oc_carve(char* base, int cap) {
    OcDev d;
    char* p = base;
    d.hdr = (int*)p; p += oc_al(8 * sizeof(int));
#define X(name, type, n) d.name = (type*)p; p += oc_al((size_t)cap * (n) * sizeof(type));
    OC_FIELDS(X)
#undef X
    return d;
}

// ---- 7-d SORT Kalman filter (ocsort.py:75-84, kalmanfilter.py:339-379,488-526) ---------------------------
__device__ void kf7_predict(double* x, double* P) {
    // x = F x ; P = F P F^T + Q  (F has 0/1 entries: sums of two terms, single rounding each)
    x[0] = x[0] + x[4]; x[1] = x[1] + x[5]; x[2] = x[2] + x[6];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 7; ++j) P[i * 7 + j] = P[i * 7 + j] + P[(i + 4) * 7 + j];
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 3; ++j) P[i * 7 + j] = P[i * 7 + j] + P[i * 7 + j + 4];
    const double q[7] = {1.0, 1.0, 1.0, 1.0, 0.01, 0.01, 0.01 * 0.01};   // ocsort.py:83-84
    for (int i = 0; i < 7; ++i) P[i * 8] = P[i * 8] + q[i];
}

// ---- group-cooperative measurement update: 8 consecutive lanes own one track, lane j < 7 holds row j of P and x[j] ----
// KalmanFilterNew.update (kalmanfilter.py:488-526): S = H P H^T + R, K = P H^T S^-1 (S^-1 through its Cholesky factor),
// x += K y, P = (I-KH) P (I-KH)^T + K R K^T. A single thread doing the two 7x7x7 products of the Joseph form keeps ~150
// doubles live and spills to local memory (~55 us per frame incl. the ORU replays); row-per-lane keeps everything in registers.
__device__ __forceinline__ double grp_bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src, 8); }

// row[7] / xj: this lane's row of P and element of x (identity rows on idle lanes); the result is committed when `on`.
__device__ __forceinline__ bool kf7_group_correct(double (&row_io)[7], double& xj_io, bool on, const double* z) {
    const int j = threadIdx.x & 7;
    double row[7], xj = xj_io;
#pragma unroll
    for (int c = 0; c < 7; ++c) row[c] = row_io[c];
    const double R[4] = {1.0, 1.0, 10.0, 10.0};
    double S[16], L[16], Li[16], SI[16];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { const double v = grp_bcast(row[b], a); S[a * 4 + b] = (a == b) ? v + R[a] : v; }
    // S^-1 through the Cholesky factor with reciprocal pivots: one rsqrt per column is the only long-latency operation on the
    // dependency chain (the reference inverts S with LAPACK's LU, so no operation order can be bit-identical anyway; the
    // result differs from it by a few ulp either way, tests compare boxes at 1e-6 px).
    bool ok = true;
    double invd[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double d = S[c * 4 + c];
#pragma unroll
        for (int k = 0; k < c; ++k) d -= L[c * 4 + k] * L[c * 4 + k];
        ok = ok && d > 0.0;
        invd[c] = rsqrt(d);
        L[c * 4 + c] = d * invd[c];
#pragma unroll
        for (int i = c + 1; i < 4; ++i) {
            double t = S[i * 4 + c];
#pragma unroll
            for (int k = 0; k < c; ++k) t -= L[i * 4 + k] * L[c * 4 + k];
            L[i * 4 + c] = t * invd[c];
        }
#pragma unroll
        for (int i = 0; i < c; ++i) L[i * 4 + c] = 0.0;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {  // Li = L^-1 (lower)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double t = (i == c) ? 1.0 : 0.0;
#pragma unroll
            for (int k = c; k < i; ++k) t -= L[i * 4 + k] * Li[k * 4 + c];
            Li[i * 4 + c] = (i < c) ? 0.0 : t * invd[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)   // SI = Li^T Li
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            double t = 0.0;
#pragma unroll
            for (int k = (i > b ? i : b); k < 4; ++k) t += Li[k * 4 + i] * Li[k * 4 + b];
            SI[i * 4 + b] = t;
        }
    double K[4], y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = z[i] - grp_bcast(xj, i);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) t += row[k] * SI[k * 4 + b];
        K[b] = t;
    }
    {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) t += K[k] * y[k];
        xj = xj + t;
    }
    // P = (I-KH) P (I-KH)^T + K R K^T   (Joseph form, kalmanfilter.py:520-521), row j on lane j
    double A[7], AP[7];
#pragma unroll
    for (int b = 0; b < 7; ++b) A[b] = ((j == b) ? 1.0 : 0.0) - (b < 4 ? K[b] : 0.0);
#pragma unroll
    for (int b = 0; b < 7; ++b) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 7; ++k) t += A[k] * grp_bcast(row[b], k);
        AP[b] = t;
    }
#pragma unroll
    for (int b = 0; b < 7; ++b) {   // column b needs row b of A and of K, i.e. K of lane b
        double Kb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) Kb[k] = grp_bcast(K[k], b);
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 7; ++k) t += AP[k] * (((b == k) ? 1.0 : 0.0) - (k < 4 ? Kb[k] : 0.0));
        double r = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) r += (K[k] * R[k]) * Kb[k];
        row[b] = t + r;
    }
    if (on) {
        xj_io = xj;
#pragma unroll
        for (int c = 0; c < 7; ++c) row_io[c] = row[c];
    }
    return ok || !on;
}

// KalmanFilterNew.predict on the group registers (kf7_predict): rows 0..2 += rows 4..6, columns 0..2 += columns 4..6, + Q
__device__ __forceinline__ void kf7_group_predict(double (&row)[7], double& xj, bool on) {
    const int j = threadIdx.x & 7;
    const double xh = grp_bcast(xj, (j + 4) & 7);
    double hi[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) hi[c] = grp_bcast(row[c], (j + 4) & 7);
    if (!on) return;
    if (j < 3) {
        xj = xj + xh;
#pragma unroll
        for (int c = 0; c < 7; ++c) row[c] = row[c] + hi[c];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) row[c] = row[c] + row[c + 4];
    const double q[7] = {1.0, 1.0, 1.0, 1.0, 0.01, 0.01, 0.01 * 0.01};
#pragma unroll
    for (int c = 0; c < 7; ++c) if (c == j) row[c] = row[c] + q[c];
}

__device__ __forceinline__ void box_to_z(const double* b, double* z) {  // ocsort.py:21-33
    const double w = b[2] - b[0], h = b[3] - b[1];
    z[0] = b[0] + w / 2.0; z[1] = b[1] + h / 2.0; z[2] = __dmul_rn(w, h); z[3] = w / (h + 1e-6);
}

__device__ __forceinline__ void x_to_box(const double* x, double* b) {  // ocsort.py:36-46
    const double w = sqrt(__dmul_rn(x[2], x[3]));
    const double h = x[2] / w;
    b[0] = x[0] - w / 2.0; b[1] = x[1] - h / 2.0; b[2] = x[0] + w / 2.0; b[3] = x[1] + h / 2.0;
}

// KalmanBoxTracker.update(bbox) (ocsort.py:103-148): book-keeping; KalmanFilterNew.update incl. the ORU replay when the track
// was frozen. Returns the Kalman work left for oc_apply_updates: 1 = plain correction with z_out, 2 = ORU replay (parameters in
// oru_par / oru_gap) followed by the correction.
// i-th virtual measurement of the ORU line (kalmanfilter.py:412-421), par = {x1, y1, w1, h1, dx, dy, dw, dh}
__device__ __forceinline__ void oru_virtual_z(const double* par, int i, double* vz) {
    const double t = (double)(i + 1);
    const double ww = __dadd_rn(par[2], __dmul_rn(t, par[6])), hh = __dadd_rn(par[3], __dmul_rn(t, par[7]));
    vz[0] = __dadd_rn(par[0], __dmul_rn(t, par[4])); vz[1] = __dadd_rn(par[1], __dmul_rn(t, par[5]));
    vz[2] = __dmul_rn(ww, hh); vz[3] = ww / hh;
}

__device__ int oc_track_update(OcDev& S, int s, const double* bbox5, double cls, double det_id, int delta_t, double* z_out, double* oru_par,
                               int* oru_gap) {
    double* lo = S.last_obs + (size_t)s * 5;
    const int age = S.age[s];
    S.conf[s] = bbox5[4];
    S.cls[s] = cls;
    if (lo[0] + lo[1] + lo[2] + lo[3] + lo[4] >= 0) {   // has a previous observation
        const double* prev = nullptr;
        for (int i = 0; i < delta_t; ++i) {
            const int want = age - (delta_t - i);
            const int r = ((want % RING) + RING) % RING;
            if (want >= 0 && S.ring_age[(size_t)s * RING + r] == want) { prev = S.ring_obs + ((size_t)s * RING + r) * 5; break; }
        }
        if (!prev) prev = lo;
        // speed_direction (ocsort.py:49-54)
        const double cx1 = (prev[0] + prev[2]) / 2.0, cy1 = (prev[1] + prev[3]) / 2.0;
        const double cx2 = (bbox5[0] + bbox5[2]) / 2.0, cy2 = (bbox5[1] + bbox5[3]) / 2.0;
        const double dy = cy2 - cy1, dx = cx2 - cx1;
        const double norm = sqrt(__dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx))) + 1e-6;
        S.vel[(size_t)s * 2 + 0] = dy / norm;
        S.vel[(size_t)s * 2 + 1] = dx / norm;
        S.has_vel[s] = 1;
    }
    for (int k = 0; k < 5; ++k) lo[k] = bbox5[k];
    {
        const int r = age % RING;
        S.ring_age[(size_t)s * RING + r] = age;
        for (int k = 0; k < 5; ++k) S.ring_obs[((size_t)s * RING + r) * 5 + k] = bbox5[k];
    }
    S.tsu[s] = 0;
    S.hits[s] += 1;
    S.streak[s] += 1;
    S.det_id[s] = det_id;
    // ---- KalmanFilterNew.update(z) (kalmanfilter.py:437-526)
    double z[4];
    box_to_z(bbox5, z);
    if (z_out) { for (int k = 0; k < 4; ++k) z_out[k] = z[k]; }
    int hist = S.hist_len[s] + 1;   // history_obs.append(z)
    if (!S.observed[s] && S.frozen[s]) {
        // unfreeze (kalmanfilter.py:390-434): the filter is restored to the frozen state and replayed over a straight line of
        // virtual boxes between the last stored observation and z (ORU); the replay itself runs in oc_apply_updates.
        const int i1 = S.last_z_idx[s], i2 = hist - 1;
        const double* b1 = S.last_z + (size_t)s * 4;
        const double x1 = b1[0], y1 = b1[1], s1 = b1[2], r1 = b1[3];
        const double w1 = sqrt(__dmul_rn(s1, r1)), h1 = sqrt(s1 / r1);
        const double w2 = sqrt(__dmul_rn(z[2], z[3])), h2 = sqrt(z[2] / z[3]);
        const int gap = i2 - i1;
        const double dg = (double)gap;
        const double dx = (z[0] - x1) / dg, dy = (z[1] - y1) / dg, dw = (w2 - w1) / dg, dh = (h2 - h1) / dg;
        oru_par[0] = x1; oru_par[1] = y1; oru_par[2] = w1; oru_par[3] = h1; oru_par[4] = dx; oru_par[5] = dy; oru_par[6] = dw; oru_par[7] = dh;
        *oru_gap = gap;
        double vz[4];
        oru_virtual_z(oru_par, gap - 1, vz);
        hist = S.frozen_n[s] - 1 + gap;
        S.frozen[s] = 0;
        for (int k = 0; k < 4; ++k) S.last_z[(size_t)s * 4 + k] = vz[k];
        S.last_z_idx[s] = hist - 1;
        S.observed[s] = 1;
        S.hist_len[s] = hist;
        return 2;
    }
    for (int k = 0; k < 4; ++k) S.last_z[(size_t)s * 4 + k] = z[k];
    S.last_z_idx[s] = hist - 1;
    S.observed[s] = 1;
    S.hist_len[s] = hist;
    return 1;   // plain correction
}

// Apply the matched (track, detection) pairs of one association round. Pass A, one thread per pair: book-keeping; pass B,
// 8 lanes per pair: the Kalman work (ORU replay where the track was frozen, then the correction with the real measurement).
// slot < 0 skips entry i. Shared scratch: zbuf 4 doubles, oru 8 doubles, gap 1 int, need 1 byte, slot_of 1 int per entry.
template <class Get>
__device__ void oc_apply_updates(OcDev& S, int n, Get get, int delta_t, int* status, double* zbuf, double* oru, int* gap_of,
                                 unsigned char* need, int* slot_of) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int slot; const double* db;
        get(i, slot, db);
        slot_of[i] = slot;
        gap_of[i] = 0;
        need[i] = slot >= 0 ? (unsigned char)oc_track_update(S, slot, db, db[5], db[6], delta_t, zbuf + 4 * i, oru + 8 * i, gap_of + i) : 0;
    }
    __syncthreads();
    // order the entries: replays first (their latency is gap x (correct + predict); packing them into one pass keeps the
    // plain corrections from waiting behind them twice), then the plain corrections
    int* order = slot_of + n;   // scratch behind slot_of (2 * n <= its size, see the caller)
    if (warp_id() == 0) {
        int m = warp_compact(n, 0, [&](int i) { return need[i] == 2; }, [&](int i, int p) { order[p] = i; });
        m = warp_compact(n, m, [&](int i) { return need[i] == 1; }, [&](int i, int p) { order[p] = i; });
        if (lane_id() == 0) gap_of[n] = m;
    }
    __syncthreads();
    const int m = gap_of[n];
    for (int base = 0; base < m; base += blockDim.x / 8) {
        const int e = base + (threadIdx.x >> 3), j = threadIdx.x & 7;
        const int i = e < m ? order[e] : 0;
        const int code = e < m ? need[i] : 0;
        const bool act = code != 0, replay = code == 2;
        const int slot = act ? slot_of[i] : 0;
        const bool own = act && j < 7;
        double* gx = S.x + (size_t)slot * 7;
        double* gP = S.P + (size_t)slot * 49;
        const double* sx = replay ? S.fx + (size_t)slot * 7 : gx;        // ORU restarts from the frozen state
        const double* sP = replay ? S.fP + (size_t)slot * 49 : gP;
        double row[7], xj = own ? sx[j] : 0.0;
#pragma unroll
        for (int c = 0; c < 7; ++c) row[c] = own ? sP[j * 7 + c] : (c == j ? 1.0 : 0.0);
        bool ok = true;
        const int gap = replay ? gap_of[i] : 0;
        int gmax = gap;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gmax = max(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        for (int it = 0; it < gmax; ++it) {
            const bool on = it < gap;
            double vz[4] = {0.0, 0.0, 1.0, 1.0};
            if (on) oru_virtual_z(oru + 8 * i, it, vz);
            ok = kf7_group_correct(row, xj, on, vz) && ok;
            kf7_group_predict(row, xj, on && it != gap - 1);
        }
        double z[4] = {0.0, 0.0, 0.0, 0.0};
        if (act) { z[0] = zbuf[4 * i]; z[1] = zbuf[4 * i + 1]; z[2] = zbuf[4 * i + 2]; z[3] = zbuf[4 * i + 3]; }
        ok = kf7_group_correct(row, xj, act, z) && ok;
        if (own) {
            gx[j] = xj;
#pragma unroll
            for (int c = 0; c < 7; ++c) gP[j * 7 + c] = row[c];
        }
        if (!ok) atomicOr(status, TK_DEV_BAD_CHOLESKY);
    }
    __syncthreads();
}

// KalmanBoxTracker.update(None): freeze on the observed -> unobserved transition (kalmanfilter.py:465-477)
__device__ void oc_track_miss(OcDev& S, int s) {
    const int hist = S.hist_len[s] + 1;
    if (S.observed[s]) {
        for (int k = 0; k < 7; ++k) S.fx[(size_t)s * 7 + k] = S.x[(size_t)s * 7 + k];
        for (int k = 0; k < 49; ++k) S.fP[(size_t)s * 49 + k] = S.P[(size_t)s * 49 + k];
        S.frozen_n[s] = hist;
        S.frozen[s] = 1;
    }
    S.observed[s] = 0;
    S.hist_len[s] = hist;
}

struct OcShared {
    int lap_ok;
    int nd, nlo, nt, shortcut, n_ud, n_ut, n_pairs, sorted_ud, n_births, n_out;
    int rowmax, colmax, maxflag;
    unsigned long long dmax_bits;   // ct_dist: largest centre distance of the current matrix (non-negative double bits are ordered)
};

// Solve the min-cost full assignment of the smaller side (no limit). match_d[d] = t or -1.
// C is stored with the smaller side as rows and leading dimension lap_pitch(cols).
__device__ void oc_solve(const double* C, int nd, int nt, int* match_d, double* u, int* col4row, int* row4col,
                         int* path, int* ok_flag, int* status) {
    for (int i = threadIdx.x; i < nd; i += blockDim.x) match_d[i] = -1;
    __syncthreads();
    const bool d_rows = nd <= nt;
    const int nr = d_rows ? nd : nt, nc = d_rows ? nt : nd;
    const bool ok = lap_solve_cta(C, lap_pitch(nc), nr, nc, false, u, col4row, row4col, path, ok_flag);
    if (!ok) { if (threadIdx.x == 0) atomicOr(status, TK_DEV_LAP_INFEASIBLE); return; }
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        const int c = col4row[r];
        if (d_rows) match_d[r] = c; else match_d[c] = r;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(OC_THREADS)
ocsort_video_kernel(OcParams prm, char* state_base, size_t state_stride, int cap, int capd,
                    const double* __restrict__ dets, const int* __restrict__ offsets, int n_frames,
                    double* __restrict__ out_rows, const int* __restrict__ out_start, int* __restrict__ out_frame_count,
                    int* __restrict__ out_count, double* cost_scratch, size_t cost_stride, int cost_in_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int seq = blockIdx.x, tid = threadIdx.x;
    OcDev S = oc_carve(state_base + (size_t)seq * state_stride, cap);
    unsigned char* sp = smem_raw;
    auto take = [&](size_t bytes) { unsigned char* p = sp; sp += (bytes + 15) & ~(size_t)15; return p; };
    const int side = cap > capd ? cap : capd;
    double* lap_u = (double*)take(sizeof(double) * side);
    double* trk_box = (double*)take(sizeof(double) * 4 * cap);   // predicted boxes, by list position
    double* kobs = (double*)take(sizeof(double) * 5 * cap);
    int* d_hi = (int*)take(sizeof(int) * capd);
    int* d_lo = (int*)take(sizeof(int) * capd);
    int* match_d = (int*)take(sizeof(int) * side);
    int* col4row = (int*)take(sizeof(int) * side);
    int* row4col = (int*)take(sizeof(int) * side);
    int* path = (int*)take(sizeof(int) * side);
    int* un_d = (int*)take(sizeof(int) * capd);
    int* un_t = (int*)take(sizeof(int) * cap);
    int* tmp_d = (int*)take(sizeof(int) * capd);
    int* tmp_t = (int*)take(sizeof(int) * cap);
    int* rowcnt = (int*)take(sizeof(int) * (2 * side + 2));   // also: slot / order scratch of oc_apply_updates
    int* colcnt = (int*)take(sizeof(int) * (2 * side + 2));   // also: ORU gap scratch of oc_apply_updates
    unsigned char* flag_d = (unsigned char*)take(capd);
    unsigned char* flag_t = (unsigned char*)take(cap);
    double* zbuf = (double*)take(sizeof(double) * 4 * capd);        // measurements of the pairs being applied (oc_apply_updates)
    unsigned char* upd_need = (unsigned char*)take(capd);
    double* oru_buf = (double*)take(sizeof(double) * 8 * capd);     // ORU line parameters of the pairs being applied
    double* tprm = (double*)take(sizeof(double) * 5 * cap);         // per-track terms of the round-1 direction cost
    OcShared* sh = (OcShared*)take(sizeof(OcShared));
    // per-slot scalars and the small observation records are mirrored in shared memory for the whole launch
    // (the frame loop is a chain of short serial list edits; global round trips there are pure latency)
    const OcDev G = S;
#define OC_MIRROR(X) X(last_obs, double, 5) X(vel, double, 2) X(last_z, double, 4) X(conf, double, 1) X(cls, double, 1) \
    X(det_id, double, 1) X(tsu, int, 1) X(uid, int, 1) X(hits, int, 1) X(streak, int, 1) X(age, int, 1) X(hist_len, int, 1) \
    X(frozen_n, int, 1) X(last_z_idx, int, 1) X(list, int, 1) X(free_list, int, 1) X(has_vel, unsigned char, 1)           \
    X(observed, unsigned char, 1) X(frozen, unsigned char, 1)
    {
        int* m_hdr = (int*)take(8 * sizeof(int));
        if (tid < 8) m_hdr[tid] = G.hdr[tid];
        S.hdr = m_hdr;
#define X(name, type, n) { type* m = (type*)take(sizeof(type) * (size_t)cap * (n)); \
        for (int i = tid; i < cap * (n); i += OC_THREADS) m[i] = G.name[i]; S.name = m; }
        OC_MIRROR(X)
#undef X
        __syncthreads();
    }
    double* iou_m = cost_in_smem ? (double*)take(sizeof(double) * (size_t)(cap + 1) * (capd + 1)) : cost_scratch + (size_t)seq * cost_stride * 2;
    double* cost = cost_in_smem ? (double*)take(0) : iou_m + cost_stride;

    int* status = &S.hdr[4];
    const int F1 = n_frames + 1;
    const int out_base = out_start[seq];
    int out_n = out_count[seq];

#ifdef TK_PHASE_PROF
    long long ph_t0 = clock64();
#endif
    for (int f = 0; f < n_frames; ++f) {
        const int r0 = offsets[seq * F1 + f], r1 = offsets[seq * F1 + f + 1];
        const int nraw = r1 - r0;
        if (nraw == 0) { if (tid == 0) out_frame_count[seq * n_frames + f] = 0; continue; }   // oc_sort_api.py:51-52
        if (nraw > capd) { if (tid == 0) atomicOr(status, TK_DEV_OVERFLOW_DETS); break; }
        const double* D = dets + (size_t)r0 * 7;

        if (warp_id() == 0) {   // wrapper filter (oc_sort_api.py:54), high / low score split (ocsort.py:226-231), order kept
            const int nh = warp_compact(nraw, 0, [&](int i) { const double c = D[i * 7 + 4]; return c > prm.min_conf && c > prm.det_thresh; },
                                        [&](int i, int p) { d_hi[p] = i; });
            const int nl = warp_compact(nraw, 0, [&](int i) { const double c = D[i * 7 + 4]; return c > prm.min_conf && !(c > prm.det_thresh) && c > 0.1 && c < prm.det_thresh; },
                                        [&](int i, int p) { d_lo[p] = i; });
            if (lane_id() == 0) { S.hdr[0] += 1; sh->nd = nh; sh->nlo = nl; sh->nt = S.hdr[2]; }
        }
        __syncthreads();
        const int frame_count = S.hdr[0];
        int nt = sh->nt;
        const int nd = sh->nd, nlo = sh->nlo;
        PH(0);

        // ---- predict every tracker (ocsort.py:234-244, :150-163) ------------------------------------
        for (int k = tid; k < nt; k += OC_THREADS) {
            const int s = S.list[k];
            double x[7], P[49], b[4];
            for (int i = 0; i < 7; ++i) x[i] = S.x[(size_t)s * 7 + i];
            for (int i = 0; i < 49; ++i) P[i] = S.P[(size_t)s * 49 + i];
            if (x[6] + x[2] <= 0) x[6] *= 0.0;
            kf7_predict(x, P);
            for (int i = 0; i < 7; ++i) S.x[(size_t)s * 7 + i] = x[i];
            for (int i = 0; i < 49; ++i) S.P[(size_t)s * 49 + i] = P[i];
            S.age[s] += 1;
            if (S.tsu[s] > 0) S.streak[s] = 0;
            S.tsu[s] += 1;
            x_to_box(x, b);
            for (int i = 0; i < 4; ++i) trk_box[4 * k + i] = b[i];
            flag_t[k] = (isnan(b[0]) || isnan(b[1]) || isnan(b[2]) || isnan(b[3])) ? 1 : 0;
        }
        __syncthreads();
        PH(1);
        if (tid == 0) sh->maxflag = 0;
        __syncthreads();
        for (int k = tid; k < nt; k += OC_THREADS) if (flag_t[k]) sh->maxflag = 1;
        __syncthreads();
        if (tid == 0 && sh->maxflag) {   // drop trackers whose prediction is not finite (ocsort.py:241-244); practically never taken
            int n = 0, nfree = S.hdr[5];
            for (int k = 0; k < nt; ++k) {
                const int s = S.list[k];
                if (flag_t[k]) { S.free_list[nfree++] = s; continue; }
                if (n != k) { for (int i = 0; i < 4; ++i) trk_box[4 * n + i] = trk_box[4 * k + i]; }
                S.list[n++] = s;
            }
            S.hdr[5] = nfree; S.hdr[2] = n; sh->nt = n;
        }
        __syncthreads();
        nt = sh->nt;
        // k_previous_obs (ocsort.py:10-18)
        PH(2);
        for (int k = tid; k < nt; k += OC_THREADS) {
            const int s = S.list[k];
            const double* src = nullptr;
            const double* lo = S.last_obs + (size_t)s * 5;
            const bool has = (lo[0] + lo[1] + lo[2] + lo[3] + lo[4]) >= 0;   // observations dict non-empty <=> observed once
            if (has) {
                const int age = S.age[s];
                for (int i = 0; i < prm.delta_t; ++i) {
                    const int want = age - (prm.delta_t - i);
                    const int r = ((want % RING) + RING) % RING;
                    if (want >= 0 && S.ring_age[(size_t)s * RING + r] == want) { src = S.ring_obs + ((size_t)s * RING + r) * 5; break; }
                }
                if (!src) src = lo;   // observations[max(keys)] is the last observation
            }
            for (int i = 0; i < 5; ++i) kobs[5 * k + i] = src ? src[i] : -1.0;
        }
        __syncthreads();

        PH(3);
        // ---- first round: associate() (association.py:242-298) -------------------------------------
        if (tid == 0) { sh->rowmax = 0; sh->colmax = 0; sh->n_pairs = 0; }
        for (int i = tid; i < nd; i += OC_THREADS) { rowcnt[i] = 0; match_d[i] = -1; }
        for (int i = tid; i < nt; i += OC_THREADS) colcnt[i] = 0;
        __syncthreads();
        const bool d_rows = nd <= nt;
        const int ld = lap_pitch(d_rows ? nt : nd);
        if (nt > 0) {
            // per-detection and per-track terms of the velocity-direction consistency (association.py:175-184,246-266) once per frame
            for (int d = tid; d < nd; d += OC_THREADS) {
                const double* db = D + (size_t)d_hi[d] * 7;
                zbuf[2 * d] = (db[0] + db[2]) / 2.0; zbuf[2 * d + 1] = (db[1] + db[3]) / 2.0;
            }
            for (int t = tid; t < nt; t += OC_THREADS) {
                const int s = S.list[t];
                const double* ko = kobs + 5 * t;
                double* tp = tprm + 5 * t;
                tp[0] = (ko[0] + ko[2]) / 2.0; tp[1] = (ko[1] + ko[3]) / 2.0;
                tp[2] = S.has_vel[s] ? S.vel[(size_t)s * 2 + 1] : 0.0;   // vx
                tp[3] = S.has_vel[s] ? S.vel[(size_t)s * 2 + 0] : 0.0;   // vy
                tp[4] = ko[4] < 0 ? 0.0 : 1.0;
            }
            __syncthreads();
            PH(12);
            // two independent entries per trip: the chain sqrt -> 2 divisions -> acos -> division is pure FP64 latency with 8 warps
#pragma unroll 2
            for (int e = tid; e < nd * nt; e += OC_THREADS) {
                const int d = e / nt, t = e - d * nt;
                const double* db = D + (size_t)d_hi[d] * 7;
                const double iou = iou_plain(db, trk_box + 4 * t);
                const double* tp = tprm + 5 * t;
                double dx = zbuf[2 * d] - tp[0], dy = zbuf[2 * d + 1] - tp[1];
                const double norm = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) + 1e-6;
                dx = dx / norm; dy = dy / norm;
                double c = __dadd_rn(__dmul_rn(tp[2], dx), __dmul_rn(tp[3], dy));
                c = fmin(fmax(c, -1.0), 1.0);
                const double pi = 3.141592653589793;
                const double ang = (pi / 2.0 - fabs(acos(c))) / pi;
                const double vdc = __dmul_rn(__dmul_rn(__dmul_rn(tp[4], ang), prm.inertia), db[5]);   // x class column (q1)
                iou_m[(size_t)d * nt + t] = iou;
                const double cc = -__dadd_rn(iou, vdc);
                if (d_rows) cost[(size_t)d * ld + t] = cc; else cost[(size_t)t * ld + d] = cc;
                if (iou > prm.iou_threshold) { atomicAdd(&rowcnt[d], 1); atomicAdd(&colcnt[t], 1); }
            }
            PH(13);
            __syncthreads();
            PH(4);
            for (int i = tid; i < nd; i += OC_THREADS) atomicMax(&sh->rowmax, rowcnt[i]);
            for (int i = tid; i < nt; i += OC_THREADS) atomicMax(&sh->colmax, colcnt[i]);
            __syncthreads();
            if (nd > 0) {
                if (sh->rowmax == 1 && sh->colmax == 1) {   // thresholded IoU is a partial permutation: use it (association.py:268-270)
                    for (int e = tid; e < nd * nt; e += OC_THREADS)
                        if (iou_m[e] > prm.iou_threshold) match_d[e / nt] = e % nt;
                    __syncthreads();
                } else {
                    oc_solve(cost, nd, nt, match_d, lap_u, col4row, row4col, path, &sh->lap_ok, status);
                }
            }
        }
        __syncthreads();
        PH(5);
        // unmatched lists in the reference's order: not-in-pairs ascending, then low-IoU pairs in pair order
        for (int t = tid; t < nt; t += OC_THREADS) flag_t[t] = 0;
        __syncthreads();
        for (int d = tid; d < nd; d += OC_THREADS) if (match_d[d] >= 0) flag_t[match_d[d]] = 1;
        __syncthreads();
        if (warp_id() == 0) {
            const int nud0 = warp_compact(nd, 0, [&](int d) { return match_d[d] < 0; }, [&](int d, int p) { un_d[p] = d; });
            const int nut0 = warp_compact(nt, 0, [&](int t) { return !flag_t[t]; }, [&](int t, int p) { un_t[p] = t; });
            const int nlow = warp_compact(nd, 0, [&](int d) { const int t = match_d[d]; return t >= 0 && iou_m[(size_t)d * nt + t] < prm.iou_threshold; },
                                          [&](int d, int p) { un_d[nud0 + p] = d; un_t[nut0 + p] = match_d[d]; match_d[d] = -1; });
            if (lane_id() == 0) { sh->n_ud = nud0 + nlow; sh->n_ut = nut0 + nlow; }
        }
        __syncthreads();
        PH(10);
        oc_apply_updates(S, nd, [&](int d, int& slot, const double*& db) {   // ocsort.py:257-258
            const int t = match_d[d];
            slot = t >= 0 ? S.list[t] : -1; db = D + (size_t)d_hi[d] * 7;
        }, prm.delta_t, status, zbuf, oru_buf, colcnt, upd_need, rowcnt);

        PH(6);
        // ---- BYTE round on low-score detections (ocsort.py:264-282) -----------------------------------
        if (prm.use_byte && nlo > 0 && sh->n_ut > 0) {
            const int nut = sh->n_ut;
            if (tid == 0) { sh->maxflag = 0; sh->dmax_bits = 0ull; }
            __syncthreads();
            const bool dr = nlo <= nut;
            const int l2 = lap_pitch(dr ? nut : nlo);
            if (prm.asso == TK_ASSO_CT_DIST) {
                for (int e = tid; e < nlo * nut; e += OC_THREADS) {
                    const double dd = centre_dist(D + (size_t)d_lo[e / nut] * 7, trk_box + 4 * un_t[e % nut]);
                    iou_m[e] = dd;
                    atomicMax(&sh->dmax_bits, (unsigned long long)__double_as_longlong(dd));
                }
                __syncthreads();
            }
            const double dmax = __longlong_as_double((long long)sh->dmax_bits);
            for (int e = tid; e < nlo * nut; e += OC_THREADS) {
                const int d = e / nut, t = e % nut;
                const double v = prm.asso == TK_ASSO_CT_DIST ? 1.0 - iou_m[e] / dmax
                                                             : asso_value(prm.asso, D + (size_t)d_lo[d] * 7, trk_box + 4 * un_t[t]);
                iou_m[(size_t)d * nut + t] = v;
                if (dr) cost[(size_t)d * l2 + t] = -v; else cost[(size_t)t * l2 + d] = -v;
                if (v > prm.iou_threshold) sh->maxflag = 1;
            }
            __syncthreads();
            if (sh->maxflag) {
                oc_solve(cost, nlo, nut, match_d, lap_u, col4row, row4col, path, &sh->lap_ok, status);
                for (int t = tid; t < nut; t += OC_THREADS) flag_t[t] = 0;
                __syncthreads();
                oc_apply_updates(S, nlo, [&](int d, int& slot, const double*& db) {
                    const int t = match_d[d];
                    db = D + (size_t)d_lo[d] * 7;
                    if (t < 0 || iou_m[(size_t)d * nut + t] < prm.iou_threshold) { slot = -1; return; }
                    slot = S.list[un_t[t]];
                    flag_t[t] = 1;
                }, prm.delta_t, status, zbuf, oru_buf, colcnt, upd_need, rowcnt);
                if (tid == 0) {   // np.setdiff1d: sorted remaining tracker indices
                    int n = 0;
                    for (int t = 0; t < nut; ++t) if (!flag_t[t]) tmp_t[n++] = un_t[t];
                    for (int i = 1; i < n; ++i) { const int v = tmp_t[i]; int j = i - 1; while (j >= 0 && tmp_t[j] > v) { tmp_t[j + 1] = tmp_t[j]; --j; } tmp_t[j + 1] = v; }
                    for (int i = 0; i < n; ++i) un_t[i] = tmp_t[i];
                    sh->n_ut = n;
                }
                __syncthreads();
            }
        }

        PH(7);
        // ---- OCR round on the last observations (ocsort.py:284-306) -----------------------------------
        if (sh->n_ud > 0 && sh->n_ut > 0) {
            const int nud = sh->n_ud, nut = sh->n_ut;
            if (tid == 0) { sh->maxflag = 0; sh->dmax_bits = 0ull; }
            __syncthreads();
            const bool dr = nud <= nut;
            const int l2 = lap_pitch(dr ? nut : nud);
            if (prm.asso == TK_ASSO_CT_DIST) {
                for (int e = tid; e < nud * nut; e += OC_THREADS) {
                    const double dd = centre_dist(D + (size_t)d_hi[un_d[e / nut]] * 7, S.last_obs + (size_t)S.list[un_t[e % nut]] * 5);
                    iou_m[e] = dd;
                    atomicMax(&sh->dmax_bits, (unsigned long long)__double_as_longlong(dd));
                }
                __syncthreads();
            }
            const double dmax = __longlong_as_double((long long)sh->dmax_bits);
            for (int e = tid; e < nud * nut; e += OC_THREADS) {
                const int d = e / nut, t = e % nut;
                const double v = prm.asso == TK_ASSO_CT_DIST ? 1.0 - iou_m[e] / dmax
                                                             : asso_value(prm.asso, D + (size_t)d_hi[un_d[d]] * 7, S.last_obs + (size_t)S.list[un_t[t]] * 5);
                iou_m[(size_t)d * nut + t] = v;
                if (dr) cost[(size_t)d * l2 + t] = -v; else cost[(size_t)t * l2 + d] = -v;
                if (v > prm.iou_threshold) sh->maxflag = 1;
            }
            __syncthreads();
            if (sh->maxflag) {
                oc_solve(cost, nud, nut, match_d, lap_u, col4row, row4col, path, &sh->lap_ok, status);
                for (int t = tid; t < nut; t += OC_THREADS) flag_t[t] = 0;
                for (int d = tid; d < nud; d += OC_THREADS) flag_d[d] = 0;
                __syncthreads();
                oc_apply_updates(S, nud, [&](int d, int& slot, const double*& db) {
                    const int t = match_d[d];
                    db = D + (size_t)d_hi[un_d[d]] * 7;
                    if (t < 0 || iou_m[(size_t)d * nut + t] < prm.iou_threshold) { slot = -1; return; }
                    slot = S.list[un_t[t]];
                    flag_t[t] = 1; flag_d[d] = 1;
                }, prm.delta_t, status, zbuf, oru_buf, colcnt, upd_need, rowcnt);
                if (tid == 0) {   // np.setdiff1d on both lists (sorted)
                    int n = 0;
                    for (int t = 0; t < nut; ++t) if (!flag_t[t]) tmp_t[n++] = un_t[t];
                    for (int i = 1; i < n; ++i) { const int v = tmp_t[i]; int j = i - 1; while (j >= 0 && tmp_t[j] > v) { tmp_t[j + 1] = tmp_t[j]; --j; } tmp_t[j + 1] = v; }
                    for (int i = 0; i < n; ++i) un_t[i] = tmp_t[i];
                    sh->n_ut = n;
                    n = 0;
                    for (int d = 0; d < nud; ++d) if (!flag_d[d]) tmp_d[n++] = un_d[d];
                    for (int i = 1; i < n; ++i) { const int v = tmp_d[i]; int j = i - 1; while (j >= 0 && tmp_d[j] > v) { tmp_d[j + 1] = tmp_d[j]; --j; } tmp_d[j + 1] = v; }
                    for (int i = 0; i < n; ++i) un_d[i] = tmp_d[i];
                    sh->n_ud = n;
                }
                __syncthreads();
            }
        }

        PH(8);
        // ---- unmatched trackers freeze; unmatched detections start trackers (ocsort.py:308-314) --------------
        for (int k = tid; k < sh->n_ut; k += OC_THREADS) oc_track_miss(S, S.list[un_t[k]]);
        __syncthreads();
        if (warp_id() == 0) {
            const int nfree = S.hdr[5], n = S.hdr[2], uid0 = S.hdr[1];
            const int nb = sh->n_ud < nfree ? sh->n_ud : nfree;
            if (sh->n_ud > nfree && lane_id() == 0) atomicOr(status, TK_DEV_OVERFLOW_TRACKS);
            for (int k = lane_id(); k < nb; k += 32) {
                const int s = S.free_list[nfree - 1 - k];
                S.list[n + k] = s;
                tmp_d[k] = un_d[k]; tmp_t[k] = s;
                S.uid[s] = uid0 + k;
            }
            __syncwarp();
            if (lane_id() == 0) { S.hdr[5] = nfree - nb; S.hdr[2] = n + nb; S.hdr[1] = uid0 + nb; sh->n_births = nb; }
        }
        __syncthreads();
        for (int k = tid; k < sh->n_births; k += OC_THREADS) {   // KalmanBoxTracker.__init__ (ocsort.py:63-101)
            const int s = tmp_t[k];
            const double* db = D + (size_t)d_hi[tmp_d[k]] * 7;
            double z[4];
            box_to_z(db, z);
            double* x = S.x + (size_t)s * 7;
            double* P = S.P + (size_t)s * 49;
            for (int i = 0; i < 7; ++i) x[i] = i < 4 ? z[i] : 0.0;
            for (int i = 0; i < 49; ++i) P[i] = 0.0;
            for (int i = 0; i < 7; ++i) P[i * 8] = i < 4 ? 10.0 : 10000.0;
            S.tsu[s] = 0; S.hits[s] = 0; S.streak[s] = 0; S.age[s] = 0;
            S.conf[s] = db[4]; S.cls[s] = db[5]; S.det_id[s] = db[6];
            for (int i = 0; i < 5; ++i) S.last_obs[(size_t)s * 5 + i] = -1.0;
            for (int i = 0; i < RING; ++i) S.ring_age[(size_t)s * RING + i] = -1;
            S.has_vel[s] = 0; S.observed[s] = 0; S.frozen[s] = 0; S.hist_len[s] = 0; S.last_z_idx[s] = -1;
        }
        __syncthreads();

        PH(9);
        // ---- output rows + death (ocsort.py:315-334), walking the tracker list backwards ------------------------
        if (warp_id() == 0) {
            const int n = S.hdr[2];
            for (int k = lane_id(); k < n; k += 32) rowcnt[k] = -1;
            __syncwarp();
            const int cnt = warp_compact(n, 0, [&](int i) { const int s = S.list[n - 1 - i];
                                                           return (S.tsu[s] < 1) && (S.streak[s] >= prm.min_hits || frame_count <= prm.min_hits); },
                                         [&](int i, int p) { rowcnt[n - 1 - i] = p; });
            if (lane_id() == 0) { sh->n_out = cnt; out_frame_count[seq * n_frames + f] = cnt; }
        }
        __syncthreads();
        {
            const int n = S.hdr[2];
            for (int k = tid; k < n; k += OC_THREADS) {
                if (rowcnt[k] < 0) continue;
                const int s = S.list[k];
                double* o = out_rows + (size_t)(out_base + out_n + rowcnt[k]) * 8;
                const double* lo = S.last_obs + (size_t)s * 5;
                if (lo[0] + lo[1] + lo[2] + lo[3] + lo[4] < 0) {
                    double b[4];
                    x_to_box(S.x + (size_t)s * 7, b);
                    o[0] = b[0]; o[1] = b[1]; o[2] = b[2]; o[3] = b[3];
                } else { o[0] = lo[0]; o[1] = lo[1]; o[2] = lo[2]; o[3] = lo[3]; }
                o[4] = (double)(S.uid[s] + 1); o[5] = S.cls[s]; o[6] = S.conf[s]; o[7] = S.det_id[s];
            }
            out_n += sh->n_out;
        }
        __syncthreads();
        if (warp_id() == 0) {
            const int n0 = S.hdr[2];
            const int nfree = warp_compact(n0, S.hdr[5], [&](int k) { return S.tsu[S.list[k]] > prm.max_age; }, [&](int k, int p) { S.free_list[p] = S.list[k]; });
            const int n = warp_compact(n0, 0, [&](int k) { return !(S.tsu[S.list[k]] > prm.max_age); }, [&](int k, int p) { tmp_t[p] = S.list[k]; });
            for (int k = lane_id(); k < n; k += 32) S.list[k] = tmp_t[k];
            __syncwarp();
            if (lane_id() == 0) { S.hdr[2] = n; S.hdr[5] = nfree; }
        }
        __syncthreads();
        PH(11);
    }
    if (tid == 0) out_count[seq] = out_n;
    __syncthreads();
    if (tid < 8) G.hdr[tid] = S.hdr[tid];
#define X(name, type, n) for (int i = tid; i < cap * (n); i += OC_THREADS) G.name[i] = S.name[i];
    OC_MIRROR(X)
#undef X
}

struct OcHandle {
    OcParams prm;
    int n_seq, cap, capd;
    char* state;
    size_t state_stride;
    double* cost;
    size_t cost_stride, smem_bytes;
    int cost_in_smem;
};

__global__ void ocsort_reset_kernel(char* base, size_t stride, int cap) {
    OcDev S = oc_carve(base + (size_t)blockIdx.x * stride, cap);
    if (threadIdx.x == 0) { S.hdr[0] = 0; S.hdr[1] = 0; S.hdr[2] = 0; S.hdr[3] = 0; S.hdr[4] = 0; S.hdr[5] = cap; }
    for (int i = threadIdx.x; i < cap; i += blockDim.x) S.free_list[i] = cap - 1 - i;
}

size_t oc_smem_fixed(int cap, int capd) {
    const int side = cap > capd ? cap : capd;
    auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
    size_t s = al(sizeof(double) * side) + al(sizeof(double) * 4 * cap) + al(sizeof(double) * 5 * cap);
    s += 2 * al(sizeof(int) * capd) + 4 * al(sizeof(int) * side);
    s += al(sizeof(int) * capd) + al(sizeof(int) * cap) + al(sizeof(int) * capd) + al(sizeof(int) * cap);
    s += 2 * al(sizeof(int) * (2 * side + 2)) + al((size_t)capd) + al((size_t)cap) + al(sizeof(OcShared));
    s += al(sizeof(double) * 4 * capd) + al((size_t)capd) + al(sizeof(double) * 8 * capd) + al(sizeof(double) * 5 * cap);
    // book-keeping mirror: hdr + per-slot records (see OC_MIRROR in the kernel)
    s += al(8 * sizeof(int)) + al(sizeof(double) * cap * 5) + al(sizeof(double) * cap * 2) + al(sizeof(double) * cap * 4)
       + 3 * al(sizeof(double) * cap) + 10 * al(sizeof(int) * cap) + 3 * al((size_t)cap);
    return s;
}

}  // namespace

extern "C" {

int tk_ocsort_create(const tk_ocsort_params* p, int n_seq, int cap_tracks, int cap_dets, void** handle) {
    if (!p || !handle || n_seq <= 0 || cap_tracks <= 0 || cap_dets <= 0) return TK_ERR_ARG;
    if (cap_tracks > tk::LAP_MAX_COLS || cap_dets > tk::LAP_MAX_COLS) return TK_ERR_CAPACITY;
    if (p->delta_t < 1 || p->delta_t > RING || p->asso_func < 0 || p->asso_func > TK_ASSO_CT_DIST) return TK_ERR_ARG;
    OcHandle* h = new OcHandle();
    h->prm.det_thresh = p->det_thresh; h->prm.iou_threshold = p->iou_threshold; h->prm.inertia = p->inertia;
    h->prm.min_conf = p->min_confidence; h->prm.max_age = p->max_age; h->prm.min_hits = p->min_hits;
    h->prm.delta_t = p->delta_t; h->prm.asso = p->asso_func; h->prm.use_byte = p->use_byte;
    h->n_seq = n_seq; h->cap = cap_tracks; h->capd = cap_dets;
    h->state_stride = (oc_state_bytes(cap_tracks) + 255) & ~(size_t)255;
    h->state = nullptr; h->cost = nullptr;
    const size_t fixed = oc_smem_fixed(cap_tracks, cap_dets);
    const size_t mat = (size_t)(cap_tracks + 1) * (cap_dets + 1) * sizeof(double);
    h->cost_in_smem = (fixed + 2 * mat <= 200 * 1024) ? 1 : 0;
    h->smem_bytes = fixed + (h->cost_in_smem ? 2 * mat : 0);
    h->cost_stride = (size_t)(cap_tracks + 1) * (cap_dets + 1);
    cudaError_t e = cudaMalloc((void**)&h->state, h->state_stride * n_seq);
    if (e == cudaSuccess && !h->cost_in_smem) e = cudaMalloc((void**)&h->cost, 2 * h->cost_stride * sizeof(double) * n_seq);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(ocsort_video_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    if (e != cudaSuccess) {
        tk_set_last_cuda_error((int)e);
        if (h->state) cudaFree(h->state);
        if (h->cost) cudaFree(h->cost);
        delete h;
        return TK_ERR_CUDA;
    }
    *handle = h;
    return tk_ocsort_reset(h, 0, nullptr);
}

int tk_ocsort_reset(void* handle, int keep_id_counter, void* stream) {
    (void)keep_id_counter;  // KalmanBoxTracker.count is reset by every OCSort() (ocsort.py:201)
    if (!handle) return TK_ERR_ARG;
    OcHandle* h = (OcHandle*)handle;
    ocsort_reset_kernel<<<h->n_seq, 128, 0, (cudaStream_t)stream>>>(h->state, h->state_stride, h->cap);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_ocsort_run(void* handle, const double* dets, const int* offsets, int n_frames, double* out_rows,
                  const int* out_start, int* out_frame_count, int* out_count, void* stream) {
    if (!handle || !offsets || !out_rows || !out_start || !out_frame_count || !out_count || n_frames < 0) return TK_ERR_ARG;
    OcHandle* h = (OcHandle*)handle;
    if (n_frames == 0) return TK_OK;
    // the attribute is per kernel function, not per handle: another handle with smaller capacities may have lowered it
    TK_CUDA_TRY(cudaFuncSetAttribute(ocsort_video_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    ocsort_video_kernel<<<h->n_seq, OC_THREADS, h->smem_bytes, (cudaStream_t)stream>>>(
        h->prm, h->state, h->state_stride, h->cap, h->capd, dets, offsets, n_frames, out_rows, out_start,
        out_frame_count, out_count, h->cost, h->cost_stride, h->cost_in_smem);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_ocsort_status(void* handle, int* status_host, void* stream) {
    if (!handle || !status_host) return TK_ERR_ARG;
    OcHandle* h = (OcHandle*)handle;
    for (int s = 0; s < h->n_seq; ++s)
        TK_CUDA_TRY(cudaMemcpyAsync(status_host + s, h->state + (size_t)s * h->state_stride + 4 * sizeof(int), sizeof(int),
                                    cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    TK_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return TK_OK;
}

#ifdef TK_PHASE_PROF
int tk_debug_ocsort_phases(unsigned long long* host_out64, int reset) {
    cudaDeviceSynchronize();
    if (host_out64) cudaMemcpyFromSymbol(host_out64, g_oc_prof, sizeof(unsigned long long) * 64);
    if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(g_oc_prof, z, sizeof(z)); }
    return 0;
}
#endif

int tk_ocsort_destroy(void* handle) {
    if (!handle) return TK_ERR_ARG;
    OcHandle* h = (OcHandle*)handle;
    cudaFree(h->state);
    if (h->cost) cudaFree(h->cost);
    delete h;
    return TK_OK;
}

}  // extern "C"
