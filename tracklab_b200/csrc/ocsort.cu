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
#include "tk_common.cuh"
#include "trackkern.h"

namespace {

using namespace tk;

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

// ---- IoU family (association.py:5-171), boxes x1y1x2y2, float64 -----------------------------------------
__device__ __forceinline__ double iou_plain(const double* a, const double* b) {
    const double w = fmax(0.0, fmin(a[2], b[2]) - fmax(a[0], b[0]));
    const double h = fmax(0.0, fmin(a[3], b[3]) - fmax(a[1], b[1]));
    const double wh = __dmul_rn(w, h);
    const double ua = __dsub_rn(__dadd_rn(__dmul_rn(a[2] - a[0], a[3] - a[1]), __dmul_rn(b[2] - b[0], b[3] - b[1])), wh);
    return wh / ua;
}

__device__ double asso_value(int kind, const double* a, const double* b) {
    const double iou = iou_plain(a, b);
    if (kind == 0) return iou;
    const double w = fmax(0.0, fmin(a[2], b[2]) - fmax(a[0], b[0]));
    const double h = fmax(0.0, fmin(a[3], b[3]) - fmax(a[1], b[1]));
    const double wh = __dmul_rn(w, h);
    const double wc = fmax(a[2], b[2]) - fmin(a[0], b[0]);
    const double hc = fmax(a[3], b[3]) - fmin(a[1], b[1]);
    if (kind == 1) {  // giou (association.py:24-55)
        const double hull = __dmul_rn(wc, hc);
        const double g = __dsub_rn(iou, __dsub_rn(hull, wh) / hull);
        return __dadd_rn(g, 1.0) / 2.0;
    }
    const double dcx = __dsub_rn((a[0] + a[2]) / 2.0, (b[0] + b[2]) / 2.0);
    const double dcy = __dsub_rn((a[1] + a[3]) / 2.0, (b[1] + b[3]) / 2.0);
    const double inner = __dadd_rn(__dmul_rn(dcx, dcx), __dmul_rn(dcy, dcy));
    const double outer = __dadd_rn(__dmul_rn(wc, wc), __dmul_rn(hc, hc));
    if (kind == 2) return __dadd_rn(__dsub_rn(iou, inner / outer), 1.0) / 2.0;  // diou (:58-95)
    // ciou (:97-147)
    const double w1 = a[2] - a[0], h1 = (a[3] - a[1]) + 1.0, w2 = b[2] - b[0], h2 = (b[3] - b[1]) + 1.0;
    const double at = __dsub_rn(atan(w2 / h2), atan(w1 / h1));
    const double pi = 3.141592653589793;
    const double v = __dmul_rn(4.0 / __dmul_rn(pi, pi), __dmul_rn(at, at));
    const double alpha = v / __dadd_rn(__dsub_rn(1.0, iou), v);
    return __dadd_rn(__dsub_rn(__dsub_rn(iou, inner / outer), __dmul_rn(alpha, v)), 1.0) / 2.0;
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

__device__ bool kf7_correct(double* x, double* P, const double* z) {
    const double R[4] = {1.0, 1.0, 10.0, 10.0};
    double S[16], L[16], Li[16], SI[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) S[i * 4 + j] = (i == j) ? P[i * 7 + j] + R[i] : P[i * 7 + j];
    bool ok = true;
    for (int j = 0; j < 4; ++j) {  // Cholesky S = L L^T
        double d = S[j * 4 + j];
        for (int k = 0; k < j; ++k) d -= L[j * 4 + k] * L[j * 4 + k];
        ok = ok && d > 0.0;
        const double ljj = sqrt(d);
        L[j * 4 + j] = ljj;
        for (int i = j + 1; i < 4; ++i) {
            double s = S[i * 4 + j];
            for (int k = 0; k < j; ++k) s -= L[i * 4 + k] * L[j * 4 + k];
            L[i * 4 + j] = s / ljj;
        }
        for (int i = 0; i < j; ++i) L[i * 4 + j] = 0.0;
    }
    for (int c = 0; c < 4; ++c) {  // Li = L^-1 (lower)
        for (int i = 0; i < 4; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) s -= L[i * 4 + k] * Li[k * 4 + c];
            Li[i * 4 + c] = (i < c) ? 0.0 : s / L[i * 4 + i];
        }
    }
    for (int i = 0; i < 4; ++i)   // SI = Li^T Li
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = (i > j ? i : j); k < 4; ++k) s += Li[k * 4 + i] * Li[k * 4 + j];
            SI[i * 4 + j] = s;
        }
    double K[28], y[4];
    for (int i = 0; i < 4; ++i) y[i] = z[i] - x[i];
    for (int a = 0; a < 7; ++a)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += P[a * 7 + k] * SI[k * 4 + j];
            K[a * 4 + j] = s;
        }
    for (int a = 0; a < 7; ++a) {
        double s = 0.0;
        for (int k = 0; k < 4; ++k) s += K[a * 4 + k] * y[k];
        x[a] = x[a] + s;
    }
    // P = (I-KH) P (I-KH)^T + K R K^T   (Joseph form, kalmanfilter.py:520-521)
    double A[49], AP[49];
    for (int a = 0; a < 7; ++a)
        for (int b = 0; b < 7; ++b) A[a * 7 + b] = ((a == b) ? 1.0 : 0.0) - (b < 4 ? K[a * 4 + b] : 0.0);
    for (int a = 0; a < 7; ++a)
        for (int b = 0; b < 7; ++b) {
            double s = 0.0;
            for (int k = 0; k < 7; ++k) s += A[a * 7 + k] * P[k * 7 + b];
            AP[a * 7 + b] = s;
        }
    for (int a = 0; a < 7; ++a)
        for (int b = 0; b < 7; ++b) {
            double s = 0.0;
            for (int k = 0; k < 7; ++k) s += AP[a * 7 + k] * A[b * 7 + k];
            double r = 0.0;
            for (int k = 0; k < 4; ++k) r += (K[a * 4 + k] * R[k]) * K[b * 4 + k];
            P[a * 7 + b] = s + r;
        }
    return ok;
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

// KalmanBoxTracker.update(bbox) incl. KalmanFilterNew.update with the ORU replay (ocsort.py:103-148)
__device__ void oc_track_update(OcDev& S, int s, const double* bbox5, double cls, double det_id, int delta_t, int* status) {
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
    double z[4], x[7], P[49];
    box_to_z(bbox5, z);
    double* gx = S.x + (size_t)s * 7;
    double* gP = S.P + (size_t)s * 49;
    int hist = S.hist_len[s] + 1;   // history_obs.append(z)
    bool ok = true;
    if (!S.observed[s] && S.frozen[s]) {
        // unfreeze (kalmanfilter.py:390-434): restore, interpolate between the last stored observation and z
        for (int k = 0; k < 7; ++k) x[k] = S.fx[(size_t)s * 7 + k];
        for (int k = 0; k < 49; ++k) P[k] = S.fP[(size_t)s * 49 + k];
        const int i1 = S.last_z_idx[s], i2 = hist - 1;
        const double* b1 = S.last_z + (size_t)s * 4;
        const double x1 = b1[0], y1 = b1[1], s1 = b1[2], r1 = b1[3];
        const double w1 = sqrt(__dmul_rn(s1, r1)), h1 = sqrt(s1 / r1);
        const double w2 = sqrt(__dmul_rn(z[2], z[3])), h2 = sqrt(z[2] / z[3]);
        const int gap = i2 - i1;
        const double dg = (double)gap;
        const double dx = (z[0] - x1) / dg, dy = (z[1] - y1) / dg, dw = (w2 - w1) / dg, dh = (h2 - h1) / dg;
        double vz[4];
        for (int i = 0; i < gap; ++i) {
            const double t = (double)(i + 1);
            const double ww = __dadd_rn(w1, __dmul_rn(t, dw)), hh = __dadd_rn(h1, __dmul_rn(t, dh));
            vz[0] = __dadd_rn(x1, __dmul_rn(t, dx)); vz[1] = __dadd_rn(y1, __dmul_rn(t, dy));
            vz[2] = __dmul_rn(ww, hh); vz[3] = ww / hh;
            ok = kf7_correct(x, P, vz) && ok;
            if (i != gap - 1) kf7_predict(x, P);
        }
        hist = S.frozen_n[s] - 1 + gap;
        S.frozen[s] = 0;
        for (int k = 0; k < 4; ++k) S.last_z[(size_t)s * 4 + k] = vz[k];
        S.last_z_idx[s] = hist - 1;
        ok = kf7_correct(x, P, z) && ok;   // the real measurement is applied on top (not appended again)
    } else {
        for (int k = 0; k < 7; ++k) x[k] = gx[k];
        for (int k = 0; k < 49; ++k) P[k] = gP[k];
        ok = kf7_correct(x, P, z);
        for (int k = 0; k < 4; ++k) S.last_z[(size_t)s * 4 + k] = z[k];
        S.last_z_idx[s] = hist - 1;
    }
    S.observed[s] = 1;
    S.hist_len[s] = hist;
    for (int k = 0; k < 7; ++k) gx[k] = x[k];
    for (int k = 0; k < 49; ++k) gP[k] = P[k];
    if (!ok) atomicOr(status, TK_DEV_BAD_CHOLESKY);
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

// ct_dist (association.py:150-171): centre distance d, then (d / d.max()).max() - d / d.max() = 1 - d / d.max()
// (NaN everywhere when all centres coincide, exactly like the 0/0 of the reference). Two passes over the matrix.
__device__ __forceinline__ double centre_dist(const double* a, const double* b) {
    const double dx = (a[0] + a[2]) / 2.0 - (b[0] + b[2]) / 2.0, dy = (a[1] + a[3]) / 2.0 - (b[1] + b[3]) / 2.0;
    return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

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
    int* rowcnt = (int*)take(sizeof(int) * side);
    int* colcnt = (int*)take(sizeof(int) * side);
    unsigned char* flag_d = (unsigned char*)take(capd);
    unsigned char* flag_t = (unsigned char*)take(cap);
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

    for (int f = 0; f < n_frames; ++f) {
        const int r0 = offsets[seq * F1 + f], r1 = offsets[seq * F1 + f + 1];
        const int nraw = r1 - r0;
        if (nraw == 0) { if (tid == 0) out_frame_count[seq * n_frames + f] = 0; continue; }   // oc_sort_api.py:51-52
        if (nraw > capd) { if (tid == 0) atomicOr(status, TK_DEV_OVERFLOW_DETS); break; }
        const double* D = dets + (size_t)r0 * 7;

        if (tid == 0) {
            S.hdr[0] += 1;
            int nh = 0, nl = 0;
            for (int i = 0; i < nraw; ++i) {
                const double c = D[i * 7 + 4];
                if (!(c > prm.min_conf)) continue;                             // oc_sort_api.py:54
                if (c > prm.det_thresh) d_hi[nh++] = i;                          // ocsort.py:230-231
                else if (c > 0.1 && c < prm.det_thresh) d_lo[nl++] = i;          // ocsort.py:226-229
            }
            sh->nd = nh; sh->nlo = nl; sh->nt = S.hdr[2];
        }
        __syncthreads();
        const int frame_count = S.hdr[0];
        int nt = sh->nt;
        const int nd = sh->nd, nlo = sh->nlo;

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
        if (tid == 0) {   // drop trackers whose prediction is not finite (ocsort.py:241-244)
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

        // ---- first round: associate() (association.py:242-298) -------------------------------------
        if (tid == 0) { sh->rowmax = 0; sh->colmax = 0; sh->n_pairs = 0; }
        for (int i = tid; i < nd; i += OC_THREADS) { rowcnt[i] = 0; match_d[i] = -1; }
        for (int i = tid; i < nt; i += OC_THREADS) colcnt[i] = 0;
        __syncthreads();
        const bool d_rows = nd <= nt;
        const int ld = lap_pitch(d_rows ? nt : nd);
        if (nt > 0) {
            for (int e = tid; e < nd * nt; e += OC_THREADS) {
                const int d = e / nt, t = e % nt;
                const double* db = D + (size_t)d_hi[d] * 7;
                const int s = S.list[t];
                const double iou = iou_plain(db, trk_box + 4 * t);
                // velocity-direction consistency (association.py:175-184,246-266)
                const double* ko = kobs + 5 * t;
                const double cx1 = (db[0] + db[2]) / 2.0, cy1 = (db[1] + db[3]) / 2.0;
                const double cx2 = (ko[0] + ko[2]) / 2.0, cy2 = (ko[1] + ko[3]) / 2.0;
                double dx = cx1 - cx2, dy = cy1 - cy2;
                const double norm = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) + 1e-6;
                dx = dx / norm; dy = dy / norm;
                const double vy = S.has_vel[s] ? S.vel[(size_t)s * 2 + 0] : 0.0;
                const double vx = S.has_vel[s] ? S.vel[(size_t)s * 2 + 1] : 0.0;
                double c = __dadd_rn(__dmul_rn(vx, dx), __dmul_rn(vy, dy));
                c = fmin(fmax(c, -1.0), 1.0);
                const double pi = 3.141592653589793;
                const double ang = (pi / 2.0 - fabs(acos(c))) / pi;
                const double valid = ko[4] < 0 ? 0.0 : 1.0;
                const double vdc = __dmul_rn(__dmul_rn(__dmul_rn(valid, ang), prm.inertia), db[5]);   // x class column (q1)
                iou_m[(size_t)d * nt + t] = iou;
                const double cc = -__dadd_rn(iou, vdc);
                if (d_rows) cost[(size_t)d * ld + t] = cc; else cost[(size_t)t * ld + d] = cc;
                if (iou > prm.iou_threshold) { atomicAdd(&rowcnt[d], 1); atomicAdd(&colcnt[t], 1); }
            }
            __syncthreads();
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
        if (tid == 0) {
            // unmatched lists in the reference's order: not-in-pairs ascending, then low-IoU pairs in pair order
            for (int t = 0; t < nt; ++t) flag_t[t] = 0;
            int nud = 0, nut = 0;
            for (int d = 0; d < nd; ++d) { if (match_d[d] < 0) un_d[nud++] = d; else flag_t[match_d[d]] = 1; }
            if (nt == 0) { nud = 0; for (int d = 0; d < nd; ++d) un_d[nud++] = d; }
            for (int t = 0; t < nt; ++t) if (!flag_t[t]) un_t[nut++] = t;
            for (int d = 0; d < nd; ++d) {
                const int t = match_d[d];
                if (t >= 0 && iou_m[(size_t)d * nt + t] < prm.iou_threshold) { un_d[nud++] = d; un_t[nut++] = t; match_d[d] = -1; }
            }
            sh->n_ud = nud; sh->n_ut = nut;
        }
        __syncthreads();
        for (int d = tid; d < nd; d += OC_THREADS) {   // ocsort.py:257-258
            const int t = match_d[d];
            if (t >= 0) { const double* db = D + (size_t)d_hi[d] * 7; oc_track_update(S, S.list[t], db, db[5], db[6], prm.delta_t, status); }
        }
        __syncthreads();

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
                for (int d = tid; d < nlo; d += OC_THREADS) {
                    const int t = match_d[d];
                    if (t < 0 || iou_m[(size_t)d * nut + t] < prm.iou_threshold) continue;
                    const double* db = D + (size_t)d_lo[d] * 7;
                    oc_track_update(S, S.list[un_t[t]], db, db[5], db[6], prm.delta_t, status);
                    flag_t[t] = 1;
                }
                __syncthreads();
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
                for (int d = tid; d < nud; d += OC_THREADS) {
                    const int t = match_d[d];
                    if (t < 0 || iou_m[(size_t)d * nut + t] < prm.iou_threshold) continue;
                    const double* db = D + (size_t)d_hi[un_d[d]] * 7;
                    oc_track_update(S, S.list[un_t[t]], db, db[5], db[6], prm.delta_t, status);
                    flag_t[t] = 1; flag_d[d] = 1;
                }
                __syncthreads();
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

        // ---- unmatched trackers freeze; unmatched detections start trackers (ocsort.py:308-314) --------------
        for (int k = tid; k < sh->n_ut; k += OC_THREADS) oc_track_miss(S, S.list[un_t[k]]);
        __syncthreads();
        if (tid == 0) {
            int nfree = S.hdr[5], n = S.hdr[2], nb = 0;
            for (int k = 0; k < sh->n_ud; ++k) {
                if (nfree == 0) { atomicOr(status, TK_DEV_OVERFLOW_TRACKS); break; }
                const int s = S.free_list[--nfree];
                S.list[n++] = s;
                tmp_d[nb] = un_d[k]; tmp_t[nb] = s; ++nb;
                S.uid[s] = S.hdr[1]++;
            }
            S.hdr[5] = nfree; S.hdr[2] = n; sh->n_births = nb;
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

        // ---- output rows + death (ocsort.py:315-334), walking the tracker list backwards ------------------------
        if (tid == 0) {
            const int n = S.hdr[2];
            int cnt = 0;
            for (int k = n - 1; k >= 0; --k) {
                const int s = S.list[k];
                const bool emit = (S.tsu[s] < 1) && (S.streak[s] >= prm.min_hits || frame_count <= prm.min_hits);
                rowcnt[k] = emit ? cnt++ : -1;
            }
            sh->n_out = cnt;
            out_frame_count[seq * n_frames + f] = cnt;
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
        if (tid == 0) {
            int n = 0, nfree = S.hdr[5];
            const int n0 = S.hdr[2];
            for (int k = 0; k < n0; ++k) {
                const int s = S.list[k];
                if (S.tsu[s] > prm.max_age) S.free_list[nfree++] = s; else S.list[n++] = s;
            }
            S.hdr[2] = n; S.hdr[5] = nfree;
        }
        __syncthreads();
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
    s += 2 * al(sizeof(int) * side) + al((size_t)capd) + al((size_t)cap) + al(sizeof(OcShared));
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

int tk_ocsort_destroy(void* handle) {
    if (!handle) return TK_ERR_ARG;
    OcHandle* h = (OcHandle*)handle;
    cudaFree(h->state);
    if (h->cost) cudaFree(h->cost);
    delete h;
    return TK_OK;
}

}  // extern "C"
