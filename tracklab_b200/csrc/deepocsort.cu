// Deep OC-SORT association, whole video per launch, one CTA per video.
//
// Device restatement of
//   /root/reference/plugins/track/deep_oc_sort/ocsort.py:22-93,96-304,324-542     (helpers, KalmanBoxTracker, OCSort.update)
//   /root/reference/plugins/track/deep_oc_sort/association.py:202-212,263-360     (linear_assignment, adaptive weighting, associate)
//   /root/reference/plugins/track/deep_oc_sort/kalmanfilter.py:340-379,383-481,483-569 (predict, freeze, affine correction, ORU, update)
// and of the wrapper filter /root/reference/tracklab/wrappers/track/deep_oc_sort_api.py:63-67, for the default 8-d filter
// (new_kf_off = false). The in-tracker ReID forward and the camera-motion estimator are separate stages here: the kernel takes the
// per-detection embeddings (float32) and one 2x3 affine per frame, what `_get_features` / `CMCComputer.compute_affine` return.
//
// Same execution shape as ocsort.cu: one launch walks the frames of a video, state resident in L2, per-frame scratch in shared memory;
// the assignment is scipy's solver on lap's published extension (lsap_scipy.cuh, see doc_solve).
// The behaviour of the reference that decides ids is kept operation by operation (oracle/deepocsort_np.py q1..q6):
//   q1  linear_assignment keeps `[y[i], i] for i in x` for unassigned rows too -> pairs (detection of the last tracker, -1); NumPy's
//       negative indices re-validate them, so the last tracker is updated once more per unassigned detection. Indices are kept RAW
//       (possibly -1) in every list and wrapped on access, like the Python lists.
//   q2  last_observation and observations[age] are one array object: the affine correction warps the newest observation twice while
//       it is at most delta_t frames old. The ring of recent observations therefore aliases the entry of `last_age` to last_obs.
//   q3  frame_count stays 0; the reported confidence is the one of the detection that created the track.
//   q4  the ORU replay reads (x, y, w, h) as (x, y, s, r), runs with R = I4 / Q = I8, and the measurement noise of the real update
//       comes from the state before the replay.
//   q5  first round: plain IoU, VDC term multiplied by the class column.
//   q6  embedding EMA, appearance matrix and adaptive weights in float32; everything else float64.
#include "lap.cuh"          // LAP_MAX_COLS
#include "lsap_scipy.cuh"
#include "oc_boxes.cuh"
#include "tk_common.cuh"
#include "trackkern.h"

namespace {

using namespace tk;

// Optional per-phase cycle accounting (build with -DTK_PHASE_PROF), read with tk_debug_deepocsort_phases().
#ifdef TK_PHASE_PROF
__device__ unsigned long long g_doc_prof[32];
#define PH(k) do { __syncthreads(); if (threadIdx.x == 0) { const long long _t = clock64(); g_doc_prof[k] += (unsigned long long)(_t - ph_t0); ph_t0 = _t; } } while (0)
#else
#define PH(k) do { } while (0)
#endif

constexpr int DOC_THREADS = 256;
constexpr int DOC_LSAP_MAX = 512;   // rows + columns of the extended assignment problem
constexpr int DRING = 8;   // observations of ages age-delta_t .. age (delta_t <= 7)

struct DocParams {
    double det_thresh, iou_threshold, inertia, min_conf, w_emb, alpha_fixed, aw_param;
    int max_age, min_hits, delta_t, asso, embedding_off, cmc_off, aw_off, emb_dim;
};

struct DocDev {
    int* hdr;  // 0 frame_count (never incremented, q3), 1 next uid, 2 n_trk, 4 status, 5 n_free
    double *x, *P, *sx, *sP, *s_last, *last_entry, *last_obs, *vel, *ring_obs, *conf, *cls, *det_id;
    int *ring_age, *tsu, *uid, *streak, *age, *hist_len, *saved_n, *last_age, *list, *free_list;
    unsigned char *has_vel, *observed, *has_saved, *frozen_flag;
};

__host__ __device__ inline size_t doc_al(size_t x) { return (x + 15) & ~(size_t)15; }

#define DOC_FIELDS(X)                                                                                                        \
    X(x, double, 8) X(P, double, 64) X(sx, double, 8) X(sP, double, 64) X(s_last, double, 4) X(last_entry, double, 4)        \
    X(last_obs, double, 5) X(vel, double, 2) X(ring_obs, double, DRING * 5) X(conf, double, 1) X(cls, double, 1)              \
    X(det_id, double, 1) X(ring_age, int, DRING) X(tsu, int, 1) X(uid, int, 1) X(streak, int, 1) X(age, int, 1)              \
    X(hist_len, int, 1) X(saved_n, int, 1) X(last_age, int, 1) X(list, int, 1) X(free_list, int, 1)                           \
    X(has_vel, unsigned char, 1) X(observed, unsigned char, 1) X(has_saved, unsigned char, 1) X(frozen_flag, unsigned char, 1)

__host__ __device__ inline size_t doc_state_bytes(int cap) {
    size_t s = doc_al(8 * sizeof(int));
#define X(name, type, n) s += doc_al((size_t)cap * (n) * sizeof(type));
    DOC_FIELDS(X)
#undef X
    return s;
}

__host__ __device__ inline DocDev doc_carve(char* base, int cap) {
    DocDev d;
    char* p = base;
    d.hdr = (int*)p; p += doc_al(8 * sizeof(int));
#define X(name, type, n) d.name = (type*)p; p += doc_al((size_t)cap * (n) * sizeof(type));
    DOC_FIELDS(X)
#undef X
    return d;
}

// per-frame scratch of one video (global memory, L2 resident)
struct DocScratch {
    double *iou, *cost, *trk_box, *kobs, *last_snap, *alpha;
    float *emb, *rw, *cw;
    int *d_idx, *un_d, *un_t, *p0, *p1, *m0, *m1, *match, *gd, *gt, *tmp;
    unsigned char* flag_t;
};

// The unmatched lists of the second round keep raw duplicates (q1): up to 2 capd detections x (cap + capd) trackers.
__host__ __device__ inline size_t doc_mat(int cap, int capd) { return (size_t)(2 * capd + 2) * (cap + capd + 2); }

// Small per-frame arrays always live in shared memory; the three matrices (iou / cost float64, emb float32) do when the frame's
// problem fits the rest of the CTA's shared memory (B200: ~800-cycle L2 round trips dominated the per-frame latency when every
// phase went through global scratch), else in the global scratch of the video.
__host__ __device__ inline size_t doc_small_bytes(int cap, int capd) {
    const size_t lst = (size_t)2 * (cap + capd) + 8;
    size_t s = 0;
    s += doc_al((size_t)cap * 4 * sizeof(double)) + 2 * doc_al((size_t)cap * 5 * sizeof(double)) + doc_al((size_t)capd * sizeof(double));
    s += doc_al((size_t)capd * sizeof(float)) + doc_al((size_t)cap * sizeof(float));
    s += 11 * doc_al(lst * sizeof(int));
    s += doc_al((size_t)cap);
    return s;
}

__host__ __device__ inline size_t doc_scratch_bytes(int cap, int capd) {      // global fallback for the matrices
    const size_t mat = doc_mat(cap, capd);
    return 2 * doc_al(mat * sizeof(double)) + doc_al(mat * sizeof(float));
}

__device__ inline DocScratch doc_scratch_carve(char* small, char* big, int cap, int capd) {
    DocScratch d;
    const size_t mat = doc_mat(cap, capd);
    const size_t lst = (size_t)2 * (cap + capd) + 8;
    char* p = big;
    d.iou = (double*)p; p += doc_al(mat * sizeof(double));
    d.cost = (double*)p; p += doc_al(mat * sizeof(double));
    d.emb = (float*)p;
    p = small;
    d.trk_box = (double*)p; p += doc_al((size_t)cap * 4 * sizeof(double));
    d.kobs = (double*)p; p += doc_al((size_t)cap * 5 * sizeof(double));
    d.last_snap = (double*)p; p += doc_al((size_t)cap * 5 * sizeof(double));
    d.alpha = (double*)p; p += doc_al((size_t)capd * sizeof(double));
    d.rw = (float*)p; p += doc_al((size_t)capd * sizeof(float));
    d.cw = (float*)p; p += doc_al((size_t)cap * sizeof(float));
    int** lists[11] = {&d.d_idx, &d.un_d, &d.un_t, &d.p0, &d.p1, &d.m0, &d.m1, &d.match, &d.gd, &d.gt, &d.tmp};
    for (int i = 0; i < 11; ++i) { *lists[i] = (int*)p; p += doc_al(lst * sizeof(int)); }
    d.flag_t = (unsigned char*)p;
    return d;
}

// matrices of an (rows x cols) problem: shared memory when they fit `sm_entries` (20 bytes per entry), else the global scratch
__device__ __forceinline__ void doc_place_matrices(DocScratch& W, const DocScratch& G, char* sm_mat, size_t sm_entries, size_t entries) {
    if (entries <= sm_entries) {
        W.iou = (double*)sm_mat; W.cost = W.iou + sm_entries; W.emb = (float*)(W.cost + sm_entries);
    } else { W.iou = G.iou; W.cost = G.cost; W.emb = G.emb; }
}

// ---- 8-d filter (ocsort.py:82-93, kalmanfilter.py:340-379, 531-569) --------------------------------------------------------------
__device__ __forceinline__ void process_noise_diag(double w, double h, double* q) {   // ocsort.py:82-86
    const double p = 1.0 / 20, v = 1.0 / 160;
    const double pw = __dmul_rn(p, w), ph = __dmul_rn(p, h), vw = __dmul_rn(v, w), vh = __dmul_rn(v, h);
    q[0] = __dmul_rn(pw, pw); q[1] = __dmul_rn(ph, ph); q[2] = q[0]; q[3] = q[1];
    q[4] = __dmul_rn(vw, vw); q[5] = __dmul_rn(vh, vh); q[6] = q[4]; q[7] = q[5];
}

// x = F x, P = F P F^T + diag(q) with F = [[I, I], [0, I]]; every sum has one rounding like the dense products of the reference
__device__ void kf8_predict(double* x, double* P, const double* q) {
    for (int i = 0; i < 4; ++i) x[i] = __dadd_rn(x[i], x[i + 4]);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 8; ++j) P[i * 8 + j] = __dadd_rn(P[i * 8 + j], P[(i + 4) * 8 + j]);       // F P
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) P[i * 8 + j] = __dadd_rn(P[i * 8 + j], P[i * 8 + j + 4]);         // (F P) F^T
    for (int i = 0; i < 8; ++i) P[i * 9] = __dadd_rn(P[i * 9], q[i]);
}

// inverse of a general 4x4 (LU with partial pivoting, like LAPACK getrf/getri behind np.linalg.inv); false when singular
__device__ bool inv4(const double* A, double* inv) {
    double a[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { a[i][j] = A[i * 4 + j]; a[i][j + 4] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        double best = fabs(a[c][c]);
        for (int r = c + 1; r < 4; ++r) if (fabs(a[r][c]) > best) { best = fabs(a[r][c]); piv = r; }
        if (!(best > 0.0)) return false;
        if (piv != c) for (int j = 0; j < 8; ++j) { const double t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
        const double d = a[c][c];
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            const double f = a[r][c] / d;
            if (f != 0.0) for (int j = c; j < 8; ++j) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) inv[i * 4 + j] = a[i][j + 4] / a[i][i];
    return true;
}

// big_m = kron(I4, m): x <- big_m x (+ t on the first two), P <- big_m P big_m^T   (kalmanfilter.py:393-397)
__device__ void kf8_affine(double* x, double* P, const double* A) {
    const double m00 = A[0], m01 = A[1], t0 = A[2], m10 = A[3], m11 = A[4], t1 = A[5];
    for (int b = 0; b < 4; ++b) {
        const double u = x[2 * b], v = x[2 * b + 1];
        x[2 * b] = m00 * u + m01 * v;
        x[2 * b + 1] = m10 * u + m11 * v;
    }
    x[0] += t0; x[1] += t1;
    for (int bi = 0; bi < 4; ++bi)          // rows: big_m P
        for (int j = 0; j < 8; ++j) {
            const double u = P[(2 * bi) * 8 + j], v = P[(2 * bi + 1) * 8 + j];
            P[(2 * bi) * 8 + j] = m00 * u + m01 * v;
            P[(2 * bi + 1) * 8 + j] = m10 * u + m11 * v;
        }
    for (int i = 0; i < 8; ++i)             // columns: (.) big_m^T
        for (int bj = 0; bj < 4; ++bj) {
            const double u = P[i * 8 + 2 * bj], v = P[i * 8 + 2 * bj + 1];
            P[i * 8 + 2 * bj] = u * m00 + v * m01;
            P[i * 8 + 2 * bj + 1] = u * m10 + v * m11;
        }
}

__device__ __forceinline__ void warp_points(double* b, const double* A) {   // ocsort.py:256-259: both corners through m, t
    const double x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
    b[0] = A[0] * x1 + A[1] * y1 + A[2]; b[1] = A[3] * x1 + A[4] * y1 + A[5];
    b[2] = A[0] * x2 + A[1] * y2 + A[2]; b[3] = A[3] * x2 + A[4] * y2 + A[5];
}

__device__ __forceinline__ double sum5(const double* b) {   // ndarray.sum() of 5 values: sequential from 0
    return __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(0.0, b[0]), b[1]), b[2]), b[3]), b[4]);
}

__device__ __forceinline__ void box_to_z8(const double* b, double* z) {   // ocsort.py:48-53
    const double w = b[2] - b[0], h = b[3] - b[1];
    z[0] = b[0] + w / 2.0; z[1] = b[1] + h / 2.0; z[2] = w; z[3] = h;
}

__device__ __forceinline__ void x_to_box8(const double* x, double* b) {   // ocsort.py:56-58
    b[0] = x[0] - x[2] / 2; b[1] = x[1] - x[3] / 2; b[2] = x[0] + x[2] / 2; b[3] = x[1] + x[3] / 2;
}

// observation of age `a` of slot s (dict lookup `a in observations`), nullptr when absent; the newest entry lives in last_obs (q2)
__device__ __forceinline__ double* obs_at(DocDev& S, int s, int a) {
    if (a < 0) return nullptr;
    if (S.last_age[s] == a) return S.last_obs + (size_t)s * 5;
    const int r = a % DRING;
    if (S.ring_age[(size_t)s * DRING + r] == a) return S.ring_obs + ((size_t)s * DRING + r) * 5;
    return nullptr;
}

// KalmanBoxTracker.apply_affine_correction (ocsort.py:252-271) + KalmanFilterNew.apply_affine_correction (kalmanfilter.py:388-405)
__device__ void doc_track_affine(DocDev& S, int s, const double* A, int delta_t) {
    double* lo = S.last_obs + (size_t)s * 5;
    if (sum5(lo) > 0) warp_points(lo, A);
    const int age = S.age[s];
    for (int dt = delta_t; dt >= 0; --dt) {
        double* o = obs_at(S, s, age - dt);
        if (o) warp_points(o, A);                 // the entry of last_age IS last_obs: warped a second time (q2)
    }
    kf8_affine(S.x + (size_t)s * 8, S.P + (size_t)s * 64, A);
    if (!S.observed[s] && S.has_saved[s]) {
        kf8_affine(S.sx + (size_t)s * 8, S.sP + (size_t)s * 64, A);
        double* lm = S.s_last + (size_t)s * 4;
        const double a0 = lm[0], a1 = lm[1], a2 = lm[2], a3 = lm[3];
        lm[0] = A[0] * a0 + A[1] * a1 + A[2]; lm[1] = A[3] * a0 + A[4] * a1 + A[5];
        lm[2] = A[0] * a2 + A[1] * a3; lm[3] = A[3] * a2 + A[4] * a3;
    }
}

// ---- warp-cooperative filter arithmetic: x[8], P[64] and the temporaries live in a per-warp shared buffer, every lane owns two
// entries of the 8x8 products. A single thread doing these products keeps ~200 doubles live in local memory (335 us per frame for
// ~36 tracks, measured with -DTK_PHASE_PROF); the cooperative form takes a few microseconds.
struct KfBuf { double x[8], P[64], T[64], K[32], S[16], SI[16], y[4]; int ok; };

__device__ void kf8_predict_warp(KfBuf& B, const double* q) {
    const int l = lane_id();
    __syncwarp();
    for (int e = l; e < 64; e += 32) {
        const int i = e >> 3, c = e & 7;
        double a = B.P[e];
        if (i < 4) a = __dadd_rn(a, B.P[(i + 4) * 8 + c]);                       // (F P)[i][c]
        if (c < 4) {
            double b2 = B.P[i * 8 + c + 4];
            if (i < 4) b2 = __dadd_rn(b2, B.P[(i + 4) * 8 + c + 4]);             // (F P)[i][c+4]
            a = __dadd_rn(a, b2);
        }
        if (i == c) a = __dadd_rn(a, q[i]);
        B.T[e] = a;
    }
    double xn = 0.0;
    if (l < 8) xn = l < 4 ? __dadd_rn(B.x[l], B.x[l + 4]) : B.x[l];
    __syncwarp();
    for (int e = l; e < 64; e += 32) B.P[e] = B.T[e];
    if (l < 8) B.x[l] = xn;
    __syncwarp();
}

__device__ bool kf8_correct_warp(KfBuf& B, const double* z, const double* rdiag) {
    const int l = lane_id();
    __syncwarp();
    if (l < 16) { const int a = l >> 2, b = l & 3; B.S[l] = B.P[a * 8 + b] + (a == b ? rdiag[a] : 0.0); }
    if (l < 4) B.y[l] = __dsub_rn(z[l], B.x[l]);
    __syncwarp();
    if (l == 0) B.ok = inv4(B.S, B.SI) ? 1 : 0;
    __syncwarp();
    if (!B.ok) return false;
    {   // K = P H^T S^-1: lane -> (i, j)
        const int i = l >> 2, j = l & 3;
        double acc = 0.0;
        for (int k = 0; k < 4; ++k) acc += B.P[i * 8 + k] * B.SI[k * 4 + j];
        B.K[l] = acc;
    }
    __syncwarp();
    double xn = 0.0;
    if (l < 8) { double acc = 0.0; for (int k = 0; k < 4; ++k) acc += B.K[l * 4 + k] * B.y[k]; xn = B.x[l] + acc; }
    for (int e = l; e < 64; e += 32) {          // T = (I - K H) P
        const int i = e >> 3, c = e & 7;
        double acc = B.P[e];
        for (int k = 0; k < 4; ++k) acc -= B.K[i * 4 + k] * B.P[k * 8 + c];
        B.T[e] = acc;
    }
    __syncwarp();
    if (l < 8) B.x[l] = xn;
    for (int e = l; e < 64; e += 32) {          // P = T (I - K H)^T + K R K^T
        const int i = e >> 3, c = e & 7;
        double acc = B.T[e], krk = 0.0;
        for (int k = 0; k < 4; ++k) acc -= B.T[i * 8 + k] * B.K[c * 4 + k];
        for (int k = 0; k < 4; ++k) krk += B.K[i * 4 + k] * rdiag[k] * B.K[c * 4 + k];
        B.P[e] = acc + krk;
    }
    __syncwarp();
    return true;
}

// KalmanBoxTracker.update(bbox) (ocsort.py:203-238) incl. KalmanFilterNew.update with the ORU replay (kalmanfilter.py:432-481, 483-569).
// One warp: lane 0 does the book-keeping, all lanes the filter arithmetic.
__device__ void doc_track_update(DocDev& S, int s, const double* bbox5, double cls, double det_id, int delta_t, int* status, KfBuf& B) {
    const int l = lane_id();
    double* lo = S.last_obs + (size_t)s * 5;
    if (l == 0) {
        const int age = S.age[s];
        S.frozen_flag[s] = 0;
        S.cls[s] = cls;
        if (sum5(lo) >= 0) {
            const double* prev = nullptr;
            for (int dt = delta_t; dt >= 1; --dt) { prev = obs_at(S, s, age - dt); if (prev) break; }
            if (!prev) prev = lo;
            const double cx1 = (prev[0] + prev[2]) / 2.0, cy1 = (prev[1] + prev[3]) / 2.0;
            const double cx2 = (bbox5[0] + bbox5[2]) / 2.0, cy2 = (bbox5[1] + bbox5[3]) / 2.0;
            const double dy = cy2 - cy1, dx = cx2 - cx1;
            const double norm = sqrt(__dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx))) + 1e-6;
            S.vel[(size_t)s * 2] = dy / norm; S.vel[(size_t)s * 2 + 1] = dx / norm; S.has_vel[s] = 1;
        }
        // last_observation = bbox; observations[age] = bbox (one object). The previous newest observation keeps living in the dict.
        const int la = S.last_age[s];
        if (la >= 0 && la != age) {
            const int r = la % DRING;
            S.ring_age[(size_t)s * DRING + r] = la;
            for (int i = 0; i < 5; ++i) S.ring_obs[((size_t)s * DRING + r) * 5 + i] = lo[i];
        }
        for (int i = 0; i < 5; ++i) lo[i] = bbox5[i];
        S.last_age[s] = age;
        S.tsu[s] = 0;
        S.streak[s] += 1;
        S.hist_len[s] += 1;                                     // history_obs.append(z)
        S.det_id[s] = det_id;
    }
    __syncwarp();
    double z[4], rdiag[4];
    box_to_z8(bbox5, z);
    {   // R from the state BEFORE the replay (q4)  ocsort.py:234
        const double m = 1.0 / 20, mw = __dmul_rn(m, S.x[(size_t)s * 8 + 2]), mh = __dmul_rn(m, S.x[(size_t)s * 8 + 3]);
        rdiag[0] = __dmul_rn(mw, mw); rdiag[1] = __dmul_rn(mh, mh); rdiag[2] = rdiag[0]; rdiag[3] = rdiag[1];
    }
    bool ok = true;
    const bool replay = !S.observed[s] && S.has_saved[s];
    if (replay) {               // unfreeze (kalmanfilter.py:432-481)
        const int full_len = S.hist_len[s], n = S.saved_n[s];
        if (l < 8) B.x[l] = S.sx[(size_t)s * 8 + l];
        for (int e = l; e < 64; e += 32) B.P[e] = S.sP[(size_t)s * 64 + e];
        const double* lm = S.s_last + (size_t)s * 4;
        const double x1 = lm[0], y1 = lm[1], w1 = sqrt(lm[2] * lm[3]), h1 = sqrt(lm[2] / lm[3]);
        const double x2 = z[0], y2 = z[1], w2 = sqrt(z[2] * z[3]), h2 = sqrt(z[2] / z[3]);
        const int gap = (full_len - 1) - (n - 2);
        const double g = (double)gap;
        const double dx = (x2 - x1) / g, dy = (y2 - y1) / g, dw = (w2 - w1) / g, dh = (h2 - h1) / g;
        const double one8[8] = {1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
        double vz[4] = {z[0], z[1], z[2], z[3]};
        for (int i = 0; i < gap; ++i) {
            const double k = (double)(i + 1);
            const double ww = w1 + k * dw, hh = h1 + k * dh;
            vz[0] = x1 + k * dx; vz[1] = y1 + k * dy; vz[2] = ww * hh; vz[3] = ww / hh;
            ok = kf8_correct_warp(B, vz, one8) && ok;
            if (i != gap - 1) kf8_predict_warp(B, one8);
        }
        if (l == 0) {
            S.hist_len[s] = (n - 1) + gap;
            for (int i = 0; i < 4; ++i) S.last_entry[(size_t)s * 4 + i] = vz[i];      // history_obs[-1] is the last virtual box
            S.has_saved[s] = 0;
        }
    } else {
        if (l < 8) B.x[l] = S.x[(size_t)s * 8 + l];
        for (int e = l; e < 64; e += 32) B.P[e] = S.P[(size_t)s * 64 + e];
        if (l < 4) S.last_entry[(size_t)s * 4 + l] = z[l];
    }
    ok = kf8_correct_warp(B, z, rdiag) && ok;
    if (!ok && l == 0) atomicOr(status, TK_DEV_BAD_CHOLESKY);
    if (l < 8) S.x[(size_t)s * 8 + l] = B.x[l];
    for (int e = l; e < 64; e += 32) S.P[(size_t)s * 64 + e] = B.P[e];
    if (l == 0) S.observed[s] = 1;
    __syncwarp();
}

// KalmanBoxTracker.update(None) (ocsort.py:239-241): freeze on the observed -> unobserved transition (kalmanfilter.py:506-518)
__device__ void doc_track_miss(DocDev& S, int s) {
    S.hist_len[s] += 1;
    if (S.observed[s]) {
        for (int i = 0; i < 8; ++i) S.sx[(size_t)s * 8 + i] = S.x[(size_t)s * 8 + i];
        for (int i = 0; i < 64; ++i) S.sP[(size_t)s * 64 + i] = S.P[(size_t)s * 64 + i];
        for (int i = 0; i < 4; ++i) S.s_last[(size_t)s * 4 + i] = S.last_entry[(size_t)s * 4 + i];    // history_obs[-2]
        S.saved_n[s] = S.hist_len[s];
        S.has_saved[s] = 1;
    }
    S.observed[s] = 0;
    S.frozen_flag[s] = 1;
}

// update_emb (ocsort.py:246-248), float32 (q6): emb = f32(alpha) * emb + f32(1 - alpha) * det, emb /= |emb|. One warp per call.
__device__ void doc_update_emb_warp(float* emb, const float* det, double alpha, int E) {
    const float a = (float)alpha, b = (float)(1.0 - alpha);
    float ss = 0.0f;
    for (int e = lane_id(); e < E; e += 32) {
        const float v = __fadd_rn(__fmul_rn(a, emb[e]), __fmul_rn(b, det[e]));
        emb[e] = v;
        ss = fmaf(v, v, ss);
    }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float n = sqrtf(ss);
    for (int e = lane_id(); e < E; e += 32) emb[e] = __fdiv_rn(emb[e], n);
    __syncwarp();
}

struct DocShared {
    int nd, nt, n_ud, n_ut, n_pairs, n_match, maxflag, lap_ok, n_gd, out_n;
    unsigned long long dmax_bits, cmax_key;
    double u[DOC_LSAP_MAX], v[DOC_LSAP_MAX], spc[DOC_LSAP_MAX];
    int path[DOC_LSAP_MAX], col4row[DOC_LSAP_MAX], row4col[DOC_LSAP_MAX], remaining[DOC_LSAP_MAX];
    unsigned char SR[DOC_LSAP_MAX], SC[DOC_LSAP_MAX];
};

// lap.lapjv(cost, extend_cost=True) of association.py:206 on the nd x nt matrix C (row-major): lap 0.5.12 is not vendored, so - like
// the goldens' stand-in (oracle/ref_shims/lap) - its published extension is solved: the (nd+nt)^2 matrix with C in the top-left block,
// cost.max() + 1 in the off-diagonal blocks and 0 in the bottom-right block, by scipy's solver INCLUDING its tie-breaking
// (lsap_scipy.cuh). Ties are real here: all zero-overlap pairs of tracks without a velocity cost exactly 0, which of them receives the
// discarded dummy pair orders `unmatched_detections`, i.e. the birth order, i.e. which tracker is `trackers[-1]` for q1.
// match[d] = t or -1; *ylast = detection assigned to the last tracker or -1. `cmax` = C.max().
__device__ void doc_solve(const double* C, int nd, int nt, double cmax, int* match, int* ylast, DocShared* sh, int* status) {
    if (nd + nt > DOC_LSAP_MAX) {     // cannot happen for cap_tracks + cap_dets <= DOC_LSAP_MAX unless the second-round lists are full of duplicates
        if (threadIdx.x == 0) { atomicOr(status, TK_DEV_OVERFLOW_ASSIGN); *ylast = -1; }
        for (int d = threadIdx.x; d < nd; d += blockDim.x) match[d] = -1;
        __syncthreads();
        return;
    }
    if (warp_id() == 0) {
        const int N = nd + nt;
        const double fill = cmax + 1.0;
        const bool ok = lsap_scipy_warp(N, N, [&](int i, int j) {
            return (i < nd && j < nt) ? C[(size_t)i * nt + j] : ((i >= nd && j >= nt) ? 0.0 : fill);
        }, sh->u, sh->v, sh->spc, sh->path, sh->col4row, sh->row4col, sh->remaining, sh->SR, sh->SC);
        __syncwarp();
        if (!ok) { if (lane_id() == 0) atomicOr(status, TK_DEV_LAP_INFEASIBLE); }
        for (int d = lane_id(); d < nd; d += 32) match[d] = (ok && sh->col4row[d] < nt) ? sh->col4row[d] : -1;
        if (lane_id() == 0) *ylast = (ok && sh->row4col[nt - 1] < nd) ? sh->row4col[nt - 1] : -1;
    }
    __syncthreads();
}

__device__ __forceinline__ double key_to_double(unsigned long long k) {   // inverse of tk::ordered_key
    const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ int wrap(int i, int n) { return i < 0 ? i + n : i; }

// Apply (tracker, detection) updates given as raw index pairs, in list order per tracker (the same tracker may appear several times,
// q1): one warp per tracker walks the list; Kalman update and embedding EMA are warp-cooperative.
template <class Pair>
__device__ void doc_apply_updates(DocDev& S, const DocParams& prm, int n_pairs, Pair pair, int nt, int nd, const double* D,
                                  const int* d_idx, const float* det_embs, float* trk_embs, const double* alpha, int* status, KfBuf* bufs) {
    const int nw = blockDim.x >> 5;
    KfBuf& B = bufs[warp_id()];
    for (int k = warp_id(); k < nt; k += nw) {
        const int s = S.list[k];
        for (int i = 0; i < n_pairs; ++i) {
            int pd, pt;
            if (!pair(i, pd, pt) || wrap(pt, nt) != k) continue;
            const int d = wrap(pd, nd);
            const double* db = D + (size_t)d_idx[d] * 7;
            doc_track_update(S, s, db, db[5], db[6], prm.delta_t, status, B);
            if (!prm.embedding_off && det_embs)
                doc_update_emb_warp(trk_embs + (size_t)s * prm.emb_dim, det_embs + (size_t)d_idx[d] * prm.emb_dim, alpha[d], prm.emb_dim);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(DOC_THREADS)
deepocsort_video_kernel(DocParams prm, char* state_base, size_t state_stride, char* scratch_base, size_t scratch_stride,
                        float* trk_emb_base, int cap, int capd, const double* __restrict__ dets, const float* __restrict__ embs,
                        const double* __restrict__ affines, const int* __restrict__ offsets, int n_frames,
                        double* __restrict__ out_rows, const int* __restrict__ out_start, int* __restrict__ out_frame_count,
                        int* __restrict__ out_count, int out_capacity_rows, int sm_mat_entries) {
    __shared__ DocShared shs;
    __shared__ KfBuf kfbufs[DOC_THREADS / 32];
    DocShared* sh = &shs;
    const int seq = blockIdx.x, tid = threadIdx.x;
    DocDev S = doc_carve(state_base + (size_t)seq * state_stride, cap);
    extern __shared__ __align__(16) char doc_dsm[];
    const size_t small_bytes = doc_small_bytes(cap, capd);
    const DocScratch G = doc_scratch_carve(doc_dsm, scratch_base + (size_t)seq * scratch_stride, cap, capd);
    DocScratch W = G;
    char* sm_mat = doc_dsm + small_bytes;
    const size_t sm_entries = sm_mat_entries;
    float* trk_embs = trk_emb_base + (size_t)seq * cap * (prm.embedding_off ? 1 : prm.emb_dim);
    int* status = &S.hdr[4];
    const int F1 = n_frames + 1, E = prm.emb_dim;
    const int out_base = out_start[seq];
    if (tid == 0) sh->out_n = out_count[seq];
    __syncthreads();

#ifdef TK_PHASE_PROF
    long long ph_t0 = clock64();
#endif
    for (int f = 0; f < n_frames; ++f) {
        const int r0 = offsets[seq * F1 + f], r1 = offsets[seq * F1 + f + 1];
        const int nraw = r1 - r0;
        if (nraw == 0) { if (tid == 0) out_frame_count[seq * n_frames + f] = 0; continue; }   // deep_oc_sort_api.py:61-62
        if (nraw > capd) { if (tid == 0) atomicOr(status, TK_DEV_OVERFLOW_DETS); break; }
        const double* D = dets + (size_t)r0 * 7;
        const float* DE = embs ? embs + (size_t)r0 * E : nullptr;
        if (warp_id() == 0) {   // wrapper filter (deep_oc_sort_api.py:65) and det_thresh (ocsort.py:388-389), order kept
            const int n = warp_compact(nraw, 0, [&](int i) { const double c = D[i * 7 + 4]; return c > prm.min_conf && c > prm.det_thresh; },
                                       [&](int i, int p) { W.d_idx[p] = i; });
            if (lane_id() == 0) { sh->nd = n; sh->nt = S.hdr[2]; }
        }
        __syncthreads();
        const int nd = sh->nd;
        int nt = sh->nt;

        PH(0);
        // ---- CMC (ocsort.py:405-408) -----------------------------------------------------------------------------------------
        if (!prm.cmc_off && affines) {
            const double* A = affines + ((size_t)seq * n_frames + f) * 6;
            for (int k = tid; k < nt; k += DOC_THREADS) doc_track_affine(S, S.list[k], A, prm.delta_t);
            __syncthreads();
        }
        // dets_alpha (ocsort.py:410-413)
        for (int d = tid; d < nd; d += DOC_THREADS) {
            const double trust = (D[(size_t)W.d_idx[d] * 7 + 4] - prm.det_thresh) / (1 - prm.det_thresh);
            W.alpha[d] = prm.alpha_fixed + (1 - prm.alpha_fixed) * (1 - trust);
        }
        PH(1);
        // ---- predict (ocsort.py:273-299, 420-427) -----------------------------------------------------------------------------
        for (int k = tid; k < nt; k += DOC_THREADS) {
            const int s = S.list[k];
            double x[8], P[64], q[8], b[4];
            for (int i = 0; i < 8; ++i) x[i] = S.x[(size_t)s * 8 + i];
            for (int i = 0; i < 64; ++i) P[i] = S.P[(size_t)s * 64 + i];
            if (x[2] + x[6] <= 0) x[6] = 0;
            if (x[3] + x[7] <= 0) x[7] = 0;
            if (S.frozen_flag[s]) { x[6] = 0; x[7] = 0; }
            process_noise_diag(x[2], x[3], q);
            kf8_predict(x, P, q);
            for (int i = 0; i < 8; ++i) S.x[(size_t)s * 8 + i] = x[i];
            for (int i = 0; i < 64; ++i) S.P[(size_t)s * 64 + i] = P[i];
            S.age[s] += 1;
            if (S.tsu[s] > 0) S.streak[s] = 0;
            S.tsu[s] += 1;
            x_to_box8(x, b);
            for (int i = 0; i < 4; ++i) W.trk_box[4 * k + i] = b[i];
            W.flag_t[k] = (isnan(b[0]) || isnan(b[1]) || isnan(b[2]) || isnan(b[3])) ? 1 : 0;
        }
        if (tid == 0) sh->maxflag = 0;
        __syncthreads();
        for (int k = tid; k < nt; k += DOC_THREADS) if (W.flag_t[k]) sh->maxflag = 1;
        __syncthreads();
        if (tid == 0 && sh->maxflag) {   // trackers with a non-finite prediction are dropped (ocsort.py:424-436)
            int n = 0, nfree = S.hdr[5];
            for (int k = 0; k < nt; ++k) {
                const int s = S.list[k];
                if (W.flag_t[k]) { S.free_list[nfree++] = s; continue; }
                if (n != k) for (int i = 0; i < 4; ++i) W.trk_box[4 * n + i] = W.trk_box[4 * k + i];
                S.list[n++] = s;
            }
            S.hdr[5] = nfree; S.hdr[2] = n; sh->nt = n;
        }
        __syncthreads();
        nt = sh->nt;
        PH(2);
        // velocities / last_boxes snapshot / k_previous_obs (ocsort.py:438-440, 22-30)
        for (int k = tid; k < nt; k += DOC_THREADS) {
            const int s = S.list[k];
            for (int i = 0; i < 5; ++i) W.last_snap[5 * k + i] = S.last_obs[(size_t)s * 5 + i];
            const double* o = nullptr;
            if (S.last_age[s] >= 0) {
                for (int dt = prm.delta_t; dt >= 1; --dt) { o = obs_at(S, s, S.age[s] - dt); if (o) break; }
                if (!o) o = S.last_obs + (size_t)s * 5;       // observations[max key] = the newest observation
            }
            for (int i = 0; i < 5; ++i) W.kobs[5 * k + i] = o ? o[i] : -1.0;
        }
        __syncthreads();

        PH(3);
        // ---- first round (association.py:291-360) ---------------------------------------------------------------------------
        if (tid == 0) { sh->n_pairs = 0; sh->n_match = 0; sh->n_ud = 0; sh->n_ut = 0; }
        __syncthreads();
        if (nt == 0) {
            for (int d = tid; d < nd; d += DOC_THREADS) W.un_d[d] = d;
            if (tid == 0) sh->n_ud = nd;
            __syncthreads();
        } else if (nd > 0) {
            doc_place_matrices(W, G, sm_mat, sm_entries, (size_t)nd * nt);
            // IoU, thresholded row / column counts
            for (int k = tid; k < nd + nt; k += DOC_THREADS) W.tmp[k] = 0;
            __syncthreads();
            for (int e = tid; e < nd * nt; e += DOC_THREADS) {
                const int d = e / nt, t = e - d * nt;
                const double v = iou_plain(D + (size_t)W.d_idx[d] * 7, W.trk_box + 4 * t);
                W.iou[e] = v;
                if (v > prm.iou_threshold) { atomicAdd(&W.tmp[d], 1); atomicAdd(&W.tmp[nd + t], 1); }
            }
            if (tid == 0) sh->maxflag = 0;      // here: 1 when a row / column holds more than one candidate, 2 when none holds one
            __syncthreads();
            {
                int mx = 0;
                for (int k = tid; k < nd + nt; k += DOC_THREADS) mx = max(mx, W.tmp[k]);
                if (mx > 1) atomicOr(&sh->maxflag, 1);
                if (mx >= 1) atomicOr(&sh->maxflag, 4);
            }
            __syncthreads();
            // a.sum(1).max() == 1 and a.sum(0).max() == 1  <=> some candidate exists and no row / column has two (association.py:325-327)
            const bool shortcut = (sh->maxflag & 4) && !(sh->maxflag & 1);
            if (shortcut) {
                if (warp_id() == 0) {           // np.stack(np.where(a)): row-major order
                    const int n = warp_compact(nd * nt, 0, [&](int e) { return W.iou[e] > prm.iou_threshold; },
                                               [&](int e, int p) { W.p0[p] = e / nt; W.p1[p] = e % nt; });
                    if (lane_id() == 0) sh->n_pairs = n;
                }
                __syncthreads();
            } else {
                const bool use_emb = !prm.embedding_off && DE != nullptr;
                if (use_emb) {
                    for (int e = warp_id(); e < nd * nt; e += DOC_THREADS / 32) {    // dets_embs @ trk_embs.T (float32), zero where IoU <= 0
                        const int d = e / nt, t = e - d * nt;
                        float acc = 0.0f;
                        if (!(W.iou[e] <= 0)) {          // only the overlapping pairs need the dot product
                            const float* a = DE + (size_t)W.d_idx[d] * E;
                            const float* b = trk_embs + (size_t)S.list[t] * E;
                            for (int k = lane_id(); k < E; k += 32) acc = fmaf(a[k], b[k], acc);
                            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                        }
                        if (lane_id() == 0) W.emb[e] = acc;
                    }
                    __syncthreads();
                    if (!prm.aw_off) {          // compute_aw_max_metric (association.py:263-288), float32
                        const float bottom = (float)prm.aw_param, denom = (float)(1.0 - prm.aw_param);
                        for (int d = tid; d < nd; d += DOC_THREADS) {
                            float w = 1.0f;
                            if (nt >= 2) {
                                float m1 = -INFINITY, m2 = -INFINITY;
                                for (int t = 0; t < nt; ++t) { const float v = W.emb[d * nt + t]; if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) m2 = v; }
                                if (m1 == 0.0f) w = 0.0f;
                                else { const float r = __fsub_rn(__fdiv_rn(m2, m1), bottom); w = (r > 0.0f) ? __fsub_rn(1.0f, __fdiv_rn(r, denom)) : 1.0f; }
                            }
                            W.rw[d] = w;
                        }
                        for (int t = tid; t < nt; t += DOC_THREADS) {
                            float w = 1.0f;
                            if (nd >= 2) {
                                float m1 = -INFINITY, m2 = -INFINITY;
                                for (int d = 0; d < nd; ++d) { const float v = W.emb[d * nt + t]; if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) m2 = v; }
                                if (m1 == 0.0f) w = 0.0f;
                                else { const float r = __fsub_rn(__fdiv_rn(m2, m1), bottom); w = (r > 0.0f) ? __fsub_rn(1.0f, __fdiv_rn(r, denom)) : 1.0f; }
                            }
                            W.cw[t] = w;
                        }
                        __syncthreads();
                    }
                }
                // final_cost = -(iou + angle_diff_cost + emb_cost)
                const float w0 = (float)prm.w_emb;
                if (tid == 0) sh->cmax_key = 0ull;
                __syncthreads();
                for (int e = tid; e < nd * nt; e += DOC_THREADS) {
                    const int d = e / nt, t = e - d * nt;
                    const double* db = D + (size_t)W.d_idx[d] * 7;
                    const double* pv = W.kobs + 5 * t;
                    const int s = S.list[t];
                    // speed_direction_batch (association.py:215-225) and the angle term (:303-320)
                    const double cx1 = (db[0] + db[2]) / 2.0, cy1 = (db[1] + db[3]) / 2.0;
                    const double cx2 = (pv[0] + pv[2]) / 2.0, cy2 = (pv[1] + pv[3]) / 2.0;
                    const double dx = cx1 - cx2, dy = cy1 - cy2;
                    const double norm = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) + 1e-6;
                    const double X = dx / norm, Y = dy / norm;
                    const double iy = S.has_vel[s] ? S.vel[(size_t)s * 2] : 0.0, ix = S.has_vel[s] ? S.vel[(size_t)s * 2 + 1] : 0.0;
                    double c = __dadd_rn(__dmul_rn(ix, X), __dmul_rn(iy, Y));
                    c = fmin(fmax(c, -1.0), 1.0);
                    const double pi = 3.141592653589793;
                    const double ang = (pi / 2.0 - fabs(acos(c))) / pi;
                    const double valid = pv[4] < 0 ? 0.0 : 1.0;
                    const double vdc = __dmul_rn(__dmul_rn(__dmul_rn(valid, ang), prm.inertia), db[5]);      // x class column (q5)
                    double embc = 0.0;
                    if (use_emb) {
                        float w = prm.aw_off ? w0 : __fmul_rn(__fmul_rn(w0, W.rw[d]), W.cw[t]);
                        embc = (double)__fmul_rn(w, W.emb[e]);
                    }
                    const double v = -__dadd_rn(__dadd_rn(W.iou[e], vdc), embc);
                    W.cost[e] = v;
                    atomicMax(&sh->cmax_key, ordered_key(v));
                }
                __syncthreads();
                PH(10);
                doc_solve(W.cost, nd, nt, key_to_double(sh->cmax_key), W.match, &sh->lap_ok, sh, status);
                PH(11);
                if (tid == 0) {                 // [[y[i], i] for i in x] (q1)
                    const int ylast = sh->lap_ok;
                    for (int d = 0; d < nd; ++d) {
                        const int i = W.match[d];
                        W.p0[d] = i >= 0 ? d : ylast;
                        W.p1[d] = i >= 0 ? i : -1;
                    }
                    sh->n_pairs = nd;
                }
                __syncthreads();
            }
            // unmatched lists, then the IoU re-validation of every pair (association.py:338-357), serial order kept by one warp
            if (warp_id() == 0) {
                const int np = sh->n_pairs;
                for (int k = lane_id(); k < nd + nt; k += 32) W.tmp[k] = 0;
                __syncwarp();
                for (int i = lane_id(); i < np; i += 32) {
                    if (W.p0[i] >= 0) W.tmp[W.p0[i]] = 1;
                    if (W.p1[i] >= 0) W.tmp[nd + W.p1[i]] = 1;
                }
                __syncwarp();
                int nud = warp_compact(nd, 0, [&](int d) { return W.tmp[d] == 0; }, [&](int d, int p) { W.un_d[p] = d; });
                int nut = warp_compact(nt, 0, [&](int t) { return W.tmp[nd + t] == 0; }, [&](int t, int p) { W.un_t[p] = t; });
                auto low = [&](int i) { return W.iou[(size_t)wrap(W.p0[i], nd) * nt + wrap(W.p1[i], nt)] < prm.iou_threshold; };
                const int nud2 = warp_compact(np, nud, low, [&](int i, int p) { W.un_d[p] = W.p0[i]; });
                const int nut2 = warp_compact(np, nut, low, [&](int i, int p) { W.un_t[p] = W.p1[i]; });
                const int nm = warp_compact(np, 0, [&](int i) { return !low(i); }, [&](int i, int p) { W.m0[p] = W.p0[i]; W.m1[p] = W.p1[i]; });
                if (lane_id() == 0) { sh->n_ud = nud2; sh->n_ut = nut2; sh->n_match = nm; }
            }
            __syncthreads();
        } else {   // no detections: every tracker is unmatched (association.py:329-344)
            for (int t = tid; t < nt; t += DOC_THREADS) W.un_t[t] = t;
            if (tid == 0) sh->n_ut = nt;
            __syncthreads();
        }
        PH(4);
        // matched updates (ocsort.py:467-469)
        {
            const int nm = sh->n_match;
            doc_apply_updates(S, prm, nm, [&](int i, int& pd, int& pt) { pd = W.m0[i]; pt = W.m1[i]; return true; }, nt, nd, D, W.d_idx,
                              DE, trk_embs, W.alpha, status, kfbufs);
        }

        PH(5);
        // ---- second round: OCR on the last observations (ocsort.py:474-508) ---------------------------------------------------
        if (sh->n_ud > 0 && sh->n_ut > 0) {
            const int nud = sh->n_ud, nut = sh->n_ut;
            doc_place_matrices(W, G, sm_mat, sm_entries, (size_t)nud * nut);
            if (tid == 0) { sh->maxflag = 0; sh->dmax_bits = 0ull; sh->cmax_key = 0ull; }
            __syncthreads();
            if (prm.asso == TK_ASSO_CT_DIST) {
                for (int e = tid; e < nud * nut; e += DOC_THREADS) {
                    const double dd = centre_dist(D + (size_t)W.d_idx[wrap(W.un_d[e / nut], nd)] * 7, W.last_snap + 5 * wrap(W.un_t[e % nut], nt));
                    W.iou[e] = dd;
                    atomicMax(&sh->dmax_bits, (unsigned long long)__double_as_longlong(dd));
                }
                __syncthreads();
            }
            const double dmax = __longlong_as_double((long long)sh->dmax_bits);
            for (int e = tid; e < nud * nut; e += DOC_THREADS) {
                const int a = e / nut, b = e % nut;
                const double v = prm.asso == TK_ASSO_CT_DIST ? 1.0 - W.iou[e] / dmax
                                                             : asso_value(prm.asso, D + (size_t)W.d_idx[wrap(W.un_d[a], nd)] * 7, W.last_snap + 5 * wrap(W.un_t[b], nt));
                W.iou[e] = v;
                W.cost[e] = -v;
                atomicMax(&sh->cmax_key, ordered_key(-v));
                if (v > prm.iou_threshold) sh->maxflag = 1;
            }
            __syncthreads();
            if (sh->maxflag) {
                doc_solve(W.cost, nud, nut, key_to_double(sh->cmax_key), W.match, &sh->lap_ok, sh, status);
                if (tid == 0) {
                    const int ylast = sh->lap_ok;
                    int ng = 0;
                    for (int a = 0; a < nud; ++a) {
                        const int i = W.match[a];
                        const int q0 = i >= 0 ? a : ylast, q1 = i >= 0 ? i : -1;
                        W.p0[a] = W.un_d[wrap(q0, nud)];       // det_ind, trk_ind (raw values of the unmatched lists)
                        W.p1[a] = W.un_t[wrap(q1, nut)];
                        W.tmp[a] = (W.iou[(size_t)wrap(q0, nud) * nut + wrap(q1, nut)] < prm.iou_threshold) ? 0 : 1;
                        if (W.tmp[a]) { W.gd[ng] = W.p0[a]; W.gt[ng] = W.p1[a]; ++ng; }
                    }
                    sh->n_gd = ng;
                }
                __syncthreads();
                doc_apply_updates(S, prm, nud, [&](int i, int& pd, int& pt) { pd = W.p0[i]; pt = W.p1[i]; return W.tmp[i] != 0; }, nt, nd, D,
                                  W.d_idx, DE, trk_embs, W.alpha, status, kfbufs);
                if (tid == 0) {   // np.setdiff1d: sorted unique values of the list that are not in the removed set
                    const int ng = sh->n_gd;
                    for (int pass = 0; pass < 2; ++pass) {
                        int* lst = pass == 0 ? W.un_d : W.un_t;
                        const int* rem = pass == 0 ? W.gd : W.gt;
                        const int n = pass == 0 ? nud : nut;
                        int m = 0;
                        for (int i = 0; i < n; ++i) {
                            const int v = lst[i];
                            bool drop = false;
                            for (int j = 0; j < ng && !drop; ++j) drop = rem[j] == v;
                            for (int j = 0; j < m && !drop; ++j) drop = W.tmp[j] == v;
                            if (!drop) W.tmp[m++] = v;
                        }
                        for (int i = 1; i < m; ++i) { const int v = W.tmp[i]; int j = i - 1; while (j >= 0 && W.tmp[j] > v) { W.tmp[j + 1] = W.tmp[j]; --j; } W.tmp[j + 1] = v; }
                        for (int i = 0; i < m; ++i) lst[i] = W.tmp[i];
                        if (pass == 0) sh->n_ud = m; else sh->n_ut = m;
                    }
                }
                __syncthreads();
            }
        }

        PH(6);
        // ---- unmatched trackers: update(None) once per list entry (ocsort.py:510-511) -----------------------------------------
        {
            const int nut = sh->n_ut;
            for (int k = tid; k < nt; k += DOC_THREADS) {
                int c = 0;
                for (int i = 0; i < nut; ++i) c += (wrap(W.un_t[i], nt) == k);
                for (int i = 0; i < c; ++i) doc_track_miss(S, S.list[k]);
            }
            __syncthreads();
        }
        PH(7);
        // ---- births (ocsort.py:513-519) ---------------------------------------------------------------------------------------
        {
            const int nud = sh->n_ud;
            if (tid == 0) {
                int ntrk = S.hdr[2], nfree = S.hdr[5], uid = S.hdr[1];
                for (int i = 0; i < nud; ++i) {
                    if (nfree <= 0) { atomicOr(status, TK_DEV_OVERFLOW_TRACKS); W.tmp[i] = -1; continue; }
                    const int s = S.free_list[--nfree];
                    S.list[ntrk++] = s;
                    S.uid[s] = uid++;
                    W.tmp[i] = s;
                }
                S.hdr[2] = ntrk; S.hdr[5] = nfree; S.hdr[1] = uid;
            }
            __syncthreads();
            for (int i = warp_id(); i < nud; i += DOC_THREADS / 32) {
                const int s = W.tmp[i];
                if (s < 0) continue;
                const int d = wrap(W.un_d[i], nd);
                const double* db = D + (size_t)W.d_idx[d] * 7;
                if (lane_id() == 0) {
                    double z[4], q[8];
                    box_to_z8(db, z);
                    process_noise_diag(z[2], z[3], q);
                    for (int a = 0; a < 64; ++a) S.P[(size_t)s * 64 + a] = 0.0;
                    for (int a = 0; a < 8; ++a) S.P[(size_t)s * 64 + a * 9] = q[a] * (a < 4 ? 4.0 : 100.0);
                    for (int a = 0; a < 8; ++a) S.x[(size_t)s * 8 + a] = a < 4 ? z[a] : 0.0;
                    S.tsu[s] = 0; S.streak[s] = 0; S.age[s] = 0; S.hist_len[s] = 0; S.saved_n[s] = 0; S.last_age[s] = -1;
                    S.has_vel[s] = 0; S.observed[s] = 0; S.has_saved[s] = 0; S.frozen_flag[s] = 0;
                    for (int a = 0; a < 5; ++a) S.last_obs[(size_t)s * 5 + a] = -1.0;
                    for (int a = 0; a < DRING; ++a) S.ring_age[(size_t)s * DRING + a] = -1;
                    S.conf[s] = db[4]; S.cls[s] = db[5]; S.det_id[s] = db[6];
                }
                if (!prm.embedding_off && DE) for (int e = lane_id(); e < E; e += 32) trk_embs[(size_t)s * E + e] = DE[(size_t)W.d_idx[d] * E + e];
            }
            __syncthreads();
        }
        PH(8);
        // ---- output + removal (ocsort.py:520-540), reversed list order --------------------------------------------------------
        if (warp_id() == 0) {
            const int ntrk = S.hdr[2];
            const int frame_count = S.hdr[0];
            const int base = sh->out_n;
            const int n_out = warp_compact(ntrk, 0, [&](int r) {
                const int s = S.list[ntrk - 1 - r];
                return S.tsu[s] < 1 && (S.streak[s] >= prm.min_hits || frame_count <= prm.min_hits);
            }, [&](int r, int p) {
                const int s = S.list[ntrk - 1 - r];
                if (base + p >= out_capacity_rows) return;
                double* o = out_rows + (size_t)(out_base + base + p) * 8;
                const double* lo = S.last_obs + (size_t)s * 5;
                if (sum5(lo) < 0) { double b[4]; x_to_box8(S.x + (size_t)s * 8, b); o[0] = b[0]; o[1] = b[1]; o[2] = b[2]; o[3] = b[3]; }
                else { o[0] = lo[0]; o[1] = lo[1]; o[2] = lo[2]; o[3] = lo[3]; }
                o[4] = (double)(S.uid[s] + 1); o[5] = S.cls[s]; o[6] = S.conf[s]; o[7] = S.det_id[s];
            });
            if (lane_id() == 0) {
                if (base + n_out > out_capacity_rows) atomicOr(status, TK_DEV_OVERFLOW_OUT);
                out_frame_count[seq * n_frames + f] = n_out;
                sh->out_n = min(base + n_out, out_capacity_rows);
            }
            __syncwarp();
            // remove dead tracklets: ordered compaction of the survivors, freed slots back to the free list
            int nfree = S.hdr[5];
            const int nfree2 = warp_compact(ntrk, nfree, [&](int k) { return S.tsu[S.list[k]] > prm.max_age; },
                                            [&](int k, int p) { S.free_list[p] = S.list[k]; });
            // in-place ordered compaction is safe chunk by chunk only through a temporary
            const int keep = warp_compact(ntrk, 0, [&](int k) { return !(S.tsu[S.list[k]] > prm.max_age); }, [&](int k, int p) { W.tmp[p] = S.list[k]; });
            for (int k = lane_id(); k < keep; k += 32) S.list[k] = W.tmp[k];
            if (lane_id() == 0) { S.hdr[2] = keep; S.hdr[5] = nfree2; }
        }
        __syncthreads();
        PH(9);
    }
    if (tid == 0) out_count[seq] = sh->out_n;
}


struct DocHandle {
    DocParams prm;
    int n_seq, cap, capd;
    size_t state_stride, scratch_stride, smem_bytes;
    int sm_mat_entries;
    char *state, *scratch;
    float* trk_emb;
};

__global__ void deepocsort_reset_kernel(char* base, size_t stride, int cap) {
    DocDev S = doc_carve(base + (size_t)blockIdx.x * stride, cap);
    for (int i = threadIdx.x; i < 8; i += blockDim.x) S.hdr[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < cap; i += blockDim.x) { S.free_list[i] = cap - 1 - i; S.list[i] = -1; }
    if (threadIdx.x == 0) S.hdr[5] = cap;
}

}  // namespace

extern "C" {

int tk_deepocsort_create(const tk_deepocsort_params* p, int n_seq, int cap_tracks, int cap_dets, void** handle) {
    if (!p || !handle || n_seq <= 0 || cap_tracks <= 0 || cap_dets <= 0) return TK_ERR_ARG;
    if (cap_dets + cap_tracks > DOC_LSAP_MAX) return TK_ERR_CAPACITY;   // rows + columns of the extended assignment problem
    if (p->delta_t < 1 || p->delta_t >= DRING || p->asso_func < 0 || p->asso_func > TK_ASSO_CT_DIST) return TK_ERR_ARG;
    if (!p->embedding_off && p->feature_dim <= 0) return TK_ERR_ARG;
    DocHandle* h = new DocHandle();
    h->prm.det_thresh = p->det_thresh; h->prm.iou_threshold = p->iou_threshold; h->prm.inertia = p->inertia;
    h->prm.min_conf = p->min_confidence; h->prm.w_emb = p->w_association_emb; h->prm.alpha_fixed = p->alpha_fixed_emb;
    h->prm.aw_param = p->aw_param; h->prm.max_age = p->max_age; h->prm.min_hits = p->min_hits; h->prm.delta_t = p->delta_t;
    h->prm.asso = p->asso_func; h->prm.embedding_off = p->embedding_off; h->prm.cmc_off = p->cmc_off; h->prm.aw_off = p->aw_off;
    h->prm.emb_dim = p->embedding_off ? 1 : p->feature_dim;
    h->n_seq = n_seq; h->cap = cap_tracks; h->capd = cap_dets;
    h->state_stride = (doc_state_bytes(cap_tracks) + 255) & ~(size_t)255;
    h->scratch_stride = (doc_scratch_bytes(cap_tracks, cap_dets) + 255) & ~(size_t)255;
    h->state = nullptr; h->scratch = nullptr; h->trk_emb = nullptr;
    {   // shared memory: the small arrays + as many matrix entries (20 bytes each) as fit ~190 KB together with the static part
        const size_t small = doc_small_bytes(cap_tracks, cap_dets);
        const size_t budget = 190 * 1024 - sizeof(DocShared) - sizeof(KfBuf) * (DOC_THREADS / 32);
        if (small + 20 * 64 > budget) { delete h; return TK_ERR_CAPACITY; }
        size_t ent = (budget - small) / 20;
        const size_t mat = doc_mat(cap_tracks, cap_dets);
        if (ent > mat) ent = mat;
        ent &= ~(size_t)1;
        h->sm_mat_entries = (int)ent;
        h->smem_bytes = small + ent * 20;
    }
    cudaError_t e = cudaMalloc((void**)&h->state, h->state_stride * n_seq);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->scratch, h->scratch_stride * n_seq);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->trk_emb, sizeof(float) * (size_t)n_seq * cap_tracks * h->prm.emb_dim);
    if (e != cudaSuccess) {
        tk_set_last_cuda_error((int)e);
        if (h->state) cudaFree(h->state);
        if (h->scratch) cudaFree(h->scratch);
        if (h->trk_emb) cudaFree(h->trk_emb);
        delete h;
        return TK_ERR_CUDA;
    }
    *handle = h;
    return tk_deepocsort_reset(h, nullptr);
}

int tk_deepocsort_reset(void* handle, void* stream) {   // KalmanBoxTracker.count = 0 in every OCSort() (ocsort.py:365)
    if (!handle) return TK_ERR_ARG;
    DocHandle* h = (DocHandle*)handle;
    deepocsort_reset_kernel<<<h->n_seq, 128, 0, (cudaStream_t)stream>>>(h->state, h->state_stride, h->cap);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_deepocsort_run(void* handle, const double* dets, const float* embeddings, const double* affines, const int* offsets, int n_frames,
                      double* out_rows, const int* out_start, int* out_frame_count, int* out_count, int out_capacity_rows, void* stream) {
    if (!handle || !offsets || !out_rows || !out_start || !out_frame_count || !out_count || n_frames < 0 || out_capacity_rows < 0)
        return TK_ERR_ARG;
    DocHandle* h = (DocHandle*)handle;
    if (!h->prm.embedding_off && !embeddings) return TK_ERR_ARG;
    if (!h->prm.cmc_off && !affines) return TK_ERR_ARG;
    if (n_frames == 0) return TK_OK;
    // the attribute is per kernel function, not per handle: set it before every launch
    TK_CUDA_TRY(cudaFuncSetAttribute(deepocsort_video_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    deepocsort_video_kernel<<<h->n_seq, DOC_THREADS, h->smem_bytes, (cudaStream_t)stream>>>(
        h->prm, h->state, h->state_stride, h->scratch, h->scratch_stride, h->trk_emb, h->cap, h->capd, dets, embeddings, affines, offsets,
        n_frames, out_rows, out_start, out_frame_count, out_count, out_capacity_rows, h->sm_mat_entries);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_deepocsort_status(void* handle, int* status_host, void* stream) {
    if (!handle || !status_host) return TK_ERR_ARG;
    DocHandle* h = (DocHandle*)handle;
    for (int s = 0; s < h->n_seq; ++s)
        TK_CUDA_TRY(cudaMemcpyAsync(status_host + s, h->state + (size_t)s * h->state_stride + 4 * sizeof(int), sizeof(int),
                                    cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    TK_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return TK_OK;
}

#ifdef TK_PHASE_PROF
int tk_debug_deepocsort_phases(unsigned long long* host_out32, int reset) {
    cudaDeviceSynchronize();
    if (host_out32) cudaMemcpyFromSymbol(host_out32, g_doc_prof, sizeof(unsigned long long) * 32);
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_doc_prof, z, sizeof(z)); }
    return 0;
}
#endif

int tk_deepocsort_destroy(void* handle) {
    if (!handle) return TK_ERR_ARG;
    DocHandle* h = (DocHandle*)handle;
    cudaFree(h->state); cudaFree(h->scratch); cudaFree(h->trk_emb);
    delete h;
    return TK_OK;
}

}  // extern "C"
