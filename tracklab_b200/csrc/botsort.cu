// BoT-SORT association, whole video per launch, one CTA per video. Same execution shape and list logic as
// botsort.cu (the plugin is ByteTrack's state machine); what differs is restated below.
//
// Device restatement of
//   /root/reference/plugins/track/bot_sort/bot_sort.py:15-240,243-485,507-545   (STrack, BoTSORT.update, list helpers)
//   /root/reference/plugins/track/bot_sort/matching.py:37-48,72-89,127-195,198-233 (assignment, IoU / embedding distances, fusions)
//   /root/reference/plugins/track/bot_sort/kalman_filter.py:23-268               (xywh filter)
// and of the wrapper's per-frame filter /root/reference/tracklab/wrappers/track/bot_sort_api.py:63-65. The in-tracker ReID forward
// (`_get_features`) and the camera-motion estimator (`GMC.apply`) are separate stages: the kernel takes per-detection embeddings
// (float32, un-normalised like the backbone returns them) and one 2x3 warp per frame.
//   * first association = JDE fusion: lambda * cosine distance (scipy cdist: float64, sequential sums) + (1 - lambda) * squared
//     Mahalanobis distance of the xywh filter, infeasible above chi2inv95[4] (matching.py:165-176);
//   * xywh filter: noise from (w, h); a freshly initiated mean AND covariance are float32, the process noise of multi_predict is
//     evaluated in float32 only while the whole pool is still float32 (NumPy promotion), multi_gmc / update make a track float64;
//   * GMC: mean <- kron(I4, R) mean (+ t), cov <- R8 cov R8^T for the pool and the unconfirmed tracks (bot_sort.py:93-106,343-346);
//   * appearance: detections normalise their feature once, a track's smooth feature is the float32 EMA (alpha 0.9) of the
//     re-normalised detection features, re-normalised (bot_sort.py:43-51);
//   * unconfirmed tracks: min(fuse_score(IoU distance), cosine distance / 2 gated by appearance_thresh and proximity_thresh);
//   * detection boxes: centre form stored as tlwh (q1 of oracle/botsort_np.py), low-score boxes through tlbr_to_tlwh of a centre box.
// Class histogram (update_cls, bot_sort.py:53-71): reduced to "class of the last matched detection" - identical whenever all
// detections of a track share a class (TrackLab feeds one category).
#include "kf_xyah.cuh"
#include "lap.cuh"
#include "trackkern.h"

namespace {

using namespace tk;

constexpr int BO_THREADS = 256;   // 32 octets: one Kalman update per 8 lanes (kf_xyah.cuh)

// Optional per-phase cycle accounting (build with -DTK_PHASE_PROF): thread 0 accumulates clock64() deltas between
// the barriers of the frame loop into g_bo_prof[]; read back with tk_debug_botsort_phases().
#ifdef TK_PHASE_PROF
__device__ unsigned long long g_bo_prof[64];
#define PH(k) do { if (threadIdx.x == 0) { const long long _t = clock64(); g_bo_prof[k] += (unsigned long long)(_t - ph_t0); ph_t0 = _t; } } while (0)
#else
#define PH(k) do { } while (0)
#endif
enum : unsigned char { ST_NEW = 0, ST_TRACKED = 1, ST_LOST = 2, ST_REMOVED = 3 };

struct BoDev {
    // persistent per-sequence state in global memory
    int* hdr;              // [8]: frame_id, next_id, n_tracked, n_lost, status, n_free, -, -
    double* mean;          // [cap][8]
    double* cov;           // [cap][64]
    double* score;         // [cap]
    double* cls;           // [cap]
    double* det_id;        // [cap]
    int* track_id;         // [cap]
    int* frame_id;         // [cap]
    int* start_frame;      // [cap]
    unsigned char* state;      // [cap]
    unsigned char* activated;  // [cap]
    unsigned char* mean_f32;   // [cap]
    unsigned char* in_removed; // [cap]
    int* tracked;          // [cap]
    int* lost;             // [cap]
    int* free_list;        // [cap]
};

__host__ __device__ inline size_t bo_align(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline size_t bo_state_bytes(int cap) {
    size_t s = 0;
    s += bo_align(8 * sizeof(int));
    s += bo_align((size_t)cap * 8 * sizeof(double));
    s += bo_align((size_t)cap * 64 * sizeof(double));
    s += 3 * bo_align((size_t)cap * sizeof(double));
    s += 3 * bo_align((size_t)cap * sizeof(int));
    s += 4 * bo_align((size_t)cap);
    s += 3 * bo_align((size_t)cap * sizeof(int));
    return s;
}

__host__ __device__ inline BoDev bo_carve(char* base, int cap) {
    BoDev d;
    char* p = base;
    d.hdr = (int*)p; p += bo_align(8 * sizeof(int));
    d.mean = (double*)p; p += bo_align((size_t)cap * 8 * sizeof(double));
    d.cov = (double*)p; p += bo_align((size_t)cap * 64 * sizeof(double));
    d.score = (double*)p; p += bo_align((size_t)cap * sizeof(double));
    d.cls = (double*)p; p += bo_align((size_t)cap * sizeof(double));
    d.det_id = (double*)p; p += bo_align((size_t)cap * sizeof(double));
    d.track_id = (int*)p; p += bo_align((size_t)cap * sizeof(int));
    d.frame_id = (int*)p; p += bo_align((size_t)cap * sizeof(int));
    d.start_frame = (int*)p; p += bo_align((size_t)cap * sizeof(int));
    d.state = (unsigned char*)p; p += bo_align((size_t)cap);
    d.activated = (unsigned char*)p; p += bo_align((size_t)cap);
    d.mean_f32 = (unsigned char*)p; p += bo_align((size_t)cap);
    d.in_removed = (unsigned char*)p; p += bo_align((size_t)cap);
    d.tracked = (int*)p; p += bo_align((size_t)cap * sizeof(int));
    d.lost = (int*)p; p += bo_align((size_t)cap * sizeof(int));
    d.free_list = (int*)p;
    return d;
}

struct BoParams {
    double track_thresh, match_thresh, det_thresh, min_conf, proximity_thresh, appearance_thresh, lambda_;
    int max_time_lost, emb_dim;
};
constexpr double CHI2INV95_4 = 9.4877;   // kalman_filter.py:11-20

// ---- float32 box helpers: every operation is a single IEEE fp32 op (no FMA contraction) ----------
// STrack.tlwh / tlbr (byte_tracker.py:100-120) for a track whose mean is float32 or float64.
__device__ __forceinline__ void track_tlwh(const double* m, bool f32, double* out) {
    if (f32) {   // bot_sort.py:169-178 on a float32 mean: ret[:2] -= ret[2:] / 2
        const float x = (float)m[0], y = (float)m[1], w = (float)m[2], h = (float)m[3];
        out[0] = (double)__fsub_rn(x, __fdiv_rn(w, 2.0f));
        out[1] = (double)__fsub_rn(y, __fdiv_rn(h, 2.0f));
        out[2] = (double)w;
        out[3] = (double)h;
    } else {
        out[0] = m[0] - m[2] / 2;
        out[1] = m[1] - m[3] / 2;
        out[2] = m[2];
        out[3] = m[3];
    }
}

__device__ __forceinline__ void track_tlbr32(const double* m, bool f32, float* o) {
    double t[4];
    track_tlwh(m, f32, t);
    if (f32) {
        o[0] = (float)t[0]; o[1] = (float)t[1];
        o[2] = __fadd_rn((float)t[2], (float)t[0]);
        o[3] = __fadd_rn((float)t[3], (float)t[1]);
    } else {
        o[0] = (float)t[0]; o[1] = (float)t[1];
        o[2] = (float)(t[2] + t[0]);
        o[3] = (float)(t[3] + t[1]);
    }
}

// bbox_ious (matching.py:182-218): +1-pixel IoU in float32, returns the DISTANCE 1 - iou in float32
__device__ __forceinline__ float iou_dist_p1(const float* a, const float* b) {
    float ov = 0.0f;
    const float iw = __fadd_rn(__fsub_rn(fminf(a[2], b[2]), fmaxf(a[0], b[0])), 1.0f);
    if (iw > 0.0f) {
        const float ih = __fadd_rn(__fsub_rn(fminf(a[3], b[3]), fmaxf(a[1], b[1])), 1.0f);
        if (ih > 0.0f) {
            const float area_b = __fmul_rn(__fadd_rn(__fsub_rn(b[2], b[0]), 1.0f), __fadd_rn(__fsub_rn(b[3], b[1]), 1.0f));
            const float area_a = __fmul_rn(__fadd_rn(__fsub_rn(a[2], a[0]), 1.0f), __fadd_rn(__fsub_rn(a[3], a[1]), 1.0f));
            const float inter = __fmul_rn(iw, ih);
            const float ua = __fsub_rn(__fadd_rn(area_a, area_b), inter);
            ov = __fdiv_rn(inter, ua);
        }
    }
    return __fsub_rn(1.0f, ov);
}

// measurement (x, y, a, h) of a detection whose box is the float32 "tlwh" record (byte_tracker.py:124-131)
__device__ __forceinline__ void det_xyah(const float* b, double* z) {   // tlwh_to_xywh on the float32 record (bot_sort.py:209-218)
    z[0] = (double)__fadd_rn(b[0], __fdiv_rn(b[2], 2.0f));
    z[1] = (double)__fadd_rn(b[1], __fdiv_rn(b[3], 2.0f));
    z[2] = (double)b[2];
    z[3] = (double)b[3];
}

constexpr double W_POS = 1.0 / 20;
constexpr double W_VEL = 1.0 / 160;

// KalmanFilter.update (kalman_filter.py:194-226, project :126-153) and multi_predict (:155-192); q is float32 when the
// whole pool still carries float32 means (NumPy promotion of np.asarray([...float32 means...])).
// Octet-cooperative (8 lanes per track, kf_xyah.cuh), register resident.
__device__ __forceinline__ void bo_octet_update(BoDev& S, int slot, bool active, const float* detbox) {
    double z[4] = {0, 0, 0, 0}, r[4] = {1, 1, 1, 1};
    double* gm = S.mean + (size_t)(active ? slot : 0) * 8;
    double* gP = S.cov + (size_t)(active ? slot : 0) * 64;
    if (active) {
        det_xyah(detbox, z);
        const double sw = W_POS * gm[2], sh = W_POS * gm[3];
        r[0] = sw * sw; r[1] = sh * sh; r[2] = sw * sw; r[3] = sh * sh;
    }
    if (!kf8_octet_update(gm, gP, active, z, r)) atomicOr(&S.hdr[4], TK_DEV_BAD_CHOLESKY);
}

__device__ __forceinline__ void bo_octet_predict(BoDev& S, int slot, bool active, bool pool_f32) {
    const int j = threadIdx.x & 7;
    double* gm = S.mean + (size_t)(active ? slot : 0) * 8;
    double* gP = S.cov + (size_t)(active ? slot : 0) * 64;
    double qj = 0.0;
    bool zero_vh = false;
    if (active) {
        const double wh = gm[2 + (j & 1)];          // std order (w, h, w, h | w, h, w, h), kalman_filter.py:173-182
        const bool pos = j < 4;
        if (pool_f32) {
            const float sd = __fmul_rn(pos ? (float)W_POS : (float)W_VEL, (float)wh);
            qj = (double)__fmul_rn(sd, sd);
        } else {
            const double sd = (pos ? W_POS : W_VEL) * wh;
            qj = sd * sd;
        }
    }
    kf8_octet_predict(gm, gP, active, zero_vh, qj);
}

// KalmanFilter.initiate (kalman_filter.py:55-86): mean stays float32-valued, std are float32 products
__device__ void bo_kf_initiate(BoDev& S, int slot, const float* detbox) {
    double z[4];
    det_xyah(detbox, z);
    double* gm = S.mean + (size_t)slot * 8;
    double* gP = S.cov + (size_t)slot * 64;
    for (int i = 0; i < 4; ++i) { gm[i] = z[i]; gm[i + 4] = 0.0; }
    // every std is a NumPy float32 scalar, np.square of the list stays float32: the covariance is float32-valued too
    const float w = (float)z[2], h = (float)z[3];
    const float spw = __fmul_rn((float)(2 * W_POS), w), sph = __fmul_rn((float)(2 * W_POS), h);
    const float svw = __fmul_rn((float)(10 * W_VEL), w), svh = __fmul_rn((float)(10 * W_VEL), h);
    const double d[8] = {(double)__fmul_rn(spw, spw), (double)__fmul_rn(sph, sph), (double)__fmul_rn(spw, spw), (double)__fmul_rn(sph, sph),
                         (double)__fmul_rn(svw, svw), (double)__fmul_rn(svh, svh), (double)__fmul_rn(svw, svw), (double)__fmul_rn(svh, svh)};
    for (int i = 0; i < 64; ++i) gP[i] = 0.0;
    for (int i = 0; i < 8; ++i) gP[i * 9] = d[i];
    S.mean_f32[slot] = 1;
}

// STrack.multi_gmc (bot_sort.py:93-106): mean <- kron(I4, R) mean, mean[:2] += t, cov <- (R8 cov) R8^T. One thread per track.
__device__ void bo_gmc(double* m, double* P, const double* H) {
    const double r00 = H[0], r01 = H[1], t0 = H[2], r10 = H[3], r11 = H[4], t1 = H[5];
    for (int b = 0; b < 4; ++b) {
        const double u = m[2 * b], v = m[2 * b + 1];
        m[2 * b] = r00 * u + r01 * v;
        m[2 * b + 1] = r10 * u + r11 * v;
    }
    m[0] += t0; m[1] += t1;
    for (int bi = 0; bi < 4; ++bi)
        for (int j = 0; j < 8; ++j) {
            const double u = P[(2 * bi) * 8 + j], v = P[(2 * bi + 1) * 8 + j];
            P[(2 * bi) * 8 + j] = r00 * u + r01 * v;
            P[(2 * bi + 1) * 8 + j] = r10 * u + r11 * v;
        }
    for (int i = 0; i < 8; ++i)
        for (int bj = 0; bj < 4; ++bj) {
            const double u = P[i * 8 + 2 * bj], v = P[i * 8 + 2 * bj + 1];
            P[i * 8 + 2 * bj] = u * r00 + v * r01;
            P[i * 8 + 2 * bj + 1] = u * r10 + v * r11;
        }
}

// projected state of a track for the gating distance (kalman_filter.py:126-153,228-268): g[0..3] = H mean, g[4..19] = lower Cholesky
// factor of H P H^T + R (row-major), g[20..23] = 1 / diagonal. Returns false when it is not positive definite.
__device__ bool bo_gate_prepare(const double* m, const double* P, double* g) {
    const double sw = W_POS * m[2], sh = W_POS * m[3];
    const double r[4] = {sw * sw, sh * sh, sw * sw, sh * sh};
    double S[16], L[16];
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) S[a * 4 + b] = P[a * 8 + b] + (a == b ? r[a] : 0.0);
    bool ok = true;
    for (int i = 0; i < 16; ++i) L[i] = 0.0;
    for (int c = 0; c < 4; ++c) {
        double d = S[c * 4 + c];
        for (int k = 0; k < c; ++k) d -= L[c * 4 + k] * L[c * 4 + k];
        ok = ok && (d > 0.0);
        const double lcc = sqrt(d);
        L[c * 4 + c] = lcc;
        g[20 + c] = 1.0 / lcc;
        for (int i = c + 1; i < 4; ++i) {
            double t = S[i * 4 + c];
            for (int k = 0; k < c; ++k) t -= L[i * 4 + k] * L[c * 4 + k];
            L[i * 4 + c] = t / lcc;
        }
    }
    for (int i = 0; i < 4; ++i) g[i] = m[i];
    for (int i = 0; i < 16; ++i) g[4 + i] = L[i];
    return ok;
}

__device__ __forceinline__ double bo_gate_dist(const double* g, const double* z) {   // sum((L^-1 (z - mean))^2)
    double y[4], acc = 0.0;
    for (int i = 0; i < 4; ++i) {
        double t = z[i] - g[i];
        for (int k = 0; k < i; ++k) t -= g[4 + i * 4 + k] * y[k];
        y[i] = t / g[4 + i * 4 + i];
        acc = __dadd_rn(acc, __dmul_rn(y[i], y[i]));
    }
    return acc;
}

// scipy cdist(..., 'cosine') of two float32 feature rows in float64 (matching.py:127-145): 1 - u.v / (|u| |v|), clipped at 0
__device__ __forceinline__ double bo_dot64(const float* a, const float* b, int E) {
    double s = 0.0;
    for (int e = 0; e < E; ++e) s = __dadd_rn(s, __dmul_rn((double)a[e], (double)b[e]));
    return s;
}
// one warp per pair: coalesced loads, 32 float64 partial sums combined by a butterfly (scipy's sequential sum and this
// one differ by ~1 ulp of the dot product; the assignment costs are continuous, no decision hangs on that)
__device__ __forceinline__ double bo_cos_dist_warp(const float* tf, double tn, const float* df, double dn, int E) {
    double s = 0.0;
    for (int e = lane_id(); e < E; e += 32) s += (double)tf[e] * (double)df[e];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    double c = s / (tn * dn);
    if (fabs(c) > 1.0) c = copysign(1.0, c);
    return fmax(0.0, 1.0 - c);
}

// update_features on a matched track (bot_sort.py:43-51), float32, one warp: feat = det / |det| (the detection's feature was already
// normalised once at its creation), smooth = 0.9 smooth + 0.1 feat, smooth /= |smooth|
__device__ void bo_update_feat_warp(float* smooth, const float* det, int E) {
    const int lane = lane_id();
    float ss = 0.0f;
    for (int e = lane; e < E; e += 32) ss = fmaf(det[e], det[e], ss);
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float dn = sqrtf(ss);
    const float a = 0.9f, b = (float)(1.0 - 0.9);
    float s2 = 0.0f;
    for (int e = lane; e < E; e += 32) {
        const float f = __fdiv_rn(det[e], dn);
        const float v = __fadd_rn(__fmul_rn(a, smooth[e]), __fmul_rn(b, f));
        smooth[e] = v;
        s2 = fmaf(v, v, s2);
    }
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    const float sn = sqrtf(s2);
    for (int e = lane; e < E; e += 32) smooth[e] = __fdiv_rn(smooth[e], sn);
    __syncwarp();
}

struct BoShared {
    // sizes
    int lap_ok;
    int nd, nh, nl, npool, nconf, nunc, nrest, nleft;
    int n_udet1, n_utrk1, n_births, n_lostnow, all_f32;
    int n_tracked, n_lost;
};

// Solve one association. cost (already limit-reduced) is stored with the smaller side as rows and leading
// dimension lap_pitch(cols). On return match_a[i] = j or -1, match_b[j] = i or -1.
__device__ void solve_assignment(const double* C, int na, int nb, int* match_a, int* match_b,
                                 double* u, int* col4row, int* row4col, int* path, int* ok_flag, int* status) {
    for (int i = threadIdx.x; i < na; i += blockDim.x) match_a[i] = -1;
    for (int j = threadIdx.x; j < nb; j += blockDim.x) match_b[j] = -1;
    __syncthreads();
    if (na == 0 || nb == 0) return;
    const bool a_rows = na <= nb;
    const int nr = a_rows ? na : nb, nc = a_rows ? nb : na;
    const int ld = lap_pitch(nc);
    const bool ok = lap_solve_cta(C, ld, nr, nc, true, u, col4row, row4col, path, ok_flag);
    if (!ok) { if (threadIdx.x == 0) atomicOr(status, TK_DEV_LAP_INFEASIBLE); return; }
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        const int c = col4row[r];
        if (c >= 0 && C[(size_t)r * ld + c] < 0.0) {  // pairs at/above the limit carry cost 0
            if (a_rows) { match_a[r] = c; match_b[c] = r; }
            else { match_a[c] = r; match_b[r] = c; }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(BO_THREADS)
botsort_video_kernel(BoParams prm, char* state_base, size_t state_stride, int cap, int capd,
                       const double* __restrict__ dets, const int* __restrict__ offsets, int n_frames,
                       double* __restrict__ out_rows, const int* __restrict__ out_start,
                       int* __restrict__ out_frame_count, int* __restrict__ out_count,
                       double* cost_scratch, size_t cost_stride, int cost_in_smem,
                       const float* __restrict__ embs, const double* __restrict__ warps, float* tfeat_base, float* dfeat_base,
                       double* gate_base) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int seq = blockIdx.x;
    const int tid = threadIdx.x;
    BoDev S = bo_carve(state_base + (size_t)seq * state_stride, cap);

    // ---- shared-memory carve-up -----------------------------------------------------------------
    unsigned char* sp = smem_raw;
    auto take = [&](size_t bytes) { unsigned char* p = sp; sp += (bytes + 15) & ~(size_t)15; return p; };
    const int side = cap > capd ? cap : capd;
    double* lap_u = (double*)take(sizeof(double) * side);
    double* d_score = (double*)take(sizeof(double) * capd);
    double* d_cls = (double*)take(sizeof(double) * capd);
    double* d_id = (double*)take(sizeof(double) * capd);
    float* d_box = (float*)take(sizeof(float) * 4 * capd);     // (cx, cy, w, h) float32 == STrack._tlwh
    float* d_tlbr = (float*)take(sizeof(float) * 4 * capd);
    float* d_boxl = (float*)take(sizeof(float) * 4 * capd);    // low-score form: tlbr_to_tlwh of the centre box (bot_sort.py:389-390)
    float* d_tlbrl = (float*)take(sizeof(float) * 4 * capd);
    float* t_tlbr = (float*)take(sizeof(float) * 4 * cap);     // indexed by list position
    float* t_tlbr2 = (float*)take(sizeof(float) * 4 * cap);
    int* d_high = (int*)take(sizeof(int) * capd);
    int* d_low = (int*)take(sizeof(int) * capd);
    int* d_left = (int*)take(sizeof(int) * capd);
    int* pool = (int*)take(sizeof(int) * cap);
    int* unconf = (int*)take(sizeof(int) * cap);
    int* rest = (int*)take(sizeof(int) * cap);
    int* lostnow = (int*)take(sizeof(int) * cap);
    int* births = (int*)take(sizeof(int) * capd);              // raw det index of each birth
    int* birth_slot = (int*)take(sizeof(int) * capd);
    int* match_a = (int*)take(sizeof(int) * side);
    int* match_b = (int*)take(sizeof(int) * side);
    int* col4row = (int*)take(sizeof(int) * side);
    int* row4col = (int*)take(sizeof(int) * side);
    int* path = (int*)take(sizeof(int) * side);
    int* newlist = (int*)take(sizeof(int) * cap);
    int* out_pos = (int*)take(sizeof(int) * cap);
    unsigned char* dup_a = (unsigned char*)take(cap);
    unsigned char* dup_b = (unsigned char*)take(cap);
    unsigned char* in_tracked = (unsigned char*)take(cap);
    BoShared* sh = (BoShared*)take(sizeof(BoShared));
    // Book-keeping of the tracker (lists, per-slot flags/ids/scores) lives in shared memory for the whole launch:
    // the frame loop below is a chain of short serial list edits, and every global round trip in it is pure latency.
    // Only the Kalman means/covariances (576 B per track) stay in global memory (L1/L2 resident).
    const BoDev G = S;
    {
        int* m_hdr = (int*)take(8 * sizeof(int));
        double* m_score = (double*)take(sizeof(double) * cap);
        double* m_cls = (double*)take(sizeof(double) * cap);
        double* m_det = (double*)take(sizeof(double) * cap);
        int* m_tid = (int*)take(sizeof(int) * cap);
        int* m_fid = (int*)take(sizeof(int) * cap);
        int* m_sf = (int*)take(sizeof(int) * cap);
        int* m_trk = (int*)take(sizeof(int) * cap);
        int* m_lost = (int*)take(sizeof(int) * cap);
        int* m_free = (int*)take(sizeof(int) * cap);
        unsigned char* m_state = (unsigned char*)take(cap);
        unsigned char* m_act = (unsigned char*)take(cap);
        unsigned char* m_f32 = (unsigned char*)take(cap);
        unsigned char* m_rem = (unsigned char*)take(cap);
        if (tid < 8) m_hdr[tid] = G.hdr[tid];
        for (int i = tid; i < cap; i += BO_THREADS) {
            m_score[i] = G.score[i]; m_cls[i] = G.cls[i]; m_det[i] = G.det_id[i];
            m_tid[i] = G.track_id[i]; m_fid[i] = G.frame_id[i]; m_sf[i] = G.start_frame[i];
            m_trk[i] = G.tracked[i]; m_lost[i] = G.lost[i]; m_free[i] = G.free_list[i];
            m_state[i] = G.state[i]; m_act[i] = G.activated[i]; m_f32[i] = G.mean_f32[i]; m_rem[i] = G.in_removed[i];
        }
        S.hdr = m_hdr; S.score = m_score; S.cls = m_cls; S.det_id = m_det; S.track_id = m_tid; S.frame_id = m_fid;
        S.start_frame = m_sf; S.tracked = m_trk; S.lost = m_lost; S.free_list = m_free; S.state = m_state;
        S.activated = m_act; S.mean_f32 = m_f32; S.in_removed = m_rem;
        __syncthreads();
    }
    double* cost = cost_in_smem ? (double*)take(0) : cost_scratch + (size_t)seq * cost_stride;
    const int E = prm.emb_dim;
    float* tfeat = tfeat_base + (size_t)seq * cap * E;          // smooth feature of every track slot
    float* dfeat = dfeat_base + (size_t)seq * capd * E;         // normalised features of this frame's detections
    double* gate = gate_base + (size_t)seq * ((size_t)cap * 25 + capd + (size_t)cap * capd);  // per pool position: projected mean, Cholesky factor, feature norm
    double* dnorm = gate + (size_t)cap * 25;                    // |feature| of every detection (float64, for cdist)
    double* gdm = dnorm + capd;                                 // [npool x nh] squared Mahalanobis distances of the first association

    int* status = &S.hdr[4];
    const int F1 = n_frames + 1;
    int out_base = out_start[seq];
    int out_n = out_count[seq];

#ifdef TK_PHASE_PROF
    long long ph_t0 = clock64();
#endif
    for (int f = 0; f < n_frames; ++f) {
        const int r0 = offsets[seq * F1 + f], r1 = offsets[seq * F1 + f + 1];
        const int nraw = r1 - r0;
        if (nraw == 0) {  // byte_track_api.py:51-52: frames without detections never reach update()
            if (tid == 0) out_frame_count[seq * n_frames + f] = 0;
            continue;
        }
        if (nraw > capd) { if (tid == 0) atomicOr(status, TK_DEV_OVERFLOW_DETS); break; }

        // ---- A. detections: xyxy -> centre xywh (float64) -> float32 record (byte_tracker.py:174-203)
        for (int i = tid; i < nraw; i += BO_THREADS) {
            const double* d = dets + (size_t)(r0 + i) * 7;
            const double x1 = d[0], y1 = d[1], x2 = d[2], y2 = d[3];
            float* b = d_box + 4 * i;
            b[0] = (float)((x1 + x2) / 2); b[1] = (float)((y1 + y2) / 2);
            b[2] = (float)(x2 - x1); b[3] = (float)(y2 - y1);
            float* t = d_tlbr + 4 * i;  // STrack.tlbr of a detection: float32 tlwh + float32 adds
            t[0] = b[0]; t[1] = b[1]; t[2] = __fadd_rn(b[2], b[0]); t[3] = __fadd_rn(b[3], b[1]);
            d_score[i] = d[4]; d_cls[i] = d[5]; d_id[i] = d[6];
            const double cx = (x1 + x2) / 2, cy = (y1 + y2) / 2, w = x2 - x1, h = y2 - y1;
            float* bl = d_boxl + 4 * i;     // ret[2:] -= ret[:2] on (cx, cy, w, h) in float64, then float32
            bl[0] = (float)cx; bl[1] = (float)cy; bl[2] = (float)(w - cx); bl[3] = (float)(h - cy);
            float* tl = d_tlbrl + 4 * i;
            tl[0] = bl[0]; tl[1] = bl[1]; tl[2] = __fadd_rn(bl[2], bl[0]); tl[3] = __fadd_rn(bl[3], bl[1]);
        }
        // STrack.__init__ -> update_features: feat /= |feat| (float32), for the high-score detections (features_keep, bot_sort.py:305-313)
        for (int i = warp_id(); i < nraw; i += BO_THREADS / 32) {
            const double c = dets[(size_t)(r0 + i) * 7 + 4];
            if (!(c > prm.min_conf && c > prm.track_thresh)) continue;
            const float* src = embs + (size_t)(r0 + i) * E;
            float* dst = dfeat + (size_t)i * E;
            float ss = 0.0f;
            for (int e = lane_id(); e < E; e += 32) ss = fmaf(src[e], src[e], ss);
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            const float nrm = sqrtf(ss);
            for (int e = lane_id(); e < E; e += 32) dst[e] = __fdiv_rn(src[e], nrm);
        }
        __syncthreads();
        for (int i = tid; i < nraw; i += BO_THREADS) {
            const double c = d_score[i];
            if (c > prm.min_conf && c > prm.track_thresh) dnorm[i] = sqrt(bo_dot64(dfeat + (size_t)i * E, dfeat + (size_t)i * E, E));
        }
        __syncthreads();
        PH(1);
        if (warp_id() == 0) {
            const int lane = lane_id();
            if (lane == 0) S.hdr[0] += 1;
            // wrapper filter (byte_track_api.py:54) + score split (byte_tracker.py:186-203)
            const int nh_ = warp_compact(nraw, 0, [&](int i) { const double c = d_score[i]; return c > prm.min_conf && c > prm.track_thresh; },
                                         [&](int i, int p) { d_high[p] = i; });
            const int nl_ = warp_compact(nraw, 0, [&](int i) { const double c = d_score[i]; return c > prm.min_conf && !(c > prm.track_thresh) && c > 0.1 && c < prm.track_thresh; },
                                         [&](int i, int p) { d_low[p] = i; });
            // ---- B. split tracked list, build the pool = confirmed + lost (byte_tracker.py:207-218)
            const int nt = S.hdr[2], nlost = S.hdr[3];
            int nc = warp_compact(nt, 0, [&](int k) { return S.activated[S.tracked[k]] != 0; }, [&](int k, int p) { pool[p] = S.tracked[k]; });
            const int nu = warp_compact(nt, 0, [&](int k) { return S.activated[S.tracked[k]] == 0; }, [&](int k, int p) { unconf[p] = S.tracked[k]; });
            for (int k = lane; k < nlost; k += 32) pool[nc + k] = S.lost[k];
            nc += nlost;
            __syncwarp();
            int allf = 1;
            for (int k = lane; k < nc; k += 32) allf &= S.mean_f32[pool[k]];
            allf = __all_sync(0xffffffffu, allf);
            if (lane == 0) {
                sh->nh = nh_; sh->nl = nl_; sh->npool = nc; sh->nunc = nu; sh->all_f32 = allf;
                sh->n_tracked = nt; sh->n_lost = nlost;
            }
        }
        __syncthreads();
        PH(2);
        const int frame_id = S.hdr[0];
        const int nh = sh->nh, nl = sh->nl, npool = sh->npool, nunc = sh->nunc;

        // ---- C. multi_predict over the pool ----------------------------------------------------
        for (int k = tid; k < npool; k += BO_THREADS) {     // mean_state[6] = mean_state[7] = 0 for non-Tracked tracks (bot_sort.py:84-87)
            const int s = pool[k];
            if (S.state[s] != ST_TRACKED) { S.mean[(size_t)s * 8 + 6] = 0.0; S.mean[(size_t)s * 8 + 7] = 0.0; }
        }
        __syncthreads();
        {
            const bool pf32 = sh->all_f32 != 0;
            for (int base = 0; base < npool; base += BO_THREADS / 8) {
                const int k = base + (tid >> 3);
                const bool act = k < npool;
                bo_octet_predict(S, act ? pool[k] : 0, act, pf32);
            }
            __syncthreads();
            for (int k = tid; k < npool; k += BO_THREADS) S.mean_f32[pool[k]] = 0;
        }
        __syncthreads();
        // camera motion (bot_sort.py:343-346): the pool and the unconfirmed tracks; an unconfirmed (float32) track becomes float64
        {
            const double* H = warps + ((size_t)seq * n_frames + f) * 6;
            for (int k = tid; k < npool + nunc; k += BO_THREADS) {
                const int s = k < npool ? pool[k] : unconf[k - npool];
                bo_gmc(S.mean + (size_t)s * 8, S.cov + (size_t)s * 64, H);
                S.mean_f32[s] = 0;
            }
        }
        __syncthreads();
        for (int k = tid; k < npool; k += BO_THREADS) {          // gating projection + feature norm of every pool track
            const int s = pool[k];
            if (!bo_gate_prepare(S.mean + (size_t)s * 8, S.cov + (size_t)s * 64, gate + (size_t)k * 25)) atomicOr(status, TK_DEV_BAD_CHOLESKY);
            gate[(size_t)k * 25 + 24] = sqrt(bo_dot64(tfeat + (size_t)s * E, tfeat + (size_t)s * E, E));
        }
        __syncthreads();
        PH(3);
        for (int k = tid; k < npool; k += BO_THREADS) {
            const int s = pool[k];
            track_tlbr32(S.mean + (size_t)s * 8, S.mean_f32[s] != 0, t_tlbr + 4 * k);
        }
        __syncthreads();
        PH(4);

        // ---- D. first association: fused IoU/score cost, limit match_thresh ----------------------
        {
            const bool a_rows = npool <= nh;
            const int ld = lap_pitch(a_rows ? nh : npool);
            for (int e = tid; e < npool * nh; e += BO_THREADS) {
                const int it = e / nh, jd = e % nh;
                const int di = d_high[jd];
                // fuse_motion (matching.py:165-176), part 1: squared Mahalanobis distance; pairs above chi2inv95[4] are infeasible
                double z[4];
                det_xyah(d_box + 4 * di, z);
                gdm[e] = bo_gate_dist(gate + (size_t)it * 25, z);
            }
            __syncthreads();
            for (int e = warp_id(); e < npool * nh; e += BO_THREADS / 32) {      // part 2, one warp per pair: the cosine distance of the feasible pairs only
                const int it = e / nh, jd = e % nh;
                const int di = d_high[jd];
                const double gd = gdm[e];
                double red = 0.0;                                                 // inf entries can never match
                if (!(gd > CHI2INV95_4)) {
                    const double emb = bo_cos_dist_warp(tfeat + (size_t)pool[it] * E, gate[(size_t)it * 25 + 24], dfeat + (size_t)di * E, dnorm[di], E);
                    const double fused = __dadd_rn(__dmul_rn(prm.lambda_, emb), __dmul_rn(1 - prm.lambda_, gd));
                    red = fmin(fused - prm.match_thresh, 0.0);
                }
                if (lane_id() == 0) { if (a_rows) cost[(size_t)it * ld + jd] = red; else cost[(size_t)jd * ld + it] = red; }
            }
            __syncthreads();
            PH(5);
            solve_assignment(cost, npool, nh, match_a, match_b, lap_u, col4row, row4col, path, &sh->lap_ok, status);
        }
        // matched pool tracks: update / re_activate (byte_tracker.py:229-237)
        for (int base = 0; base < npool; base += BO_THREADS / 8) {
            const int k = base + (tid >> 3);
            const int j = k < npool ? match_a[k] : -1;
            const bool act = j >= 0;
            const int s = act ? pool[k] : 0, di = act ? d_high[j] : 0;
            bo_octet_update(S, s, act, d_box + 4 * di);
            if (act && (tid & 7) == 0) {
                S.cls[s] = d_cls[di];
                S.state[s] = ST_TRACKED; S.activated[s] = 1; S.frame_id[s] = frame_id; S.mean_f32[s] = 0;
                S.score[s] = d_score[di]; S.det_id[s] = d_id[di];
            }
        }
        for (int k = warp_id(); k < npool; k += BO_THREADS / 32)      // update_features of the matched tracks (bot_sort.py:128-129,155-156)
            if (match_a[k] >= 0) bo_update_feat_warp(tfeat + (size_t)pool[k] * E, dfeat + (size_t)d_high[match_a[k]] * E, E);
        __syncthreads();
        PH(6);
        if (warp_id() == 0) {
            const int nr = warp_compact(npool, 0, [&](int k) { return match_a[k] < 0 && S.state[pool[k]] == ST_TRACKED; },
                                        [&](int k, int p) { rest[p] = pool[k]; });
            const int nleft_ = warp_compact(nh, 0, [&](int j) { return match_b[j] < 0; }, [&](int j, int p) { d_left[p] = d_high[j]; });
            if (lane_id() == 0) { sh->nrest = nr; sh->nleft = nleft_; }
        }
        __syncthreads();
        PH(7);
        const int nrest = sh->nrest, nleft = sh->nleft;

        // ---- E. second association: remaining Tracked tracks vs low-score boxes, limit 0.5 -------
        for (int k = tid; k < nrest; k += BO_THREADS) {
            const int s = rest[k];
            track_tlbr32(S.mean + (size_t)s * 8, S.mean_f32[s] != 0, t_tlbr + 4 * k);
        }
        __syncthreads();
        PH(8);
        {
            const bool a_rows = nrest <= nl;
            const int ld = lap_pitch(a_rows ? nl : nrest);
            for (int e = tid; e < nrest * nl; e += BO_THREADS) {
                const int it = e / nl, jd = e % nl;
                const float dist = iou_dist_p1(t_tlbr + 4 * it, d_tlbrl + 4 * d_low[jd]);
                const double red = fmin((double)dist - 0.5, 0.0);
                if (a_rows) cost[(size_t)it * ld + jd] = red; else cost[(size_t)jd * ld + it] = red;
            }
            __syncthreads();
            PH(9);
            solve_assignment(cost, nrest, nl, match_a, match_b, lap_u, col4row, row4col, path, &sh->lap_ok, status);
        }
        for (int base = 0; base < nrest; base += BO_THREADS / 8) {
            const int k = base + (tid >> 3);
            const int j = k < nrest ? match_a[k] : -1;
            const bool act = j >= 0;
            const int s = act ? rest[k] : 0, di = act ? d_low[j] : 0;
            bo_octet_update(S, s, act, d_boxl + 4 * di);
            if (act && (tid & 7) == 0) {
                S.cls[s] = d_cls[di];
                S.state[s] = ST_TRACKED; S.activated[s] = 1; S.frame_id[s] = frame_id; S.mean_f32[s] = 0;
                S.score[s] = d_score[di]; S.det_id[s] = d_id[di];
            }
        }
        __syncthreads();
        PH(10);
        if (warp_id() == 0) {
            const int n = warp_compact(nrest, 0, [&](int k) { return match_a[k] < 0; },
                                       [&](int k, int p) { S.state[rest[k]] = ST_LOST; lostnow[p] = rest[k]; });
            if (lane_id() == 0) sh->n_lostnow = n;
        }
        __syncthreads();
        PH(11);

        // ---- F. unconfirmed tracks vs leftover high boxes, limit 0.7 (byte_tracker.py:266-278) ---
        for (int k = tid; k < nunc; k += BO_THREADS) {
            const int s = unconf[k];
            track_tlbr32(S.mean + (size_t)s * 8, S.mean_f32[s] != 0, t_tlbr + 4 * k);
        }
        __syncthreads();
        PH(12);
        {
            const bool a_rows = nunc <= nleft;
            const int ld = lap_pitch(a_rows ? nleft : nunc);
            for (int k = tid; k < nunc; k += BO_THREADS) {
                const int su = unconf[k];
                gate[(size_t)k * 25 + 24] = sqrt(bo_dot64(tfeat + (size_t)su * E, tfeat + (size_t)su * E, E));
            }
            __syncthreads();
            for (int e = warp_id(); e < nunc * nleft; e += BO_THREADS / 32) {     // one warp per pair
                const int it = e / nleft, jd = e % nleft;
                const int di = d_left[jd];
                // bot_sort.py:412-422: min(fuse_score(IoU distance), cosine distance / 2 gated by appearance and proximity)
                const float dist = iou_dist_p1(t_tlbr + 4 * it, d_tlbr + 4 * di);
                const float sim = __fsub_rn(1.0f, dist);
                const double iou_c = __dsub_rn(1.0, __dmul_rn((double)sim, d_score[di]));
                const int su = unconf[it];
                double emb = 1.0;
                if (!(dist > (float)prm.proximity_thresh)) {
                    emb = bo_cos_dist_warp(tfeat + (size_t)su * E, gate[(size_t)it * 25 + 24], dfeat + (size_t)di * E, dnorm[di], E) / 2.0;
                    if (emb > prm.appearance_thresh) emb = 1.0;
                }
                const double fused = fmin(iou_c, emb);
                const double red = fmin(fused - 0.7, 0.0);
                if (lane_id() != 0) continue;
                if (a_rows) cost[(size_t)it * ld + jd] = red; else cost[(size_t)jd * ld + it] = red;
            }
            __syncthreads();
            PH(13);
            solve_assignment(cost, nunc, nleft, match_a, match_b, lap_u, col4row, row4col, path, &sh->lap_ok, status);
        }
        for (int base = 0; base < nunc; base += BO_THREADS / 8) {
            const int k = base + (tid >> 3);
            const int j = k < nunc ? match_a[k] : -1;
            const bool act = j >= 0;
            const int s = k < nunc ? unconf[k] : 0, di = act ? d_left[j] : 0;
            bo_octet_update(S, s, act, d_box + 4 * di);
            if (k < nunc && (tid & 7) == 0) {
                if (act) {
                    S.cls[s] = d_cls[di];
                    S.state[s] = ST_TRACKED; S.activated[s] = 1; S.frame_id[s] = frame_id; S.mean_f32[s] = 0;
                    S.score[s] = d_score[di]; S.det_id[s] = d_id[di];
                } else {
                    S.state[s] = ST_REMOVED;   // mark_removed; it leaves `tracked` below and is never looked at again
                }
            }
        }
        for (int k = warp_id(); k < nunc; k += BO_THREADS / 32)
            if (match_a[k] >= 0) bo_update_feat_warp(tfeat + (size_t)unconf[k] * E, dfeat + (size_t)d_left[match_a[k]] * E, E);
        __syncthreads();
        PH(14);

        // ---- G. births (byte_tracker.py:280-286) + ageing (:288-291) + list maintenance (:293-299)
        if (warp_id() == 0) {
            const int nfree = S.hdr[5];
            const int id0 = S.hdr[1];
            const int nb = warp_compact(nleft, 0, [&](int j) { return match_b[j] < 0 && !(d_score[d_left[j]] < prm.det_thresh); },
                                        [&](int j, int p) {
                                            if (p >= nfree) { atomicOr(status, TK_DEV_OVERFLOW_TRACKS); return; }
                                            const int di = d_left[j];
                                            const int s = S.free_list[nfree - 1 - p];
                                            births[p] = di; birth_slot[p] = s;
                                            S.track_id[s] = id0 + 1 + p;
                                            S.state[s] = ST_TRACKED; S.activated[s] = (frame_id == 1) ? 1 : 0;
                                            S.frame_id[s] = frame_id; S.start_frame[s] = frame_id; S.in_removed[s] = 0;
                                            S.score[s] = d_score[di]; S.cls[s] = d_cls[di]; S.det_id[s] = d_id[di];
                                        });
            const int nb_ok = nb < nfree ? nb : nfree;
            if (lane_id() == 0) { S.hdr[5] = nfree - nb_ok; S.hdr[1] = id0 + nb_ok; sh->n_births = nb_ok; }
        }
        __syncthreads();
        PH(15);
        for (int k = tid; k < sh->n_births; k += BO_THREADS) bo_kf_initiate(S, birth_slot[k], d_box + 4 * births[k]);
        for (int k = warp_id(); k < sh->n_births; k += BO_THREADS / 32)       // smooth_feat = the detection's normalised feature
            for (int e = lane_id(); e < E; e += 32) tfeat[(size_t)birth_slot[k] * E + e] = dfeat[(size_t)births[k] * E + e];
        if (warp_id() == 0) {
            const int lane = lane_id();
            const int nt = sh->n_tracked, nlost = sh->n_lost, nb = sh->n_births;
            // ageing of the OLD lost list; removed_now membership is applied to in_removed after the subtraction
            for (int k = lane; k < nlost; k += 32) {
                const int s = S.lost[k];
                const bool aged = frame_id - S.frame_id[s] > prm.max_time_lost;
                if (aged) S.state[s] = ST_REMOVED;
                dup_b[k] = aged ? 1 : 0;
            }
            for (int k = lane; k < cap; k += 32) in_tracked[k] = 0;
            __syncwarp();
            // tracked' = [old tracked still Tracked] + births + refinds(lost order)
            int n = warp_compact(nt, 0, [&](int k) { return S.state[S.tracked[k]] == ST_TRACKED; },
                                 [&](int k, int p) { const int s = S.tracked[k]; newlist[p] = s; in_tracked[s] = 1; });
            for (int k = lane; k < nb; k += 32) { newlist[n + k] = birth_slot[k]; in_tracked[birth_slot[k]] = 1; }
            n += nb;
            __syncwarp();
            n = warp_compact(nlost, n, [&](int k) { const int s = S.lost[k]; return S.state[s] == ST_TRACKED && !in_tracked[s]; },
                             [&](int k, int p) { const int s = S.lost[k]; newlist[p] = s; in_tracked[s] = 1; });
            // free-list pushes (slot numbers are anonymous, their order is irrelevant)
            int nfree = S.hdr[5];
            nfree = warp_compact(nlost, nfree, [&](int k) { const int s = S.lost[k]; return !in_tracked[s] && S.in_removed[s]; },
                                 [&](int k, int p) { S.free_list[p] = S.lost[k]; });
            nfree = warp_compact(sh->n_lostnow, nfree, [&](int k) { return S.in_removed[lostnow[k]] != 0; },
                                 [&](int k, int p) { S.free_list[p] = lostnow[k]; });
            nfree = warp_compact(nt, nfree, [&](int k) { const int s = S.tracked[k]; return S.state[s] == ST_REMOVED && !S.activated[s]; },
                                 [&](int k, int p) { S.free_list[p] = S.tracked[k]; });
            // lost' = (old lost - tracked') + lost_now, minus everything that was in `removed` BEFORE this frame
            int m = warp_compact(nlost, 0, [&](int k) { const int s = S.lost[k]; return !in_tracked[s] && !S.in_removed[s]; },
                                 [&](int k, int p) { const int s = S.lost[k]; if (dup_b[k]) S.in_removed[s] = 1; rest[p] = s; });
            m = warp_compact(sh->n_lostnow, m, [&](int k) { return S.in_removed[lostnow[k]] == 0; }, [&](int k, int p) { rest[p] = lostnow[k]; });
            for (int k = lane; k < n; k += 32) S.tracked[k] = newlist[k];
            for (int k = lane; k < m; k += 32) S.lost[k] = rest[k];
            if (lane == 0) { S.hdr[5] = nfree; sh->n_tracked = n; sh->n_lost = m; }
        }
        __syncthreads();
        PH(16);

        // ---- H. remove_duplicate_stracks (byte_tracker.py:348-361) ------------------------------
        {
            const int nt = sh->n_tracked, nlost = sh->n_lost;
            for (int k = tid; k < nt; k += BO_THREADS) {
                const int s = S.tracked[k];
                track_tlbr32(S.mean + (size_t)s * 8, S.mean_f32[s] != 0, t_tlbr + 4 * k);
                dup_a[k] = 0;
            }
            for (int k = tid; k < nlost; k += BO_THREADS) {
                const int s = S.lost[k];
                track_tlbr32(S.mean + (size_t)s * 8, S.mean_f32[s] != 0, t_tlbr2 + 4 * k);
                dup_b[k] = 0;
            }
            __syncthreads();
            PH(17);
            for (int e = tid; e < nt * nlost; e += BO_THREADS) {
                const int p = e / nlost, q = e % nlost;
                const float dist = iou_dist_p1(t_tlbr + 4 * p, t_tlbr2 + 4 * q);
                if (dist < 0.15f) {
                    const int sp_ = S.tracked[p], sq = S.lost[q];
                    const int tp = S.frame_id[sp_] - S.start_frame[sp_];
                    const int tq = S.frame_id[sq] - S.start_frame[sq];
                    if (tp > tq) dup_b[q] = 1; else dup_a[p] = 1;
                }
            }
            __syncthreads();
            PH(18);
            if (warp_id() == 0) {
                int nfree = S.hdr[5];
                nfree = warp_compact(nt, nfree, [&](int k) { return dup_a[k] != 0; }, [&](int k, int p) { S.free_list[p] = S.tracked[k]; });
                nfree = warp_compact(nlost, nfree, [&](int k) { return dup_b[k] != 0; }, [&](int k, int p) { S.free_list[p] = S.lost[k]; });
                // in-place ordered compaction is safe: every lane reads its element before the ballot, writes land at <= its index
                const int n = warp_compact(nt, 0, [&](int k) { return dup_a[k] == 0; },
                                           [&](int k, int p) { const int s = S.tracked[k]; newlist[p] = s; });
                for (int k = lane_id(); k < n; k += 32) S.tracked[k] = newlist[k];
                const int m = warp_compact(nlost, 0, [&](int k) { return dup_b[k] == 0; },
                                           [&](int k, int p) { const int s = S.lost[k]; rest[p] = s; });
                for (int k = lane_id(); k < m; k += 32) S.lost[k] = rest[k];
                __syncwarp();
                // ---- I. output rows of activated tracks (byte_tracker.py:301-318)
                for (int k = lane_id(); k < n; k += 32) out_pos[k] = -1;
                __syncwarp();
                const int cnt = warp_compact(n, 0, [&](int k) { return S.activated[S.tracked[k]] != 0; }, [&](int k, int p) { out_pos[k] = p; });
                if (lane_id() == 0) {
                    S.hdr[5] = nfree; S.hdr[2] = n; S.hdr[3] = m;
                    sh->n_tracked = n; sh->nd = cnt;
                    out_frame_count[seq * n_frames + f] = cnt;
                }
            }
            __syncthreads();
            PH(19);
        }
        {
            const int n = sh->n_tracked;
            for (int k = tid; k < n; k += BO_THREADS) {
                if (out_pos[k] < 0) continue;
                const int s = S.tracked[k];
                double t[4];
                const bool f32 = S.mean_f32[s] != 0;
                track_tlwh(S.mean + (size_t)s * 8, f32, t);
                double* o = out_rows + (size_t)(out_base + out_n + out_pos[k]) * 8;
                if (f32) {  // xywh2xyxy on a float32 row (byte_tracker.py:311)
                    const float x = (float)t[0], y = (float)t[1];
                    const float hw = __fdiv_rn((float)t[2], 2.0f), hh = __fdiv_rn((float)t[3], 2.0f);
                    o[0] = (double)__fsub_rn(x, hw); o[1] = (double)__fsub_rn(y, hh);
                    o[2] = (double)__fadd_rn(x, hw); o[3] = (double)__fadd_rn(y, hh);
                } else {
                    const double hw = t[2] / 2, hh = t[3] / 2;
                    o[0] = t[0] - hw; o[1] = t[1] - hh; o[2] = t[0] + hw; o[3] = t[1] + hh;
                }
                o[4] = (double)S.track_id[s]; o[5] = S.cls[s]; o[6] = S.score[s]; o[7] = S.det_id[s];
            }
            out_n += sh->nd;
        }
        __syncthreads();
        PH(20);
    }
    if (tid == 0) out_count[seq] = out_n;
    // write the book-keeping back for the next chunk of frames
    __syncthreads();
    if (tid < 8) G.hdr[tid] = S.hdr[tid];
    for (int i = tid; i < cap; i += BO_THREADS) {
        G.score[i] = S.score[i]; G.cls[i] = S.cls[i]; G.det_id[i] = S.det_id[i];
        G.track_id[i] = S.track_id[i]; G.frame_id[i] = S.frame_id[i]; G.start_frame[i] = S.start_frame[i];
        G.tracked[i] = S.tracked[i]; G.lost[i] = S.lost[i]; G.free_list[i] = S.free_list[i];
        G.state[i] = S.state[i]; G.activated[i] = S.activated[i]; G.mean_f32[i] = S.mean_f32[i]; G.in_removed[i] = S.in_removed[i];
    }
}

struct BoHandle {
    BoParams prm;
    int n_seq, cap, capd, first_id;
    char* state;
    size_t state_stride;
    float *tfeat, *dfeat;
    double* gate;
    double* cost;
    size_t cost_stride;
    size_t smem_bytes;
    int cost_in_smem;
};

__global__ void botsort_reset_kernel(char* base, size_t stride, int cap, int first_id, int keep_ids) {
    BoDev S = bo_carve(base + (size_t)blockIdx.x * stride, cap);
    if (threadIdx.x == 0) {
        // BaseTrack.clear_count() runs in every BoTSORT() (bot_sort.py:262): the numbering restarts per video
        (void)keep_ids;
        S.hdr[0] = 0; S.hdr[1] = first_id - 1; S.hdr[2] = 0; S.hdr[3] = 0; S.hdr[4] = 0; S.hdr[5] = cap;
    }
    for (int i = threadIdx.x; i < cap; i += blockDim.x) {
        S.free_list[i] = cap - 1 - i;  // pop order = slot 0, 1, 2, ...
        S.state[i] = ST_NEW; S.activated[i] = 0; S.mean_f32[i] = 0; S.in_removed[i] = 0;
    }
}

size_t bo_smem_fixed(int cap, int capd) {
    const int side = cap > capd ? cap : capd;
    auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
    size_t s = 0;
    s += al(sizeof(double) * side);
    s += 3 * al(sizeof(double) * capd);
    s += 4 * al(sizeof(float) * 4 * capd);
    s += 2 * al(sizeof(float) * 4 * cap);
    s += 3 * al(sizeof(int) * capd);
    s += 4 * al(sizeof(int) * cap);
    s += 2 * al(sizeof(int) * capd);
    s += 5 * al(sizeof(int) * side);
    s += 2 * al(sizeof(int) * cap);
    s += 3 * al((size_t)cap);
    s += al(sizeof(BoShared));
    s += al(8 * sizeof(int)) + 3 * al(sizeof(double) * cap) + 6 * al(sizeof(int) * cap) + 4 * al((size_t)cap);   // book-keeping mirror
    return s;
}

}  // namespace

extern "C" {

int tk_botsort_create(const tk_botsort_params* p, int n_seq, int cap_tracks, int cap_dets, void** handle) {
    if (!p || !handle || n_seq <= 0 || cap_tracks <= 0 || cap_dets <= 0 || p->feature_dim <= 0) return TK_ERR_ARG;
    if (cap_tracks > tk::LAP_MAX_COLS || cap_dets > tk::LAP_MAX_COLS) return TK_ERR_CAPACITY;
    BoHandle* h = new BoHandle();
    h->prm.track_thresh = p->track_high_thresh;
    h->prm.match_thresh = p->match_thresh;
    h->prm.det_thresh = p->new_track_thresh;
    h->prm.min_conf = p->min_confidence;
    h->prm.proximity_thresh = p->proximity_thresh; h->prm.appearance_thresh = p->appearance_thresh; h->prm.lambda_ = p->lambda_;
    h->prm.max_time_lost = (int)((double)p->frame_rate / 30.0 * p->track_buffer);  // bot_sort.py:268-269
    h->prm.emb_dim = p->feature_dim;
    h->n_seq = n_seq; h->cap = cap_tracks; h->capd = cap_dets; h->first_id = 1;      // BaseTrack.clear_count() in every BoTSORT()
    h->state_stride = (bo_state_bytes(cap_tracks) + 255) & ~(size_t)255;
    h->state = nullptr; h->cost = nullptr; h->tfeat = nullptr; h->dfeat = nullptr; h->gate = nullptr;
    const size_t fixed = bo_smem_fixed(cap_tracks, cap_dets);
    const size_t cost_bytes = (size_t)(cap_tracks + 1) * (cap_dets + 1) * sizeof(double);
    h->cost_in_smem = (fixed + cost_bytes <= 200 * 1024) ? 1 : 0;
    h->smem_bytes = fixed + (h->cost_in_smem ? cost_bytes : 0);
    h->cost_stride = (size_t)(cap_tracks + 1) * (cap_dets + 1);
    cudaError_t e = cudaMalloc((void**)&h->state, h->state_stride * n_seq);
    if (e == cudaSuccess && !h->cost_in_smem) e = cudaMalloc((void**)&h->cost, h->cost_stride * sizeof(double) * n_seq);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->tfeat, sizeof(float) * (size_t)n_seq * cap_tracks * p->feature_dim);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->dfeat, sizeof(float) * (size_t)n_seq * cap_dets * p->feature_dim);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->gate, sizeof(double) * (size_t)n_seq * ((size_t)cap_tracks * 25 + cap_dets + (size_t)cap_tracks * cap_dets));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(botsort_video_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    if (e != cudaSuccess) {
        tk_set_last_cuda_error((int)e);
        if (h->state) cudaFree(h->state);
        if (h->cost) cudaFree(h->cost);
        if (h->tfeat) cudaFree(h->tfeat);
        if (h->dfeat) cudaFree(h->dfeat);
        if (h->gate) cudaFree(h->gate);
        delete h;
        return TK_ERR_CUDA;
    }
    *handle = h;
    return tk_botsort_reset(h, 0, nullptr);
}

int tk_botsort_reset(void* handle, int keep_id_counter, void* stream) {
    if (!handle) return TK_ERR_ARG;
    BoHandle* h = (BoHandle*)handle;
    botsort_reset_kernel<<<h->n_seq, 128, 0, (cudaStream_t)stream>>>(h->state, h->state_stride, h->cap, h->first_id,
                                                                      keep_id_counter);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_botsort_run(void* handle, const double* dets, const float* embeddings, const double* warps, const int* offsets, int n_frames,
                   double* out_rows, const int* out_start, int* out_frame_count, int* out_count, void* stream) {
    if (!handle || !offsets || !out_rows || !out_start || !out_frame_count || !out_count || n_frames < 0 || !embeddings || !warps)
        return TK_ERR_ARG;
    BoHandle* h = (BoHandle*)handle;
    if (n_frames == 0) return TK_OK;
    // the attribute is per kernel function, not per handle: another handle with smaller capacities may have lowered it
    TK_CUDA_TRY(cudaFuncSetAttribute(botsort_video_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    botsort_video_kernel<<<h->n_seq, BO_THREADS, h->smem_bytes, (cudaStream_t)stream>>>(
        h->prm, h->state, h->state_stride, h->cap, h->capd, dets, offsets, n_frames, out_rows, out_start,
        out_frame_count, out_count, h->cost, h->cost_stride, h->cost_in_smem, embeddings, warps, h->tfeat, h->dfeat, h->gate);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_botsort_status(void* handle, int* status_host, void* stream) {
    if (!handle || !status_host) return TK_ERR_ARG;
    BoHandle* h = (BoHandle*)handle;
    for (int s = 0; s < h->n_seq; ++s) {
        TK_CUDA_TRY(cudaMemcpyAsync(status_host + s, h->state + (size_t)s * h->state_stride + 4 * sizeof(int), sizeof(int),
                                    cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    }
    TK_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return TK_OK;
}

#ifdef TK_PHASE_PROF
int tk_debug_botsort_phases(unsigned long long* host_out64, int reset) {
    cudaDeviceSynchronize();
    if (host_out64) cudaMemcpyFromSymbol(host_out64, g_bo_prof, sizeof(unsigned long long) * 64);
    if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(g_bo_prof, z, sizeof(z)); }
    return 0;
}
#endif

int tk_botsort_destroy(void* handle) {
    if (!handle) return TK_ERR_ARG;
    BoHandle* h = (BoHandle*)handle;
    cudaFree(h->state);
    if (h->cost) cudaFree(h->cost);
    cudaFree(h->tfeat); cudaFree(h->dfeat); cudaFree(h->gate);
    delete h;
    return TK_OK;
}

}  // extern "C"
