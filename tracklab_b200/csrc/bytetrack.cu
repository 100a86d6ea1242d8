// ByteTrack association, whole video per launch, one CTA per video (sequence).
//
// Device restatement of BYTETracker.update and the STrack life cycle
//   /root/reference/plugins/track/byte_track/byte_tracker.py:10-148,167-361
//   /root/reference/plugins/track/byte_track/matching.py:37-48,171-218
//   /root/reference/plugins/track/byte_track/kalman_filter.py:55-226
// and of the wrapper's per-frame filter /root/reference/tracklab/wrappers/track/byte_track_api.py:50-56.
//
// The reference is a per-frame Python state machine; the frames of a video form a strict dependency
// chain, so the B200 shape is: detections of the whole video (or a chunk) resident in HBM, ONE kernel
// launch walks the frames with zero host round trips, tracker state stays on chip between frames
// (bookkeeping in shared memory, filter state in L1/L2-resident global memory), the matrix steps
// (IoU, cost fusion, Kalman predict/update) are spread over the CTA and the assignment is solved by a
// single warp (lap.cuh). Videos are independent: grid = number of videos.
#include "kf_xyah.cuh"
#include "lap.cuh"
#include "trackkern.h"

namespace {

using namespace tk;

constexpr int BT_THREADS = 256;   // 32 octets: one Kalman update per 8 lanes (kf_xyah.cuh)

// Optional per-phase cycle accounting (build with -DTK_PHASE_PROF): thread 0 accumulates clock64() deltas between
// the barriers of the frame loop into g_bt_prof[]; read back with tk_debug_bytetrack_phases().
#ifdef TK_PHASE_PROF
__device__ unsigned long long g_bt_prof[64];
#define PH(k) do { if (threadIdx.x == 0) { const long long _t = clock64(); g_bt_prof[k] += (unsigned long long)(_t - ph_t0); ph_t0 = _t; } } while (0)
#else
#define PH(k) do { } while (0)
#endif
enum : unsigned char { ST_NEW = 0, ST_TRACKED = 1, ST_LOST = 2, ST_REMOVED = 3 };

struct BtDev {
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

__host__ __device__ inline size_t bt_align(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline size_t bt_state_bytes(int cap) {
    size_t s = 0;
    s += bt_align(8 * sizeof(int));
    s += bt_align((size_t)cap * 8 * sizeof(double));
    s += bt_align((size_t)cap * 64 * sizeof(double));
    s += 3 * bt_align((size_t)cap * sizeof(double));
    s += 3 * bt_align((size_t)cap * sizeof(int));
    s += 4 * bt_align((size_t)cap);
    s += 3 * bt_align((size_t)cap * sizeof(int));
    return s;
}

__host__ __device__ inline BtDev bt_carve(char* base, int cap) {
    BtDev d;
    char* p = base;
    d.hdr = (int*)p; p += bt_align(8 * sizeof(int));
    d.mean = (double*)p; p += bt_align((size_t)cap * 8 * sizeof(double));
    d.cov = (double*)p; p += bt_align((size_t)cap * 64 * sizeof(double));
    d.score = (double*)p; p += bt_align((size_t)cap * sizeof(double));
    d.cls = (double*)p; p += bt_align((size_t)cap * sizeof(double));
    d.det_id = (double*)p; p += bt_align((size_t)cap * sizeof(double));
    d.track_id = (int*)p; p += bt_align((size_t)cap * sizeof(int));
    d.frame_id = (int*)p; p += bt_align((size_t)cap * sizeof(int));
    d.start_frame = (int*)p; p += bt_align((size_t)cap * sizeof(int));
    d.state = (unsigned char*)p; p += bt_align((size_t)cap);
    d.activated = (unsigned char*)p; p += bt_align((size_t)cap);
    d.mean_f32 = (unsigned char*)p; p += bt_align((size_t)cap);
    d.in_removed = (unsigned char*)p; p += bt_align((size_t)cap);
    d.tracked = (int*)p; p += bt_align((size_t)cap * sizeof(int));
    d.lost = (int*)p; p += bt_align((size_t)cap * sizeof(int));
    d.free_list = (int*)p;
    return d;
}

struct BtParams {
    double track_thresh, match_thresh, det_thresh, min_conf;
    int max_time_lost;
};

// ---- float32 box helpers: every operation is a single IEEE fp32 op (no FMA contraction) ----------
// STrack.tlwh / tlbr (byte_tracker.py:100-120) for a track whose mean is float32 or float64.
__device__ __forceinline__ void track_tlwh(const double* m, bool f32, double* out) {
    if (f32) {
        const float x = (float)m[0], y = (float)m[1], a = (float)m[2], h = (float)m[3];
        const float w = __fmul_rn(a, h);
        out[0] = (double)__fsub_rn(x, __fdiv_rn(w, 2.0f));
        out[1] = (double)__fsub_rn(y, __fdiv_rn(h, 2.0f));
        out[2] = (double)w;
        out[3] = (double)h;
    } else {
        const double w = m[2] * m[3];
        out[0] = m[0] - w / 2;
        out[1] = m[1] - m[3] / 2;
        out[2] = w;
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
__device__ __forceinline__ void det_xyah(const float* b, double* z) {
    z[0] = (double)__fadd_rn(b[0], __fdiv_rn(b[2], 2.0f));
    z[1] = (double)__fadd_rn(b[1], __fdiv_rn(b[3], 2.0f));
    z[2] = (double)__fdiv_rn(b[2], b[3]);
    z[3] = (double)b[3];
}

constexpr double W_POS = 1.0 / 20;
constexpr double W_VEL = 1.0 / 160;

// KalmanFilter.update (kalman_filter.py:194-226, project :126-153) and multi_predict (:155-192); q is float32 when the
// whole pool still carries float32 means (NumPy promotion of np.asarray([...float32 means...])).
// Octet-cooperative (8 lanes per track, kf_xyah.cuh), register resident.
__device__ __forceinline__ void bt_octet_update(BtDev& S, int slot, bool active, const float* detbox) {
    double z[4] = {0, 0, 0, 0}, r[4] = {1, 1, 1, 1};
    double* gm = S.mean + (size_t)(active ? slot : 0) * 8;
    double* gP = S.cov + (size_t)(active ? slot : 0) * 64;
    if (active) {
        det_xyah(detbox, z);
        const double h = gm[3];
        const double sp = S.mean_f32[slot] ? (double)__fmul_rn((float)W_POS, (float)h) : W_POS * h;
        r[0] = sp * sp; r[1] = sp * sp; r[2] = 1e-1 * 1e-1; r[3] = sp * sp;
    }
    if (!kf8_octet_update(gm, gP, active, z, r)) atomicOr(&S.hdr[4], TK_DEV_BAD_CHOLESKY);
}

__device__ __forceinline__ void bt_octet_predict(BtDev& S, int slot, bool active, bool pool_f32) {
    const int j = threadIdx.x & 7;
    double* gm = S.mean + (size_t)(active ? slot : 0) * 8;
    double* gP = S.cov + (size_t)(active ? slot : 0) * 64;
    double qj = 0.0;
    bool zero_vh = false;
    if (active) {
        zero_vh = S.state[slot] != ST_TRACKED;
        const double h = gm[3];
        const bool pos = j < 4;
        if (pool_f32) {
            const float hf = (float)h;
            const float sd = (j == 2) ? (float)1e-2 : (j == 6) ? (float)1e-5 : __fmul_rn(pos ? (float)W_POS : (float)W_VEL, hf);
            qj = (double)__fmul_rn(sd, sd);
        } else {
            const double sd = (j == 2) ? 1e-2 : (j == 6) ? 1e-5 : (pos ? W_POS : W_VEL) * h;
            qj = sd * sd;
        }
    }
    kf8_octet_predict(gm, gP, active, zero_vh, qj);
}

// KalmanFilter.initiate (kalman_filter.py:55-86): mean stays float32-valued, std are float32 products
__device__ void bt_kf_initiate(BtDev& S, int slot, const float* detbox) {
    double z[4];
    det_xyah(detbox, z);
    double* gm = S.mean + (size_t)slot * 8;
    double* gP = S.cov + (size_t)slot * 64;
    for (int i = 0; i < 4; ++i) { gm[i] = z[i]; gm[i + 4] = 0.0; }
    const float h = (float)z[3];
    const double sp = (double)__fmul_rn((float)(2 * W_POS), h);
    const double sv = (double)__fmul_rn((float)(10 * W_VEL), h);
    const double d[8] = {sp * sp, sp * sp, 1e-2 * 1e-2, sp * sp, sv * sv, sv * sv, 1e-5 * 1e-5, sv * sv};
    for (int i = 0; i < 64; ++i) gP[i] = 0.0;
    for (int i = 0; i < 8; ++i) gP[i * 9] = d[i];
    S.mean_f32[slot] = 1;
}

struct BtShared {
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

__global__ void __launch_bounds__(BT_THREADS)
bytetrack_video_kernel(BtParams prm, char* state_base, size_t state_stride, int cap, int capd,
                       const double* __restrict__ dets, const int* __restrict__ offsets, int n_frames,
                       double* __restrict__ out_rows, const int* __restrict__ out_start,
                       int* __restrict__ out_frame_count, int* __restrict__ out_count,
                       double* cost_scratch, size_t cost_stride, int cost_in_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int seq = blockIdx.x;
    const int tid = threadIdx.x;
    BtDev S = bt_carve(state_base + (size_t)seq * state_stride, cap);

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
    BtShared* sh = (BtShared*)take(sizeof(BtShared));
    // Book-keeping of the tracker (lists, per-slot flags/ids/scores) lives in shared memory for the whole launch:
    // the frame loop below is a chain of short serial list edits, and every global round trip in it is pure latency.
    // Only the Kalman means/covariances (576 B per track) stay in global memory (L1/L2 resident).
    const BtDev G = S;
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
        for (int i = tid; i < cap; i += BT_THREADS) {
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
        for (int i = tid; i < nraw; i += BT_THREADS) {
            const double* d = dets + (size_t)(r0 + i) * 7;
            const double x1 = d[0], y1 = d[1], x2 = d[2], y2 = d[3];
            float* b = d_box + 4 * i;
            b[0] = (float)((x1 + x2) / 2); b[1] = (float)((y1 + y2) / 2);
            b[2] = (float)(x2 - x1); b[3] = (float)(y2 - y1);
            float* t = d_tlbr + 4 * i;  // STrack.tlbr of a detection: float32 tlwh + float32 adds
            t[0] = b[0]; t[1] = b[1]; t[2] = __fadd_rn(b[2], b[0]); t[3] = __fadd_rn(b[3], b[1]);
            d_score[i] = d[4]; d_cls[i] = d[5]; d_id[i] = d[6];
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
        {
            const bool pf32 = sh->all_f32 != 0;
            for (int base = 0; base < npool; base += BT_THREADS / 8) {
                const int k = base + (tid >> 3);
                const bool act = k < npool;
                bt_octet_predict(S, act ? pool[k] : 0, act, pf32);
            }
            __syncthreads();
            for (int k = tid; k < npool; k += BT_THREADS) S.mean_f32[pool[k]] = 0;
        }
        __syncthreads();
        PH(3);
        for (int k = tid; k < npool; k += BT_THREADS) {
            const int s = pool[k];
            track_tlbr32(S.mean + (size_t)s * 8, S.mean_f32[s] != 0, t_tlbr + 4 * k);
        }
        __syncthreads();
        PH(4);

        // ---- D. first association: fused IoU/score cost, limit match_thresh ----------------------
        {
            const bool a_rows = npool <= nh;
            const int ld = lap_pitch(a_rows ? nh : npool);
            for (int e = tid; e < npool * nh; e += BT_THREADS) {
                const int it = e / nh, jd = e % nh;
                const int di = d_high[jd];
                const float dist = iou_dist_p1(t_tlbr + 4 * it, d_tlbr + 4 * di);
                const float sim = __fsub_rn(1.0f, dist);                       // fuse_score (matching.py:171-179)
                const double fused = __dsub_rn(1.0, __dmul_rn((double)sim, d_score[di]));
                const double red = fmin(fused - prm.match_thresh, 0.0);
                if (a_rows) cost[(size_t)it * ld + jd] = red; else cost[(size_t)jd * ld + it] = red;
            }
            __syncthreads();
            PH(5);
            solve_assignment(cost, npool, nh, match_a, match_b, lap_u, col4row, row4col, path, &sh->lap_ok, status);
        }
        // matched pool tracks: update / re_activate (byte_tracker.py:229-237)
        for (int base = 0; base < npool; base += BT_THREADS / 8) {
            const int k = base + (tid >> 3);
            const int j = k < npool ? match_a[k] : -1;
            const bool act = j >= 0;
            const int s = act ? pool[k] : 0, di = act ? d_high[j] : 0;
            bt_octet_update(S, s, act, d_box + 4 * di);
            if (act && (tid & 7) == 0) {
                if (S.state[s] != ST_TRACKED) S.cls[s] = d_cls[di];   // re_activate also copies cls
                S.state[s] = ST_TRACKED; S.activated[s] = 1; S.frame_id[s] = frame_id; S.mean_f32[s] = 0;
                S.score[s] = d_score[di]; S.det_id[s] = d_id[di];
            }
        }
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
        for (int k = tid; k < nrest; k += BT_THREADS) {
            const int s = rest[k];
            track_tlbr32(S.mean + (size_t)s * 8, S.mean_f32[s] != 0, t_tlbr + 4 * k);
        }
        __syncthreads();
        PH(8);
        {
            const bool a_rows = nrest <= nl;
            const int ld = lap_pitch(a_rows ? nl : nrest);
            for (int e = tid; e < nrest * nl; e += BT_THREADS) {
                const int it = e / nl, jd = e % nl;
                const float dist = iou_dist_p1(t_tlbr + 4 * it, d_tlbr + 4 * d_low[jd]);
                const double red = fmin((double)dist - 0.5, 0.0);
                if (a_rows) cost[(size_t)it * ld + jd] = red; else cost[(size_t)jd * ld + it] = red;
            }
            __syncthreads();
            PH(9);
            solve_assignment(cost, nrest, nl, match_a, match_b, lap_u, col4row, row4col, path, &sh->lap_ok, status);
        }
        for (int base = 0; base < nrest; base += BT_THREADS / 8) {
            const int k = base + (tid >> 3);
            const int j = k < nrest ? match_a[k] : -1;
            const bool act = j >= 0;
            const int s = act ? rest[k] : 0, di = act ? d_low[j] : 0;
            bt_octet_update(S, s, act, d_box + 4 * di);
            if (act && (tid & 7) == 0) {
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
        for (int k = tid; k < nunc; k += BT_THREADS) {
            const int s = unconf[k];
            track_tlbr32(S.mean + (size_t)s * 8, S.mean_f32[s] != 0, t_tlbr + 4 * k);
        }
        __syncthreads();
        PH(12);
        {
            const bool a_rows = nunc <= nleft;
            const int ld = lap_pitch(a_rows ? nleft : nunc);
            for (int e = tid; e < nunc * nleft; e += BT_THREADS) {
                const int it = e / nleft, jd = e % nleft;
                const int di = d_left[jd];
                const float dist = iou_dist_p1(t_tlbr + 4 * it, d_tlbr + 4 * di);
                const float sim = __fsub_rn(1.0f, dist);
                const double fused = __dsub_rn(1.0, __dmul_rn((double)sim, d_score[di]));
                const double red = fmin(fused - 0.7, 0.0);
                if (a_rows) cost[(size_t)it * ld + jd] = red; else cost[(size_t)jd * ld + it] = red;
            }
            __syncthreads();
            PH(13);
            solve_assignment(cost, nunc, nleft, match_a, match_b, lap_u, col4row, row4col, path, &sh->lap_ok, status);
        }
        for (int base = 0; base < nunc; base += BT_THREADS / 8) {
            const int k = base + (tid >> 3);
            const int j = k < nunc ? match_a[k] : -1;
            const bool act = j >= 0;
            const int s = k < nunc ? unconf[k] : 0, di = act ? d_left[j] : 0;
            bt_octet_update(S, s, act, d_box + 4 * di);
            if (k < nunc && (tid & 7) == 0) {
                if (act) {
                    S.state[s] = ST_TRACKED; S.activated[s] = 1; S.frame_id[s] = frame_id; S.mean_f32[s] = 0;
                    S.score[s] = d_score[di]; S.det_id[s] = d_id[di];
                } else {
                    S.state[s] = ST_REMOVED;   // mark_removed; it leaves `tracked` below and is never looked at again
                }
            }
        }
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
        for (int k = tid; k < sh->n_births; k += BT_THREADS) bt_kf_initiate(S, birth_slot[k], d_box + 4 * births[k]);
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
            for (int k = tid; k < nt; k += BT_THREADS) {
                const int s = S.tracked[k];
                track_tlbr32(S.mean + (size_t)s * 8, S.mean_f32[s] != 0, t_tlbr + 4 * k);
                dup_a[k] = 0;
            }
            for (int k = tid; k < nlost; k += BT_THREADS) {
                const int s = S.lost[k];
                track_tlbr32(S.mean + (size_t)s * 8, S.mean_f32[s] != 0, t_tlbr2 + 4 * k);
                dup_b[k] = 0;
            }
            __syncthreads();
            PH(17);
            for (int e = tid; e < nt * nlost; e += BT_THREADS) {
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
            for (int k = tid; k < n; k += BT_THREADS) {
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
    for (int i = tid; i < cap; i += BT_THREADS) {
        G.score[i] = S.score[i]; G.cls[i] = S.cls[i]; G.det_id[i] = S.det_id[i];
        G.track_id[i] = S.track_id[i]; G.frame_id[i] = S.frame_id[i]; G.start_frame[i] = S.start_frame[i];
        G.tracked[i] = S.tracked[i]; G.lost[i] = S.lost[i]; G.free_list[i] = S.free_list[i];
        G.state[i] = S.state[i]; G.activated[i] = S.activated[i]; G.mean_f32[i] = S.mean_f32[i]; G.in_removed[i] = S.in_removed[i];
    }
}

struct BtHandle {
    BtParams prm;
    int n_seq, cap, capd, first_id;
    char* state;
    size_t state_stride;
    double* cost;
    size_t cost_stride;
    size_t smem_bytes;
    int cost_in_smem;
};

__global__ void bytetrack_reset_kernel(char* base, size_t stride, int cap, int first_id, int keep_ids) {
    BtDev S = bt_carve(base + (size_t)blockIdx.x * stride, cap);
    if (threadIdx.x == 0) {
        // BaseTrack._count is process-global in the reference (basetrack.py:13): keep_ids continues the numbering
        S.hdr[0] = 0; if (!keep_ids) S.hdr[1] = first_id - 1; S.hdr[2] = 0; S.hdr[3] = 0; S.hdr[4] = 0; S.hdr[5] = cap;
    }
    for (int i = threadIdx.x; i < cap; i += blockDim.x) {
        S.free_list[i] = cap - 1 - i;  // pop order = slot 0, 1, 2, ...
        S.state[i] = ST_NEW; S.activated[i] = 0; S.mean_f32[i] = 0; S.in_removed[i] = 0;
    }
}

size_t bt_smem_fixed(int cap, int capd) {
    const int side = cap > capd ? cap : capd;
    auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
    size_t s = 0;
    s += al(sizeof(double) * side);
    s += 3 * al(sizeof(double) * capd);
    s += 2 * al(sizeof(float) * 4 * capd);
    s += 2 * al(sizeof(float) * 4 * cap);
    s += 3 * al(sizeof(int) * capd);
    s += 4 * al(sizeof(int) * cap);
    s += 2 * al(sizeof(int) * capd);
    s += 5 * al(sizeof(int) * side);
    s += 2 * al(sizeof(int) * cap);
    s += 3 * al((size_t)cap);
    s += al(sizeof(BtShared));
    s += al(8 * sizeof(int)) + 3 * al(sizeof(double) * cap) + 6 * al(sizeof(int) * cap) + 4 * al((size_t)cap);   // book-keeping mirror
    return s;
}

}  // namespace

extern "C" {

int tk_bytetrack_create(const tk_bytetrack_params* p, int n_seq, int cap_tracks, int cap_dets, void** handle) {
    if (!p || !handle || n_seq <= 0 || cap_tracks <= 0 || cap_dets <= 0) return TK_ERR_ARG;
    if (cap_tracks > tk::LAP_MAX_COLS || cap_dets > tk::LAP_MAX_COLS) return TK_ERR_CAPACITY;
    BtHandle* h = new BtHandle();
    h->prm.track_thresh = p->track_thresh;
    h->prm.match_thresh = p->match_thresh;
    h->prm.det_thresh = p->track_thresh + 0.1;                                   // byte_tracker.py:161
    h->prm.min_conf = p->min_confidence;
    h->prm.max_time_lost = (int)((double)p->frame_rate / 30.0 * p->track_buffer);  // byte_tracker.py:162
    h->n_seq = n_seq; h->cap = cap_tracks; h->capd = cap_dets; h->first_id = p->first_id;
    h->state_stride = (bt_state_bytes(cap_tracks) + 255) & ~(size_t)255;
    h->state = nullptr; h->cost = nullptr;
    const size_t fixed = bt_smem_fixed(cap_tracks, cap_dets);
    const size_t cost_bytes = (size_t)(cap_tracks + 1) * (cap_dets + 1) * sizeof(double);
    h->cost_in_smem = (fixed + cost_bytes <= 200 * 1024) ? 1 : 0;
    h->smem_bytes = fixed + (h->cost_in_smem ? cost_bytes : 0);
    h->cost_stride = (size_t)(cap_tracks + 1) * (cap_dets + 1);
    cudaError_t e = cudaMalloc((void**)&h->state, h->state_stride * n_seq);
    if (e == cudaSuccess && !h->cost_in_smem) e = cudaMalloc((void**)&h->cost, h->cost_stride * sizeof(double) * n_seq);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(bytetrack_video_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    if (e != cudaSuccess) {
        tk_set_last_cuda_error((int)e);
        if (h->state) cudaFree(h->state);
        if (h->cost) cudaFree(h->cost);
        delete h;
        return TK_ERR_CUDA;
    }
    *handle = h;
    return tk_bytetrack_reset(h, 0, nullptr);
}

int tk_bytetrack_reset(void* handle, int keep_id_counter, void* stream) {
    if (!handle) return TK_ERR_ARG;
    BtHandle* h = (BtHandle*)handle;
    bytetrack_reset_kernel<<<h->n_seq, 128, 0, (cudaStream_t)stream>>>(h->state, h->state_stride, h->cap, h->first_id,
                                                                      keep_id_counter);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_bytetrack_run(void* handle, const double* dets, const int* offsets, int n_frames, double* out_rows,
                     const int* out_start, int* out_frame_count, int* out_count, void* stream) {
    if (!handle || !offsets || !out_rows || !out_start || !out_frame_count || !out_count || n_frames < 0) return TK_ERR_ARG;
    BtHandle* h = (BtHandle*)handle;
    if (n_frames == 0) return TK_OK;
    // the attribute is per kernel function, not per handle: another handle with smaller capacities may have lowered it
    TK_CUDA_TRY(cudaFuncSetAttribute(bytetrack_video_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    bytetrack_video_kernel<<<h->n_seq, BT_THREADS, h->smem_bytes, (cudaStream_t)stream>>>(
        h->prm, h->state, h->state_stride, h->cap, h->capd, dets, offsets, n_frames, out_rows, out_start,
        out_frame_count, out_count, h->cost, h->cost_stride, h->cost_in_smem);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_bytetrack_status(void* handle, int* status_host, void* stream) {
    if (!handle || !status_host) return TK_ERR_ARG;
    BtHandle* h = (BtHandle*)handle;
    for (int s = 0; s < h->n_seq; ++s) {
        TK_CUDA_TRY(cudaMemcpyAsync(status_host + s, h->state + (size_t)s * h->state_stride + 4 * sizeof(int), sizeof(int),
                                    cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    }
    TK_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return TK_OK;
}

#ifdef TK_PHASE_PROF
int tk_debug_bytetrack_phases(unsigned long long* host_out64, int reset) {
    cudaDeviceSynchronize();
    if (host_out64) cudaMemcpyFromSymbol(host_out64, g_bt_prof, sizeof(unsigned long long) * 64);
    if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(g_bt_prof, z, sizeof(z)); }
    return 0;
}
#endif

int tk_bytetrack_destroy(void* handle) {
    if (!handle) return TK_ERR_ARG;
    BtHandle* h = (BtHandle*)handle;
    cudaFree(h->state);
    if (h->cost) cudaFree(h->cost);
    delete h;
    return TK_OK;
}

}  // extern "C"
