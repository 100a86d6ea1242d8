// StrongSORT (DeepSORT lineage) association with externally supplied appearance features; whole video per launch.
//
// Device restatement of
//   /root/reference/plugins/track/strong_sort/strong_sort.py:41-85,88-121            (update, box conversions, output rule)
//   /root/reference/plugins/track/strong_sort/sort/tracker.py:53-59,80-115,151-193   (predict, update, _match, _initiate_track)
//   /root/reference/plugins/track/strong_sort/sort/track.py:65-95,97-108,245-322     (Track life cycle, EMA feature)
//   /root/reference/plugins/track/strong_sort/sort/kalman_filter.py:47-214            (x/y/a/h-scaled noise, confidence-scaled R)
//   /root/reference/plugins/track/strong_sort/sort/nn_matching.py:30-49,73-91,127-161 (cosine metric, gallery with budget)
//   /root/reference/plugins/track/strong_sort/sort/linear_assignment.py:11-72,131-174 (matching, Mahalanobis gating, fusion)
//   /root/reference/plugins/track/strong_sort/sort/iou_matching.py:7-82               (IoU cost)
// and of the wrapper filter /root/reference/tracklab/wrappers/track/strong_sort_api.py:66-93 (ecc off, max_unmatched_preds 0).
//
// Execution shape. The appearance term is min over a gallery of up to `budget` (100) EMA features per confirmed track:
// T x budget x D x E multiply-adds per frame (164 MFLOP at 40 x 100 x 40 x 512, 4x that with ResNet-50's 2048-d features)
// inside the per-frame dependency chain, of which the Mahalanobis gate (linear_assignment.py:166-174) then overwrites all
// but a few entries per detection with 1e5. So the gate goes first and each video gets a GROUP of CTAs launched
// cooperatively: per frame the master CTA predicts, lists and gates (pairs that survive are published); meanwhile the
// other CTAs L2-normalise the frame's detection features; after a group barrier ALL CTAs evaluate the cosine distance of
// the surviving pairs only, one warp per (pair, 8 gallery rows), min-reduced with ordered-int atomics; a second barrier,
// then the master fuses, solves both assignments (lap.cuh), updates filters (octets), EMA features and galleries, and
// emits rows. Still one launch per chunk of frames, no host round trips.
#include <cooperative_groups.h>
#include "kf_xyah.cuh"
#include "lap.cuh"
#include "lsap_scipy.cuh"
#include "trackkern.h"

namespace {

using namespace tk;

constexpr int SS_THREADS = 256;
enum : unsigned char { SS_FREE = 0, SS_TENTATIVE = 1, SS_CONFIRMED = 2, SS_DELETED = 3 };
constexpr double W_POS = 1.0 / 20, W_VEL = 1.0 / 160, INFTY_COST = 1e5, CHI2_4 = 9.4877;

struct SsParams {
    double max_dist, max_iou_dist, mc_lambda, min_conf;
    float ema_alpha, ema_beta;   // float32(alpha), float32(1 - alpha): python float * float32 array (track.py:286)
    int max_age, n_init, budget, width, height, E;
};

struct SsDev {
    int* hdr;            // 0 next_id, 1 n_tracks, 2 -, 3 -, 4 status, 5 n_free, 6 nd (frame), 7 nconf (frame)
    double *mean, *cov, *conf, *det_id;
    int *hits, *age, *tsu, *track_id, *cls, *g_count, *g_head, *list, *free_list, *conf_list, *det_rows;
    unsigned char *state, *fresh;
    float *smooth, *gallery;   // [cap][E], [cap][budget][E] (gallery rows are stored re-normalised)
    int* app_key;              // [cap][capd] appearance cost of the current frame as ordered-int float keys (row = position in conf_list)
    int* ppack;                // [cap][capd] (confirmed-track position << 8 | detection) of the pairs inside the gate
    float* dfeatn;             // [capd][E] L2-normalised detection features of the frame
    int *task_slot, *task_row, *task_flag;   // feature work of the frame published for the worker CTAs (hdr[8] = count)
    unsigned* bar;             // group barrier: count, generation
};

__host__ __device__ inline size_t ss_al(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ __device__ inline size_t ss_state_bytes(int cap, int capd, int budget, int E) {
    size_t s = ss_al(8 * sizeof(int)) + ss_al(64);
    s += ss_al((size_t)cap * 8 * 8) + ss_al((size_t)cap * 64 * 8) + 2 * ss_al((size_t)cap * 8);
    s += 10 * ss_al((size_t)cap * 4) + ss_al((size_t)capd * 4) + 2 * ss_al((size_t)cap);
    s += ss_al((size_t)cap * E * 4) + ss_al((size_t)cap * budget * E * 4) + ss_al((size_t)cap * capd * 8);
    s += ss_al((size_t)cap * capd * 4) + ss_al((size_t)capd * E * 4) + 3 * ss_al((size_t)cap * 4);
    return s;
}

__host__ __device__ inline SsDev ss_carve(char* base, int cap, int capd, int budget, int E) {
    SsDev d;
    char* p = base;
    d.hdr = (int*)p; p += ss_al(8 * sizeof(int));
    d.bar = (unsigned*)p; p += ss_al(64);
    d.mean = (double*)p; p += ss_al((size_t)cap * 8 * 8);
    d.cov = (double*)p; p += ss_al((size_t)cap * 64 * 8);
    d.conf = (double*)p; p += ss_al((size_t)cap * 8);
    d.det_id = (double*)p; p += ss_al((size_t)cap * 8);
    int** ints[10] = {&d.hits, &d.age, &d.tsu, &d.track_id, &d.cls, &d.g_count, &d.g_head, &d.list, &d.free_list, &d.conf_list};
    for (int k = 0; k < 10; ++k) { *ints[k] = (int*)p; p += ss_al((size_t)cap * 4); }
    d.det_rows = (int*)p; p += ss_al((size_t)capd * 4);
    d.state = (unsigned char*)p; p += ss_al((size_t)cap);
    d.fresh = (unsigned char*)p; p += ss_al((size_t)cap);
    d.smooth = (float*)p; p += ss_al((size_t)cap * E * 4);
    d.gallery = (float*)p; p += ss_al((size_t)cap * budget * E * 4);
    d.app_key = (int*)p; p += ss_al((size_t)cap * capd * 8);
    d.ppack = (int*)p; p += ss_al((size_t)cap * capd * 4);
    d.dfeatn = (float*)p; p += ss_al((size_t)capd * E * 4);
    d.task_slot = (int*)p; p += ss_al((size_t)cap * 4);
    d.task_row = (int*)p; p += ss_al((size_t)cap * 4);
    d.task_flag = (int*)p;
    return d;
}

// float32 L2 norm of a length-E vector by one warp (np.linalg.norm on float32), result in every lane
__device__ __forceinline__ float warp_norm(const float* x, int E) {
    float s = 0.0f;
    for (int k = lane_id(); k < E; k += 32) s = fmaf(x[k], x[k], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return sqrtf(s);
}

// float32 dot product of two length-E vectors by one warp (result in every lane); float4 path when E % 4 == 0
__device__ __forceinline__ float warp_dot(const float* __restrict__ a, const float* __restrict__ b, int E) {
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    if ((E & 3) == 0) {
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll 4
        for (int k = lane_id(); k < (E >> 2); k += 32) {
            const float4 x = a4[k], y = b4[k];
            s0 = fmaf(x.x, y.x, s0); s1 = fmaf(x.y, y.y, s1); s2 = fmaf(x.z, y.z, s2); s3 = fmaf(x.w, y.w, s3);
        }
    } else {
#pragma unroll 4
        for (int k = lane_id(); k < E; k += 32) s0 = fmaf(a[k], b[k], s0);
    }
    float s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// b = features / ||features|| (nn_matching.py:46-48) for the frame's detections into S.dfeatn. The calling CTA filters the
// frame itself (strong_sort_api.py:69) into `rows` (shared memory) so that it does not wait for the master.
__device__ void ss_det_normalise(const SsDev& S, const SsParams& prm, const double* __restrict__ D, int nraw, const float* __restrict__ feats,
                                 int r0, int* rows, int* nd_smem, int wg, int wn) {
    if (warp_id() == 0) {
        const int nd_ = warp_compact(nraw, 0, [&](int i) { return D[i * 7 + 4] > prm.min_conf; }, [&](int i, int p) { rows[p] = i; });
        if (lane_id() == 0) *nd_smem = nd_;
    }
    __syncthreads();
    const int nd = *nd_smem, E = prm.E;
    for (int d = wg; d < nd; d += wn) {
        const float* fv = feats + (size_t)(r0 + rows[d]) * E;
        float* dn = S.dfeatn + (size_t)d * E;
        const float nf = sqrtf(warp_dot(fv, fv, E));
#pragma unroll 4
        for (int k = lane_id(); k < E; k += 32) dn[k] = __fdiv_rn(fv[k], nf);
    }
}

// Feature side of the frame, one warp per track: birth feature (track.py:84), EMA update (track.py:284-288) and, for
// confirmed tracks, the gallery append of the re-normalised EMA feature (nn_matching.py:127-142, what _cosine_distance
// normalises again on every call). Runs on the worker CTAs while the master already predicts the next frame.
constexpr int SS_TASK_BIRTH = 1, SS_TASK_EMA = 2, SS_TASK_APPEND = 4;
__device__ void ss_feature_tasks(const SsDev& S, const SsParams& prm, const float* __restrict__ feats, int wg, int wn) {
    const int n = S.hdr[8], E = prm.E, lane = lane_id();
    for (int t = wg; t < n; t += wn) {
        const int s = S.task_slot[t], fl = S.task_flag[t];
        float* sm = S.smooth + (size_t)s * E;
        if (fl & (SS_TASK_BIRTH | SS_TASK_EMA)) {
            const float* fv = feats + (size_t)S.task_row[t] * E;
            const float nf = warp_norm(fv, E);
            if (fl & SS_TASK_BIRTH) {
                for (int k = lane; k < E; k += 32) sm[k] = __fdiv_rn(fv[k], nf);
            } else {
                for (int k = lane; k < E; k += 32)
                    sm[k] = __fadd_rn(__fmul_rn(prm.ema_alpha, sm[k]), __fmul_rn(prm.ema_beta, __fdiv_rn(fv[k], nf)));
                __syncwarp();
                const float ns = warp_norm(sm, E);
                for (int k = lane; k < E; k += 32) sm[k] = __fdiv_rn(sm[k], ns);
            }
            __syncwarp();
        }
        if (fl & SS_TASK_APPEND) {
            const int h = S.g_head[s];
            float* dst = S.gallery + ((size_t)s * prm.budget + h) * E;
            const float ns = warp_norm(sm, E);
            for (int e = lane; e < E; e += 32) dst[e] = __fdiv_rn(sm[e], ns);
            if (lane == 0) { S.g_head[s] = (h + 1) % prm.budget; if (S.g_count[s] < prm.budget) S.g_count[s] += 1; }
        }
    }
}

// Detection.to_xyah (sort/detection.py:45-52) from the float64 wrapper row: tlwh float32, then float32 ops
__device__ __forceinline__ void det_boxes(const double* d, float* tlwh, double* z) {
    const double cx = (d[0] + d[2]) / 2, cy = (d[1] + d[3]) / 2, w = d[2] - d[0], h = d[3] - d[1];   // xyxy2xywh
    tlwh[0] = (float)(cx - w / 2.0); tlwh[1] = (float)(cy - h / 2.0); tlwh[2] = (float)w; tlwh[3] = (float)h;   // _xywh_to_tlwh
    if (z) {
        z[0] = (double)__fadd_rn(tlwh[0], __fdiv_rn(tlwh[2], 2.0f));
        z[1] = (double)__fadd_rn(tlwh[1], __fdiv_rn(tlwh[3], 2.0f));
        z[2] = (double)__fdiv_rn(tlwh[2], tlwh[3]);
        z[3] = (double)tlwh[3];
    }
}

struct SsShared { int lap_ok, nd, nconf, ncand, nud, npairs, n_out, n_ut, nap, nd_w; };

__global__ void __launch_bounds__(SS_THREADS)
strongsort_video_kernel(SsParams prm, char* state_base, size_t state_stride, int cap, int capd, int ncta,
                        const double* __restrict__ dets, const float* __restrict__ feats, const int* __restrict__ offsets,
                        int n_frames, double* __restrict__ out_rows, const int* __restrict__ out_start,
                        int* __restrict__ out_frame_count, int* __restrict__ out_count, int out_cap, const float* __restrict__ warps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int seq = blockIdx.x / ncta, cta = blockIdx.x % ncta, tid = threadIdx.x;
    const bool master = cta == 0;
    const int E = prm.E;
    SsDev S = ss_carve(state_base + (size_t)seq * state_stride, cap, capd, prm.budget, E);
    unsigned char* sp = smem_raw;
    auto take = [&](size_t bytes) { unsigned char* p = sp; sp += (bytes + 15) & ~(size_t)15; return p; };
    const int side = cap > capd ? cap : capd;
    // master only (workers: tmp_d for their own copy of the filtered detection list)
    double* cost = (double*)take(sizeof(double) * (size_t)(cap + 1) * (capd + 1));
    double* lap_u = (double*)take(sizeof(double) * side);
    double* d_z = (double*)take(sizeof(double) * 4 * capd);
    float* d_tlwh = (float*)take(sizeof(float) * 4 * capd);
    double* chol = (double*)take(sizeof(double) * 24 * cap);        // per confirmed row: mean4, L(16), invd(4)
    int* match_a = (int*)take(sizeof(int) * side);
    int* match_b = (int*)take(sizeof(int) * side);
    int* col4row = (int*)take(sizeof(int) * side);
    int* row4col = (int*)take(sizeof(int) * side);
    int* path = (int*)take(sizeof(int) * side);
    int* cand = (int*)take(sizeof(int) * cap);
    int* un_d = (int*)take(sizeof(int) * capd);
    int* tmp_d = (int*)take(sizeof(int) * capd);
    int* pair_t = (int*)take(sizeof(int) * cap);
    int* pair_d = (int*)take(sizeof(int) * cap);
    int* out_pos = (int*)take(sizeof(int) * cap);
    int* upd_row = (int*)take(sizeof(int) * cap);                   // absolute detection row that updated / created the slot this frame
    unsigned char* t_flag = (unsigned char*)take(cap);
    unsigned char* d_flag = (unsigned char*)take(capd);
    // scratch of the scipy-exact assignment (lsap_scipy.cuh) + bookkeeping of the pairs it returns
    double* lap_v = (double*)take(sizeof(double) * side);
    double* lap_spc = (double*)take(sizeof(double) * side);
    int* lap_rem = (int*)take(sizeof(int) * side);
    int* rej_t = (int*)take(sizeof(int) * side);                    // detection of a row's rejected pair (cost above the threshold) or -1
    unsigned char* lap_sr = (unsigned char*)take(side);
    unsigned char* lap_sc = (unsigned char*)take(side);
    unsigned char* touch_d = (unsigned char*)take(side);            // 1 when the solver paired the detection column (accepted or not)
    SsShared* sh = (SsShared*)take(sizeof(SsShared));

    int* status = &S.hdr[4];
    const int F1 = n_frames + 1;
    const int out_base = out_start[seq];
    int out_n = out_count[seq];
    const double L_app = prm.max_dist + 1e-5, L_iou = prm.max_iou_dist + 1e-5;

    const bool worker = !master || ncta == 1;
    const int wg = (ncta == 1 ? 0 : cta - 1) * (SS_THREADS / 32) + warp_id(), wn = (ncta == 1 ? 1 : ncta - 1) * (SS_THREADS / 32);
    bool pending = false;   // feature tasks of the last processed frame still to be applied

    for (int f = 0; f < n_frames; ++f) {
        const int r0 = offsets[seq * F1 + f], r1 = offsets[seq * F1 + f + 1];
        const int nraw = r1 - r0;
        // ---- camera motion compensation (strong_sort_api.py:62-65 -> tracker.camera_update -> track.py:224-239): every track's
        // box corners go through the frame pair's warp (tk_ecc_euclidean) before anything else, also on frames without detections.
        // A NaN first entry = no previous frame / ECC failed: the reference leaves the tracks alone.
        if (master && warps != nullptr) {
            const float* wm = warps + ((size_t)seq * n_frames + f) * 6;
            if (wm[0] == wm[0]) {
                double a[6];
                for (int i = 0; i < 6; ++i) a[i] = (double)wm[i];
                const double dist = sqrt((1 - a[0]) * (1 - a[0]) + a[1] * a[1] + a[2] * a[2] + a[3] * a[3] + (1 - a[4]) * (1 - a[4]) + a[5] * a[5]);
                if (!(dist < 100)) { a[0] = 1; a[1] = 0; a[2] = 0; a[3] = 0; a[4] = 1; a[5] = 0; }   // get_matrix (track.py:216-222)
                const int ntk = S.hdr[1];
                for (int k = tid; k < ntk; k += SS_THREADS) {
                    const int s = S.list[k];
                    double* m = S.mean + (size_t)s * 8;
                    double x1, y1, x2, y2;
                    if (S.fresh[s]) {   // mean is a float32 array until the first predict (kalman_filter.py:47-78): float32 box arithmetic
                        const float cx = (float)m[0], cy = (float)m[1], w = __fmul_rn((float)m[2], (float)m[3]), hh = (float)m[3];
                        const float l = __fsub_rn(cx, __fdiv_rn(w, 2.0f)), t = __fsub_rn(cy, __fdiv_rn(hh, 2.0f));
                        x1 = (double)l; y1 = (double)t; x2 = (double)__fadd_rn(l, w); y2 = (double)__fadd_rn(t, hh);
                    } else {
                        const double w = m[2] * m[3];
                        const double l = m[0] - w / 2, t = m[1] - m[3] / 2;   // to_tlwh (track.py:97-101), to_tlbr (:114-126)
                        x1 = l; y1 = t; x2 = l + w; y2 = t + m[3];
                    }
                    const double x1_ = a[0] * x1 + a[1] * y1 + a[2], y1_ = a[3] * x1 + a[4] * y1 + a[5];
                    const double x2_ = a[0] * x2 + a[1] * y2 + a[2], y2_ = a[3] * x2 + a[4] * y2 + a[5];
                    const double w = x2_ - x1_, hh = y2_ - y1_;
                    double o0 = x1_ + w / 2, o1 = y1_ + hh / 2, o2 = w / hh, o3 = hh;
                    if (S.fresh[s]) { o0 = (double)(float)o0; o1 = (double)(float)o1; o2 = (double)(float)o2; o3 = (double)(float)o3; }
                    m[0] = o0; m[1] = o1; m[2] = o2; m[3] = o3;
                }
            }
            __syncthreads();
        }
        if (nraw == 0) { if (master && tid == 0) out_frame_count[seq * n_frames + f] = 0; continue; }   // strong_sort_api.py:66-67
        if (nraw > capd) { if (master && tid == 0) atomicOr(status, TK_DEV_OVERFLOW_DETS); break; }
        const double* D = dets + (size_t)r0 * 7;

        if (!master) {   // workers: last frame's feature tasks, this frame's normalised detection features (overlaps the master's predict + gate)
            if (pending) ss_feature_tasks(S, prm, feats, wg, wn);
            ss_det_normalise(S, prm, D, nraw, feats, r0, tmp_d, &sh->nd_w, wg, wn);
            __threadfence();
        }
        // ================= master: predict, detections, lists, gate =================
        if (master) {
            const int nt = S.hdr[1];
            for (int base = 0; base < nt; base += SS_THREADS / 8) {   // Track.predict (track.py:245-249, kalman_filter.py:80-112)
                const int k = base + (tid >> 3), j = tid & 7;
                const bool act = k < nt;
                const int s = act ? S.list[k] : 0;
                double* gm = S.mean + (size_t)s * 8;
                double qj = 0.0;
                if (act) {
                    const double ref = gm[j & 3];           // noise of x,y,a,h (and their velocities) scales with x,y,a,h
                    if (S.fresh[s]) {                         // float32 mean right after initiate
                        const float r32 = (float)ref;
                        const float w = (j & 3) == 2 ? (j < 4 ? 1.0f : (float)0.1) : (j < 4 ? (float)W_POS : (float)W_VEL);
                        const float sd = ((j & 3) == 2 && j < 4) ? r32 : __fmul_rn(w, r32);
                        qj = (double)__fmul_rn(sd, sd);
                    } else {
                        const double w = (j & 3) == 2 ? (j < 4 ? 1.0 : 0.1) : (j < 4 ? W_POS : W_VEL);
                        const double sd = w * ref;
                        qj = sd * sd;
                    }
                }
                kf8_octet_predict(gm, S.cov + (size_t)s * 64, act, false, qj);
                if (act && j == 0) { S.age[s] += 1; S.tsu[s] += 1; S.fresh[s] = 0; }
            }
            if (warp_id() == 0) {   // wrapper filter (strong_sort_api.py:69), detection order kept
                const int nd_ = warp_compact(nraw, 0, [&](int i) { return D[i * 7 + 4] > prm.min_conf; }, [&](int i, int p) { S.det_rows[p] = i; });
                if (lane_id() == 0) { sh->nd = nd_; S.hdr[6] = nd_; }
            }
            __syncthreads();
            const int nd = sh->nd;
            for (int i = tid; i < nd; i += SS_THREADS) det_boxes(D + (size_t)S.det_rows[i] * 7, d_tlwh + 4 * i, d_z + 4 * i);
            if (warp_id() == 0) {
                const int nc = warp_compact(nt, 0, [&](int k) { return S.state[S.list[k]] == SS_CONFIRMED; },
                                            [&](int k, int p) { S.conf_list[p] = S.list[k]; });
                if (lane_id() == 0) { sh->nconf = nc; S.hdr[7] = nc; sh->nap = 0; }
            }
            if (ncta == 1) {
                if (pending) ss_feature_tasks(S, prm, feats, wg, wn);
                ss_det_normalise(S, prm, D, nraw, feats, r0, tmp_d, &sh->nd_w, wg, wn);
            }
            __syncthreads();
            {
                const int nc = sh->nconf, nd2 = sh->nd;
                for (int r = tid; r < nc; r += SS_THREADS) {   // projected distribution of each confirmed track (confidence 0)
                    const int s = S.conf_list[r];
                    const double* m = S.mean + (size_t)s * 8;
                    const double* P = S.cov + (size_t)s * 64;
                    const double sp_ = W_POS * m[3];
                    const double rr[4] = {sp_ * sp_, sp_ * sp_, 1e-1 * 1e-1, sp_ * sp_};
                    double Pl[64];
                    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) Pl[i * 8 + j] = P[i * 8 + j];
                    double L[16], Sm[16], invd[4];
                    if (!kf8_chol4(Pl, rr, L, Sm, invd)) atomicOr(status, TK_DEV_BAD_CHOLESKY);
                    double* c = chol + 24 * r;
                    for (int i = 0; i < 4; ++i) c[i] = m[i];
                    for (int i = 0; i < 16; ++i) c[4 + i] = L[i];
                    for (int i = 0; i < 4; ++i) c[20 + i] = invd[i];
                }
                __syncthreads();
                // Mahalanobis gate first (linear_assignment.py:166-174): only pairs inside it need an appearance distance
                for (int e0 = 0; e0 < nc * nd2; e0 += SS_THREADS) {
                    const int e = e0 + tid;
                    bool in = false;
                    int r = 0, d = 0;
                    if (e < nc * nd2) {
                        r = e / nd2; d = e - r * nd2;
                        const double* c = chol + 24 * r;
                        in = !(kf8_maha(c, c + 4, c + 20, d_z + 4 * d) > CHI2_4);
                        S.app_key[(size_t)r * capd + d] = 0x7fffffff;
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, in);
                    if (m) {
                        int base = 0;
                        if (lane_id() == 0) base = atomicAdd(&sh->nap, __popc(m));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (in) S.ppack[base + __popc(m & ((1u << lane_id()) - 1u))] = (r << 8) | d;
                    }
                }
                __syncthreads();
                if (tid == 0) S.hdr[2] = sh->nap;
            }
            __threadfence();
        }
        group_barrier(S.bar, ncta);

        // ================= all CTAs: appearance cost of the gated-in pairs (nn_matching.py:30-49,73-91,144-161) =================
        // item = (pair, 8 gallery rows of its track): 1 - <gallery row, normalised detection>, min over the rows with an
        // ordered-int atomic per item (the gallery rows are stored re-normalised, as _cosine_distance does on every call).
        {
            const int nap = S.hdr[2], budget = prm.budget;
            const int nchunk = (budget + 7) >> 3;
            const int gw = cta * (SS_THREADS / 32) + warp_id(), nw = ncta * (SS_THREADS / 32);
            for (int it = gw; it < nap * nchunk; it += nw) {
                const int p = it / nchunk, ch = it - p * nchunk;
                const int pk = S.ppack[p], r = pk >> 8, d = pk & 255;
                const int slot = S.conf_list[r];
                const int j0 = ch * 8, j1 = min(S.g_count[slot], j0 + 8);
                if (j0 >= j1) continue;
                const float* dn = S.dfeatn + (size_t)d * E;
                const float* g = S.gallery + ((size_t)slot * budget + j0) * E;
                float best = __int_as_float(0x7f800000);
                for (int j = j0; j < j1; ++j, g += E) best = fminf(best, __fsub_rn(1.0f, warp_dot(g, dn, E)));
                if (lane_id() == 0) {
                    int key = __float_as_int(best);
                    key = key >= 0 ? key : key ^ 0x7fffffff;
                    atomicMin(&S.app_key[(size_t)r * capd + d], key);
                }
            }
            __threadfence();
        }
        group_barrier(S.bar, ncta);
        pending = true;
        if (!master) { group_barrier(S.bar, ncta); continue; }   // third barrier: the master has published the frame's feature tasks

        // ================= master: gating, fusion, assignments, updates =================
        const int nd = sh->nd, nconf = sh->nconf, nt = S.hdr[1];
        // ---- stage A: confirmed tracks x all detections (tracker.py:152-170, linear_assignment.py:131-174)
        {
            const bool a_rows = nconf <= nd;
            const int ld = lap_pitch(a_rows ? nd : nconf);
            for (int e = tid; e < nconf * nd; e += SS_THREADS) {
                const int r = e / nd, d = e - r * nd;
                const double* c = chol + 24 * r;
                const double g = kf8_maha(c, c + 4, c + 20, d_z + 4 * d);
                int akey = S.app_key[(size_t)r * capd + d];
                akey = akey >= 0 ? akey : akey ^ 0x7fffffff;
                double a = (double)__int_as_float(akey);
                if (g > CHI2_4) a = INFTY_COST;
                const double fused = __dadd_rn(__dmul_rn(prm.mc_lambda, a), __dmul_rn(1.0 - prm.mc_lambda, g));
                const double cl = fused > prm.max_dist ? L_app : fused;   // cost_matrix[cost_matrix > max_distance] = max_distance + 1e-5
                if (a_rows) cost[(size_t)r * ld + d] = cl; else cost[(size_t)d * ld + r] = cl;
            }
            for (int i = tid; i < nconf; i += SS_THREADS) { match_a[i] = -1; rej_t[i] = -1; }
            for (int i = tid; i < nd; i += SS_THREADS) { match_b[i] = -1; touch_d[i] = 0; }
            if (tid == 0) sh->lap_ok = 1;
            __syncthreads();
            if (nconf > 0 && nd > 0) {
                // scipy.optimize.linear_sum_assignment on the full clamped matrix (linear_assignment.py:53-55), same tie-breaking:
                // the solver pairs every row of the smaller side; pairs above the threshold are "rejected" (:62-68)
                const int nr = a_rows ? nconf : nd, nc = a_rows ? nd : nconf;
                if (warp_id() == 0) {
                    const bool ok = lsap_scipy_warp(nr, nc, [&](int i, int j) { return cost[(size_t)i * ld + j]; }, lap_u, lap_v, lap_spc, path,
                                                    col4row, row4col, lap_rem, lap_sr, lap_sc);
                    if (!ok && lane_id() == 0) { atomicOr(status, TK_DEV_LAP_INFEASIBLE); sh->lap_ok = 0; }
                }
                __syncthreads();
                if (sh->lap_ok) for (int i = tid; i < nr; i += SS_THREADS) {
                    const int j = col4row[i];
                    const int r = a_rows ? i : j, d = a_rows ? j : i;
                    touch_d[d] = 1;
                    if (cost[(size_t)i * ld + j] > prm.max_dist) rej_t[r] = d; else { match_a[r] = d; match_b[d] = r; }
                }
            }
            __syncthreads();
        }
        // record stage-A pairs, build stage-B candidates: unconfirmed + unmatched confirmed with tsu == 1 (tracker.py:171-178)
        if (warp_id() == 0) {
            const int np_ = warp_compact(nconf, 0, [&](int r) { return match_a[r] >= 0; },
                                         [&](int r, int p) { pair_t[p] = S.conf_list[r]; pair_d[p] = match_a[r]; });
            int nc = warp_compact(nt, 0, [&](int k) { return S.state[S.list[k]] != SS_CONFIRMED; }, [&](int k, int p) { cand[p] = S.list[k]; });
            nc = warp_compact(nconf, nc, [&](int r) { return match_a[r] < 0 && S.tsu[S.conf_list[r]] == 1; },
                              [&](int r, int p) { cand[p] = S.conf_list[r]; });
            // unmatched_detections in the reference's order (linear_assignment.py:57-68): columns the solver left alone, in
            // detection order, then the detections of the rejected pairs in ROW (track list) order
            int nu = warp_compact(nd, 0, [&](int d) { return !touch_d[d]; }, [&](int d, int p) { un_d[p] = d; });
            nu = warp_compact(nconf, nu, [&](int r) { return rej_t[r] >= 0; }, [&](int r, int p) { un_d[p] = rej_t[r]; });
            if (lane_id() == 0) { sh->npairs = np_; sh->ncand = nc; sh->nud = nu; }
        }
        for (int k = tid; k < cap; k += SS_THREADS) t_flag[k] = 0;     // t_flag[slot] = 1 when the track got a detection this frame
        __syncthreads();
        for (int p = tid; p < sh->npairs; p += SS_THREADS) t_flag[pair_t[p]] = 1;
        // ---- stage B: IoU cost on the candidates (iou_matching.py:42-82, linear_assignment.py:11-72)
        {
            const int ncand = sh->ncand, nud = sh->nud;
            const bool a_rows = ncand <= nud;
            const int ld = lap_pitch(a_rows ? nud : ncand);
            for (int e = tid; e < ncand * nud; e += SS_THREADS) {
                const int r = e / nud, c = e - r * nud;
                const int s = cand[r];
                double v;
                if (S.tsu[s] > 1) v = INFTY_COST;
                else {
                    const double* m = S.mean + (size_t)s * 8;
                    const double w = m[2] * m[3];
                    const double bx = m[0] - w / 2, by = m[1] - m[3] / 2;            // Track.to_tlwh (track.py:97-101)
                    const float* cb = d_tlwh + 4 * un_d[c];
                    const double x0 = fmax(bx, (double)cb[0]), y0 = fmax(by, (double)cb[1]);
                    const double x1 = fmin(bx + w, (double)__fadd_rn(cb[0], cb[2])), y1 = fmin(by + m[3], (double)__fadd_rn(cb[1], cb[3]));
                    const double iw = fmax(0.0, x1 - x0), ih = fmax(0.0, y1 - y0);
                    const double inter = __dmul_rn(iw, ih);
                    const double uni = __dsub_rn(__dadd_rn(__dmul_rn(w, m[3]), (double)__fmul_rn(cb[2], cb[3])), inter);
                    v = 1.0 - inter / uni;
                }
                const double cl = v > prm.max_iou_dist ? L_iou : v;
                if (a_rows) cost[(size_t)r * ld + c] = cl; else cost[(size_t)c * ld + r] = cl;
            }
            for (int i = tid; i < ncand; i += SS_THREADS) { match_a[i] = -1; rej_t[i] = -1; }
            for (int i = tid; i < nud; i += SS_THREADS) { match_b[i] = -1; touch_d[i] = 0; }
            if (tid == 0) sh->lap_ok = 1;
            __syncthreads();
            if (ncand > 0 && nud > 0) {
                const int nr = a_rows ? ncand : nud, nc = a_rows ? nud : ncand;
                if (warp_id() == 0) {
                    const bool ok = lsap_scipy_warp(nr, nc, [&](int i, int j) { return cost[(size_t)i * ld + j]; }, lap_u, lap_v, lap_spc, path,
                                                    col4row, row4col, lap_rem, lap_sr, lap_sc);
                    if (!ok && lane_id() == 0) { atomicOr(status, TK_DEV_LAP_INFEASIBLE); sh->lap_ok = 0; }
                }
                __syncthreads();
                if (sh->lap_ok) for (int i = tid; i < nr; i += SS_THREADS) {
                    const int j = col4row[i];
                    const int r = a_rows ? i : j, c = a_rows ? j : i;
                    touch_d[c] = 1;
                    if (cost[(size_t)i * ld + j] > prm.max_iou_dist) rej_t[r] = c; else { match_a[r] = c; match_b[c] = r; }
                }
            }
            __syncthreads();
            if (warp_id() == 0) {
                const int np_ = warp_compact(ncand, sh->npairs, [&](int r) { return match_a[r] >= 0; },
                                             [&](int r, int p) { pair_t[p] = cand[r]; pair_d[p] = un_d[match_a[r]]; });
                // detections that stay unmatched, in the reference's order: untouched columns first, then the rejected pairs in row order
                int nu = warp_compact(nud, 0, [&](int c) { return !touch_d[c]; }, [&](int c, int p) { tmp_d[p] = un_d[c]; });
                nu = warp_compact(ncand, nu, [&](int r) { return rej_t[r] >= 0; }, [&](int r, int p) { tmp_d[p] = un_d[rej_t[r]]; });
                if (lane_id() == 0) { sh->npairs = np_; sh->nud = nu; }
            }
            __syncthreads();
            for (int p = tid; p < sh->npairs; p += SS_THREADS) t_flag[pair_t[p]] = 1;
            __syncthreads();
        }
        // ---- Track.update for every pair (track.py:272-301, kalman_filter.py:114-174)
        {
            const int np_ = sh->npairs;
            for (int base = 0; base < np_; base += SS_THREADS / 8) {
                const int p = base + (tid >> 3);
                const bool act = p < np_;
                const int s = act ? pair_t[p] : 0, d = act ? pair_d[p] : 0;
                double z[4] = {0, 0, 0, 0}, r[4] = {1, 1, 1, 1};
                if (act) {
                    const double* dz = d_z + 4 * d;
                    z[0] = dz[0]; z[1] = dz[1]; z[2] = dz[2]; z[3] = dz[3];
                    const double cf = D[(size_t)S.det_rows[d] * 7 + 4];
                    const double h = S.mean[(size_t)s * 8 + 3];
                    const double sp_ = (1 - cf) * (W_POS * h), sa = (1 - cf) * 1e-1;
                    r[0] = sp_ * sp_; r[1] = sp_ * sp_; r[2] = sa * sa; r[3] = sp_ * sp_;
                }
                if (!kf8_octet_update(S.mean + (size_t)s * 8, S.cov + (size_t)s * 64, act, z, r)) atomicOr(status, TK_DEV_BAD_CHOLESKY);
                if (act && (tid & 7) == 0) {
                    const double* dr = D + (size_t)S.det_rows[d] * 7;
                    S.conf[s] = dr[4]; S.cls[s] = (int)dr[5]; S.det_id[s] = dr[6];
                    S.hits[s] += 1; S.tsu[s] = 0; upd_row[s] = r0 + S.det_rows[d];
                    if (S.state[s] == SS_TENTATIVE && S.hits[s] >= prm.n_init) S.state[s] = SS_CONFIRMED;
                }
            }
            __syncthreads();
        }
        // ---- mark_missed (track.py:303-309), births (tracker.py:189-193), prune, gallery append (nn_matching.py:127-142)
        if (warp_id() == 0) {
            const int lane = lane_id();
            for (int k = lane; k < nt; k += 32) {
                const int s = S.list[k];
                if (!t_flag[s]) { if (S.state[s] == SS_TENTATIVE || S.tsu[s] > prm.max_age) S.state[s] = SS_DELETED; }
            }
            __syncwarp();
            // births in the order of the reference's unmatched_detections (tmp_d: untouched columns, then the rejected pairs in
            // row order — the pairs scipy's tie-breaking picks among the equal clamped costs, reproduced by lsap_scipy.cuh)
            const int nfree = S.hdr[5];
            const int nb = sh->nud < nfree ? sh->nud : nfree;
            if (sh->nud > nfree && lane == 0) atomicOr(status, TK_DEV_OVERFLOW_TRACKS);
            const int id0 = S.hdr[0];
            int n = nt;
            for (int k = lane; k < nb; k += 32) {
                const int s = S.free_list[nfree - 1 - k];
                const int d = tmp_d[k];
                const double* dr = D + (size_t)S.det_rows[d] * 7;
                S.list[nt + k] = s;
                S.track_id[s] = id0 + k; S.cls[s] = (int)dr[5]; S.conf[s] = dr[4]; S.det_id[s] = dr[6];
                S.hits[s] = 1; S.age[s] = 1; S.tsu[s] = 0; S.state[s] = SS_TENTATIVE; S.fresh[s] = 1;
                S.g_count[s] = 0; S.g_head[s] = 0; t_flag[s] = 2; upd_row[s] = r0 + S.det_rows[d];
                pair_t[k] = s; pair_d[k] = d;     // reuse as the birth list for the parallel initialisation below
            }
            n += nb;
            if (lane == 0) { S.hdr[0] = id0 + nb; S.hdr[5] = nfree - nb; sh->n_ut = nb; sh->npairs = n; }
        }
        __syncthreads();
        {
            const int nb = sh->n_ut;
            for (int k = tid; k < nb; k += SS_THREADS) {   // KalmanFilter.initiate on the float32 measurement (kalman_filter.py:47-78)
                const int s = pair_t[k];
                const double* z = d_z + 4 * pair_d[k];
                double* m = S.mean + (size_t)s * 8;
                double* P = S.cov + (size_t)s * 64;
                for (int i = 0; i < 4; ++i) { m[i] = z[i]; m[4 + i] = 0.0; }
                for (int i = 0; i < 64; ++i) P[i] = 0.0;
                const float zf[4] = {(float)z[0], (float)z[1], (float)z[2], (float)z[3]};
                const float sd[8] = {__fmul_rn((float)(2 * W_POS), zf[0]), __fmul_rn((float)(2 * W_POS), zf[1]), zf[2], __fmul_rn((float)(2 * W_POS), zf[3]),
                                     __fmul_rn((float)(10 * W_VEL), zf[0]), __fmul_rn((float)(10 * W_VEL), zf[1]), __fmul_rn((float)0.1, zf[2]),
                                     __fmul_rn((float)(10 * W_VEL), zf[3])};
                for (int i = 0; i < 8; ++i) P[i * 9] = (double)__fmul_rn(sd[i], sd[i]);
            }
        }
        __syncthreads();
        if (warp_id() == 0) {   // tracks = [t for t in tracks if not deleted] (order kept); free the deleted slots
            const int n0 = sh->npairs;
            int nfree = S.hdr[5];
            nfree = warp_compact(n0, nfree, [&](int k) { return S.state[S.list[k]] == SS_DELETED; },
                                 [&](int k, int p) { const int s = S.list[k]; S.free_list[p] = s; });
            const int n = warp_compact(n0, 0, [&](int k) { return S.state[S.list[k]] != SS_DELETED; },
                                       [&](int k, int p) { const int s = S.list[k]; cand[p] = s; });
            for (int k = lane_id(); k < n; k += 32) S.list[k] = cand[k];
            __syncwarp();
            for (int k = lane_id(); k < n; k += 32) out_pos[k] = -1;
            __syncwarp();
            const int cnt = warp_compact(n, 0, [&](int k) { const int s = S.list[k]; return S.state[s] == SS_CONFIRMED && S.tsu[s] <= 1; },
                                         [&](int k, int p) { out_pos[k] = p; });
            const int ntask = warp_compact(n, 0, [&](int k) { const int s = S.list[k]; return t_flag[s] != 0 || S.state[s] == SS_CONFIRMED; },
                                           [&](int k, int p) {
                                               const int s = S.list[k];
                                               S.task_slot[p] = s; S.task_row[p] = upd_row[s];
                                               S.task_flag[p] = (t_flag[s] == 2 ? SS_TASK_BIRTH : (t_flag[s] == 1 ? SS_TASK_EMA : 0)) | (S.state[s] == SS_CONFIRMED ? SS_TASK_APPEND : 0);
                                           });
            if (lane_id() == 0) {
                S.hdr[8] = ntask;
                S.hdr[1] = n; S.hdr[5] = nfree; sh->n_out = cnt;
                out_frame_count[seq * n_frames + f] = cnt;
                if (out_n + cnt > out_cap) atomicOr(status, TK_DEV_OVERFLOW_OUT);
            }
        }
        __syncthreads();
        {
            const int n = S.hdr[1];
            // output rows (strong_sort.py:70-82,110-121)
            if (out_n + sh->n_out <= out_cap)
                for (int k = tid; k < n; k += SS_THREADS) {
                    if (out_pos[k] < 0) continue;
                    const int s = S.list[k];
                    const double* m = S.mean + (size_t)s * 8;
                    const double w = m[2] * m[3];
                    const double x = m[0] - w / 2, y = m[1] - m[3] / 2;
                    double* o = out_rows + (size_t)(out_base + out_n + out_pos[k]) * 8;
                    o[0] = (double)max((int)x, 0); o[2] = (double)min((int)(x + w), prm.width - 1);
                    o[1] = (double)max((int)y, 0); o[3] = (double)min((int)(y + m[3]), prm.height - 1);
                    o[4] = (double)S.track_id[s]; o[5] = (double)S.cls[s]; o[6] = S.conf[s]; o[7] = S.det_id[s];
                }
            out_n += sh->n_out;
        }
        __syncthreads();
        __threadfence();
        group_barrier(S.bar, ncta);   // third barrier of the frame: the task list is visible to the workers
    }
    if (worker && pending) ss_feature_tasks(S, prm, feats, wg, wn);   // tasks of the last frame of the launch
    if (master && tid == 0) out_count[seq] = out_n;
}

struct SsHandle {
    SsParams prm;
    int n_seq, cap, capd, ncta;
    char* state;
    size_t state_stride, smem_bytes;
};

__global__ void strongsort_reset_kernel(char* base, size_t stride, int cap, int capd, int budget, int E) {
    SsDev S = ss_carve(base + (size_t)blockIdx.x * stride, cap, capd, budget, E);
    if (threadIdx.x == 0) { S.hdr[0] = 1; S.hdr[1] = 0; S.hdr[4] = 0; S.hdr[5] = cap; S.hdr[6] = 0; S.hdr[7] = 0; S.bar[0] = 0; S.bar[1] = 0; }
    for (int i = threadIdx.x; i < cap; i += blockDim.x) { S.free_list[i] = cap - 1 - i; S.state[i] = SS_FREE; S.fresh[i] = 0; S.g_count[i] = 0; S.g_head[i] = 0; }
}

size_t ss_smem(int cap, int capd) {
    const int side = cap > capd ? cap : capd;
    auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
    size_t s = 0;
    s += al(8 * (size_t)(cap + 1) * (capd + 1)) + al(8 * side) + al(8 * 4 * capd) + al(4 * 4 * capd) + al(8 * 24 * cap);
    s += 5 * al(4 * side) + al(4 * cap) + 2 * al(4 * capd) + 4 * al(4 * cap) + al(cap) + al(capd) + al(sizeof(SsShared));
    s += 2 * al(8 * side) + 2 * al(4 * side) + 3 * al(side);   // lsap_scipy scratch
    return s;
}

}  // namespace

extern "C" {

int tk_strongsort_create(const tk_strongsort_params* p, int n_seq, int cap_tracks, int cap_dets, void** handle) {
    if (!p || !handle || n_seq <= 0 || cap_tracks <= 0 || cap_dets <= 0 || p->feature_dim <= 0 || p->nn_budget <= 0) return TK_ERR_ARG;
    if (cap_tracks > tk::LAP_MAX_COLS || cap_dets > tk::LAP_MAX_COLS) return TK_ERR_CAPACITY;
    if (p->max_unmatched_preds != 0) return TK_ERR_ARG;   // only the reference configuration (strong_sort.yaml:19)
    SsHandle* h = new SsHandle();
    h->prm.max_dist = p->max_dist; h->prm.max_iou_dist = p->max_iou_dist; h->prm.mc_lambda = p->mc_lambda;
    h->prm.min_conf = p->min_confidence; h->prm.ema_alpha = (float)p->ema_alpha; h->prm.ema_beta = (float)(1 - p->ema_alpha);
    h->prm.max_age = p->max_age; h->prm.n_init = p->n_init; h->prm.budget = p->nn_budget;
    h->prm.width = p->image_width; h->prm.height = p->image_height; h->prm.E = p->feature_dim;
    h->n_seq = n_seq; h->cap = cap_tracks; h->capd = cap_dets;
    h->ncta = p->ctas_per_video > 0 ? p->ctas_per_video : 8;
    h->state_stride = (ss_state_bytes(cap_tracks, cap_dets, p->nn_budget, p->feature_dim) + 255) & ~(size_t)255;
    h->smem_bytes = ss_smem(cap_tracks, cap_dets);
    h->state = nullptr;
    if (h->smem_bytes > 220 * 1024) { delete h; return TK_ERR_CAPACITY; }
    cudaError_t e = cudaMalloc((void**)&h->state, h->state_stride * n_seq);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(strongsort_video_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    if (e != cudaSuccess) { tk_set_last_cuda_error((int)e); if (h->state) cudaFree(h->state); delete h; return TK_ERR_CUDA; }
    *handle = h;
    return tk_strongsort_reset(h, 0, nullptr);
}

int tk_strongsort_reset(void* handle, int keep_id_counter, void* stream) {
    (void)keep_id_counter;   // Tracker._next_id restarts with every StrongSORT() (tracker.py:51)
    if (!handle) return TK_ERR_ARG;
    SsHandle* h = (SsHandle*)handle;
    strongsort_reset_kernel<<<h->n_seq, 128, 0, (cudaStream_t)stream>>>(h->state, h->state_stride, h->cap, h->capd, h->prm.budget, h->prm.E);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_strongsort_run(void* handle, const double* dets, const float* features, const int* offsets, int n_frames, double* out_rows,
                      const int* out_start, int* out_frame_count, int* out_count, int out_capacity_rows, void* stream) {
    return tk_strongsort_run_cmc(handle, dets, features, offsets, n_frames, nullptr, out_rows, out_start, out_frame_count, out_count,
                                 out_capacity_rows, stream);
}

int tk_strongsort_run_cmc(void* handle, const double* dets, const float* features, const int* offsets, int n_frames, const float* warps,
                          double* out_rows, const int* out_start, int* out_frame_count, int* out_count, int out_capacity_rows, void* stream) {
    if (!handle || !offsets || !out_rows || !out_start || !out_frame_count || !out_count || n_frames < 0) return TK_ERR_ARG;
    SsHandle* h = (SsHandle*)handle;
    if (n_frames == 0) return TK_OK;
    int cap = h->cap, capd = h->capd, ncta = h->ncta;
    void* args[] = {&h->prm, &h->state, &h->state_stride, &cap, &capd, &ncta, &dets, &features, &offsets, &n_frames,
                    &out_rows, &out_start, &out_frame_count, &out_count, &out_capacity_rows, &warps};
    // cooperative launch: every CTA of a video group must be co-resident for the group barrier
    TK_CUDA_TRY(cudaFuncSetAttribute(strongsort_video_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));   // per function, not per handle
    TK_CUDA_TRY(cudaLaunchCooperativeKernel((void*)strongsort_video_kernel, dim3(h->n_seq * ncta), dim3(SS_THREADS), args,
                                            h->smem_bytes, (cudaStream_t)stream));
    return TK_OK;
}

int tk_strongsort_status(void* handle, int* status_host, void* stream) {
    if (!handle || !status_host) return TK_ERR_ARG;
    SsHandle* h = (SsHandle*)handle;
    for (int s = 0; s < h->n_seq; ++s)
        TK_CUDA_TRY(cudaMemcpyAsync(status_host + s, h->state + (size_t)s * h->state_stride + 4 * sizeof(int), sizeof(int),
                                    cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    TK_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return TK_OK;
}

int tk_strongsort_destroy(void* handle) {
    if (!handle) return TK_ERR_ARG;
    SsHandle* h = (SsHandle*)handle;
    cudaFree(h->state);
    delete h;
    return TK_OK;
}

}  // extern "C"
