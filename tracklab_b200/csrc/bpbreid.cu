// BPBReID-StrongSORT association (part-based appearance, visibility-aware EMA); whole video per launch.
//
// Device restatement of
//   /root/reference/plugins/track/bpbreid_strong_sort/strong_sort.py:53-141             (update, filter, output rule)
//   /root/reference/plugins/track/bpbreid_strong_sort/sort/tracker.py:92-99,123-167,242-333,409-441
//   /root/reference/plugins/track/bpbreid_strong_sort/sort/track.py:66-92,112-195       (life cycle, visibility-aware EMA)
//   /root/reference/plugins/track/bpbreid_strong_sort/sort/kalman_filter.py:47-227       (every noise term scales with the height)
//   /root/reference/plugins/track/bpbreid_strong_sort/sort/nn_matching.py:99-135,171-209 (part-based distance, 1 sample per track)
//   /root/reference/plugins/track/bpbreid_strong_sort/sort/linear_assignment.py:11-73,132-175
//   /root/reference/plugins/track/bpbreid_strong_sort/sort/iou_matching.py:7-78          (no time_since_update mask)
// for matching_strategy "strong_sort_matching", motion_criterium "iou" (configs/modules/track/bpbreid_strong_sort.yaml).
// The per-part distance itself lives in the un-vendored torchreid fork: restated as the visibility-weighted mean over
// parts of the Euclidean distance between L2-normalised part embeddings (PARITY UNPINNED, oracle/bpbreid_np.py).
// The `costs` visualisation dictionaries (tracker.py:365-407, three dense T x D matrices per frame) are not produced.
//
// Execution shape. With the reference configuration (n_init 0, max_age 300) every false positive stays a confirmed track
// for 300 frames, so T is several hundred while D is ~40, and the reference evaluates T x D x K part distances of
// length E per frame although the Mahalanobis gate then overwrites nearly all of them with 1e5. Here the gate goes
// first:
//   master CTA : predict only the filters still being predicted (octets); a per-track float32 gate rectangle
//                (|dx| <= sqrt(chi2 * S_xx), a necessary condition of the 4-d gate by Cauchy-Schwarz, widened by a margin)
//                prunes T x D with two compares per pair out of shared memory (count / scan / write, no atomics);
//   worker CTAs: meanwhile apply the previous frame's feature updates (visibility-aware EMA, births) and L2-normalise
//                the frame's detection parts;
//   all CTAs   : (group barrier) 4 rectangle hits per warp step get the exact Mahalanobis distance from the per-track
//                cached Cholesky factor; the survivors' part-based distance is evaluated by the whole warp on the
//                pre-normalised part embeddings, software-pipelined over the parts;
//   master CTA : (group barrier) fuse, keep the rows that still have a feasible entry, solve both assignments on that
//                compacted problem (lap.cuh: entries above the threshold are "unmatched at cost 0", so all-zero rows cannot
//                change the optimum), update filters, births, deletions, 14-column rows, publish the feature updates
//                (third group barrier).
// Ages and time_since_update are differences of a per-video tick, so stale tracks cost nothing per frame; the hot
// per-track integers and gate rectangles are mirrored in shared memory for the launch.
#include <cooperative_groups.h>
#include "kf_xyah.cuh"
#include "lap.cuh"
#include "lsap_scipy.cuh"
#include "trackkern.h"

namespace {

using namespace tk;

// Optional per-phase cycle accounting of the master CTA (build with -DTK_PHASE_PROF), read with tk_debug_bpbreid_phases().
#ifdef TK_PHASE_PROF
__device__ unsigned long long g_bp_prof[64];
#define PH(k) do { if (master && threadIdx.x == 0) { const long long _t = clock64(); g_bp_prof[k] += (unsigned long long)(_t - ph_t0); ph_t0 = _t; } } while (0)
#else
#define PH(k) do { } while (0)
#endif

constexpr int BP_THREADS = 256;
enum : unsigned char { BP_FREE = 0, BP_TENTATIVE = 1, BP_CONFIRMED = 2, BP_DELETED = 3 };
constexpr double W_POS = 1.0 / 20, W_VEL = 1.0 / 160, CHI2_4 = 9.4877;
constexpr int BP_COLS = 14;
constexpr int BP_NARR = 37;

struct BpParams {
    double max_dist, max_iou_dist, mc_lambda, min_conf;
    float ema_alpha, ema_beta;   // np.float32(alpha), np.float32(1 - alpha) (track.py:155-158)
    int max_age, n_init, max_pred, K, E;
    int bot_sort;                 // matching_strategy: 0 strong_sort_matching (two stages), 1 bot_sort_matching (one weighted stage)
    double gtf, w_k, w_r, w_s;    // gating_thres_factor, w_kfgd, w_reid, w_st (tracker.py:169-240)
};

struct BpDev {
    int* hdr;   // 0 next_id, 1 n_tracks, 2 tick, 4 status, 5 n_free, 6 nd (frame), 7 n appearance pairs (frame), 8 n_ema, 9 n_born
    unsigned* bar;
    double *mean, *cov, *gate, *chol, *dz, *det_id, *m_dist, *pred, *pg, *pf;   // chol: per track mean4, L(16), 1/diag(4) of the gating projection
    int *hits, *birth, *last, *track_id, *list, *list_tmp, *free_list, *conf_list, *det_rows, *ppack, *cpack;
    int *ema_slot, *ema_row, *born_slot, *born_row;   // feature updates of the frame, published for the worker CTAs
    unsigned char *state, *has_pred, *m_code;
    float *feat, *featn, *dfeatn, *vis, *dvis, *pa;   // featn / dfeatn: L2-normalised parts of the tracks / of the frame's detections
    unsigned char* lsap_ws;   // column-side scratch of the scipy-exact assignment (up to max(cap, capd) columns), see bp_lsap_ws
};

__host__ __device__ inline size_t bp_al(size_t x) { return (x + 255) & ~(size_t)255; }

template <class F>
__host__ __device__ inline void bp_layout(int cap, int capd, int K, int E, F&& f) {
    const size_t np = (size_t)cap * capd;
    int i = 0;
    f(i++, 16 * sizeof(int)); f(i++, 64);
    f(i++, (size_t)cap * 8 * 8); f(i++, (size_t)cap * 64 * 8); f(i++, (size_t)cap * 4 * 8); f(i++, (size_t)cap * 24 * 8); f(i++, (size_t)capd * 4 * 8);   // mean cov gate chol dz
    f(i++, (size_t)cap * 8); f(i++, (size_t)cap * 8); f(i++, (size_t)cap * 4 * 8);                    // det_id m_dist pred
    f(i++, np * 8); f(i++, np * 8);                                                                    // pg pf
    for (int k = 0; k < 8; ++k) f(i++, (size_t)cap * 4);                                               // hits .. conf_list
    f(i++, (size_t)capd * 4); f(i++, np * 4); f(i++, np * 4);                                          // det_rows ppack cpack
    for (int k = 0; k < 3; ++k) f(i++, (size_t)cap);                                                   // state has_pred m_code
    f(i++, (size_t)cap * K * E * 4); f(i++, (size_t)cap * K * E * 4); f(i++, (size_t)capd * K * E * 4); f(i++, (size_t)cap * K * 4); f(i++, (size_t)capd * K * 4);   // feat featn dfeatn vis dvis
    f(i++, np * 4);                                                                                    // pa
    for (int k = 0; k < 4; ++k) f(i++, (size_t)capd * 4);                                              // ema_slot ema_row born_slot born_row
    f(i++, (size_t)(cap > capd ? cap : capd) * 40 + 64);                                               // lsap_ws
}

__host__ __device__ inline size_t bp_state_bytes(int cap, int capd, int K, int E) {
    size_t s = 0;
    bp_layout(cap, capd, K, E, [&](int, size_t b) { s += bp_al(b); });
    return s;
}

__host__ __device__ inline BpDev bp_carve(char* base, int cap, int capd, int K, int E) {
    BpDev d;
    void** slots[BP_NARR] = {
        (void**)&d.hdr, (void**)&d.bar, (void**)&d.mean, (void**)&d.cov, (void**)&d.gate, (void**)&d.chol, (void**)&d.dz, (void**)&d.det_id, (void**)&d.m_dist,
        (void**)&d.pred, (void**)&d.pg, (void**)&d.pf, (void**)&d.hits, (void**)&d.birth, (void**)&d.last, (void**)&d.track_id,
        (void**)&d.list, (void**)&d.list_tmp, (void**)&d.free_list, (void**)&d.conf_list, (void**)&d.det_rows, (void**)&d.ppack,
        (void**)&d.cpack, (void**)&d.state, (void**)&d.has_pred, (void**)&d.m_code, (void**)&d.feat, (void**)&d.featn,
        (void**)&d.dfeatn, (void**)&d.vis, (void**)&d.dvis, (void**)&d.pa, (void**)&d.ema_slot, (void**)&d.ema_row, (void**)&d.born_slot, (void**)&d.born_row,
        (void**)&d.lsap_ws};
    char* p = base;
    bp_layout(cap, capd, K, E, [&](int i, size_t b) { *slots[i] = (void*)p; p += bp_al(b); });
    return d;
}

// float32 copy of a gate rectangle, widened so that the float test can only pass more pairs than the double one
// (coordinates up to 16k pixels: three roundings of at most 2^-9 each)
__device__ __forceinline__ float4 bp_gate_f32(const double* g) {
    return make_float4((float)g[0], (float)g[1], __double2float_ru(g[2]) + 0.01f, __double2float_ru(g[3]) + 0.01f);
}

__device__ __forceinline__ float warp_sum(float s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// float32 L2 norm of a length-4*E4 vector by one warp, clamped like F.normalize (eps 1e-12); result in every lane.
// float4 loads, four independent chains: these loops are latency-bound (one warp, L2-resident operands).
__device__ __forceinline__ float warp_part_norm(const float4* x, int E4) {
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll 4
    for (int i = lane_id(); i < E4; i += 32) {
        const float4 v = x[i];
        s0 = fmaf(v.x, v.x, s0); s1 = fmaf(v.y, v.y, s1); s2 = fmaf(v.z, v.z, s2); s3 = fmaf(v.w, v.w, s3);
    }
    return fmaxf(sqrtf(warp_sum((s0 + s1) + (s2 + s3))), 1e-12f);
}

// Per-track gating cache, refreshed whenever mean / covariance change: the rectangle (centre, half extents) such that
// |z_x - cx| > rx or |z_y - cy| > ry implies d^2 > chi2, and the Cholesky factor of the projected covariance with
// confidence 0 (kalman_filter.py:106-136,168-227) for the exact test. Returns false when the projection is not PD.
__device__ __forceinline__ bool bp_track_cache(const double* m, const double* P, double* g, double* c) {
    const double sp = W_POS * m[3];
    const double rr[4] = {sp * sp, sp * sp, sp * sp, sp * sp};
    g[0] = m[0]; g[1] = m[1];
    g[2] = sqrt(CHI2_4 * (P[0] + rr[0])) * (1.0 + 1e-6);
    g[3] = sqrt(CHI2_4 * (P[9] + rr[1])) * (1.0 + 1e-6);
    double Pl[64];
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) Pl[a * 8 + b] = P[a * 8 + b];
    double L[16], Sm[16], invd[4];
    const bool ok = kf8_chol4(Pl, rr, L, Sm, invd);
    for (int i = 0; i < 4; ++i) c[i] = m[i];
    for (int i = 0; i < 16; ++i) c[4 + i] = L[i];
    for (int i = 0; i < 4; ++i) c[20 + i] = invd[i];
    return ok;
}

// Feature side of Track.update / _initiate_track for the pairs and births the master published (track.py:149-165,
// tracker.py:423-441): visibility-aware EMA, float32, no contraction; one warp per (track, part). Runs on the worker
// CTAs while the master already predicts the next frame (the features are next read after the following barrier).
__device__ void bp_feature_updates(const BpDev& S, const BpParams& prm, const float* __restrict__ feats, const float* __restrict__ viss,
                                   int wg, int wn) {
    const int K = prm.K, E4 = prm.E >> 2, KE = prm.K * prm.E, lane = lane_id();
    const int n_ema = S.hdr[8], n_born = S.hdr[9];
    for (int i = wg; i < n_ema * K; i += wn) {
        const int p = i / K, k = i - p * K;
        const int s = S.ema_slot[p];
        const size_t drow = (size_t)S.ema_row[p];
        const float4* fv = reinterpret_cast<const float4*>(feats + drow * KE + (size_t)k * prm.E);
        float4* sm = reinterpret_cast<float4*>(S.feat + (size_t)s * KE + (size_t)k * prm.E);
        const float tv = S.vis[s * K + k], dv = viss[drow * K + k];
        const bool x = (tv != 0.0f) != (dv != 0.0f);
        const float both = __fmul_rn(tv, dv);
        const float et = __fadd_rn(__fmul_rn(both, prm.ema_alpha), x ? tv : 0.0f);
        const float ed = __fadd_rn(__fmul_rn(both, prm.ema_beta), x ? dv : 0.0f);
        const bool none = et == 0.0f && ed == 0.0f;
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll 4
        for (int e = lane; e < E4; e += 32) {
            const float4 a = sm[e], b = fv[e];
            float4 v;
            v.x = none ? 1.0f : __fadd_rn(__fmul_rn(et, a.x), __fmul_rn(ed, b.x));
            v.y = none ? 1.0f : __fadd_rn(__fmul_rn(et, a.y), __fmul_rn(ed, b.y));
            v.z = none ? 1.0f : __fadd_rn(__fmul_rn(et, a.z), __fmul_rn(ed, b.z));
            v.w = none ? 1.0f : __fadd_rn(__fmul_rn(et, a.w), __fmul_rn(ed, b.w));
            sm[e] = v;
            s0 = fmaf(v.x, v.x, s0); s1 = fmaf(v.y, v.y, s1); s2 = fmaf(v.z, v.z, s2); s3 = fmaf(v.w, v.w, s3);
        }
        const float nf = fmaxf(sqrtf(warp_sum((s0 + s1) + (s2 + s3))), 1e-12f);
        float4* sn = reinterpret_cast<float4*>(S.featn + (size_t)s * KE + (size_t)k * prm.E);
#pragma unroll 4
        for (int e = lane; e < E4; e += 32) {   // F.normalize (nn_matching.py:121-122), done once per update instead of once per pair
            const float4 v = sm[e];
            sn[e] = make_float4(__fdiv_rn(v.x, nf), __fdiv_rn(v.y, nf), __fdiv_rn(v.z, nf), __fdiv_rn(v.w, nf));
        }
        if (lane == 0) S.vis[s * K + k] = fmaxf(tv, dv);
    }
    for (int i = wg; i < n_born * K; i += wn) {   // the detection's parts become the track's sample as they are
        const int b = i / K, k = i - b * K;
        const int s = S.born_slot[b];
        const size_t drow = (size_t)S.born_row[b];
        const float4* fv = reinterpret_cast<const float4*>(feats + drow * KE + (size_t)k * prm.E);
        float4* sm = reinterpret_cast<float4*>(S.feat + (size_t)s * KE + (size_t)k * prm.E);
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll 4
        for (int e = lane; e < E4; e += 32) {
            const float4 v = fv[e];
            sm[e] = v;
            s0 = fmaf(v.x, v.x, s0); s1 = fmaf(v.y, v.y, s1); s2 = fmaf(v.z, v.z, s2); s3 = fmaf(v.w, v.w, s3);
        }
        const float nf = fmaxf(sqrtf(warp_sum((s0 + s1) + (s2 + s3))), 1e-12f);
        float4* sn = reinterpret_cast<float4*>(S.featn + (size_t)s * KE + (size_t)k * prm.E);
#pragma unroll 4
        for (int e = lane; e < E4; e += 32) {
            const float4 v = fv[e];
            sn[e] = make_float4(__fdiv_rn(v.x, nf), __fdiv_rn(v.y, nf), __fdiv_rn(v.z, nf), __fdiv_rn(v.w, nf));
        }
        if (lane == 0) S.vis[s * K + k] = viss[drow * K + k];
    }
}

// F.normalize of the frame's detections (nn_matching.py:121-122) into S.dfeatn. The calling CTA filters the frame itself
// (filter_detections, strong_sort.py:139-143) into `rows` (shared memory) so that it does not depend on the master.
__device__ void bp_det_normalise(const BpDev& S, const BpParams& prm, const double* __restrict__ D, int nraw, const float* __restrict__ feats,
                                 const float* __restrict__ viss, int r0, int* rows, int* nd_smem, int wg, int wn) {
    if (warp_id() == 0) {
        const int nd_ = warp_compact(nraw, 0, [&](int i) { return D[i * 7 + 4] > prm.min_conf; }, [&](int i, int p) { rows[p] = i; });
        if (lane_id() == 0) *nd_smem = nd_;
    }
    __syncthreads();
    const int nd = *nd_smem, K = prm.K, KE = prm.K * prm.E, E4 = prm.E >> 2;
    for (int i = wg; i < nd * K; i += wn) {
        const int d = i / K, k = i - d * K;
        const float4* fv = reinterpret_cast<const float4*>(feats + (size_t)(r0 + rows[d]) * KE + (size_t)k * prm.E);
        float4* dn = reinterpret_cast<float4*>(S.dfeatn + (size_t)d * KE + (size_t)k * prm.E);
        const float nf = warp_part_norm(fv, E4);
        if (lane_id() == 0) S.dvis[d * K + k] = viss[(size_t)(r0 + rows[d]) * K + k];
#pragma unroll 4
        for (int e = lane_id(); e < E4; e += 32) {
            const float4 v = fv[e];
            dn[e] = make_float4(__fdiv_rn(v.x, nf), __fdiv_rn(v.y, nf), __fdiv_rn(v.z, nf), __fdiv_rn(v.w, nf));
        }
    }
}

struct BpShared { int lap_ok, nd, nconf, ncand, nud, npairs, n_out, n_born, nkf, ncp, nap, nlive, nd_w; };

__global__ void __launch_bounds__(BP_THREADS)
bpbreid_video_kernel(BpParams prm, char* state_base, size_t state_stride, int cap, int capd, int capl, int ncta,
                     const double* __restrict__ dets, const float* __restrict__ feats, const float* __restrict__ viss,
                     const int* __restrict__ offsets, int n_frames, double* __restrict__ out_rows, const int* __restrict__ out_start,
                     int* __restrict__ out_frame_count, int* __restrict__ out_count, int out_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int seq = blockIdx.x / ncta, cta = blockIdx.x % ncta, tid = threadIdx.x;
    const bool master = cta == 0;
    const int E = prm.E, K = prm.K, KE = prm.K * prm.E;
    BpDev S = bp_carve(state_base + (size_t)seq * state_stride, cap, capd, K, E);
    unsigned char* sp = smem_raw;
    auto take = [&](size_t bytes) { unsigned char* p = sp; sp += (bytes + 15) & ~(size_t)15; return p; };
    const int side = capl > capd ? capl : capd;
    // master only (the other CTAs of the group use no shared memory)
    double* cost = (double*)take(sizeof(double) * (size_t)(capl + 1) * (capd + 1));
    double* lap_u = (double*)take(sizeof(double) * side);
    double* d_z = (double*)take(sizeof(double) * 4 * capd);
    double* d_ltwh = (double*)take(sizeof(double) * 4 * capd);
    float4* gate_f = (float4*)take(sizeof(float4) * cap);        // conservative float32 gate rectangles (cx, cy, rx, ry) per slot
    float2* dz_f = (float2*)take(sizeof(float2) * capd);         // float32 centres of the frame's detections
    int* list_m = (int*)take(sizeof(int) * cap);                  // mirror of S.list
    int* last_m = (int*)take(sizeof(int) * cap);                  // mirror of S.last (tick of the last update)
    int* scr = (int*)take(sizeof(int) * cap);                     // filters to predict / row of a confirmed position
    int* conf_m = (int*)take(sizeof(int) * cap);                  // mirror of S.conf_list (slots of the confirmed tracks, track order)
    int* match_a = (int*)take(sizeof(int) * side);
    int* match_b = (int*)take(sizeof(int) * side);
    int* col4row = (int*)take(sizeof(int) * side);
    int* row4col = (int*)take(sizeof(int) * side);
    int* path = (int*)take(sizeof(int) * side);
    int* live_rows = (int*)take(sizeof(int) * capl);
    int* cand = (int*)take(sizeof(int) * capl);
    int* un_d = (int*)take(sizeof(int) * capd);
    int* tmp_d = (int*)take(sizeof(int) * capd);
    int* pair_t = (int*)take(sizeof(int) * capd);
    int* pair_d = (int*)take(sizeof(int) * capd);
    unsigned char* state_m = (unsigned char*)take(cap);           // mirror of S.state
    unsigned char* t_flag = (unsigned char*)take(cap);
    unsigned char* t_live = (unsigned char*)take(cap);
    int* cmap = (int*)take(sizeof(int) * cap);                    // confirmed position -> row of the dense stage-A matrix, -1: no feasible entry
    unsigned char* lap_sr = (unsigned char*)take(side);
    unsigned char* touch_d = (unsigned char*)take(capd);          // 1 when the solver paired the detection column (accepted or not)
    BpShared* sh = (BpShared*)take(sizeof(BpShared));
    // column side of the scipy-exact solver (all confirmed tracks can be columns): global scratch, L1/L2 resident
    const int big = cap > capd ? cap : capd;
    double* lap_v = (double*)S.lsap_ws;
    double* lap_spc = lap_v + big;
    int* lap_path = (int*)(lap_spc + big);
    int* lap_r4c = lap_path + big;
    int* lap_rem = lap_r4c + big;
    int* rej_t = lap_rem + big;                                   // detection of a row's rejected pair (cost above the threshold) or -1
    unsigned char* lap_sc = (unsigned char*)(rej_t + big);

    int* status = &S.hdr[4];
    const int F1 = n_frames + 1;
    const int out_base = out_start[seq];
    int out_n = out_count[seq];
    const double L_app = prm.max_dist + 1e-5, L_iou = prm.max_iou_dist + 1e-5;
    int tick = S.hdr[2];

    if (master) {
        const int nt = S.hdr[1];
        for (int k = tid; k < nt; k += BP_THREADS) list_m[k] = S.list[k];
        for (int s = tid; s < cap; s += BP_THREADS) {
            last_m[s] = S.last[s]; state_m[s] = S.state[s];
            gate_f[s] = bp_gate_f32(S.gate + 4 * s);
        }
        __syncthreads();
    }

    // feature work is done by the other CTAs of the group (by the master itself when it is alone)
    const bool worker = !master || ncta == 1;
    const int wg = (ncta == 1 ? 0 : cta - 1) * (BP_THREADS / 32) + warp_id(), wn = (ncta == 1 ? 1 : ncta - 1) * (BP_THREADS / 32);
    bool pending = false;   // feature updates of the last processed frame still to be applied
#ifdef TK_PHASE_PROF
    long long ph_t0 = clock64();
#endif

    for (int f = 0; f < n_frames; ++f) {
        const int r0 = offsets[seq * F1 + f], r1 = offsets[seq * F1 + f + 1];
        const int nraw = r1 - r0;
        if (nraw == 0) { if (master && tid == 0) out_frame_count[seq * n_frames + f] = 0; continue; }   // bpbreid_strong_sort_api.py:75-84,105-106
        if (nraw > capd) { if (master && tid == 0) atomicOr(status, TK_DEV_OVERFLOW_DETS); break; }
        const double* D = dets + (size_t)r0 * 7;   // rows [l, t, w, h, conf, cls, det id]
        tick += 1;                                  // one tracker.predict per non-empty frame: age = tick - birth, tsu = tick - last

        if (!master) {   // workers: previous frame's feature updates, this frame's detection norms (overlaps the master's predict + gate)
            if (pending) bp_feature_updates(S, prm, feats, viss, wg, wn);
            bp_det_normalise(S, prm, D, nraw, feats, viss, r0, tmp_d, &sh->nd_w, wg, wn);
            __threadfence();
        }

        // ================= master: predict, detections, gate =================
        if (master) {
            const int nt = S.hdr[1];
            if (warp_id() == 0) {   // Track.predict runs the filter only while tsu < max_kalman_prediction_without_update (track.py:128-135)
                const int nk = warp_compact(nt, 0, [&](int k) { return (tick - 1) - last_m[list_m[k]] < prm.max_pred; },
                                            [&](int k, int p) { scr[p] = list_m[k]; });
                if (lane_id() == 0) sh->nkf = nk;
            } else if (warp_id() == 1) {   // filter_detections (strong_sort.py:139-143), order kept
                const int nd_ = warp_compact(nraw, 0, [&](int i) { return D[i * 7 + 4] > prm.min_conf; }, [&](int i, int p) { S.det_rows[p] = i; });
                if (lane_id() == 0) { sh->nd = nd_; S.hdr[6] = nd_; sh->ncp = 0; sh->nap = 0; }
            } else if (warp_id() == 2) {
                // bot_sort_matching associates ALL tracks in its single stage (tracker.py:343), strong_sort_matching the confirmed ones
                const int nc = warp_compact(nt, 0, [&](int k) { return prm.bot_sort || state_m[list_m[k]] == BP_CONFIRMED; },
                                            [&](int k, int p) { S.conf_list[p] = list_m[k]; conf_m[p] = list_m[k]; });
                if (lane_id() == 0) sh->nconf = nc;
            }
            __syncthreads();
            const int nkf = sh->nkf, nd = sh->nd, nconf = sh->nconf;
            PH(0);
            for (int base = 0; base < nkf; base += BP_THREADS / 8) {   // kalman_filter.py:74-104
                const int k = base + (tid >> 3), j = tid & 7;
                const bool act = k < nkf;
                const int s = act ? scr[k] : 0;
                double* gm = S.mean + (size_t)s * 8;
                double qj = 0.0;
                if (act) { const double sd = (j < 4 ? W_POS : W_VEL) * gm[3]; qj = sd * sd; }
                kf8_octet_predict(gm, S.cov + (size_t)s * 64, act, false, qj);
            }
            for (int i = tid; i < nd; i += BP_THREADS) {   // Detection.to_xyah (detection.py:58-66), float64
                const double* dr = D + (size_t)S.det_rows[i] * 7;
                double* b = d_ltwh + 4 * i;
                b[0] = dr[0]; b[1] = dr[1]; b[2] = dr[2]; b[3] = dr[3];
                double* z = d_z + 4 * i;
                z[0] = dr[0] + dr[2] / 2; z[1] = dr[1] + dr[3] / 2; z[2] = dr[2] / dr[3]; z[3] = dr[3];
                for (int c = 0; c < 4; ++c) S.dz[4 * i + c] = z[c];
                dz_f[i] = make_float2((float)z[0], (float)z[1]);
            }
            if (ncta == 1) {
                if (pending) bp_feature_updates(S, prm, feats, viss, wg, wn);
                bp_det_normalise(S, prm, D, nraw, feats, viss, r0, tmp_d, &sh->nd_w, wg, wn);
            }
            __syncthreads();
            PH(1);
            for (int k = tid; k < nkf; k += BP_THREADS) {   // gating cache of the filters that moved
                const int s = scr[k];
                double g[4];
                if (!bp_track_cache(S.mean + (size_t)s * 8, S.cov + (size_t)s * 64, g, S.chol + (size_t)s * 24)) atomicOr(status, TK_DEV_BAD_CHOLESKY);
                for (int i = 0; i < 4; ++i) S.gate[4 * s + i] = g[i];
                gate_f[s] = bp_gate_f32(g);
            }
            for (int k = tid; k < cap; k += BP_THREADS) { t_flag[k] = 0; t_live[k] = 0; }
            __syncthreads();
            PH(2);
            // Gate (linear_assignment.py:166-175), first step: rectangle test of every (confirmed track, detection) in float32
            // with a conservative margin, flattened over the CTA. Count, block scan, write: no atomics, deterministic order.
            // The exact test of the hits is done by all CTAs after the barrier.
            {
                const int total = nd > 0 ? nconf * nd : 0;
                // bot_sort_matching gates on sqrt(d2) / (sqrt(chi2) * gating_thres_factor) > 1: the rectangle grows by that factor
                const float rect_scale = prm.bot_sort ? __double2float_ru(prm.gtf) * 1.0001f : 1.0f;
                auto hit_at = [&](int r, int d, int& pk) {
                    const int slot = conf_m[r];
                    const float4 g = gate_f[slot];
                    const float2 z = dz_f[d];
                    pk = (slot << 8) | d;
                    return fabsf(z.x - g.x) <= g.z * rect_scale && fabsf(z.y - g.y) <= g.w * rect_scale;
                };
                const int per = (total + BP_THREADS - 1) / BP_THREADS, e_lo = min(total, tid * per), e_hi = min(total, e_lo + per);
                const int r_lo = nd > 0 ? e_lo / nd : 0, d_lo = e_lo - r_lo * nd;
                int cnt = 0, pk;
                for (int e = e_lo, r = r_lo, d = d_lo; e < e_hi; ++e) { cnt += hit_at(r, d, pk) ? 1 : 0; if (++d == nd) { d = 0; ++r; } }
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane_id() >= o) incl += v; }
                if (lane_id() == 31) match_a[warp_id()] = incl;
                __syncthreads();
                int base = incl - cnt;
                for (int w = 0; w < warp_id(); ++w) base += match_a[w];
                if (tid == BP_THREADS - 1) sh->ncp = base + cnt;
                for (int e = e_lo, r = r_lo, d = d_lo; e < e_hi; ++e) { if (hit_at(r, d, pk)) S.cpack[base++] = pk; if (++d == nd) { d = 0; ++r; } }
            }
            __syncthreads();
            if (tid == 0) { S.hdr[7] = sh->ncp; S.hdr[10] = 0; }
            __threadfence();
        }
        PH(3);
        group_barrier(S.bar, ncta);
        PH(4);

        // ================= all CTAs: exact gate of the hits + part distance of the survivors =================
        // A warp takes 4 hits at a time (static interleave): lanes 0-3 evaluate the Mahalanobis distance from the cached
        // factor (kalman_filter.py:168-227); survivors get a slot in the pair list (ppack, pg) and the whole warp walks their
        // K x E L2-normalised part embeddings (nn_matching.py:99-135). The part loop is software-pipelined: the next part's
        // 2 x 4 float4 per lane are in flight while the current one is reduced (one warp on L2-resident rows is latency-bound).
        {
            const int ncp = S.hdr[7];
            const int gw = cta * (BP_THREADS / 32) + warp_id(), nw = ncta * (BP_THREADS / 32), lane = lane_id(), E4 = E >> 2;
            for (int ch = gw; ch * 4 < ncp; ch += nw) {
                const int hi = ch * 4 + lane;
                int pk_l = 0;
                double g_l = 0.0;
                bool pass = false;
                if (lane < 4 && hi < ncp) {
                    pk_l = S.cpack[hi];
                    const double* c = S.chol + (size_t)(pk_l >> 8) * 24;
                    g_l = kf8_maha(c, c + 4, c + 20, S.dz + 4 * (pk_l & 255));
                    // strong_sort_matching: d2 > chi2inv95[4] (linear_assignment.py:166-175); bot_sort_matching: pos > 1 with
                    // pos = sqrt(d2) / (sqrt(chi2) * gating_thres_factor) (tracker.py:194-199,222)
                    pass = prm.bot_sort ? !(sqrt(g_l) / (sqrt(CHI2_4) * prm.gtf) > 1.0) : !(g_l > CHI2_4);
                }
                unsigned todo = __ballot_sync(0xffffffffu, pass);
                if (todo == 0u) continue;
                int q0 = 0;
                if (lane == 0) q0 = atomicAdd(&S.hdr[10], __popc(todo));
                q0 = __shfl_sync(0xffffffffu, q0, 0);
                if (pass) { const int q = q0 + __popc(todo & ((1u << lane) - 1u)); S.ppack[q] = pk_l; S.pg[q] = g_l; }
                for (int i = q0; todo; ++i) {
                    const int j = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int pk = __shfl_sync(0xffffffffu, pk_l, j), s = pk >> 8, d = pk & 255;
                    const float w_l = lane < K ? __fmul_rn(S.vis[s * K + lane], S.dvis[d * K + lane]) : 0.0f;   // lane k: weight of part k
                    const float4* a = reinterpret_cast<const float4*>(S.featn + (size_t)s * KE);
                    const float4* b = reinterpret_cast<const float4*>(S.dfeatn + (size_t)d * KE);
                    float num = 0.0f, den = 0.0f;
                    auto finish = [&](int k, float sq) {   // visibility-weighted mean over the parts, in part order
                        const float w = __shfl_sync(0xffffffffu, w_l, k);
                        num = __fadd_rn(num, __fmul_rn(sqrtf(warp_sum(sq)), w));
                        den = __fadd_rn(den, w);
                    };
                    if (E4 <= 128 && K <= 32) {
                        float4 A0[4], B0[4], A1[4], B1[4];
                        auto load = [&](int k, float4 (&A)[4], float4 (&B)[4]) {
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const int e = lane + 32 * t;
                                if (e < E4) { A[t] = a[k * E4 + e]; B[t] = b[k * E4 + e]; }
                                else { A[t] = make_float4(0.f, 0.f, 0.f, 0.f); B[t] = A[t]; }
                            }
                        };
                        auto square = [&](const float4 (&A)[4], const float4 (&B)[4]) {
                            float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, c3 = 0.0f;
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const float x0 = A[t].x - B[t].x, x1 = A[t].y - B[t].y, x2 = A[t].z - B[t].z, x3 = A[t].w - B[t].w;
                                c0 = fmaf(x0, x0, c0); c1 = fmaf(x1, x1, c1); c2 = fmaf(x2, x2, c2); c3 = fmaf(x3, x3, c3);
                            }
                            return (c0 + c1) + (c2 + c3);
                        };
                        load(0, A0, B0);
                        for (int k = 0; k < K; k += 2) {
                            if (k + 1 < K) load(k + 1, A1, B1);
                            finish(k, square(A0, B0));
                            if (k + 2 < K) load(k + 2, A0, B0);
                            if (k + 1 < K) finish(k + 1, square(A1, B1));
                        }
                    } else {
                        for (int k = 0; k < K; ++k) {
                            float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, c3 = 0.0f;
#pragma unroll 4
                            for (int e = lane; e < E4; e += 32) {
                                const float4 av = a[k * E4 + e], bv = b[k * E4 + e];
                                const float x0 = av.x - bv.x, x1 = av.y - bv.y, x2 = av.z - bv.z, x3 = av.w - bv.w;
                                c0 = fmaf(x0, x0, c0); c1 = fmaf(x1, x1, c1); c2 = fmaf(x2, x2, c2); c3 = fmaf(x3, x3, c3);
                            }
                            finish(k, (c0 + c1) + (c2 + c3));
                        }
                    }
                    if (lane == 0) S.pa[i] = __fdiv_rn(__fdiv_rn(num, den), 2.0f);   // nn_matching.py:133
                }
            }
            __threadfence();
        }
        PH(5);
        group_barrier(S.bar, ncta);
        PH(6);
        pending = true;
        if (!master) { group_barrier(S.bar, ncta); continue; }   // third barrier: the master has published the frame's feature updates

        // ================= master: fusion, assignments, updates =================
        const int nd = sh->nd, nconf = sh->nconf, nt = S.hdr[1], nap = S.hdr[10];   // pairs that passed the gate
        if (nd == 0) {   // every detection filtered out: predict only, tracker.update is not called (strong_sort.py:88-89)
            if (tid == 0) { out_frame_count[seq * n_frames + f] = 0; S.hdr[8] = 0; S.hdr[9] = 0; }
            __threadfence();
            group_barrier(S.bar, ncta);
            continue;
        }
        // ---- stage A: confirmed tracks x all detections (tracker.py:264-301, linear_assignment.py:132-175)
        for (int i = tid; i < nap; i += BP_THREADS) {
            const double a = (double)S.pa[i];
            if (!(a == a)) atomicOr(status, TK_DEV_NAN_COST);   // no commonly visible part: the reference's solver raises on NaN
            double fused;
            if (prm.bot_sort) {
                // _full_cost_metric (tracker.py:169-240): (w_kfgd * pos + w_reid * app + st * w_st) / sum(w); voided where the position
                // gate OR the appearance gate fails (the st gate never applies: np.logical_or(pos_gate, app_gate, st_gate) takes its
                // third argument as the output array). Pairs outside the position gate never reach this point.
                const int pk = S.ppack[i], sl = pk >> 8, d = pk & 255;
                const double pos = sqrt(S.pg[i]) / (sqrt(CHI2_4) * prm.gtf);
                const double* m = S.mean + (size_t)sl * 8;
                const double w = m[2] * m[3];
                const double bx = m[0] - w / 2, by = m[1] - m[3] / 2;                 // Track.to_ltwh (track.py:97-100)
                const double* cb = d_ltwh + 4 * d;
                const double x0 = fmax(bx, cb[0]), y0 = fmax(by, cb[1]);
                const double x1 = fmin(bx + w, cb[0] + cb[2]), y1 = fmin(by + m[3], cb[1] + cb[3]);
                const double iw = fmax(0.0, x1 - x0), ih = fmax(0.0, y1 - y0);
                const double inter = __dmul_rn(iw, ih);
                const double uni = __dsub_rn(__dadd_rn(__dmul_rn(w, m[3]), __dmul_rn(cb[2], cb[3])), inter);
                const double st = 1.0 - inter / uni;
                const double c = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(prm.w_k, pos), __dmul_rn(prm.w_r, a)), __dmul_rn(st, prm.w_s)),
                                           prm.w_k + prm.w_r + prm.w_s);
                fused = (prm.w_r > 0 && a > prm.max_dist) ? 1e5 : c;
            } else {
                fused = __dadd_rn(__dmul_rn(prm.mc_lambda, a), __dmul_rn(1.0 - prm.mc_lambda, S.pg[i]));
            }
            S.pf[i] = fused;
            if (!(fused > prm.max_dist)) t_live[S.ppack[i] >> 8] = 1;   // by slot
        }
        __syncthreads();
        if (warp_id() == 0) {   // rows with at least one feasible entry; the others cannot be matched (cost max_dist + 1e-5 everywhere)
            const int nl = warp_compact(nconf, 0, [&](int r) { return t_live[conf_m[r]] != 0; }, [&](int r, int p) { const int s = conf_m[r]; scr[s] = p; if (p < capl) live_rows[p] = s; });
            if (lane_id() == 0) { sh->nlive = nl; if (nl > capl) atomicOr(status, TK_DEV_OVERFLOW_ASSIGN); }
        }
        __syncthreads();
        const int nlive = sh->nlive <= capl ? sh->nlive : 0;
        PH(7);
        {
            // Dense matrix of the live rows only: dense[p * ldd + d], everything that is not feasible holds the clamped value
            // max_dist + 1e-5 (linear_assignment.py:52-53); confirmed tracks without any feasible entry are whole rows of that
            // value and exist only through cmap (-1). scipy solves the FULL confirmed x detections problem, transposed when it
            // is tall (more tracks than detections), and its tie-breaking among the equal clamped entries decides which
            // rejected pairs it returns — lsap_scipy.cuh reproduces it.
            const int ldd = capd + 1;
            for (int e = tid; e < nlive * ldd; e += BP_THREADS) cost[e] = L_app;
            for (int r = tid; r < nconf; r += BP_THREADS) { const int sl = conf_m[r]; cmap[r] = (nlive > 0 && t_live[sl]) ? scr[sl] : -1; rej_t[r] = -1; }
            for (int i = tid; i < nlive; i += BP_THREADS) match_a[i] = -1;
            for (int i = tid; i < nd; i += BP_THREADS) { match_b[i] = -1; touch_d[i] = 0; }
            if (tid == 0) sh->lap_ok = 1;
            __syncthreads();
            if (nlive > 0)
                for (int i = tid; i < nap; i += BP_THREADS) {
                    const double fused = S.pf[i];
                    if (fused > prm.max_dist) continue;
                    const int pk = S.ppack[i], r = scr[pk >> 8], d = pk & 255;
                    cost[(size_t)r * ldd + d] = fused;
                }
            __syncthreads();
            PH(8);
            if (nconf > 0 && nd > 0) {
                const bool t_rows = nconf <= nd;                   // rows = tracks unless the matrix is tall
                const int nr = t_rows ? nconf : nd, nc = t_rows ? nd : nconf;
                auto cval = [&](int r, int d) { const int p = cmap[r]; return p >= 0 ? cost[(size_t)p * ldd + d] : L_app; };
                if (warp_id() == 0) {
                    bool ok;
                    if (t_rows) ok = lsap_scipy_warp(nr, nc, [&](int i, int j) { return cval(i, j); }, lap_u, lap_v, lap_spc, lap_path, col4row,
                                                     lap_r4c, lap_rem, lap_sr, lap_sc);
                    else ok = lsap_scipy_warp(nr, nc, [&](int i, int j) { return cval(j, i); }, lap_u, lap_v, lap_spc, lap_path, col4row,
                                              lap_r4c, lap_rem, lap_sr, lap_sc);
                    if (!ok && lane_id() == 0) { atomicOr(status, TK_DEV_LAP_INFEASIBLE); sh->lap_ok = 0; }
                }
                __syncthreads();
                if (sh->lap_ok) for (int i = tid; i < nr; i += BP_THREADS) {
                    const int j = col4row[i];
                    const int r = t_rows ? i : j, d = t_rows ? j : i;
                    touch_d[d] = 1;
                    if (cval(r, d) > prm.max_dist) rej_t[r] = d; else { match_a[cmap[r]] = d; match_b[d] = cmap[r]; }
                }
            }
            __syncthreads();
        }
        if (warp_id() == 0) {
            const int np_ = warp_compact(nlive, 0, [&](int i) { return match_a[i] >= 0; },
                                         [&](int i, int p) { const int s = live_rows[i]; pair_t[p] = s; pair_d[p] = match_a[i]; t_flag[s] = 1; S.m_code[s] = 1; });
            if (lane_id() == 0) sh->npairs = np_;
        }
        if (nlive > 0)
            for (int i = tid; i < nap; i += BP_THREADS) {   // ("R", gated distance of the matched pair) tracker.py:409-421
                const int pk = S.ppack[i], sl = pk >> 8;
                if (t_live[sl] && !(S.pf[i] > prm.max_dist) && match_a[scr[sl]] == (pk & 255)) S.m_dist[sl] = S.pf[i];
            }
        __syncthreads();
        PH(9);
        // stage-B candidates: unconfirmed + unmatched confirmed with tsu == 1 (tracker.py:303-309)
        if (warp_id() == 0) {
            // bot_sort_matching has no second stage (tracker.py:335-363): no candidates, every unmatched detection is born below
            int nc = warp_compact(nt, 0, [&](int k) { return !prm.bot_sort && state_m[list_m[k]] != BP_CONFIRMED; }, [&](int k, int p) { if (p < capl) cand[p] = list_m[k]; });
            nc = warp_compact(nconf, nc, [&](int r) { const int s = conf_m[r]; return !prm.bot_sort && !t_flag[s] && tick - last_m[s] == 1; },
                              [&](int r, int p) { if (p < capl) cand[p] = conf_m[r]; });
            // unmatched_detections_a in the reference's order (linear_assignment.py:57-68): untouched columns in detection order,
            // then the detections of the rejected pairs in confirmed-track order
            int nu = warp_compact(nd, 0, [&](int d) { return !touch_d[d]; }, [&](int d, int p) { un_d[p] = d; });
            nu = warp_compact(nconf, nu, [&](int r) { return rej_t[r] >= 0; }, [&](int r, int p) { un_d[p] = rej_t[r]; });
            if (lane_id() == 0) { if (nc > capl) { atomicOr(status, TK_DEV_OVERFLOW_ASSIGN); nc = 0; } sh->ncand = nc; sh->nud = nu; }
        }
        __syncthreads();
        PH(10);
        // ---- stage B: IoU cost on the candidates (iou_matching.py:7-78, linear_assignment.py:11-73)
        {
            const int ncand = sh->ncand, nud = sh->nud;
            const bool a_rows = ncand <= nud;
            const int ld = lap_pitch(a_rows ? nud : ncand);
            double* vmat = S.pf;   // unthresholded 1 - IoU of the candidates, what matched_with reports (stage A values are consumed)
            for (int e = tid; e < ncand * nud; e += BP_THREADS) {
                const int r = e / nud, c = e - r * nud;
                const double* m = S.mean + (size_t)cand[r] * 8;
                const double w = m[2] * m[3];
                const double bx = m[0] - w / 2, by = m[1] - m[3] / 2;                 // Track.to_ltwh (track.py:97-100)
                const double* cb = d_ltwh + 4 * un_d[c];
                const double x0 = fmax(bx, cb[0]), y0 = fmax(by, cb[1]);
                const double x1 = fmin(bx + w, cb[0] + cb[2]), y1 = fmin(by + m[3], cb[1] + cb[3]);
                const double iw = fmax(0.0, x1 - x0), ih = fmax(0.0, y1 - y0);
                const double inter = __dmul_rn(iw, ih);
                const double uni = __dsub_rn(__dadd_rn(__dmul_rn(w, m[3]), __dmul_rn(cb[2], cb[3])), inter);
                const double v = 1.0 - inter / uni;
                vmat[(size_t)r * capd + c] = v;
                const double cl = v > prm.max_iou_dist ? L_iou : v;
                if (a_rows) cost[(size_t)r * ld + c] = cl; else cost[(size_t)c * ld + r] = cl;
            }
            for (int i = tid; i < ncand; i += BP_THREADS) { match_a[i] = -1; rej_t[i] = -1; }
            for (int i = tid; i < nud; i += BP_THREADS) { match_b[i] = -1; touch_d[i] = 0; }
            if (tid == 0) sh->lap_ok = 1;
            __syncthreads();
            PH(11);
            if (ncand > 0 && nud > 0) {
                const int nr = a_rows ? ncand : nud, nc = a_rows ? nud : ncand;
                if (warp_id() == 0) {
                    const bool ok = lsap_scipy_warp(nr, nc, [&](int i, int j) { return cost[(size_t)i * ld + j]; }, lap_u, lap_v, lap_spc, lap_path,
                                                    col4row, lap_r4c, lap_rem, lap_sr, lap_sc);
                    if (!ok && lane_id() == 0) { atomicOr(status, TK_DEV_LAP_INFEASIBLE); sh->lap_ok = 0; }
                }
                __syncthreads();
                if (sh->lap_ok) for (int i = tid; i < nr; i += BP_THREADS) {
                    const int j = col4row[i];
                    const int r = a_rows ? i : j, c = a_rows ? j : i;
                    touch_d[c] = 1;
                    if (cost[(size_t)i * ld + j] > prm.max_iou_dist) rej_t[r] = c; else { match_a[r] = c; match_b[c] = r; }
                }
            }
            __syncthreads();
            if (warp_id() == 0) {
                const int np_ = warp_compact(ncand, sh->npairs, [&](int r) { return match_a[r] >= 0; },
                                             [&](int r, int p) {
                                                 const int s = cand[r];
                                                 pair_t[p] = s; pair_d[p] = un_d[match_a[r]]; t_flag[s] = 1;
                                                 S.m_code[s] = 2; S.m_dist[s] = vmat[(size_t)r * capd + match_a[r]];   // ("S", dist)
                                             });
                int nu = warp_compact(nud, 0, [&](int c) { return !touch_d[c]; }, [&](int c, int p) { tmp_d[p] = un_d[c]; });
                nu = warp_compact(ncand, nu, [&](int r) { return rej_t[r] >= 0; }, [&](int r, int p) { tmp_d[p] = un_d[rej_t[r]]; });
                if (lane_id() == 0) { sh->npairs = np_; sh->nud = nu; }
            }
            __syncthreads();
        }
        PH(12);
        // ---- Track.update for every pair (track.py:137-174, kalman_filter.py:106-166)
        {
            const int np_ = sh->npairs;
            for (int p = tid; p < np_; p += BP_THREADS) {   // last_kf_pred_ltwh = to_ltwh() before the measurement update
                const int s = pair_t[p];
                const double* m = S.mean + (size_t)s * 8;
                const double w = m[2] * m[3];
                double* o = S.pred + (size_t)s * 4;
                o[0] = m[0] - w / 2; o[1] = m[1] - m[3] / 2; o[2] = w; o[3] = m[3];
                S.has_pred[s] = 1;
            }
            __syncthreads();
            for (int base = 0; base < np_; base += BP_THREADS / 8) {
                const int p = base + (tid >> 3);
                const bool act = p < np_;
                const int s = act ? pair_t[p] : 0, d = act ? pair_d[p] : 0;
                double z[4] = {0, 0, 0, 0}, r[4] = {1, 1, 1, 1};
                if (act) {
                    const double* dz = d_z + 4 * d;
                    z[0] = dz[0]; z[1] = dz[1]; z[2] = dz[2]; z[3] = dz[3];
                    const double cf = D[(size_t)S.det_rows[d] * 7 + 4];
                    const double sp_ = (1 - cf) * (W_POS * S.mean[(size_t)s * 8 + 3]);
                    r[0] = r[1] = r[2] = r[3] = sp_ * sp_;
                }
                if (!kf8_octet_update(S.mean + (size_t)s * 8, S.cov + (size_t)s * 64, act, z, r)) atomicOr(status, TK_DEV_BAD_CHOLESKY);
                if (act && (tid & 7) == 0) {
                    S.det_id[s] = D[(size_t)S.det_rows[d] * 7 + 6];
                    const int h = S.hits[s] + 1;
                    S.hits[s] = h; last_m[s] = tick;
                    if (state_m[s] == BP_TENTATIVE && h >= prm.n_init) state_m[s] = BP_CONFIRMED;
                }
            }
            for (int p = tid; p < np_; p += BP_THREADS) {   // feature side of Track.update: published for the workers
                S.ema_slot[p] = pair_t[p]; S.ema_row[p] = r0 + S.det_rows[pair_d[p]];
            }
            __syncthreads();
        }
        PH(13);
        // ---- mark_missed (track.py:181-187), births (tracker.py:423-441)
        for (int k = tid; k < nt; k += BP_THREADS) {
            const int s = list_m[k];
            if (!t_flag[s] && (state_m[s] == BP_TENTATIVE || tick - last_m[s] > prm.max_age)) state_m[s] = BP_DELETED;
        }
        if (warp_id() == 0) {
            const int lane = lane_id();
            // births in the order of unmatched_detections_b (untouched columns, then the rejected pairs scipy's tie-breaking
            // returns, in row order — reproduced by lsap_scipy.cuh)
            const int nfree = S.hdr[5];
            const int nb = sh->nud < nfree ? sh->nud : nfree;
            if (sh->nud > nfree && lane == 0) atomicOr(status, TK_DEV_OVERFLOW_TRACKS);
            const int id0 = S.hdr[0];
            for (int k = lane; k < nb; k += 32) {
                const int s = S.free_list[nfree - 1 - k];
                const int d = tmp_d[k];
                list_m[nt + k] = s;
                S.track_id[s] = id0 + k; S.det_id[s] = D[(size_t)S.det_rows[d] * 7 + 6];
                S.hits[s] = 1; S.birth[s] = tick - 1; last_m[s] = tick; state_m[s] = 1 >= prm.n_init ? BP_CONFIRMED : BP_TENTATIVE;
                S.has_pred[s] = 0; S.m_code[s] = 0;
                un_d[k] = s;      // un_d / tmp_d become the birth list (slot, detection)
            }
            if (lane == 0) { S.hdr[0] = id0 + nb; S.hdr[5] = nfree - nb; sh->n_born = nb; }
        }
        __syncthreads();
        {
            const int nb = sh->n_born, np_ = sh->npairs;
            for (int k = tid; k < nb; k += BP_THREADS) {   // KalmanFilter.initiate (kalman_filter.py:47-72)
                const int s = un_d[k];
                const double* z = d_z + 4 * tmp_d[k];
                double* m = S.mean + (size_t)s * 8;
                double* P = S.cov + (size_t)s * 64;
                for (int i = 0; i < 4; ++i) { m[i] = z[i]; m[4 + i] = 0.0; }
                for (int i = 0; i < 64; ++i) P[i] = 0.0;
                const double sp_ = 2 * W_POS * z[3], sv = 10 * W_VEL * z[3];
                for (int i = 0; i < 4; ++i) { P[i * 9] = sp_ * sp_; P[(4 + i) * 9] = sv * sv; }
            }
            for (int k = tid; k < nb; k += BP_THREADS) { S.born_slot[k] = un_d[k]; S.born_row[k] = r0 + S.det_rows[tmp_d[k]]; }
            if (tid == 0) { S.hdr[8] = np_; S.hdr[9] = nb; }
            __syncthreads();
            for (int k = tid; k < np_ + nb; k += BP_THREADS) {   // gating cache of the filters that changed (used as is when max_pred is 0)
                const int s = k < np_ ? pair_t[k] : un_d[k - np_];
                double g[4];
                if (!bp_track_cache(S.mean + (size_t)s * 8, S.cov + (size_t)s * 64, g, S.chol + (size_t)s * 24)) atomicOr(status, TK_DEV_BAD_CHOLESKY);
                for (int i = 0; i < 4; ++i) S.gate[4 * s + i] = g[i];
                gate_f[s] = bp_gate_f32(g);
            }
        }
        __syncthreads();
        PH(14);
        if (warp_id() == 0) {   // tracks = [t for t in tracks if not deleted]; free the deleted slots; emit (strong_sort.py:96-120)
            const int n0 = nt + sh->n_born;
            int nfree = S.hdr[5];
            nfree = warp_compact(n0, nfree, [&](int k) { return state_m[list_m[k]] == BP_DELETED; },
                                 [&](int k, int p) { S.free_list[p] = list_m[k]; });
            const int n = warp_compact(n0, 0, [&](int k) { return state_m[list_m[k]] != BP_DELETED; }, [&](int k, int p) { scr[p] = list_m[k]; });
            for (int k = lane_id(); k < n; k += 32) list_m[k] = scr[k];
            __syncwarp();
            const int cnt = warp_compact(n, 0, [&](int k) { const int s = list_m[k]; return state_m[s] == BP_CONFIRMED && last_m[s] == tick; },
                                         [&](int k, int p) { match_a[p] = list_m[k]; });   // at most one row per detection
            if (lane_id() == 0) {
                S.hdr[1] = n; S.hdr[5] = nfree; sh->n_out = cnt;
                out_frame_count[seq * n_frames + f] = cnt;
                if (out_n + cnt > out_cap) atomicOr(status, TK_DEV_OVERFLOW_OUT);
            }
        }
        __syncthreads();
        {
            const double qnan = __longlong_as_double(0x7ff8000000000000ll);
            const int cnt = sh->n_out;
            for (int p = tid; p < cnt; p += BP_THREADS) {
                if (out_n + p >= out_cap) break;
                const int s = match_a[p];
                const double* m = S.mean + (size_t)s * 8;
                const double w = m[2] * m[3];
                double* o = out_rows + (size_t)(out_base + out_n + p) * BP_COLS;
                o[0] = (double)S.track_id[s];
                o[1] = m[0] - w / 2; o[2] = m[1] - m[3] / 2; o[3] = w; o[4] = m[3];
                const bool hp = S.has_pred[s] != 0;
                for (int i = 0; i < 4; ++i) o[5 + i] = hp ? S.pred[(size_t)s * 4 + i] : qnan;
                const int code = S.m_code[s];
                o[9] = (double)code; o[10] = code ? S.m_dist[s] : qnan;
                o[11] = (double)S.hits[s]; o[12] = (double)(tick - S.birth[s]); o[13] = S.det_id[s];
            }
            out_n += cnt;
        }
        PH(15);
        __threadfence();
        group_barrier(S.bar, ncta);   // third barrier of the frame: ema_slot / born_slot lists are visible to the workers
        PH(16);
    }
    if (worker && pending) bp_feature_updates(S, prm, feats, viss, wg, wn);   // updates of the last frame of the launch
    if (master) {   // write the mirrors back
        __syncthreads();
        const int nt = S.hdr[1];
        for (int k = tid; k < nt; k += BP_THREADS) S.list[k] = list_m[k];
        for (int s = tid; s < cap; s += BP_THREADS) { S.last[s] = last_m[s]; S.state[s] = state_m[s]; }
        if (tid == 0) { out_count[seq] = out_n < out_cap ? out_n : out_cap; S.hdr[2] = tick; }
    }
}

struct BpHandle {
    BpParams prm;
    int n_seq, cap, capd, capl, ncta;
    char* state;
    size_t state_stride, smem_bytes;
};

__global__ void bpbreid_reset_kernel(char* base, size_t stride, int cap, int capd, int K, int E) {
    BpDev S = bp_carve(base + (size_t)blockIdx.x * stride, cap, capd, K, E);
    if (threadIdx.x == 0) { S.hdr[0] = 1; S.hdr[1] = 0; S.hdr[2] = 0; S.hdr[4] = 0; S.hdr[5] = cap; S.hdr[6] = 0; S.hdr[7] = 0; S.hdr[8] = 0; S.hdr[9] = 0; S.bar[0] = 0; S.bar[1] = 0; }
    for (int i = threadIdx.x; i < cap; i += blockDim.x) {
        S.free_list[i] = cap - 1 - i; S.state[i] = BP_FREE; S.has_pred[i] = 0; S.m_code[i] = 0; S.last[i] = 0; S.birth[i] = 0;
        for (int k = 0; k < 4; ++k) S.gate[4 * i + k] = 0.0;
    }
}

size_t bp_smem(int cap, int capd, int capl) {
    const int side = capl > capd ? capl : capd;
    auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
    size_t s = al(8 * (size_t)(capl + 1) * (capd + 1)) + al(8 * side) + 2 * al(8 * 4 * capd) + al(16 * cap) + al(8 * capd);
    s += 4 * al(4 * cap) + 5 * al(4 * side) + 2 * al(4 * capl) + 4 * al(4 * capd) + 3 * al(cap) + al(sizeof(BpShared));
    s += al(4 * cap) + al(side) + al(capd);   // cmap, lap_sr, touch_d
    return s;
}

}  // namespace

extern "C" {

int tk_bpbreid_create(const tk_bpbreid_params* p, int n_seq, int cap_tracks, int cap_dets, void** handle) {
    if (!p || !handle || n_seq <= 0 || cap_tracks <= 0 || cap_dets <= 0 || p->feature_dim <= 0 || p->n_parts <= 0) return TK_ERR_ARG;
    if (p->feature_dim % 4 != 0) return TK_ERR_ARG;   // part embeddings are moved as float4
    if (cap_dets > 256 || cap_dets > tk::LAP_MAX_COLS || cap_tracks > (1 << 22)) return TK_ERR_CAPACITY;   // (track position << 8 | detection) packing
    BpHandle* h = new BpHandle();
    h->prm.max_dist = p->max_dist; h->prm.max_iou_dist = p->max_iou_distance; h->prm.mc_lambda = p->mc_lambda;
    h->prm.min_conf = p->min_bbox_confidence; h->prm.ema_alpha = (float)p->ema_alpha; h->prm.ema_beta = (float)(1 - p->ema_alpha);
    h->prm.max_age = p->max_age; h->prm.n_init = p->n_init; h->prm.max_pred = p->max_kalman_prediction_without_update;
    h->prm.K = p->n_parts; h->prm.E = p->feature_dim;
    h->prm.bot_sort = p->matching_strategy == TK_BPBREID_BOT_SORT_MATCHING ? 1 : 0;
    h->prm.gtf = p->gating_thres_factor > 0 ? p->gating_thres_factor : 1.0;
    h->prm.w_k = p->w_kfgd; h->prm.w_r = p->w_reid; h->prm.w_s = p->w_st;
    if (p->matching_strategy != TK_BPBREID_STRONG_SORT_MATCHING && p->matching_strategy != TK_BPBREID_BOT_SORT_MATCHING) { delete h; return TK_ERR_ARG; }
    // the single-stage strategy is evaluated gate-first: it needs the Kalman position gate (w_kfgd > 0, the reference default)
    if (h->prm.bot_sort && !(p->w_kfgd > 0 && p->w_reid >= 0 && p->w_st >= 0)) { delete h; return TK_ERR_ARG; }
    h->n_seq = n_seq; h->cap = cap_tracks; h->capd = cap_dets;
    h->capl = cap_tracks < cap_dets ? cap_tracks : cap_dets;   // side of the compacted assignment problems
    h->ncta = p->ctas_per_video > 0 ? p->ctas_per_video : 8;
    h->state_stride = (bp_state_bytes(cap_tracks, cap_dets, p->n_parts, p->feature_dim) + 255) & ~(size_t)255;
    h->smem_bytes = bp_smem(cap_tracks, cap_dets, h->capl);
    h->state = nullptr;
    if (h->smem_bytes > 220 * 1024) { delete h; return TK_ERR_CAPACITY; }
    cudaError_t e = cudaMalloc((void**)&h->state, h->state_stride * n_seq);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bpbreid_video_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
    if (e != cudaSuccess) { tk_set_last_cuda_error((int)e); if (h->state) cudaFree(h->state); delete h; return TK_ERR_CUDA; }
    *handle = h;
    return tk_bpbreid_reset(h, 0, nullptr);
}

int tk_bpbreid_reset(void* handle, int keep_id_counter, void* stream) {
    (void)keep_id_counter;   // Tracker._next_id restarts with every StrongSORT() (tracker.py:87)
    if (!handle) return TK_ERR_ARG;
    BpHandle* h = (BpHandle*)handle;
    bpbreid_reset_kernel<<<h->n_seq, 128, 0, (cudaStream_t)stream>>>(h->state, h->state_stride, h->cap, h->capd, h->prm.K, h->prm.E);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_bpbreid_run(void* handle, const double* dets, const float* features, const float* visibility, const int* offsets, int n_frames,
                   double* out_rows, const int* out_start, int* out_frame_count, int* out_count, int out_capacity_rows, void* stream) {
    if (!handle || !offsets || !out_rows || !out_start || !out_frame_count || !out_count || n_frames < 0) return TK_ERR_ARG;
    BpHandle* h = (BpHandle*)handle;
    if (n_frames == 0) return TK_OK;
    int cap = h->cap, capd = h->capd, capl = h->capl, ncta = h->ncta;
    void* args[] = {&h->prm, &h->state, &h->state_stride, &cap, &capd, &capl, &ncta, &dets, &features, &visibility, &offsets, &n_frames,
                    &out_rows, &out_start, &out_frame_count, &out_count, &out_capacity_rows};
    // cooperative launch: every CTA of a video group must be co-resident for the group barrier
    TK_CUDA_TRY(cudaFuncSetAttribute(bpbreid_video_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));   // per function, not per handle
    TK_CUDA_TRY(cudaLaunchCooperativeKernel((void*)bpbreid_video_kernel, dim3(h->n_seq * ncta), dim3(BP_THREADS), args, h->smem_bytes,
                                            (cudaStream_t)stream));
    return TK_OK;
}

int tk_bpbreid_status(void* handle, int* status_host, void* stream) {
    if (!handle || !status_host) return TK_ERR_ARG;
    BpHandle* h = (BpHandle*)handle;
    for (int s = 0; s < h->n_seq; ++s)
        TK_CUDA_TRY(cudaMemcpyAsync(status_host + s, h->state + (size_t)s * h->state_stride + 4 * sizeof(int), sizeof(int),
                                    cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    TK_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return TK_OK;
}

#ifdef TK_PHASE_PROF
int tk_debug_bpbreid_phases(unsigned long long* host_out64, int reset) {
    cudaDeviceSynchronize();
    if (host_out64) cudaMemcpyFromSymbol(host_out64, g_bp_prof, sizeof(unsigned long long) * 64);
    if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(g_bp_prof, z, sizeof(z)); }
    return 0;
}
#endif

int tk_bpbreid_destroy(void* handle) {
    if (!handle) return TK_ERR_ARG;
    BpHandle* h = (BpHandle*)handle;
    cudaFree(h->state);
    delete h;
    return TK_OK;
}

}  // extern "C"
