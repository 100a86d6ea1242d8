// HOTA of one sequence on the device: the parity report next to the trackers, without a trip through MOT text files.
//
// Restates HOTA.eval_sequence of the TrackEval fork vendored in the reference
// (/root/reference/plugins/eval/PoseTrack21/posetrack21/posetrack21/trackeval/metrics/hota.py:28-154 and
// _compute_final_fields :205-221) with the box similarity of its MOT dataset class
// (trackeval/datasets/_base_dataset.py:244-282, `_calculate_box_ious(box_format='xywh')`, called from posetrack_mot.py:479):
//   pass 1  per frame: IoU matrix s, s / (row sum + column sum - s) accumulated into potential_matches_count[gt id, tracker id]
//   pass 2  per frame: scipy.optimize.linear_sum_assignment(-(global_alignment_score * s)), then per alpha the matches with
//           s >= alpha - eps: TP / FN / FP, the sum of matched similarities (LocA) and matches_counts[alpha][gt id, tracker id]
//   final   per alpha: AssA / AssRe / AssPr (np.sum over the id x id matrix), DetRe / DetPr / DetA, HOTA = sqrt(DetA * AssA)
// Everything that decides an assignment is evaluated in the reference's own floating-point order: NumPy's pairwise summation for
// s.sum(1) and np.sum (identity start, 8 accumulators per block of <= 128, halves rounded down to a multiple of 8), sequential
// sums for s.sum(0) and for the frame-by-frame `+=` of potential_matches_count and LocA, and the scipy solver of lsap_scipy.cuh
// including its tie-breaking. The FragA field of this fork (a [19, gt ids, tracker ids, frames] tensor) is not computed.
//
// Device layout: one CTA per frame for the two per-frame passes (all frames in flight), a gt-id-sliced walk over the frames for
// the ordered accumulation (each potential_matches_count cell has one owner CTA, which adds the frames in order), one CTA per
// alpha for the final fields.
#include "lsap_scipy.cuh"
#include "tk_common.cuh"
#include "trackkern.h"

namespace {

constexpr int MAX_ALPHAS = 32;
constexpr double F64_EPS = 2.220446049250313e-16;   // np.finfo('float').eps

struct HotaArgs {
    const double* gt_boxes; const int* gt_ids; const int* gt_off;
    const double* tr_boxes; const int* tr_ids; const int* tr_off;
    int n_frames, n_gt_ids, n_tr_ids, max_g, max_t, n_alphas;
    long long pairs_cap;
    double alphas[MAX_ALPHAS];
    // workspace
    long long* pair_off;        // [F+1]
    double* sims;               // [pairs]
    double* aux;                // [pairs]: s / (row + col - s) in pass 1, the assignment score in pass 2
    double* pot;                // [n_gt_ids, n_tr_ids]
    int* gcnt;                  // [n_gt_ids]
    int* tcnt;                  // [n_tr_ids]
    int* mc;                    // [n_alphas, n_gt_ids, n_tr_ids]
    double* loc_part;           // [F, n_alphas]
    unsigned long long* tot;    // [3, n_alphas] TP, FN, FP
    int* status;
    double* out;                // [TK_HOTA_FIELDS, n_alphas]
};

// NumPy's DOUBLE_pairwise_sum over f(0..n-1) (numpy/_core/src/umath/loops_utils.h.src), started from the identity like np.add.reduce
template <class F>
__device__ double np_pairwise_sum(F f, long long n) {
    // explicit stack of (start, length) ranges, combined in the recursion's order: result(range) = left + right
    // a range of length <= 128 is a leaf. Depth <= 64.
    struct Item { long long start, len; int state; double left; };
    Item st[48];
    int sp = 0;
    st[0] = {0, n, 0, 0.0};
    double ret = 0.0;
    while (sp >= 0) {
        Item& it = st[sp];
        if (it.len <= 128) {
            const long long a = it.start, m = it.len;
            double res;
            if (m < 8) {
                res = 0.0;
                for (long long i = 0; i < m; ++i) res = __dadd_rn(res, f(a + i));
            } else {
                double r[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = f(a + j);
                long long i = 8;
                for (; i < m - (m % 8); i += 8) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], f(a + i + j));
                }
                res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                                __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
                for (; i < m; ++i) res = __dadd_rn(res, f(a + i));
            }
            ret = res;
            --sp;
            continue;
        }
        long long n2 = it.len / 2;
        n2 -= n2 % 8;
        if (it.state == 0) {            // descend left
            it.state = 1;
            st[sp + 1] = {it.start, n2, 0, 0.0};
            ++sp;
        } else if (it.state == 1) {     // left done -> descend right
            it.left = ret;
            it.state = 2;
            st[sp + 1] = {it.start + n2, it.len - n2, 0, 0.0};
            ++sp;
        } else {                        // both done
            ret = __dadd_rn(it.left, ret);
            --sp;
        }
    }
    return __dadd_rn(0.0, ret);
}

// pair offsets, capacity checks, zeroing of the accumulators
__global__ void __launch_bounds__(1024) hota_setup_kernel(HotaArgs A) {
    __shared__ long long carry;
    __shared__ long long wsum[32];
    const int tid = threadIdx.x;
    if (tid == 0) { carry = 0; A.pair_off[0] = 0; }
    __syncthreads();
    bool bad = false;
    for (int f0 = 0; f0 < A.n_frames; f0 += 1024) {
        const int f = f0 + tid;
        long long v = 0;
        if (f < A.n_frames) {
            const int ng = A.gt_off[f + 1] - A.gt_off[f], nt = A.tr_off[f + 1] - A.tr_off[f];
            if (ng < 0 || nt < 0 || ng > A.max_g || nt > A.max_t) bad = true;
            v = (long long)max(ng, 0) * max(nt, 0);
        }
        long long x = v;                                    // inclusive warp scan
        for (int d = 1; d < 32; d <<= 1) { long long y = __shfl_up_sync(0xffffffffu, x, d); if ((tid & 31) >= d) x += y; }
        if ((tid & 31) == 31) wsum[tid >> 5] = x;
        __syncthreads();
        if (tid < 32) {
            long long w = wsum[tid];
            for (int d = 1; d < 32; d <<= 1) { long long y = __shfl_up_sync(0xffffffffu, w, d); if (tid >= d) w += y; }
            wsum[tid] = w;
        }
        __syncthreads();
        const long long base = carry + ((tid >> 5) ? wsum[(tid >> 5) - 1] : 0);
        if (f < A.n_frames) A.pair_off[f + 1] = base + x;
        __syncthreads();
        if (tid == 1023) carry = base + x;
        __syncthreads();
    }
    if (bad) atomicOr(A.status, TK_DEV_OVERFLOW_DETS);
    if (tid == 0 && carry > A.pairs_cap) atomicOr(A.status, TK_DEV_OVERFLOW_OUT);
    const long long cells = (long long)A.n_gt_ids * A.n_tr_ids;
    for (long long i = tid; i < cells; i += 1024) A.pot[i] = 0.0;
    for (long long i = tid; i < cells * A.n_alphas; i += 1024) A.mc[i] = 0;
    for (int i = tid; i < A.n_gt_ids; i += 1024) A.gcnt[i] = 0;
    for (int i = tid; i < A.n_tr_ids; i += 1024) A.tcnt[i] = 0;
    for (int i = tid; i < 3 * A.n_alphas; i += 1024) A.tot[i] = 0ull;
}

// pass 1, one CTA per frame: s = _calculate_box_ious(gt, tracker) and s / (s.sum(0) + s.sum(1) - s)
__global__ void __launch_bounds__(256) hota_similarity_kernel(HotaArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (*A.status) return;
    const int f = blockIdx.x;
    const int g0 = A.gt_off[f], ng = A.gt_off[f + 1] - g0, t0 = A.tr_off[f], nt = A.tr_off[f + 1] - t0;
    for (int i = threadIdx.x; i < ng; i += blockDim.x) {
        const int id = A.gt_ids[g0 + i];
        if (id < 0 || id >= A.n_gt_ids) atomicOr(A.status, TK_DEV_OVERFLOW_TRACKS); else atomicAdd(&A.gcnt[id], 1);
    }
    for (int j = threadIdx.x; j < nt; j += blockDim.x) {
        const int id = A.tr_ids[t0 + j];
        if (id < 0 || id >= A.n_tr_ids) atomicOr(A.status, TK_DEV_OVERFLOW_TRACKS); else atomicAdd(&A.tcnt[id], 1);
    }
    if (ng == 0 || nt == 0) return;
    double* rs = reinterpret_cast<double*>(smem_raw);   // [max_g] row sums
    double* cs = rs + A.max_g;                           // [max_t] column sums
    double* s = A.sims + A.pair_off[f];
    double* q = A.aux + A.pair_off[f];
    for (int idx = threadIdx.x; idx < ng * nt; idx += blockDim.x) {
        const int i = idx / nt, j = idx - i * nt;
        const double* a = A.gt_boxes + 4 * (size_t)(g0 + i);
        const double* b = A.tr_boxes + 4 * (size_t)(t0 + j);
        // xywh -> x0 y0 x1 y1 (_base_dataset.py:255-258)
        const double ax0 = a[0], ay0 = a[1], ax1 = __dadd_rn(a[0], a[2]), ay1 = __dadd_rn(a[1], a[3]);
        const double bx0 = b[0], by0 = b[1], bx1 = __dadd_rn(b[0], b[2]), by1 = __dadd_rn(b[1], b[3]);
        double inter = __dmul_rn(fmax(__dsub_rn(fmin(ax1, bx1), fmax(ax0, bx0)), 0.0),
                                 fmax(__dsub_rn(fmin(ay1, by1), fmax(ay0, by0)), 0.0));                 // :265
        const double area1 = __dmul_rn(__dsub_rn(ax1, ax0), __dsub_rn(ay1, ay0));                          // :266
        const double area2 = __dmul_rn(__dsub_rn(bx1, bx0), __dsub_rn(by1, by0));                          // :275
        double uni = __dsub_rn(__dadd_rn(area1, area2), inter);                                            // :276
        if (area1 <= F64_EPS || area2 <= F64_EPS || uni <= F64_EPS) inter = 0.0;                           // :277-279
        if (uni <= F64_EPS) uni = 1.0;                                                                     // :280
        s[idx] = __ddiv_rn(inter, uni);                                                                    // :281
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ng; i += blockDim.x) {             // similarity.sum(1): pairwise along the contiguous axis
        const double* row = s + (size_t)i * nt;
        rs[i] = np_pairwise_sum([&](long long k) { return row[k]; }, nt);
    }
    for (int j = threadIdx.x; j < nt; j += blockDim.x) {             // similarity.sum(0): row after row
        double acc = 0.0;
        for (int i = 0; i < ng; ++i) acc = __dadd_rn(acc, s[(size_t)i * nt + j]);
        cs[j] = acc;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < ng * nt; idx += blockDim.x) {   // hota.py:66-69
        const int i = idx / nt, j = idx - i * nt;
        const double den = __dsub_rn(__dadd_rn(cs[j], rs[i]), s[idx]);
        q[idx] = (den > F64_EPS) ? __ddiv_rn(s[idx], den) : 0.0;
    }
}

// ordered accumulation of potential_matches_count (hota.py:70): CTA c owns the gt ids with id % gridDim.x == c and walks the
// frames in order, so every cell is the reference's frame-by-frame sum
__global__ void __launch_bounds__(256) hota_potential_kernel(HotaArgs A) {
    if (*A.status) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int f = 0; f < A.n_frames; ++f) {
        const int g0 = A.gt_off[f], ng = A.gt_off[f + 1] - g0, t0 = A.tr_off[f], nt = A.tr_off[f + 1] - t0;
        if (ng == 0 || nt == 0) continue;
        const double* q = A.aux + A.pair_off[f];
        for (int i = warp; i < ng; i += nw) {
            const int gid = A.gt_ids[g0 + i];
            if (gid % (int)gridDim.x != (int)blockIdx.x) continue;
            double* prow = A.pot + (size_t)gid * A.n_tr_ids;
            for (int j = lane; j < nt; j += 32) {
                const int tidx = A.tr_ids[t0 + j];
                prow[tidx] = __dadd_rn(prow[tidx], q[(size_t)i * nt + j]);
            }
        }
        __syncthreads();      // the next frame may touch the same cell from another warp
    }
}

// score_mat = global_alignment_score[gt ids, tracker ids] * similarity (hota.py:91), stored negated for the solver
__global__ void __launch_bounds__(256) hota_score_kernel(HotaArgs A) {
    if (*A.status) return;
    const int f = blockIdx.x;
    const int g0 = A.gt_off[f], ng = A.gt_off[f + 1] - g0, t0 = A.tr_off[f], nt = A.tr_off[f + 1] - t0;
    if (ng == 0 || nt == 0) return;
    const double* s = A.sims + A.pair_off[f];
    double* q = A.aux + A.pair_off[f];
    for (int idx = threadIdx.x; idx < ng * nt; idx += blockDim.x) {
        const int i = idx / nt, j = idx - i * nt;
        const int gid = A.gt_ids[g0 + i], tidx = A.tr_ids[t0 + j];
        const double p = A.pot[(size_t)gid * A.n_tr_ids + tidx];
        // global_alignment_score = pot / (gt_id_count + tracker_id_count - pot)  (hota.py:77)
        const double den = __dsub_rn(__dadd_rn((double)A.gcnt[gid], (double)A.tcnt[tidx]), p);
        q[idx] = -__dmul_rn(__ddiv_rn(p, den), s[idx]);
    }
}

// pass 2, one warp per frame: the assignment and the per-alpha statistics (hota.py:80-107)
__global__ void __launch_bounds__(32) hota_match_kernel(HotaArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (*A.status) return;
    const int f = blockIdx.x, lane = threadIdx.x;
    const int g0 = A.gt_off[f], ng = A.gt_off[f + 1] - g0, t0 = A.tr_off[f], nt = A.tr_off[f + 1] - t0;
    double* loc = A.loc_part + (size_t)f * A.n_alphas;
    if (lane < A.n_alphas) loc[lane] = 0.0;
    if (ng == 0 || nt == 0) {
        if (lane < A.n_alphas) {
            if (ng == 0 && nt > 0) atomicAdd(&A.tot[2 * A.n_alphas + lane], (unsigned long long)nt);     // FP  :83-86
            if (nt == 0 && ng > 0) atomicAdd(&A.tot[1 * A.n_alphas + lane], (unsigned long long)ng);     // FN  :87-90
        }
        return;
    }
    const double* s = A.sims + A.pair_off[f];
    const double* c = A.aux + A.pair_off[f];
    const bool transpose = nt < ng;                       // scipy solves the transposed problem for tall matrices
    const int nr = transpose ? nt : ng, nc = transpose ? ng : nt;
    const int nr_max = min(A.max_g, A.max_t), nc_max = max(A.max_g, A.max_t);
    tk::LsapScratch S;
    S.carve(smem_raw, nr_max, nc_max);
    int* match = reinterpret_cast<int*>(smem_raw + tk::lsap_scipy_scratch_bytes(nr_max, nc_max));   // [max_g] column of each gt row
    bool ok;
    if (transpose) ok = tk::lsap_scipy_warp(nr, nc, [&](int i, int j) { return c[(size_t)j * nt + i]; }, S.u, S.v, S.spc, S.path,
                                            S.col4row, S.row4col, S.remaining, S.SR, S.SC);
    else ok = tk::lsap_scipy_warp(nr, nc, [&](int i, int j) { return c[(size_t)i * nt + j]; }, S.u, S.v, S.spc, S.path, S.col4row,
                                  S.row4col, S.remaining, S.SR, S.SC);
    __syncwarp();
    if (!ok) { if (lane == 0) atomicOr(A.status, TK_DEV_LAP_INFEASIBLE); return; }
    for (int i = lane; i < ng; i += 32) match[i] = transpose ? S.row4col[i] : S.col4row[i];
    __syncwarp();
    // one lane per alpha: matches in ascending gt row order (scipy returns sorted row indices), Python's sequential sum()
    if (lane < A.n_alphas) {
        const double thr = __dsub_rn(A.alphas[lane], F64_EPS);
        int* mca = A.mc + (size_t)lane * A.n_gt_ids * A.n_tr_ids;
        double acc = 0.0;
        int n = 0;
        for (int i = 0; i < ng; ++i) {
            const int j = match[i];
            if (j < 0) continue;
            const double v = s[(size_t)i * nt + j];
            if (v >= thr) {
                acc = __dadd_rn(acc, v);
                ++n;
                atomicAdd(&mca[(size_t)A.gt_ids[g0 + i] * A.n_tr_ids + A.tr_ids[t0 + j]], 1);
            }
        }
        loc[lane] = acc;
        atomicAdd(&A.tot[0 * A.n_alphas + lane], (unsigned long long)n);
        atomicAdd(&A.tot[1 * A.n_alphas + lane], (unsigned long long)(ng - n));
        atomicAdd(&A.tot[2 * A.n_alphas + lane], (unsigned long long)(nt - n));
    }
}

// final fields, one CTA per alpha (hota.py:127-154, 205-221); out rows follow TK_HOTA_* of trackkern.h
__global__ void __launch_bounds__(32) hota_final_kernel(HotaArgs A, long long num_gt, long long num_tr) {
    const int a = blockIdx.x, lane = threadIdx.x, nA = A.n_alphas;
    __shared__ double ass[3];
    double tp, fn, fp, loca;
    if (num_tr == 0 || num_gt == 0) {                      // hota.py:39-52
        tp = 0.0; fn = (num_tr == 0) ? (double)num_gt : 0.0; fp = (num_tr == 0) ? 0.0 : (double)num_tr;
        loca = 1.0;
        if (lane < 3) ass[lane] = 0.0;
        __syncwarp();
    } else {
        if (*A.status) return;
        tp = (double)A.tot[0 * nA + a]; fn = (double)A.tot[1 * nA + a]; fp = (double)A.tot[2 * nA + a];
        const long long cells = (long long)A.n_gt_ids * A.n_tr_ids;
        const int* m = A.mc + (size_t)a * cells;
        const int nt_ids = A.n_tr_ids;
        if (lane < 3) {
            const int which = lane;
            ass[which] = np_pairwise_sum([&](long long k) {
                const int gi = (int)(k / nt_ids), ti = (int)(k - (long long)gi * nt_ids);
                const double mv = (double)m[k], gc = (double)A.gcnt[gi], tc = (double)A.tcnt[ti];
                const double den = which == 0 ? fmax(1.0, __dsub_rn(__dadd_rn(gc, tc), mv)) : (which == 1 ? fmax(1.0, gc) : fmax(1.0, tc));
                return __dmul_rn(mv, __ddiv_rn(mv, den));
            }, cells);
        }
        double l = 0.0;
        if (lane == 3) { for (int f = 0; f < A.n_frames; ++f) l = __dadd_rn(l, A.loc_part[(size_t)f * nA + a]); }
        l = __shfl_sync(0xffffffffu, l, 3);
        __syncwarp();
        loca = __ddiv_rn(fmax(1e-10, l), fmax(1e-10, tp));                                  // :151
    }
    if (lane == 0) {
        const double d1 = fmax(1.0, tp);
        const double assa = __ddiv_rn(ass[0], d1), assre = __ddiv_rn(ass[1], d1), asspr = __ddiv_rn(ass[2], d1);
        const double detre = __ddiv_rn(tp, fmax(1.0, __dadd_rn(tp, fn)));
        const double detpr = __ddiv_rn(tp, fmax(1.0, __dadd_rn(tp, fp)));
        const double deta = __ddiv_rn(tp, fmax(1.0, __dadd_rn(__dadd_rn(tp, fn), fp)));
        double* o = A.out;
        o[TK_HOTA_HOTA * nA + a] = __dsqrt_rn(__dmul_rn(deta, assa));
        o[TK_HOTA_DETA * nA + a] = deta;
        o[TK_HOTA_ASSA * nA + a] = assa;
        o[TK_HOTA_DETRE * nA + a] = detre;
        o[TK_HOTA_DETPR * nA + a] = detpr;
        o[TK_HOTA_ASSRE * nA + a] = assre;
        o[TK_HOTA_ASSPR * nA + a] = asspr;
        o[TK_HOTA_LOCA * nA + a] = loca;
        o[TK_HOTA_TP * nA + a] = tp;
        o[TK_HOTA_FN * nA + a] = fn;
        o[TK_HOTA_FP * nA + a] = fp;
    }
}

inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

struct Carve {
    size_t pair_off, sims, aux, pot, gcnt, tcnt, mc, loc, tot, end;
    Carve(int F, int n_gt_ids, int n_tr_ids, int n_alphas, long long pairs_cap) {
        size_t o = 0;
        const size_t cells = (size_t)n_gt_ids * n_tr_ids;
        pair_off = o; o = align16(o + sizeof(long long) * ((size_t)F + 1));
        sims = o; o = align16(o + sizeof(double) * (size_t)pairs_cap);
        aux = o; o = align16(o + sizeof(double) * (size_t)pairs_cap);
        pot = o; o = align16(o + sizeof(double) * cells);
        gcnt = o; o = align16(o + sizeof(int) * (size_t)n_gt_ids);
        tcnt = o; o = align16(o + sizeof(int) * (size_t)n_tr_ids);
        mc = o; o = align16(o + sizeof(int) * cells * n_alphas);
        loc = o; o = align16(o + sizeof(double) * (size_t)F * n_alphas);
        tot = o; o = align16(o + sizeof(unsigned long long) * 3 * n_alphas);
        end = o;
    }
};

}  // namespace

extern "C" {

int tk_hota_workspace_bytes(int n_frames, int n_gt_ids, int n_tr_ids, int n_alphas, long long pairs_cap, long long* bytes_out) {
    if (!bytes_out || n_frames < 0 || n_gt_ids < 0 || n_tr_ids < 0 || n_alphas <= 0 || n_alphas > MAX_ALPHAS || pairs_cap < 0)
        return TK_ERR_ARG;
    *bytes_out = (long long)Carve(n_frames, n_gt_ids, n_tr_ids, n_alphas, pairs_cap).end;
    return TK_OK;
}

int tk_hota_sequence(const double* gt_boxes_xywh, const int* gt_ids, const int* gt_offsets, long long n_gt_rows,
                     const double* tr_boxes_xywh, const int* tr_ids, const int* tr_offsets, long long n_tr_rows, int n_frames,
                     int n_gt_ids, int n_tr_ids, int max_gt_per_frame, int max_tr_per_frame, const double* alphas_host,
                     int n_alphas, long long pairs_cap, void* workspace, long long workspace_bytes, double* out, int* status_dev,
                     void* stream) {
    if (!gt_offsets || !tr_offsets || !alphas_host || !workspace || !out || !status_dev || n_frames <= 0 || n_gt_ids < 0 ||
        n_tr_ids < 0 || max_gt_per_frame < 0 || max_tr_per_frame < 0 || n_alphas <= 0 || n_alphas > MAX_ALPHAS || pairs_cap < 0 ||
        n_gt_rows < 0 || n_tr_rows < 0)
        return TK_ERR_ARG;
    if ((n_gt_rows > 0 && (!gt_boxes_xywh || !gt_ids)) || (n_tr_rows > 0 && (!tr_boxes_xywh || !tr_ids))) return TK_ERR_ARG;
    const Carve cv(n_frames, n_gt_ids, n_tr_ids, n_alphas, pairs_cap);
    if ((long long)cv.end > workspace_bytes) return TK_ERR_CAPACITY;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)workspace;
    HotaArgs A;
    A.gt_boxes = gt_boxes_xywh; A.gt_ids = gt_ids; A.gt_off = gt_offsets;
    A.tr_boxes = tr_boxes_xywh; A.tr_ids = tr_ids; A.tr_off = tr_offsets;
    A.n_frames = n_frames; A.n_gt_ids = n_gt_ids; A.n_tr_ids = n_tr_ids; A.max_g = max_gt_per_frame; A.max_t = max_tr_per_frame;
    A.n_alphas = n_alphas; A.pairs_cap = pairs_cap;
    for (int i = 0; i < MAX_ALPHAS; ++i) A.alphas[i] = i < n_alphas ? alphas_host[i] : 0.0;
    A.pair_off = (long long*)(ws + cv.pair_off); A.sims = (double*)(ws + cv.sims); A.aux = (double*)(ws + cv.aux);
    A.pot = (double*)(ws + cv.pot); A.gcnt = (int*)(ws + cv.gcnt); A.tcnt = (int*)(ws + cv.tcnt); A.mc = (int*)(ws + cv.mc);
    A.loc_part = (double*)(ws + cv.loc); A.tot = (unsigned long long*)(ws + cv.tot);
    A.status = status_dev; A.out = out;
    if (n_gt_rows > 0 && n_tr_rows > 0) {
        hota_setup_kernel<<<1, 1024, 0, st>>>(A);
        const size_t smem1 = sizeof(double) * ((size_t)max_gt_per_frame + max_tr_per_frame);
        if (smem1 > 200 * 1024) return TK_ERR_CAPACITY;
        TK_CUDA_TRY(cudaFuncSetAttribute(hota_similarity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        hota_similarity_kernel<<<n_frames, 256, smem1, st>>>(A);
        int n_cta = n_gt_ids < 148 ? n_gt_ids : 148;
        if (n_cta < 1) n_cta = 1;
        hota_potential_kernel<<<n_cta, 256, 0, st>>>(A);
        hota_score_kernel<<<n_frames, 256, 0, st>>>(A);
        const int nr_max = max_gt_per_frame < max_tr_per_frame ? max_gt_per_frame : max_tr_per_frame;
        const int nc_max = max_gt_per_frame < max_tr_per_frame ? max_tr_per_frame : max_gt_per_frame;
        const size_t smem2 = tk::lsap_scipy_scratch_bytes(nr_max, nc_max) + sizeof(int) * (size_t)max_gt_per_frame;
        if (smem2 > 220 * 1024) return TK_ERR_CAPACITY;
        TK_CUDA_TRY(cudaFuncSetAttribute(hota_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        hota_match_kernel<<<n_frames, 32, smem2, st>>>(A);
    }
    hota_final_kernel<<<n_alphas, 32, 0, st>>>(A, n_gt_rows, n_tr_rows);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

}  // extern "C"
