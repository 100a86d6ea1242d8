// CTA-resident rectangular linear assignment: parallel row reduction + warp-resident shortest augmenting paths.
//
// Replaces, on device, the solvers the reference calls on the host:
//   lap.lapjv(cost, extend_cost=True, cost_limit=L)   /root/reference/plugins/track/byte_track/matching.py:37-48
//                                                      /root/reference/plugins/track/oc_sort/association.py:187-191
//   scipy.optimize.linear_sum_assignment               /root/reference/plugins/track/strong_sort/sort/linear_assignment.py:55
//
// lap's (n+m)^2 extension with fill L/2 minimises sum(c_ij - L) over partial matchings, which is the
// rectangular problem on c' = min(c - L, 0) with zero-cost pairs dropped afterwards; callers store c'
// (or c itself when there is no limit) in `C`, oriented so that rows <= cols.
//
// Design for sm_100a. The problems are 10..160 wide and one per frame, i.e. latency-bound:
//   phase 1 (whole CTA): every row finds its cheapest column in parallel (one thread per row, rows padded to
//           an odd pitch so the column walk is bank-conflict free); rows whose cheapest column is claimed by
//           nobody else are assigned outright with duals u_i = row minimum, v_j = 0 — a dual-feasible start.
//           In tracking most rows are settled here. In limit mode rows without a negative entry are left
//           unmatched straight away (they can only add zero).
//   phase 2 (one warp): classic shortest-augmenting-path search for the few contested rows; column duals and
//           shortest-path costs live in registers (K columns per lane, K a template parameter), the arg-min per
//           Dijkstra step is 3 REDUX + 1 ballot (tk::warp_argmin), no block barrier inside.
#pragma once
#include "tk_common.cuh"

namespace tk {

constexpr int LAP_MAX_COLS = 256;

// leading dimension callers must use for an `nc`-column problem (odd => conflict-free column walks)
__host__ __device__ __forceinline__ int lap_pitch(int nc) { return nc | 1; }

template <int K>
static __device__ __noinline__ bool lap_sap_warp(const double* __restrict__ C, int ld, int nr, int nc,
                                                 double* u, int* col4row, int* row4col, int* path) {
    const int lane = lane_id();
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    double v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = 0.0;
    for (int cur = 0; cur < nr; ++cur) {
        if (col4row[cur] != -1) continue;   // settled in phase 1 (or skipped)
        double spc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) spc[k] = INF;
        unsigned scanned = 0u;
        int i = cur, sink = -1;
        double minVal = 0.0;
        while (sink < 0) {
            const double ui = u[i];
            const double* row = C + (size_t)i * ld;
            double bestv = INF;
            int bestj = 0x7fffffff;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int j = lane + 32 * k;
                if (j < nc && !((scanned >> k) & 1u)) {
                    const double r = minVal + row[j] - ui - v[k];
                    if (r < spc[k]) { spc[k] = r; path[j] = i; }
                    if (spc[k] < bestv) { bestv = spc[k]; bestj = j; }
                }
            }
            double mv; int jmin;
            warp_argmin(bestv, bestj, mv, jmin);
            if (!(mv < INF)) return false;
            minVal = mv;
            if (lane == (jmin & 31)) scanned |= 1u << (jmin >> 5);
            const int r4c = row4col[jmin];
            if (r4c < 0) sink = jmin; else i = r4c;
        }
        __syncwarp();
        // dual update from the column side: rows in SR\{cur} are exactly row4col[j], j in SC\{sink}
        if (lane == 0) u[cur] += minVal;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if ((scanned >> k) & 1u) {
                const int j = lane + 32 * k;
                const double d = minVal - spc[k];
                if (j != sink) u[row4col[j]] += d;
                v[k] -= d;
            }
        }
        __syncwarp();
        if (lane == 0) {
            int j = sink;
            while (true) {
                const int r = path[j];
                row4col[j] = r;
                const int t = col4row[r];
                col4row[r] = j;
                j = t;
                if (r == cur) break;
            }
        }
        __syncwarp();
    }
    return true;
}

// Whole CTA must call this (contains __syncthreads).
//   C[i*ld + j], 0<=i<nr<=nc<=LAP_MAX_COLS, ld = lap_pitch(nc) recommended.
//   limit_mode: costs are min(c - L, 0); rows without a negative entry stay unmatched (col4row = -1).
//   u[nr] (double), col4row[nr], row4col[nc], path[nc]: shared-memory scratch.
// On return col4row[i] = column of row i (or -1 for a row skipped in limit mode). Returns false if infeasible.
static __device__ bool lap_solve_cta(const double* __restrict__ C, int ld, int nr, int nc, bool limit_mode,
                                     double* u, int* col4row, int* row4col, int* path, int* ok_flag) {
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int j = tid; j < nc; j += nth) row4col[j] = 0x7fffffff;
    if (tid == 0) *ok_flag = 1;
    __syncthreads();
    // phase 1: row minima, smallest column index on ties
    for (int i = tid; i < nr; i += nth) {
        const double* row = C + (size_t)i * ld;
        double mv = row[0];
        int mj = 0;
        for (int j = 1; j < nc; ++j) {
            const double c = row[j];
            if (c < mv) { mv = c; mj = j; }
        }
        u[i] = mv;
        path[i] = mj;
        if (!limit_mode || mv < 0.0) atomicMin(&row4col[mj], i);
    }
    __syncthreads();
    for (int i = tid; i < nr; i += nth) {
        if (limit_mode && !(u[i] < 0.0)) col4row[i] = -2;
        else col4row[i] = (row4col[path[i]] == i) ? path[i] : -1;
    }
    __syncthreads();
    for (int j = tid; j < nc; j += nth) if (row4col[j] == 0x7fffffff) row4col[j] = -1;
    __syncthreads();
    // phase 2: augment the contested rows on one warp
    if (warp_id() == 0) {
        bool ok;
        if (nc <= 32) ok = lap_sap_warp<1>(C, ld, nr, nc, u, col4row, row4col, path);
        else if (nc <= 64) ok = lap_sap_warp<2>(C, ld, nr, nc, u, col4row, row4col, path);
        else if (nc <= 128) ok = lap_sap_warp<4>(C, ld, nr, nc, u, col4row, row4col, path);
        else ok = lap_sap_warp<8>(C, ld, nr, nc, u, col4row, row4col, path);
        if (!ok && lane_id() == 0) *ok_flag = 0;
    }
    __syncthreads();
    for (int i = tid; i < nr; i += nth) if (col4row[i] == -2) col4row[i] = -1;
    __syncthreads();
    return *ok_flag != 0;
}

}  // namespace tk
