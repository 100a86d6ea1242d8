// Warp-resident rectangular linear assignment (shortest augmenting paths with dual variables).
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
// Design for sm_100a: the problems are 10..160 wide and strictly sequential per frame, so one warp owns
// the whole solve: column duals / shortest-path costs live in registers (LAP_K columns per lane), the
// row scan is a conflict-free shared/L1 row read, and the arg-min per Dijkstra step is 3 REDUX + 1 ballot
// (tk::warp_argmin) — no block barrier inside the solver.
#pragma once
#include "tk_common.cuh"

namespace tk {

constexpr int LAP_K = 8;               // columns per lane
constexpr int LAP_MAX_COLS = 32 * LAP_K;

// All 32 lanes of ONE warp must call this together.
//   C[i*ld + j]  cost of row i (0<=i<nr) and column j (0<=j<nc), nr <= nc <= LAP_MAX_COLS
//   u[nr], col4row[nr], row4col[nc], path[nc]  scratch visible to the warp (shared memory)
// On return col4row[i] is the column of row i. Returns false when no finite assignment exists.
static __device__ __noinline__ bool lap_warp(const double* __restrict__ C, int ld, int nr, int nc,
                                      double* u, int* col4row, int* row4col, int* path) {
    const int lane = lane_id();
    const int kmax = (nc + 31) >> 5;
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    double v[LAP_K];
#pragma unroll
    for (int k = 0; k < LAP_K; ++k) v[k] = 0.0;
    for (int j = lane; j < nc; j += 32) row4col[j] = -1;
    for (int i = lane; i < nr; i += 32) { u[i] = 0.0; col4row[i] = -1; }
    __syncwarp();

    for (int cur = 0; cur < nr; ++cur) {
        double spc[LAP_K];
#pragma unroll
        for (int k = 0; k < LAP_K; ++k) spc[k] = INF;
        unsigned scanned = 0u;
        int i = cur, sink = -1;
        double minVal = 0.0;
        while (sink < 0) {
            const double ui = u[i];
            const double* row = C + (size_t)i * ld;
            double bestv = INF;
            int bestj = 0x7fffffff;
#pragma unroll
            for (int k = 0; k < LAP_K; ++k) {
                if (k < kmax) {
                    const int j = lane + 32 * k;
                    if (j < nc && !((scanned >> k) & 1u)) {
                        const double r = minVal + row[j] - ui - v[k];
                        if (r < spc[k]) { spc[k] = r; path[j] = i; }
                        if (spc[k] < bestv) { bestv = spc[k]; bestj = j; }
                    }
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
        for (int k = 0; k < LAP_K; ++k) {
            if (k < kmax && ((scanned >> k) & 1u)) {
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

}  // namespace tk
