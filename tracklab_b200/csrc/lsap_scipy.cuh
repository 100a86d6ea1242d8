// scipy.optimize.linear_sum_assignment on one warp, operation by operation — INCLUDING its tie-breaking.
//
// The StrongSORT plugins solve min_cost_matching with scipy (/root/reference/plugins/track/strong_sort/sort/
// linear_assignment.py:55, bpbreid_strong_sort/sort/linear_assignment.py:56) on a matrix whose infeasible entries were all set to
// the SAME value `max_distance + 1e-5` (:53-54). The solver assigns every row of the smaller side, so which of the equal-cost
// infeasible pairs it returns is decided purely by its tie-breaking — and the reference then appends the detections of those
// rejected pairs to `unmatched_detections` in row order (:62-68), which fixes the order in which new tracks are born, i.e. the
// track ids. An optimal-but-different solver reproduces the partition of detections into tracks but not the id numbering.
// This file restates scipy's solver (rectangular_lsap.cpp: Crouse's shortest augmenting paths) so that the ids are bit-exact:
//   * rows are augmented in index order (the caller passes the problem with rows <= cols, transposed like scipy does);
//   * the not-yet-scanned columns live in a vector filled in reverse (nc-1 .. 0); the chosen column is removed by moving the
//     last element into its place;
//   * the vector is walked in order; a column replaces the current choice when its path cost is lower, or equal and the column
//     is unassigned  =>  the choice is the LAST unassigned column with the minimum cost if there is one, else the FIRST
//     column with the minimum cost (the warp evaluates this rule with three reductions instead of a serial walk);
//   * r = ((minVal + cost) - u[i]) - v[j] in this order in float64 (no contraction: additions only).
// oracle/assign_np.py::lsap_scipy_restated is the same algorithm in NumPy, pinned to scipy itself on tie-heavy matrices
// (tests/test_oracle_cpu.py); tests/test_pairwise_gpu.py pins this file to scipy through tk_lsap_scipy_batched.
#pragma once
#include "tk_common.cuh"

namespace tk {

// Shared-memory form (any nc): the `remaining` vector, duals and path costs live in the caller's scratch. Used above 512 columns.
template <class CostFn>
static __device__ __noinline__ bool lsap_scipy_warp_mem(int nr, int nc, CostFn C, double* u, double* v, double* spc, int* path,
                                                        int* col4row, int* row4col, int* remaining, unsigned char* SR,
                                                        unsigned char* SC) {
    const int lane = lane_id();
    const unsigned full = 0xffffffffu;
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    for (int i = lane; i < nr; i += 32) { u[i] = 0.0; col4row[i] = -1; }
    for (int j = lane; j < nc; j += 32) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
    __syncwarp();
    for (int cur = 0; cur < nr; ++cur) {
        for (int it = lane; it < nc; it += 32) { remaining[it] = nc - it - 1; spc[it] = INF; SC[it] = 0; }
        for (int i = lane; i < nr; i += 32) SR[i] = 0;
        __syncwarp();
        int num_remaining = nc, i = cur, sink = -1;
        double minVal = 0.0;
        while (sink < 0) {
            if (lane == 0) SR[i] = 1;
            const double ui = u[i];
            double lval = INF;
            int lfirst = 0x7fffffff, llastfree = -1;
            for (int it = lane; it < num_remaining; it += 32) {
                const int j = remaining[it];
                const double r = __dsub_rn(__dsub_rn(__dadd_rn(minVal, C(i, j)), ui), v[j]);
                double s = spc[j];
                if (r < s) { path[j] = i; spc[j] = r; s = r; }
                const bool fr = row4col[j] < 0;
                if (s < lval) { lval = s; lfirst = it; llastfree = fr ? it : -1; }
                else if (s == lval && fr) { llastfree = it; if (lfirst == 0x7fffffff) lfirst = it; }
            }
            // m = min over lanes (ordered 64-bit key -> two 32-bit REDUX)
            const unsigned long long k = ordered_key(lval);
            const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
            const unsigned mhi = __reduce_min_sync(full, hi);
            const unsigned mlo = __reduce_min_sync(full, hi == mhi ? lo : 0xffffffffu);
            const bool has = (hi == mhi) && (lo == mlo) && lfirst != 0x7fffffff;
            const double m = __shfl_sync(full, lval, __ffs(__ballot_sync(full, (hi == mhi) && (lo == mlo))) - 1);
            if (!(m < INF)) return false;
            const int glast = (int)__reduce_max_sync(full, has ? llastfree : -1);
            const unsigned gfirst = __reduce_min_sync(full, has ? (unsigned)lfirst : 0x7fffffffu);
            const int index = glast >= 0 ? glast : (int)gfirst;
            minVal = m;
            const int j = remaining[index];
            const int r4c = row4col[j];
            if (r4c < 0) sink = j; else i = r4c;
            __syncwarp();
            if (lane == 0) { SC[j] = 1; remaining[index] = remaining[num_remaining - 1]; }
            --num_remaining;
            __syncwarp();
        }
        // dual variables
        for (int i2 = lane; i2 < nr; i2 += 32) {
            if (i2 == cur) u[i2] = __dadd_rn(u[i2], minVal);
            else if (SR[i2]) u[i2] = __dadd_rn(u[i2], __dsub_rn(minVal, spc[col4row[i2]]));
        }
        for (int j2 = lane; j2 < nc; j2 += 32)
            if (SC[j2]) v[j2] = __dsub_rn(v[j2], __dsub_rn(minVal, spc[j2]));
        __syncwarp();
        if (lane == 0) {   // augment along the path
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


// Register-resident form of the same algorithm for nc <= 32 K columns: lane l owns the columns l, l + 32, ... and keeps their
// dual v, shortest-path cost, predecessor, owner row and POSITION in scipy's `remaining` vector in registers, so one Dijkstra
// step is K coalesced cost loads + register arithmetic + five warp reductions instead of a walk through shared-memory indirections
// (measured on B200: ~2000 cycles per step for the shared-memory form, tools/bench_lsap.py). scipy's tie-breaking is a statement
// about positions in `remaining` ("the LAST unassigned column among the minima, else the FIRST one", vector filled nc-1 .. 0,
// removal = move the last element into the hole): the positions are tracked exactly - the removed column gets position -1, the
// column that sat at the end of the vector takes over its position.
template <int K, class CostFn>
static __device__ __noinline__ bool lsap_scipy_warp_reg(int nr, int nc, CostFn C, double* u, double* spc_sh, int* path_sh,
                                                        int* col4row, int* row4col, unsigned char* SR) {
    const int lane = lane_id();
    const unsigned full = 0xffffffffu;
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    double vj[K], sp[K];
    int pos[K], pth[K], r4c[K];
    for (int i = lane; i < nr; i += 32) { u[i] = 0.0; col4row[i] = -1; }
#pragma unroll
    for (int c = 0; c < K; ++c) { vj[c] = 0.0; pth[c] = -1; r4c[c] = -1; }
    for (int j = lane; j < nc; j += 32) row4col[j] = -1;
    __syncwarp();
    for (int cur = 0; cur < nr; ++cur) {
#pragma unroll
        for (int c = 0; c < K; ++c) { const int j = lane + 32 * c; pos[c] = j < nc ? nc - 1 - j : -1; sp[c] = INF; }
        unsigned scanned = 0u;
        for (int i = lane; i < nr; i += 32) SR[i] = 0;
        __syncwarp();
        int num_remaining = nc, i = cur, sink = -1;
        double minVal = 0.0;
        while (sink < 0) {
            if (lane == 0) SR[i] = 1;
            const double ui = u[i];
            double lval = INF;
            int lminpos = 0x7fffffff, lmaxfree = -1;
#pragma unroll
            for (int c = 0; c < K; ++c) {
                const int p = pos[c];
                if (p >= 0) {
                    const double r = __dsub_rn(__dsub_rn(__dadd_rn(minVal, C(i, lane + 32 * c)), ui), vj[c]);
                    if (r < sp[c]) { pth[c] = i; sp[c] = r; }
                    const double s = sp[c];
                    const bool fr = r4c[c] < 0;
                    if (s < lval) { lval = s; lminpos = p; lmaxfree = fr ? p : -1; }
                    else if (s == lval) { lminpos = min(lminpos, p); if (fr) lmaxfree = max(lmaxfree, p); }
                }
            }
            const unsigned long long k = ordered_key(lval);
            const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
            const unsigned mhi = __reduce_min_sync(full, hi);
            const unsigned mlo = __reduce_min_sync(full, hi == mhi ? lo : 0xffffffffu);
            const bool has = (hi == mhi) && (lo == mlo) && lminpos != 0x7fffffff;
            const double m = __shfl_sync(full, lval, __ffs(__ballot_sync(full, (hi == mhi) && (lo == mlo))) - 1);
            if (!(m < INF)) return false;
            const int glast = (int)__reduce_max_sync(full, has ? lmaxfree : -1);
            const unsigned gfirst = __reduce_min_sync(full, has ? (unsigned)lminpos : 0x7fffffffu);
            const int index = glast >= 0 ? glast : (int)gfirst;
            // the column at that position: owner lane publishes (column, owner row + 1)
            int packed = 0;
            const int last = num_remaining - 1;
#pragma unroll
            for (int c = 0; c < K; ++c) {
                if (pos[c] == index) { packed = (lane + 32 * c) | ((r4c[c] + 1) << 10); pos[c] = -1; scanned |= 1u << c; }
                else if (pos[c] == last) pos[c] = index;             // remaining[index] = remaining[num_remaining - 1]
            }
            packed = (int)__reduce_max_sync(full, (unsigned)packed);
            const int j = packed & 1023, r4 = (packed >> 10) - 1;
            minVal = m;
            if (r4 < 0) sink = j; else i = r4;
            --num_remaining;
        }
        // duals (the shortest-path costs are needed by column index: publish them once per augmentation)
#pragma unroll
        for (int c = 0; c < K; ++c) {
            const int j = lane + 32 * c;
            if (j < nc) { spc_sh[j] = sp[c]; path_sh[j] = pth[c]; }
            if (scanned & (1u << c)) vj[c] = __dsub_rn(vj[c], __dsub_rn(minVal, sp[c]));
        }
        __syncwarp();
        for (int i2 = lane; i2 < nr; i2 += 32) {
            if (i2 == cur) u[i2] = __dadd_rn(u[i2], minVal);
            else if (SR[i2]) u[i2] = __dadd_rn(u[i2], __dsub_rn(minVal, spc_sh[col4row[i2]]));
        }
        __syncwarp();
        if (lane == 0) {   // augment along the path
            int j = sink;
            while (true) {
                const int r = path_sh[j];
                row4col[j] = r;
                const int t = col4row[r];
                col4row[r] = j;
                j = t;
                if (r == cur) break;
            }
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < K; ++c) { const int j = lane + 32 * c; if (j < nc) r4c[c] = row4col[j]; }
    }
    return true;
}

// One warp (all 32 lanes) must call this. C(i, j): cost of solver row i (< nr) and solver column j (< nc), nr <= nc.
// Scratch (shared memory): u[nr], v[nc], spc[nc] doubles; path[nc], col4row[nr], row4col[nc], remaining[nc] ints;
// SR[nr], SC[nc] bytes. On return col4row[i] is the column of row i (every row is assigned). false: infeasible (inf / NaN costs).
template <class CostFn>
static __device__ __forceinline__ bool lsap_scipy_warp(int nr, int nc, CostFn C, double* u, double* v, double* spc, int* path,
                                                       int* col4row, int* row4col, int* remaining, unsigned char* SR,
                                                       unsigned char* SC) {
    if (nc <= 32) return lsap_scipy_warp_reg<1>(nr, nc, C, u, spc, path, col4row, row4col, SR);
    if (nc <= 64) return lsap_scipy_warp_reg<2>(nr, nc, C, u, spc, path, col4row, row4col, SR);
    if (nc <= 128) return lsap_scipy_warp_reg<4>(nr, nc, C, u, spc, path, col4row, row4col, SR);
    if (nc <= 256) return lsap_scipy_warp_reg<8>(nr, nc, C, u, spc, path, col4row, row4col, SR);
    if (nc <= 512) return lsap_scipy_warp_reg<16>(nr, nc, C, u, spc, path, col4row, row4col, SR);
    return lsap_scipy_warp_mem(nr, nc, C, u, v, spc, path, col4row, row4col, remaining, SR, SC);
}

// bytes of scratch for a problem with up to `nr_max` rows and `nc_max` columns (aligned to 8)
__host__ __device__ inline size_t lsap_scipy_scratch_bytes(int nr_max, int nc_max) {
    size_t b = sizeof(double) * ((size_t)nr_max + 2 * (size_t)nc_max) + sizeof(int) * ((size_t)nr_max + 3 * (size_t)nc_max) +
               (size_t)nr_max + (size_t)nc_max;
    return (b + 15) & ~(size_t)15;
}

struct LsapScratch {
    double *u, *v, *spc;
    int *path, *col4row, *row4col, *remaining;
    unsigned char *SR, *SC;
    __device__ void carve(unsigned char* base, int nr_max, int nc_max) {
        u = (double*)base; v = u + nr_max; spc = v + nc_max;
        path = (int*)(spc + nc_max); col4row = path + nc_max; row4col = col4row + nr_max; remaining = row4col + nc_max;
        SR = (unsigned char*)(remaining + nc_max); SC = SR + nr_max;
    }
};

}  // namespace tk
