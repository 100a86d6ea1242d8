// Shared helpers for libtrackkern (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TK_OK 0
#define TK_ERR_ARG -1
#define TK_ERR_CUDA -2
#define TK_ERR_CAPACITY -3
#define TK_ERR_INFEASIBLE -4

#define TK_CUDA_TRY(expr)                                   \
    do {                                                    \
        cudaError_t _e = (expr);                            \
        if (_e != cudaSuccess) { tk_set_last_cuda_error((int)_e); return TK_ERR_CUDA; } \
    } while (0)

void tk_set_last_cuda_error(int e);

// device-side error flags written into a per-sequence status word
#define TK_DEV_OVERFLOW_TRACKS 1
#define TK_DEV_OVERFLOW_DETS 2
#define TK_DEV_LAP_INFEASIBLE 4
#define TK_DEV_OVERFLOW_OUT 8
#define TK_DEV_BAD_CHOLESKY 16
#define TK_DEV_OVERFLOW_ASSIGN 32   /* more live rows / candidates than the assignment side capacity */
#define TK_DEV_NAN_COST 64          /* appearance cost undefined (no commonly visible part) */

namespace tk {

constexpr int WARP = 32;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// IEEE double -> unsigned key with the same total order (for integer min reductions)
__device__ __forceinline__ unsigned long long ordered_key(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// warp arg-min over (value, index): 2 REDUX + 1 ballot instead of 5 shuffle rounds on 64-bit pairs.
// Ties are resolved towards the smallest index. Every lane must call it; result is warp-uniform.
__device__ __forceinline__ void warp_argmin(double val, int idx, double& out_val, int& out_idx) {
    const unsigned full = 0xffffffffu;
    unsigned long long k = ordered_key(val);
    unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    unsigned mhi = __reduce_min_sync(full, hi);
    unsigned mlo = __reduce_min_sync(full, hi == mhi ? lo : 0xffffffffu);
    bool win = (hi == mhi) && (lo == mlo);
    unsigned midx = __reduce_min_sync(full, win ? (unsigned)idx : 0xffffffffu);
    unsigned b = __ballot_sync(full, win && (unsigned)idx == midx);
    int src = __ffs(b) - 1;
    out_val = __shfl_sync(full, val, src);
    out_idx = (int)midx;
}

// Ordered compaction done by ONE warp: pred(i) for i in [0,n), emit(i, base + rank) for the passing items in index
// order; returns base + count. The serial list edits of the reference become a few ballots instead of a
// dependent-load loop on one thread (~60 cycles per element on shared memory).
template <class Pred, class Emit>
__device__ __forceinline__ int warp_compact(int n, int base, Pred pred, Emit emit) {
    const int lane = lane_id();
    int cnt = base;
    for (int c = 0; c < n; c += 32) {
        const int i = c + lane;
        const bool f = (i < n) && pred(i);
        const unsigned b = __ballot_sync(0xffffffffu, f);
        if (f) emit(i, cnt + __popc(b & ((1u << lane) - 1u)));
        cnt += __popc(b);
    }
    __syncwarp();
    return cnt;
}
__device__ __forceinline__ float4 ld_nc_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// sense-reversing barrier over the CTAs of one video group (all co-resident: cooperative launch)
__device__ __forceinline__ void group_barrier(unsigned* bar, int n) {
    __syncthreads();
    if (n > 1 && threadIdx.x == 0) {
        volatile unsigned* gen = bar + 1;
        const unsigned g = *gen;
        __threadfence();
        if (atomicAdd(bar, 1u) == (unsigned)(n - 1)) {
            bar[0] = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*gen == g) { __nanosleep(64); }
        }
        __threadfence();
    }
    __syncthreads();
}

}  // namespace tk
