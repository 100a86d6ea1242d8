// 1x1 convolution (= GEMM over NHWC activations) with the bias / activation / residual / concat-slice epilogue fused in:
//     dst[m, dst_off + n] = act( sum_k x[m, k] * w[n, k] + bias[n] ) (+ residual[m, res_off + n])       bf16 in, fp32 accumulate, bf16 out
//
// Replaces, for the 1x1 layers of the detector / ReID backbones (YOLOX CSP blocks, ResNet-50 bottlenecks: two thirds of their
// convolutions), the pair "cuDNN convolution -> tk_bias_act_nhwc": the convolution output no longer goes out to HBM/L2 through
// the library's store and comes back through a second kernel. The reference runs these layers inside third-party runtimes
// (onnxruntime / PyTorch CPU: /root/reference/tracklab/wrappers/bbox_detector/rtmlib_api.py:27-30,
// /root/reference/plugins/track/strong_sort/reid_multibackend.py:184-237).
//
// sm_100a design (hand-written, no library): persistent CTAs (one per SM), warp-specialised
//   warp 0   : TMA producer — cp.async.bulk.tensor 2-D tiles of the activation matrix [M, K] (128 rows x 64 channels) and of
//              the weight matrix [N, K] (BLOCK_N x 64) into a multi-stage 128-byte-swizzled shared-memory ring (mbarrier tx counts)
//   warp 1   : one elected thread issues tcgen05.mma (UMMA 128 x BLOCK_N x 16, bf16 -> fp32) into a TMEM accumulator;
//              tcgen05.commit releases the ring slot / publishes the accumulator
//   warp 2   : allocates / frees TMEM (two accumulator stages, so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 4-7: epilogue — tcgen05.ld the accumulator (thread = one output row, 16 columns at a time), bias + SiLU/ReLU
//              (+ residual) in fp32, one rounding to bf16, 32-byte row segments straight into the channel slice of dst
// These layers are HBM-bound (K, N = 32..768 against M = B*H*W up to millions): algorithmic bytes per output row =
// 2K (x) + 2N (dst) (+ 2N residual); the weights are L2 resident.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "tc_gemm.cuh"
#include "tk_common.cuh"
#include "trackkern.h"

namespace {

using namespace tcg;

constexpr int BM = 128;          // rows of an output tile = TMEM lanes
constexpr int UMMA_K = 16;
constexpr int C1_MAX_THREADS = 128 + 32 * 12;   // 4 control warps + up to 12 epilogue warps
constexpr int STG_SUB_BYTES = BM * 128;         // one 64-column group of the staged output tile

struct C1Params {
    long long M;
    int K, N, bk, block_n, n_blocks, stages, tmem_cols, epi_warps, stg_bufs;
    const float* bias;
    const __nv_bfloat16* res;
    int res_pitch, res_off;
    int act;
};

__global__ void __launch_bounds__(C1_MAX_THREADS, 1)
conv1x1_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ CUtensorMap map_d, const C1Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024-byte alignment is required by the 128-byte swizzle; dynamic shared memory starts 1024-aligned only by request
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row_bytes = p.bk * 2;
    const int A_STAGE_BYTES = BM * row_bytes;                         // 16 / 8 / 4 KB
    const int b_stage_bytes = p.block_n * row_bytes;
    const int stage_bytes = A_STAGE_BYTES + ((b_stage_bytes + 1023) & ~1023);   // every stage base stays 1024-aligned
    const int sub_tiles = (p.block_n + 63) / 64;                      // 64-column (128-byte) groups of the output tile
    const int stg_bytes = sub_tiles * STG_SUB_BYTES;                  // one staging buffer: sub_tiles x [128 rows][128 B], 128B-swizzled
    unsigned char* stg = smem + (size_t)p.stages * stage_bytes;       // [stg_bufs][stg_bytes], 1024-aligned
    unsigned char* tail = stg + (size_t)p.stg_bufs * stg_bytes;
    uint64_t* full_bar = (uint64_t*)tail;                             // [stages]
    uint64_t* empty_bar = full_bar + p.stages;                        // [stages]
    uint64_t* tfull_bar = empty_bar + p.stages;                       // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                             // [2]
    uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);
    float* s_bias = (float*)(tmem_slot + 4);                          // [N]
    const int n_epi = p.epi_warps * 32;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_d) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(tfull_bar + a), 1); mbar_init(smem_u32(tempty_bar + a), (uint32_t)n_epi); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // TMEM: 2 accumulator stages of block_n fp32 columns (power of two >= 32 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) s_bias[i] = p.bias ? p.bias[i] : 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long m_tiles = (p.M + BM - 1) / BM;
    const long long n_tiles = m_tiles * p.n_blocks;
    const int BK = p.bk;
    const int k_blocks = (p.K + BK - 1) / BK;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int m0 = (int)((t / p.n_blocks) * BM), n0 = (int)(t % p.n_blocks) * p.block_n;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
                    const uint32_t fb = smem_u32(full_bar + stage);
                    mbar_expect_tx(fb, (uint32_t)(A_STAGE_BYTES + b_stage_bytes));
                    unsigned char* sa = smem + (size_t)stage * stage_bytes;
                    tma_load_2d(smem_u32(sa), &map_x, fb, kb * BK, m0);                      // rows/channels past the edge arrive as zeros
                    tma_load_2d(smem_u32(sa + A_STAGE_BYTES), &map_w, fb, kb * BK, n0);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            // instruction descriptor: D fp32, A/B bf16, both K-major, N = block_n, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
                mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);      // the epilogue has drained this accumulator stage
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.block_n;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(smem_u32(full_bar + stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t da = make_desc_kmajor(sa, row_bytes), db = make_desc_kmajor(sa + A_STAGE_BYTES, row_bytes);
                    for (int k = 0; k < BK / UMMA_K; ++k)      // advance 32 bytes (16 bf16) inside the swizzle atom per UMMA_K step
                        umma_bf16(tmem_d, da + (uint64_t)(k * UMMA_K * 2 >> 4), db + (uint64_t)(k * UMMA_K * 2 >> 4), idesc, (kb | k) != 0);
                    umma_commit(smem_u32(empty_bar + stage));            // slot reusable once these MMAs have read it
                    if (kb == k_blocks - 1) umma_commit(smem_u32(tfull_bar + acc));
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> bias/activation/residual -> bf16 -> swizzled staging tile -> TMA store =====
        // Thread = one output row (TMEM lane); the 16-column chunks of the tile are dealt round-robin to the warp groups that
        // share a lane quadrant. A row-per-thread store straight to global memory touches 32 different 128-byte lines per
        // instruction (L1tex wavefront-bound, ~1 line/clock); the staging tile + cp.async.bulk.tensor store writes full lines.
        const int q = warp & 3;                       // TMEM lane quadrant this warp may access
        const int grp = (warp - 4) >> 2, n_grp = p.epi_warps >> 2;
        const int row = q * 32 + lane;
        const bool store_thread = (warp == 4 && lane == 0);
        uint32_t it = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
            const int m0 = (int)((t / p.n_blocks) * BM);
            const long long m = (long long)m0 + row;
            const int n0 = (int)(t % p.n_blocks) * p.block_n;
            unsigned char* sbuf = stg + (size_t)(p.stg_bufs == 2 ? acc : 0) * stg_bytes;
            mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
            tc_fence_after();
            // the staging buffer is free once the bulk store that last read it has finished reading shared memory
            if (store_thread) {
                if (p.stg_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            asm volatile("bar.sync 1, %0;" ::"r"(n_epi) : "memory");
            const uint32_t taddr = tmem_base + acc * (uint32_t)p.block_n + ((uint32_t)(q * 32) << 16);
            const int ncols = min(p.block_n, p.N - n0);
            const __nv_bfloat16* rrow = (p.res && m < p.M) ? p.res + (size_t)m * p.res_pitch + p.res_off + n0 : nullptr;
            for (int c = grp * 16; c < ncols; c += 16 * n_grp) {
                uint32_t v[16];
                tmem_ld16(taddr + (uint32_t)c, v);
                tmem_ld_wait();
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = act_apply(__uint_as_float(v[j]) + s_bias[n0 + c + j], p.act);
                if (rrow) {
                    const uint4 r0 = *(const uint4*)(rrow + c), r1 = *(const uint4*)(rrow + c + 8);
                    const __nv_bfloat162* rp0 = (const __nv_bfloat162*)&r0;
                    const __nv_bfloat162* rp1 = (const __nv_bfloat162*)&r1;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 a = __bfloat1622float2(rp0[j]), b = __bfloat1622float2(rp1[j]);
                        f[2 * j] += a.x; f[2 * j + 1] += a.y; f[8 + 2 * j] += b.x; f[8 + 2 * j + 1] += b.y;
                    }
                }
                if (p.act == TK_ACT_RELU_AFTER_RESIDUAL) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
                }
                uint4 o0, o1;
                __nv_bfloat162* op0 = (__nv_bfloat162*)&o0;
                __nv_bfloat162* op1 = (__nv_bfloat162*)&o1;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    op0[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                    op1[j] = __floats2bfloat162_rn(f[8 + 2 * j], f[8 + 2 * j + 1]);
                }
                // staging layout = what the SWIZZLE_128B tensor map expects: sub-tile (c / 64), row pitch 128 B, the 16-byte chunk
                // index XORed with (row mod 8) -> a quarter-warp writes 8 rows x 16 B to 32 distinct banks
                unsigned char* sub = sbuf + (size_t)(c >> 6) * STG_SUB_BYTES + (size_t)row * 128;
                const int ch = (c & 63) >> 3;                     // first of the two 16-byte chunks of these 16 columns
                *(uint4*)(sub + (((ch) ^ (row & 7)) << 4)) = o0;
                *(uint4*)(sub + (((ch + 1) ^ (row & 7)) << 4)) = o1;
            }
            tc_fence_before();
            mbar_arrive(smem_u32(tempty_bar + acc));    // this thread's accumulator reads are done -> MMA of tile it+2 may start
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA (async proxy)
            asm volatile("bar.sync 1, %0;" ::"r"(n_epi) : "memory");
            if (store_thread) {
                for (int j = 0; j * 64 < ncols; ++j)     // rows >= M and columns >= N are clipped by the tensor map
                    tma_store_2d(&map_d, smem_u32(sbuf + (size_t)j * STG_SUB_BYTES), n0 + j * 64, m0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (store_thread) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before the CTA's shared memory goes away
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// [rows, cols] bf16 matrix, row pitch `pitch` elements, box = box_rows x box_cols columns (swizzle span = box_cols * 2 bytes), zero fill outside
bool make_map(CUtensorMap* map, const void* base, unsigned long long rows, unsigned long long cols, unsigned long long pitch, unsigned box_rows,
              unsigned box_cols) {
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {pitch * 2};
    const cuuint32_t box[2] = {box_cols, box_rows};
    return make_map_nd(map, base, 2, dims, strides, box);
}

int g_sms = 0;

}  // namespace

// [rows, cols] bf16 matrix view for the output store (same box / swizzle as the loads)
int g_epi_warps = 0, g_force_one_cta = -1;

extern "C" int tk_conv1x1_bias_act_bf16(const void* x, long long M, int K, int x_pitch, const void* w, int N, const float* bias,
                                        void* dst, int dst_pitch, int dst_off, const void* residual, int res_pitch, int res_off,
                                        int act, void* stream) {
    if (!x || !w || !dst || M <= 0 || K <= 0 || N <= 0) return TK_ERR_ARG;
    if (act < TK_ACT_NONE || act > TK_ACT_RELU_AFTER_RESIDUAL) return TK_ERR_ARG;
    // 16-byte vector access / TMA stride rules: channel counts, pitches and offsets in multiples of 8; N in multiples of 16 (UMMA N)
    if ((K & 7) || (x_pitch & 7) || (N & 15) || (dst_pitch & 7) || (dst_off & 7) || x_pitch < K || dst_off + N > dst_pitch) return TK_ERR_ARG;
    if (residual && ((res_pitch & 7) || (res_off & 7) || res_off + N > res_pitch)) return TK_ERR_ARG;
    if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)dst | (uintptr_t)residual) & 15) return TK_ERR_ARG;
    if (M > 0x7fffff00ll) return TK_ERR_CAPACITY;
    if (!g_epi_warps) {   // tuning knobs for tools/bench_conv1x1.py (defaults: 8 epilogue warps, 2 CTAs per SM when they fit)
        const char* e = getenv("TK_C1_EPI_WARPS");
        g_epi_warps = e ? atoi(e) : 8;
        if (g_epi_warps != 4 && g_epi_warps != 8 && g_epi_warps != 12) g_epi_warps = 8;
        const char* o = getenv("TK_C1_ONE_CTA");
        g_force_one_cta = o ? atoi(o) : 0;
    }
    // output-channel blocking: one block when N <= 256, else the smallest number of equal blocks (multiples of 16) <= 256
    // several blocks only in multiples of 64 channels: the store boxes are 64 channels wide and are clipped by the tensor extent, not
    // by the block, so a narrower last box of block k would spill into block k+1's channels (N = 320 -> 5 x 64, not 2 x 160)
    int n_blocks = 1;
    while (N / n_blocks > 256 || N % n_blocks || (N / n_blocks) % 16 || (n_blocks > 1 && (N / n_blocks) % 64)) {
        if (++n_blocks > N / 16) return TK_ERR_ARG;
    }
    const int block_n = N / n_blocks;
    int tmem_cols = 32;
    while (tmem_cols < 2 * block_n) tmem_cols <<= 1;
    // channels per stage: no partially out-of-bounds boxes for K = 32 / 48 / 96 ... (they ran at 2.2-3.9 TB/s against 5+ for K = 64)
    const int bk = (K % 64 == 0) ? 64 : ((K % 32 == 0) ? 32 : ((K % 16 == 0) ? 16 : 64));
    const int stage_bytes = BM * bk * 2 + ((block_n * bk * 2 + 1023) & ~1023);
    const int sub_tiles = (block_n + 63) / 64;
    const int stg_bufs = block_n <= 128 ? 2 : 1;
    const size_t stg_total = (size_t)stg_bufs * sub_tiles * STG_SUB_BYTES;
    const int k_blocks = (K + bk - 1) / bk;
    auto smem_for = [&](int st) { return (size_t)1024 + (size_t)st * stage_bytes + stg_total + (2 * st + 4) * 8 + 16 + (size_t)N * 4 + 16; };
    // two CTAs per SM (two independent pipelines, twice the epilogue warps) when 3 stages + staging fit in half an SM and the
    // accumulators fit in half the TMEM; otherwise one CTA with as many stages as fit
    int stages, ctas_per_sm = 1;
    if (!g_force_one_cta && tmem_cols <= 256 && smem_for(k_blocks >= 3 ? 3 : 2) <= 110 * 1024) {
        ctas_per_sm = 2;
        stages = k_blocks >= 3 ? 3 : 2;
        const int cap2 = 6 * (64 / bk) < 12 ? 6 * (64 / bk) : 12;      // smaller stages -> a deeper ring for the same bytes in flight
        while (stages < cap2 && smem_for(stages + 1) <= 110 * 1024) ++stages;
    } else {
        stages = 2;
        const int cap1 = 8 * (64 / bk) < 16 ? 8 * (64 / bk) : 16;
        while (stages < cap1 && smem_for(stages + 1) <= 220 * 1024) ++stages;
        if (smem_for(stages) > 227 * 1024) return TK_ERR_CAPACITY;
    }
    const size_t smem = smem_for(stages);
    CUtensorMap mx, mw, md;
    if (!make_map(&mx, x, (unsigned long long)M, (unsigned long long)K, (unsigned long long)x_pitch, BM, (unsigned)bk)) return TK_ERR_CUDA;
    if (!make_map(&mw, w, (unsigned long long)N, (unsigned long long)K, (unsigned long long)K, (unsigned)block_n, (unsigned)bk)) return TK_ERR_CUDA;
    if (!make_map(&md, (const __nv_bfloat16*)dst + dst_off, (unsigned long long)M, (unsigned long long)N, (unsigned long long)dst_pitch, BM, 64)) return TK_ERR_CUDA;
    if (!g_sms) {
        int dev = 0;
        TK_CUDA_TRY(cudaGetDevice(&dev));
        TK_CUDA_TRY(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    C1Params p;
    p.M = M; p.K = K; p.N = N; p.bk = bk; p.block_n = block_n; p.n_blocks = n_blocks; p.stages = stages; p.tmem_cols = tmem_cols;
    p.epi_warps = g_epi_warps; p.stg_bufs = stg_bufs;
    p.bias = bias; p.res = (const __nv_bfloat16*)residual; p.res_pitch = res_pitch; p.res_off = res_off; p.act = act;
    const long long tiles = ((M + BM - 1) / BM) * n_blocks;
    const long long slots = (long long)g_sms * ctas_per_sm;
    const int grid = (int)(tiles < slots ? tiles : slots);
    TK_CUDA_TRY(cudaFuncSetAttribute(conv1x1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv1x1_tc_kernel<<<grid, 128 + 32 * g_epi_warps, smem, (cudaStream_t)stream>>>(mx, mw, md, p);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}
