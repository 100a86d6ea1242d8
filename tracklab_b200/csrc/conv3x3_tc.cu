// 3x3 convolution (stride 1, padding 1) over NHWC bf16 activations as an implicit GEMM on the 5th-generation tensor cores, with
// the bias / activation / residual / concat-slice epilogue fused in:
//     dst[b, y, x, dst_off + n] = act( sum_{ky,kx,c} x[b, y+ky-1, x+kx-1, c] * w[n, ky, kx, c] + bias[n] ) (+ residual[b, y, x, res_off + n])
//
// Replaces "cuDNN 3x3 convolution -> tk_bias_act_nhwc" in the YOLOX executor (the 3x3 layers carry ~80 % of the detector's FLOPs;
// their library kernels reach ~20 % of the tensor peak at the detector's channel counts and every output makes a second trip
// through the epilogue kernel). Same warp-specialised persistent structure as conv1x1_tc.cu (TMA producer / one-thread tcgen05.mma
// issuer / TMEM double buffer / 8 epilogue warps / swizzled staging tile + TMA store). What is specific to 3x3:
//   * an output tile is a SPATIAL patch of TW x TH pixels (<= 128 = the TMEM lanes) of one image;
//   * the im2col gather is done by the TMA engine: for each of the 9 taps the activation tile is one 4-D tensor-map box
//     {channels, TW, TH, 1} placed at (x0 + kx - 1, y0 + ky - 1): rows outside the image arrive as zeros (= the padding), and the box
//     lands in shared memory as a dense K-major [TW*TH rows][BK channels] tile, exactly the UMMA A operand. No halo staging, no
//     explicit im2col buffer: K loop = 9 taps x (Cin / BK) channel blocks, every step one tcgen05.mma group into the same accumulator;
//   * the weights [Cout][3][3][Cin] are a 3-D tensor map {Cin, 9, Cout}: box {BK, 1, BLOCK_N} at (c0, tap, n0);
//   * BK = 64 / 32 / 16 channels (128 / 64 / 32-byte swizzle) so that channel counts like 48 or 96 need no partially-out-of-bounds boxes;
//   * the residual tile is fetched by TMA into the staging buffer before the epilogue adds into it.
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

constexpr int BM = 128;
constexpr int UMMA_K = 16;
constexpr int C3_MAX_THREADS = 128 + 32 * 12;
constexpr int STG_SUB_BYTES = BM * 128;

struct C3Params {
    int B, H, W, Cin, N;            // output spatial size = input spatial size (stride 1, padding 1)
    int tw, th, tiles_w, tiles_h;   // spatial patch of an output tile and the number of patches per image
    int bk, cblocks;                // channels per pipeline stage, Cin / bk
    int block_n, n_blocks, stages, tmem_cols, epi_warps, stg_bufs;
    const float* bias;
    int has_res;
    int act;
};

__global__ void __launch_bounds__(C3_MAX_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ CUtensorMap map_d, const __grid_constant__ CUtensorMap map_r, const C3Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row_bytes = p.bk * 2;
    const int a_stage_bytes = BM * row_bytes;                         // 16 / 8 / 4 KB (1024-aligned in every case)
    const int b_stage_bytes = p.block_n * row_bytes;                  // block_n multiple of 16 -> multiple of 512; padded to 1024 below
    const int stage_bytes = a_stage_bytes + ((b_stage_bytes + 1023) & ~1023);
    const int sub_tiles = (p.block_n + 63) / 64;
    const int stg_bytes = sub_tiles * STG_SUB_BYTES;
    unsigned char* stg = smem + (size_t)p.stages * stage_bytes;
    unsigned char* tail = stg + (size_t)p.stg_bufs * stg_bytes;
    uint64_t* full_bar = (uint64_t*)tail;                             // [stages]
    uint64_t* empty_bar = full_bar + p.stages;                        // [stages]
    uint64_t* tfull_bar = empty_bar + p.stages;                       // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                             // [2]
    uint64_t* res_bar = tempty_bar + 2;                               // [2] residual tile landed in the staging buffer
    uint32_t* tmem_slot = (uint32_t*)(res_bar + 2);
    float* s_bias = (float*)(tmem_slot + 4);                          // [N]
    const int n_epi = p.epi_warps * 32;
    const int patch_rows = p.tw * p.th;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_d) : "memory");
        if (p.has_res) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_r) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 1); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(tfull_bar + a), 1); mbar_init(smem_u32(tempty_bar + a), (uint32_t)n_epi); mbar_init(smem_u32(res_bar + a), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) s_bias[i] = p.bias ? p.bias[i] : 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long per_image = (long long)p.tiles_w * p.tiles_h;
    const long long n_tiles = (long long)p.B * per_image * p.n_blocks;
    const int k_blocks = 9 * p.cblocks;
    auto tile_coords = [&](long long t, int& b, int& y0, int& x0, int& n0) {
        n0 = (int)(t % p.n_blocks) * p.block_n;
        const long long s = t / p.n_blocks;
        b = (int)(s / per_image);
        const int r = (int)(s % per_image);
        y0 = (r / p.tiles_w) * p.th;
        x0 = (r % p.tiles_w) * p.tw;
    };

    if (warp == 0) {
        // ===== TMA producer: 9 taps x channel blocks per tile =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = (uint32_t)(patch_rows * row_bytes + b_stage_bytes);
            for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int b, y0, x0, n0;
                tile_coords(t, b, y0, x0, n0);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    const int tap = kb / p.cblocks, c0 = (kb - tap * p.cblocks) * p.bk;
                    const int ky = tap / 3, kx = tap - ky * 3;
                    mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
                    const uint32_t fb = smem_u32(full_bar + stage);
                    mbar_expect_tx(fb, tx_bytes);
                    unsigned char* sa = smem + (size_t)stage * stage_bytes;
                    tma_load_4d(smem_u32(sa), &map_x, fb, c0, x0 + kx - 1, y0 + ky - 1, b);     // outside the image: zeros = the padding
                    tma_load_3d(smem_u32(sa + a_stage_bytes), &map_w, fb, c0, tap, n0);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const int ksteps = p.bk / UMMA_K;
            int stage = 0;
            uint32_t phase = 0, it = 0;
            for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
                mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.block_n;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(smem_u32(full_bar + stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t da = make_desc_kmajor(sa, row_bytes), db = make_desc_kmajor(sa + a_stage_bytes, row_bytes);
                    for (int k = 0; k < ksteps; ++k)
                        umma_bf16(tmem_d, da + (uint64_t)(k * UMMA_K * 2 >> 4), db + (uint64_t)(k * UMMA_K * 2 >> 4), idesc, (kb | k) != 0);
                    umma_commit(smem_u32(empty_bar + stage));
                    if (kb == k_blocks - 1) umma_commit(smem_u32(tfull_bar + acc));
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int q = warp & 3;
        const int grp = (warp - 4) >> 2, n_grp = p.epi_warps >> 2;
        const int row = q * 32 + lane;
        const bool store_thread = (warp == 4 && lane == 0);
        uint32_t it = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
            int b, y0, x0, n0;
            tile_coords(t, b, y0, x0, n0);
            const int buf = p.stg_bufs == 2 ? (int)acc : 0;
            const uint32_t use_phase = (p.stg_bufs == 2 ? (it >> 1) : it) & 1u;
            unsigned char* sbuf = stg + (size_t)buf * stg_bytes;
            const int ncols = min(p.block_n, p.N - n0);
            // staging buffer free (the bulk store that last read it is done reading) -> fetch the residual tile into it
            if (store_thread) {
                if (p.stg_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                if (p.has_res) {
                    const uint32_t rb = smem_u32(res_bar + buf);
                    mbar_expect_tx(rb, (uint32_t)(((ncols + 63) / 64) * patch_rows * 128));
                    for (int j = 0; j * 64 < ncols; ++j)
                        tma_load_4d(smem_u32(sbuf + (size_t)j * STG_SUB_BYTES), &map_r, rb, n0 + j * 64, x0, y0, b);
                }
            }
            mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
            tc_fence_after();
            asm volatile("bar.sync 1, %0;" ::"r"(n_epi) : "memory");
            if (p.has_res) mbar_wait(smem_u32(res_bar + buf), use_phase);
            const uint32_t taddr = tmem_base + acc * (uint32_t)p.block_n + ((uint32_t)(q * 32) << 16);
            for (int c = grp * 16; c < ncols; c += 16 * n_grp) {
                uint32_t v[16];
                tmem_ld16(taddr + (uint32_t)c, v);
                tmem_ld_wait();
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = act_apply(__uint_as_float(v[j]) + s_bias[n0 + c + j], p.act);
                unsigned char* sub = sbuf + (size_t)(c >> 6) * STG_SUB_BYTES + (size_t)row * 128;
                const int ch = (c & 63) >> 3;
                uint4* p0 = (uint4*)(sub + (((ch) ^ (row & 7)) << 4));
                uint4* p1 = (uint4*)(sub + (((ch + 1) ^ (row & 7)) << 4));
                if (p.has_res && row < patch_rows) {   // residual tile sits in the staging buffer in the same swizzled layout
                    const uint4 r0 = *p0, r1 = *p1;
                    const __nv_bfloat162* rp0 = (const __nv_bfloat162*)&r0;
                    const __nv_bfloat162* rp1 = (const __nv_bfloat162*)&r1;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 a = __bfloat1622float2(rp0[j]), bb = __bfloat1622float2(rp1[j]);
                        f[2 * j] += a.x; f[2 * j + 1] += a.y; f[8 + 2 * j] += bb.x; f[8 + 2 * j + 1] += bb.y;
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
                if (row < patch_rows) { *p0 = o0; *p1 = o1; }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(tempty_bar + acc));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, %0;" ::"r"(n_epi) : "memory");
            if (store_thread) {
                for (int j = 0; j * 64 < ncols; ++j)      // pixels outside the image and channels >= N are clipped by the tensor map
                    tma_store_4d(&map_d, smem_u32(sbuf + (size_t)j * STG_SUB_BYTES), n0 + j * 64, x0, y0, b);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (store_thread) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

int g_sms3 = 0, g_epi3 = 0, g_one3 = 0;

// spatial patch (tw x th <= 128 pixels) that wastes the fewest tensor-core rows / clipped pixels on a W x H image
void pick_patch(int W, int H, int& tw, int& th) {
    double best = -1.0;
    tw = 16; th = 8;
    for (int w = 1; w <= (W < 128 ? W : 128); ++w) {
        if (w > 256) break;
        int h = 128 / w;
        if (h > H) h = H;
        if (h < 1 || h > 256) continue;
        const long long covered = (long long)((W + w - 1) / w) * w * ((H + h - 1) / h) * h;
        const double eff = (double)W * H / (double)covered * ((double)(w * h) / 128.0);
        // prefer wider rows on ties (longer contiguous runs per TMA row)
        if (eff > best + 1e-9 || (eff > best - 1e-9 && w > tw)) { best = eff; tw = w; th = h; }
    }
}

}  // namespace

extern "C" int tk_conv3x3_bias_act_bf16(const void* x, int n_images, int H, int W, int Cin, const void* w, int N, const float* bias,
                                        void* dst, int dst_pitch, int dst_off, const void* residual, int res_pitch, int res_off,
                                        int act, void* stream) {
    if (!x || !w || !dst || n_images <= 0 || H <= 0 || W <= 0 || Cin <= 0 || N <= 0) return TK_ERR_ARG;
    if (act < TK_ACT_NONE || act > TK_ACT_RELU_AFTER_RESIDUAL) return TK_ERR_ARG;
    if ((Cin & 15) || (N & 15) || (dst_pitch & 7) || (dst_off & 7) || dst_off + N > dst_pitch) return TK_ERR_ARG;
    if (residual && ((res_pitch & 7) || (res_off & 7) || res_off + N > res_pitch)) return TK_ERR_ARG;
    if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)dst | (uintptr_t)residual) & 15) return TK_ERR_ARG;
    if (!g_epi3) {
        const char* e = getenv("TK_C1_EPI_WARPS");
        g_epi3 = e ? atoi(e) : 8;
        if (g_epi3 != 4 && g_epi3 != 8 && g_epi3 != 12) g_epi3 = 8;
        const char* o = getenv("TK_C1_ONE_CTA");
        g_one3 = o ? atoi(o) : 0;
    }
    const int bk = (Cin % 64 == 0) ? 64 : ((Cin % 32 == 0) ? 32 : 16);
    // several blocks only in multiples of 64 channels: the store boxes are 64 channels wide and are clipped by the tensor extent, not
    // by the block, so a narrower last box of block k would spill into block k+1's channels (N = 320 -> 5 x 64, not 2 x 160)
    int n_blocks = 1;
    while (N / n_blocks > 256 || N % n_blocks || (N / n_blocks) % 16 || (n_blocks > 1 && (N / n_blocks) % 64)) {
        if (++n_blocks > N / 16) return TK_ERR_ARG;
    }
    const int block_n = N / n_blocks;
    int tmem_cols = 32;
    while (tmem_cols < 2 * block_n) tmem_cols <<= 1;
    int tw, th;
    pick_patch(W, H, tw, th);
    const int row_bytes = bk * 2;
    const int stage_bytes = BM * row_bytes + ((block_n * row_bytes + 1023) & ~1023);
    const int sub_tiles = (block_n + 63) / 64;
    const int stg_bufs = block_n <= 128 ? 2 : 1;
    const size_t stg_total = (size_t)stg_bufs * sub_tiles * STG_SUB_BYTES;
    auto smem_for = [&](int st) { return (size_t)1024 + (size_t)st * stage_bytes + stg_total + (2 * st + 6) * 8 + 16 + (size_t)N * 4 + 16; };
    int stages, ctas_per_sm = 1;
    if (!g_one3 && tmem_cols <= 256 && smem_for(4) <= 110 * 1024) {
        ctas_per_sm = 2;
        stages = 4;
        while (stages < 8 && smem_for(stages + 1) <= 110 * 1024) ++stages;
    } else {
        stages = 3;
        while (stages < 10 && smem_for(stages + 1) <= 220 * 1024) ++stages;
        if (smem_for(stages) > 227 * 1024) return TK_ERR_CAPACITY;
    }
    const size_t smem = smem_for(stages);
    CUtensorMap mx, mw, md, mr;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_images};
        const cuuint64_t str[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        const cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)tw, (cuuint32_t)th, 1};
        if (!make_map_nd(&mx, x, 4, dims, str, box)) return TK_ERR_CUDA;
    }
    {
        const cuuint64_t dims[3] = {(cuuint64_t)Cin, 9, (cuuint64_t)N};
        const cuuint64_t str[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)9 * Cin * 2};
        const cuuint32_t box[3] = {(cuuint32_t)bk, 1, (cuuint32_t)block_n};
        if (!make_map_nd(&mw, w, 3, dims, str, box)) return TK_ERR_CUDA;
    }
    {
        const cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_images};
        const cuuint64_t str[3] = {(cuuint64_t)dst_pitch * 2, (cuuint64_t)W * dst_pitch * 2, (cuuint64_t)H * W * dst_pitch * 2};
        const cuuint32_t box[4] = {64, (cuuint32_t)tw, (cuuint32_t)th, 1};
        if (!make_map_nd(&md, (const __nv_bfloat16*)dst + dst_off, 4, dims, str, box)) return TK_ERR_CUDA;
        if (residual) {
            const cuuint64_t rstr[3] = {(cuuint64_t)res_pitch * 2, (cuuint64_t)W * res_pitch * 2, (cuuint64_t)H * W * res_pitch * 2};
            if (!make_map_nd(&mr, (const __nv_bfloat16*)residual + res_off, 4, dims, rstr, box)) return TK_ERR_CUDA;
        } else {
            mr = md;
        }
    }
    if (!g_sms3) {
        int dev = 0;
        TK_CUDA_TRY(cudaGetDevice(&dev));
        TK_CUDA_TRY(cudaDeviceGetAttribute(&g_sms3, cudaDevAttrMultiProcessorCount, dev));
    }
    C3Params p;
    p.B = n_images; p.H = H; p.W = W; p.Cin = Cin; p.N = N;
    p.tw = tw; p.th = th; p.tiles_w = (W + tw - 1) / tw; p.tiles_h = (H + th - 1) / th;
    p.bk = bk; p.cblocks = Cin / bk;
    p.block_n = block_n; p.n_blocks = n_blocks; p.stages = stages; p.tmem_cols = tmem_cols; p.epi_warps = g_epi3; p.stg_bufs = stg_bufs;
    p.bias = bias; p.has_res = residual ? 1 : 0; p.act = act;
    const long long tiles = (long long)n_images * p.tiles_w * p.tiles_h * n_blocks;
    const long long slots = (long long)g_sms3 * ctas_per_sm;
    const int grid = (int)(tiles < slots ? tiles : slots);
    TK_CUDA_TRY(cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3x3_tc_kernel<<<grid, 128 + 32 * g_epi3, smem, (cudaStream_t)stream>>>(mx, mw, md, mr, p);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}
