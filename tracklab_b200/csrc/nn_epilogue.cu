// Channels-last (NHWC) bf16 epilogue kernels around the cuDNN convolutions of the detector / ReID backbones.
//
// The reference runs these networks in third-party runtimes (onnxruntime / torch) where every convolution is
// followed by separate bias, activation, concat, pooling and up-sampling passes over HBM
// (/root/reference/tracklab/wrappers/bbox_detector/rtmlib_api.py:19-30 -> rtmlib/onnxruntime;
//  /root/reference/plugins/track/strong_sort/deep/models/resnet.py:342-361). On B200 the convolutions themselves
// stay in cuDNN (tensor cores, per BASELINE.json north_star); everything between them is HBM-bound byte work and
// is collapsed here into single passes that also write straight into channel slices of the concat buffers,
// so torch.cat / separate SiLU / separate bias / max-pool / upsample kernels disappear:
//   tk_bias_act_nhwc   dst[..., off:off+C] = act(src + bias) (+ residual)      16-byte vector loads/stores
//   tk_spp_nhwc        dst = [x, pool5(x), pool9(x), pool13(x)] (SPP bottleneck) one read, four slices written
//   tk_upsample2x_nhwc dst[..., off:off+C] = nearest-2x(src)
#include <cuda_bf16.h>
#include "tk_common.cuh"
#include "trackkern.h"

namespace {

struct alignas(16) bf16x8 { __nv_bfloat162 v[4]; };

// x * sigmoid(x) = h * tanh(h) + h with h = x / 2: ONE MUFU op (tanh.approx) instead of two (ex2 + rcp). The pass moves
// 32 bytes per 8 activations and the SFUs retire 4 lanes per clock per sub-partition, so two MUFU ops per activation cap the
// kernel near 9 TB/s of issue alone; the result is rounded to bf16 (relative error of tanh.approx ~2^-11 is below that).
__device__ __forceinline__ float silu_f(float x) {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

// src [n_pix, C] contiguous; dst / res rows of pitch dst_pitch / res_pitch elements, channel offsets given.
// Grid-stride over 16-byte vectors, two vectors in flight per thread; (pixel, channel-vector) coordinates are advanced
// incrementally (stride = q * c_vec + r) so the loop has no integer division — the first version spent ~200 instructions
// per vector on a 64-bit divide and was issue-bound at 45 % of HBM peak (profiles/r01_ncu_summary.md).
__device__ __forceinline__ void epi_one(const __nv_bfloat16* __restrict__ src, const float4 b0, const float4 b1,
                                        __nv_bfloat16* __restrict__ dst, const __nv_bfloat16* __restrict__ res, long long i,
                                        long long pix, int cv, int dst_pitch, int dst_off, int res_pitch, int res_off, int act) {
    const bf16x8 x = *reinterpret_cast<const bf16x8*>(src + i * 8);
    float f[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 t = __bfloat1622float2(x.v[k]); f[2 * k] = t.x; f[2 * k + 1] = t.y; }
    f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
    if (act == 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = silu_f(f[k]);
    } else if (act == 2) {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.0f);
    }
    if (res) {
        const bf16x8 r = *reinterpret_cast<const bf16x8*>(res + pix * res_pitch + res_off + cv * 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float2 t = __bfloat1622float2(r.v[k]); f[2 * k] += t.x; f[2 * k + 1] += t.y; }
        if (act == 3) {   // residual first, then ReLU (ResNet bottleneck tail)
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.0f);
        }
    }
    bf16x8 o;
#pragma unroll
    for (int k = 0; k < 4; ++k) o.v[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
    *reinterpret_cast<bf16x8*>(dst + pix * dst_pitch + dst_off + cv * 8) = o;
}

__global__ void __launch_bounds__(256, 5)
bias_act_kernel(const __nv_bfloat16* __restrict__ src, const float* __restrict__ bias, __nv_bfloat16* __restrict__ dst,
                const __nv_bfloat16* __restrict__ res, long long n_vec, int c_vec, int dst_pitch, int dst_off,
                int res_pitch, int res_off, int act) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long sq = stride / c_vec;
    const int sr = (int)(stride - sq * c_vec);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long pix = i / c_vec;
    int cv = (int)(i - pix * c_vec);
    if (sr == 0) {
        // the host sizes the grid so that the stride is a multiple of the channel vectors: every thread keeps its 8 channels,
        // the bias lives in registers (two 16-byte bias loads per vector were half of the L1 requests: 87 % l1tex throughput in ncu)
        const float4 b0 = *reinterpret_cast<const float4*>(bias + cv * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(bias + cv * 8 + 4);
        for (; i + stride < n_vec; i += 2 * stride, pix += 2 * sq) {
            epi_one(src, b0, b1, dst, res, i, pix, cv, dst_pitch, dst_off, res_pitch, res_off, act);
            epi_one(src, b0, b1, dst, res, i + stride, pix + sq, cv, dst_pitch, dst_off, res_pitch, res_off, act);
        }
        if (i < n_vec) epi_one(src, b0, b1, dst, res, i, pix, cv, dst_pitch, dst_off, res_pitch, res_off, act);
        return;
    }
    auto bias_of = [&](int c, float4& b0, float4& b1) {
        b0 = *reinterpret_cast<const float4*>(bias + c * 8);
        b1 = *reinterpret_cast<const float4*>(bias + c * 8 + 4);
    };
    float4 b0, b1;
    for (; i + stride < n_vec; i += 2 * stride) {
        long long pix2 = pix + sq; int cv2 = cv + sr;
        if (cv2 >= c_vec) { cv2 -= c_vec; ++pix2; }
        bias_of(cv, b0, b1);
        epi_one(src, b0, b1, dst, res, i, pix, cv, dst_pitch, dst_off, res_pitch, res_off, act);
        bias_of(cv2, b0, b1);
        epi_one(src, b0, b1, dst, res, i + stride, pix2, cv2, dst_pitch, dst_off, res_pitch, res_off, act);
        pix = pix2 + sq; cv = cv2 + sr;
        if (cv >= c_vec) { cv -= c_vec; ++pix; }
    }
    if (i < n_vec) { bias_of(cv, b0, b1); epi_one(src, b0, b1, dst, res, i, pix, cv, dst_pitch, dst_off, res_pitch, res_off, act); }
}

// SPP: x [B, H, W, C] -> dst [B, H, W, 4C] = [x | max5 | max9 | max13] (stride 1, -inf padding).
// max9 = max5(max5), max13 = max5(max9) exactly. One CTA per (image, 8-channel group), tile in shared memory.
__global__ void __launch_bounds__(256)
spp_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ dst, int H, int W, int C, int dst_pitch, int dst_off) {
    extern __shared__ __align__(16) unsigned char smem[];
    bf16x8* a = reinterpret_cast<bf16x8*>(smem);   // [H*W]
    bf16x8* b = a + H * W;
    const int img = blockIdx.y, cg = blockIdx.x;
    const int n = H * W;
    const __nv_bfloat16* xi = x + (size_t)img * n * C + cg * 8;
    __nv_bfloat16* di = dst + (size_t)img * n * dst_pitch + dst_off + cg * 8;
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        const bf16x8 v = *reinterpret_cast<const bf16x8*>(xi + (size_t)p * C);
        a[p] = v;
        *reinterpret_cast<bf16x8*>(di + (size_t)p * dst_pitch) = v;
    }
    __syncthreads();
    bf16x8* in = a;
    bf16x8* out = b;
    for (int stage = 1; stage <= 3; ++stage) {
        for (int p = threadIdx.x; p < n; p += blockDim.x) {
            const int y = p / W, xx = p % W;
            bf16x8 m = in[p];
            for (int dy = -2; dy <= 2; ++dy) {
                const int yy = y + dy;
                if (yy < 0 || yy >= H) continue;
                for (int dx = -2; dx <= 2; ++dx) {
                    const int xc = xx + dx;
                    if (xc < 0 || xc >= W) continue;
                    const bf16x8 t = in[yy * W + xc];
#pragma unroll
                    for (int k = 0; k < 4; ++k) m.v[k] = __hmax2(m.v[k], t.v[k]);
                }
            }
            out[p] = m;
            *reinterpret_cast<bf16x8*>(di + (size_t)p * dst_pitch + (size_t)stage * C) = m;
        }
        __syncthreads();
        bf16x8* t = in; in = out; out = t;
    }
}

// nearest 2x: src [B, h, w, C] -> dst [B, 2h, 2w, dst_pitch] at channel offset
__global__ void __launch_bounds__(256)
upsample2x_kernel(const __nv_bfloat16* __restrict__ src, int src_pitch, int src_off, __nv_bfloat16* __restrict__ dst,
                  long long n_vec, int h, int w, int c_vec, int dst_pitch, int dst_off) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % c_vec);
        long long p = i / c_vec;            // output pixel index over [B, 2h, 2w]
        const int ox = (int)(p % (2 * w)); p /= (2 * w);
        const int oy = (int)(p % (2 * h));
        const long long bimg = p / (2 * h);
        const long long sp = (bimg * h + (oy >> 1)) * w + (ox >> 1);
        const bf16x8 v = *reinterpret_cast<const bf16x8*>(src + sp * src_pitch + src_off + cv * 8);
        const long long dp = (bimg * 2 * h + oy) * (2 * w) + ox;
        *reinterpret_cast<bf16x8*>(dst + dp * dst_pitch + dst_off + cv * 8) = v;
    }
}

// bias_act: exactly one resident wave (occupancy queried once: the 48-register kernel fits 5 CTAs per SM, the previous fixed
// 8 per SM ran 1.6 waves) and a grid whose thread count is a multiple of the channel vectors (register-resident bias).
inline int bias_act_grid(long long n_vec, int c_vec) {
    static int per_sm = 0, n_sm = 0;
    if (per_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bias_act_kernel, 256, 0) != cudaSuccess || per_sm <= 0) per_sm = 4;
        if (n_sm <= 0) n_sm = 148;
    }
    long long blocks = (n_vec + 255) / 256;
    const long long cap = (long long)n_sm * per_sm;
    if (blocks > cap) blocks = cap;
    int a = c_vec, b = 256;                       // g = c_vec / gcd(c_vec, 256): blocks * 256 % c_vec == 0 <=> blocks % g == 0
    while (b) { const int t = a % b; a = b; b = t; }
    const int g = c_vec / a;
    if (blocks >= g) blocks = blocks / g * g;
    return (int)(blocks > 0 ? blocks : 1);
}

inline int grid_for(long long n_vec) {
    long long blocks = (n_vec + 255) / 256;
    const long long cap = 148LL * 8;    // 8 resident CTAs of 256 threads per SM (one full wave), grid-stride beyond that
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// 3x3 / stride 2 / pad 1 max pooling, NHWC: one thread per (output pixel, 8-channel vector); the nine 16-byte loads of
// neighbouring outputs overlap in L1/L2, HBM sees the input once.
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int H, int W, int C, int Ho, int Wo, long long n_vec) {
    const int cv_n = C >> 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % cv_n);
        long long pix = i / cv_n;
        const int xo = (int)(pix % Wo); pix /= Wo;
        const int yo = (int)(pix % Ho);
        const long long img = pix / Ho;
        const __nv_bfloat16* base = src + (size_t)img * H * W * C + cv * 8;
        bf16x8 m;
        bool first = true;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = 2 * yo + dy;
            if (y < 0 || y >= H) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int x = 2 * xo + dx;
                if (x < 0 || x >= W) continue;
                const bf16x8 t = *reinterpret_cast<const bf16x8*>(base + ((size_t)y * W + x) * C);
                if (first) { m = t; first = false; }
                else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) m.v[k] = __hmax2(m.v[k], t.v[k]);
                }
            }
        }
        *reinterpret_cast<bf16x8*>(dst + (((size_t)img * Ho + yo) * Wo + xo) * C + cv * 8) = m;
    }
}

// 3x3 / stride 2 / pad 1 max pooling, separable form: one CTA per (image, output row). Pass 1: every (input column, 8-channel
// vector) item takes the max over the <= 3 input rows (three fully coalesced 16-byte loads per item: a row of W pixels x C channels
// is one contiguous run) into shared memory; pass 2: every output item takes the max of <= 3 neighbouring column maxima. 6 loads
// per output instead of 9, all coalesced (the thread-per-output kernel above sat at 2.4 TB/s = 0.37 of the measured HBM peak on the
// ReID stem, ncu r02_forward_metrics).
__global__ void __launch_bounds__(256)
maxpool3x3s2_rows_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int H, int W, int C, int Ho, int Wo) {
    extern __shared__ __align__(16) unsigned char mp_smem[];
    bf16x8* colmax = reinterpret_cast<bf16x8*>(mp_smem);
    const int cv_n = C >> 3, yo = blockIdx.x, img = blockIdx.y;
    const int items = W * cv_n;
    const __nv_bfloat16* base = src + (size_t)img * H * W * C;
    const int y0 = 2 * yo - 1;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        bf16x8 m;
        bool first = true;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int y = y0 + dy;
            if (y < 0 || y >= H) continue;
            const bf16x8 t = *reinterpret_cast<const bf16x8*>(base + (size_t)y * W * C + (size_t)it * 8);
            if (first) { m = t; first = false; }
            else {
#pragma unroll
                for (int k = 0; k < 4; ++k) m.v[k] = __hmax2(m.v[k], t.v[k]);
            }
        }
        colmax[it] = m;
    }
    __syncthreads();
    __nv_bfloat16* out = dst + ((size_t)img * Ho + yo) * Wo * C;
    for (int it = threadIdx.x; it < Wo * cv_n; it += blockDim.x) {
        const int xo = it / cv_n, cv = it - xo * cv_n;
        const int x0 = 2 * xo - 1;
        bf16x8 m;
        bool first = true;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int x = x0 + dx;
            if (x < 0 || x >= W) continue;
            const bf16x8 t = colmax[x * cv_n + cv];
            if (first) { m = t; first = false; }
            else {
#pragma unroll
                for (int k = 0; k < 4; ++k) m.v[k] = __hmax2(m.v[k], t.v[k]);
            }
        }
        *reinterpret_cast<bf16x8*>(out + (size_t)it * 8) = m;
    }
}

// global average pool: one CTA per image; thread = (8-channel vector, slice of the positions): every position is one contiguous
// run of C channels, so the 16-byte loads of a CTA are fully coalesced (the warp-per-vector form strode 2 C bytes between lanes and
// reached 0.67 TB/s). float32 accumulation in position order within a slice, slices combined in order.
__global__ void __launch_bounds__(256)
avgpool_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int n, int HW, int C) {
    __shared__ float part[256 * 8];
    const int cv_n = C >> 3, img = blockIdx.x;
    const int parts = max(1, (int)blockDim.x / cv_n);           // position slices handled in parallel when C/8 < 256
    const __nv_bfloat16* base = src + (size_t)img * HW * C;
    for (int cv0 = 0; cv0 < cv_n; cv0 += blockDim.x) {
        const int cv = cv0 + (int)threadIdx.x % min(cv_n, (int)blockDim.x);
        const int pt = (int)threadIdx.x / min(cv_n, (int)blockDim.x);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (cv < cv_n && pt < parts) {
            for (int p = pt; p < HW; p += parts) {
                const bf16x8 t = *reinterpret_cast<const bf16x8*>(base + (size_t)p * C + cv * 8);
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(t.v[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
            }
        }
        if (parts > 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) part[threadIdx.x * 8 + k] = acc[k];
            __syncthreads();
            if (pt == 0 && cv < cv_n) {
                for (int q = 1; q < parts; ++q)
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] += part[(q * cv_n + (cv - cv0)) * 8 + k];
            }
        }
        if (pt == 0 && cv < cv_n) {
            float4* o = reinterpret_cast<float4*>(dst + (size_t)img * C + cv * 8);
            o[0] = make_float4(acc[0] / (float)HW, acc[1] / (float)HW, acc[2] / (float)HW, acc[3] / (float)HW);
            o[1] = make_float4(acc[4] / (float)HW, acc[5] / (float)HW, acc[6] / (float)HW, acc[7] / (float)HW);
        }
        if (parts > 1) __syncthreads();
    }
}

}  // namespace

extern "C" {

int tk_bias_act_nhwc(const void* src, const float* bias, void* dst, const void* residual, long long n_pixels, int channels,
                     int dst_pitch, int dst_offset, int res_pitch, int res_offset, int act, void* stream) {
    if (!src || !bias || !dst || n_pixels <= 0 || channels <= 0) return TK_ERR_ARG;
    if ((channels & 7) || (dst_pitch & 7) || (dst_offset & 7) || (residual && ((res_pitch & 7) || (res_offset & 7)))) return TK_ERR_ARG;
    if (((size_t)src & 15) || ((size_t)dst & 15) || ((size_t)bias & 15) || (residual && ((size_t)residual & 15))) return TK_ERR_ARG;
    const long long n_vec = n_pixels * (channels / 8);
    bias_act_kernel<<<bias_act_grid(n_vec, channels / 8), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)src, bias, (__nv_bfloat16*)dst, (const __nv_bfloat16*)residual, n_vec, channels / 8, dst_pitch,
        dst_offset, res_pitch, res_offset, act);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_spp_nhwc(const void* x, void* dst, int n_images, int H, int W, int channels, int dst_pitch, int dst_offset, void* stream) {
    if (!x || !dst || n_images <= 0 || H <= 0 || W <= 0 || (channels & 7) || (dst_pitch & 7) || (dst_offset & 7)) return TK_ERR_ARG;
    const size_t smem = 2 * (size_t)H * W * 16;
    if (smem > 200 * 1024) return TK_ERR_CAPACITY;
    TK_CUDA_TRY(cudaFuncSetAttribute(spp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(channels / 8, n_images);
    spp_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)dst, H, W, channels, dst_pitch,
                                                          dst_offset);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_upsample2x_nhwc(const void* src, int src_pitch, int src_offset, void* dst, int n_images, int h, int w, int channels,
                       int dst_pitch, int dst_offset, void* stream) {
    if (!src || !dst || n_images <= 0 || h <= 0 || w <= 0) return TK_ERR_ARG;
    if ((channels & 7) || (dst_pitch & 7) || (dst_offset & 7) || (src_pitch & 7) || (src_offset & 7)) return TK_ERR_ARG;
    const long long n_vec = (long long)n_images * 4 * h * w * (channels / 8);
    upsample2x_kernel<<<grid_for(n_vec), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, src_pitch, src_offset,
                                                                       (__nv_bfloat16*)dst, n_vec, h, w, channels / 8,
                                                                       dst_pitch, dst_offset);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_maxpool3x3s2_nhwc(const void* src, int n, int H, int W, int C, void* dst, void* stream) {
    if (!src || !dst || n <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7)) return TK_ERR_ARG;
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    const long long n_vec = (long long)n * Ho * Wo * (C / 8);
    const size_t smem = (size_t)W * (C / 8) * 16;
    if (smem <= 96 * 1024 && n <= 65535) {       // separable row kernel: column maxima of one output row in shared memory
        TK_CUDA_TRY(cudaFuncSetAttribute(maxpool3x3s2_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        maxpool3x3s2_rows_kernel<<<dim3(Ho, n), 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, H, W, C, Ho, Wo);
    } else {
        maxpool3x3s2_kernel<<<grid_for(n_vec), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, H, W, C, Ho, Wo, n_vec);
    }
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

int tk_avgpool_nhwc(const void* src, int n, int HW, int C, float* dst, void* stream) {
    if (!src || !dst || n <= 0 || HW <= 0 || C <= 0 || (C & 7)) return TK_ERR_ARG;
    if (((size_t)dst & 15) || ((size_t)src & 15)) return TK_ERR_ARG;
    avgpool_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, dst, n, HW, C);
    TK_CUDA_TRY(cudaGetLastError());
    return TK_OK;
}

}  // extern "C"
